"""CPU checks of the drop-in boundary: libnct.so loads and exports exactly what include/nct.h declares."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "nct.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nct_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_symbols():
    syms = header_symbols()
    assert "nct_create" in syms and "nct_patchmatch" in syms and len(syms) >= 15


def test_library_exports_every_declared_symbol(pkg):
    lib = pkg.load_library()
    for s in header_symbols():
        assert hasattr(lib, s), f"libnct.so does not export {s}"


def test_binding_table_matches_header(pkg):
    assert sorted(pkg.ABI.keys()) == header_symbols()


def test_exports_have_c_linkage():
    out = subprocess.check_output(["nm", "-D", "--defined-only", pkg_lib()]).decode()
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    for s in header_symbols():
        assert s in exported


def pkg_lib():
    return os.path.join(ROOT, "neural-color-transfer_b200", "libnct.so")


def test_no_torch_or_oracle_in_product():
    """The product library must not depend on torch, and nothing under the package may touch oracle/."""
    out = subprocess.check_output(["ldd", pkg_lib()]).decode()
    assert "torch" not in out and "c10" not in out
    pkg_dir = os.path.join(ROOT, "neural-color-transfer_b200")
    for dp, _, fns in os.walk(pkg_dir):
        if "build" in dp.split(os.sep):
            continue
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".cc")):
                txt = open(os.path.join(dp, fn), errors="ignore").read()
                for line in txt.splitlines():
                    s = line.strip()
                    if s.startswith(("//", "#", "*", "/*")) and "include" not in s and "import" not in s:
                        continue
                    assert not re.search(r"(import\s+oracle|from\s+oracle|#include\s+[\"<].*oracle|liboracle)", s), (
                        f"{fn}: product code references the oracle: {s}"
                    )


def test_create_fails_loudly_without_gpu(pkg):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg.NctError):
        pkg.Context(0)


def test_level_sizes(pkg):
    # 700 -> 700,350,175,88,44 ; 1000 -> 1000,500,250,125,63
    assert pkg.level_sizes(700) == [700, 350, 175, 88, 44]
    assert pkg.level_sizes(1000) == [1000, 500, 250, 125, 63]
    assert pkg.level_sizes(512) == [512, 256, 128, 64, 32]
    assert pkg.level_sizes(256) == [256, 128, 64, 32, 16]
