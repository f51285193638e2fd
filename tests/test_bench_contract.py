"""CPU tests of bench.py's host logic: the reference arm (`--impl reference`, the oracle "port" timed on the host cores)
prints exactly one JSON line with the contract's keys, and under torchrun only rank 0 runs and prints it."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
        "dtype", "data", "config", "cpu_baseline", "e2e"}


def _json_lines(text):
    out = []
    for line in text.splitlines():
        line = line.strip()
        if line.startswith("{") and line.endswith("}"):
            out.append(json.loads(line))
    return out


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--cpu-sample-side", "64", "--side", "96", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = _json_lines(r.stdout)
    assert len(lines) == 1
    d = lines[0]
    assert KEYS <= set(d) and d["impl"] == "reference" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    sys.path.insert(0, ROOT)
    import bench

    # the reference arm runs on OUR arm's config: the same `config` object, the sample is described in cpu_baseline
    assert d["config"] == bench.workload_config(96, 6, 1, 3) and d["vs_baseline"] is None and d["gpu_launches"] == 0
    assert "64x64" in d["cpu_baseline"]["sample"]
    full = d["cpu_baseline"]["full_size_check"]   # one full-size pair, timed once after the steps
    assert full["value"] > 0 and full["seconds"] > 0


def test_reference_arm_under_torchrun_only_rank0_prints():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29631", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--cpu-sample-side", "64",
                        "--no-full-size-check", "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=900, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = _json_lines(r.stdout)
    assert len(lines) == 1 and lines[0]["impl"] == "reference" and lines[0]["n_gpus"] == 2


def test_peak_reader_accepts_the_schemas_seen_so_far(tmp_path, monkeypatch):
    sys.path.insert(0, ROOT)
    import bench

    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    assert bench.read_peaks()[0] == 6650.0  # fallback of the profiling guide when the driver's file is absent
    for doc, want in [({"hbm_gbs": 6547.2, "bf16_tflops": 1386.1}, 6547.2),
                      ({"hbm": {"copy_gbs_burst": 6700.0, "copy_gbs_sustained": 6547.2}}, 6547.2),
                      ({"hbm_tbs": 6.5}, 6500.0)]:
        (tmp_path / "MEASURED_PEAKS.json").write_text(json.dumps(doc))
        got, src = bench.read_peaks()
        assert got == want and src.startswith("measured")
