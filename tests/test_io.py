"""CPU tests of the CLI's file layer: PNG codec against cv2, caffemodel writer/reader wire format, usage/exit codes."""
import os
import subprocess

import numpy as np
import pytest

from oracle import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "neural-color-transfer_b200", "neural_color_transfer")


def test_png_write_is_read_back_by_cv2_and_vice_versa(pkg, tmp_path):
    import cv2

    img, _ = synth.pair(0, 61, 47)
    p1 = str(tmp_path / "ours.png")
    pkg.png_write(p1, img)
    assert np.array_equal(cv2.imread(p1), img)          # cv2 decodes our file
    assert np.array_equal(pkg.png_read(p1), img)        # and so do we
    for params in ([], [cv2.IMWRITE_PNG_COMPRESSION, 9], [cv2.IMWRITE_PNG_STRATEGY, cv2.IMWRITE_PNG_STRATEGY_FILTERED]):
        p2 = str(tmp_path / "cv.png")
        cv2.imwrite(p2, img, params)                    # libpng picks adaptive filters: all five filter types occur
        assert np.array_equal(pkg.png_read(p2), img)


def test_png_alpha_grey_and_16bit_match_imread(pkg, tmp_path):
    import cv2

    rng = np.random.default_rng(0)
    rgba = rng.integers(0, 256, (20, 31, 4), dtype=np.uint8)
    grey = rng.integers(0, 256, (17, 19), dtype=np.uint8)
    for name, arr in (("rgba.png", rgba), ("grey.png", grey)):
        p = str(tmp_path / name)
        cv2.imwrite(p, arr)
        assert np.array_equal(pkg.png_read(p), cv2.imread(p))  # imread drops alpha / replicates grey (NCT/main.cu:483)


def _png_bytes(width, height, depth, ctype, rows):
    import struct
    import zlib

    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d))
    raw = b"".join(b"\x00" + bytes(r) for r in rows)
    return (b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", width, height, depth, ctype, 0, 0, 0)) +
            chunk(b"IDAT", zlib.compress(raw)) + chunk(b"IEND", b""))


@pytest.mark.parametrize("depth", [1, 2, 4])
def test_png_low_bit_depth_greyscale(pkg, tmp_path, depth):
    """colour type 0 at 1 / 2 / 4 bits per sample (valid PNGs that imread accepts): samples scaled to 8 bits"""
    import cv2

    rng = np.random.default_rng(depth)
    W, H = 13, 7
    vals = rng.integers(0, 1 << depth, (H, W))
    rows = []
    for y in range(H):
        bits = "".join(format(int(v), f"0{depth}b") for v in vals[y])
        bits += "0" * (-len(bits) % 8)
        rows.append([int(bits[i:i + 8], 2) for i in range(0, len(bits), 8)])
    p = str(tmp_path / f"g{depth}.png")
    open(p, "wb").write(_png_bytes(W, H, depth, 0, rows))
    want = (vals * 255 // ((1 << depth) - 1)).astype(np.uint8)
    got = pkg.png_read(p)
    assert np.array_equal(got, np.repeat(want[..., None], 3, axis=2))
    assert np.array_equal(got, cv2.imread(p))


def test_png_reads_the_reference_demo_inputs(pkg):
    import cv2

    demo = "/root/reference/demo/example/in"
    if not os.path.isdir(demo):
        pytest.skip("reference tree not present on this box")
    for fn in sorted(os.listdir(demo)):
        if fn.endswith(".png"):
            assert np.array_equal(pkg.png_read(os.path.join(demo, fn)), cv2.imread(os.path.join(demo, fn))), fn


def test_png_read_missing_file_raises(pkg, tmp_path):
    with pytest.raises(pkg.NctError):
        pkg.png_read(str(tmp_path / "nope.png"))


def test_cli_usage_and_exit_codes():
    # -h / -? / -help print the list and exit -1; unknown flags too (NCT/main.cu:556-560, NCT/CmdLine.cpp:21-57)
    for flag in ("-h", "/?", "-help"):
        r = subprocess.run([CLI, flag], capture_output=True, text=True)
        assert r.returncode == 255 and "Running:" in r.stdout and "-bds" in r.stdout
    r = subprocess.run([CLI, "-nosuchflag", "1"], capture_output=True, text=True)
    assert r.returncode == 255 and "Unrecognized parameter: -nosuchflag" in r.stdout


def test_cli_fails_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([CLI, "-m", "/tmp", "-i", "/tmp", "-o", "/tmp/out", "-g", "0"], capture_output=True, text=True)
    assert r.returncode == 1 and "no CPU path" in r.stderr
