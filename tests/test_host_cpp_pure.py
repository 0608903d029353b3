"""The pure host logic of the C++ mirror of reference/nim/proof_input (decimal rendering, negative constants, 31-byte
chunking, log2 helpers, option parsing, the input.json layout) on CPU, under UBSan where the toolchain has it: the first
section of tests/host_cpp/test_host.cpp, which stops before it needs a GPU when built with -DCDX_TEST_PURE_ONLY."""
import os
import subprocess

import pytest

from conftest import ROOT


def test_host_mirror_pure_logic_under_ubsan(pkg, tmp_path):
    pkg_dir = os.path.join(ROOT, "codex-storage-proofs-circuits_b200")
    exe = str(tmp_path / "test_pure")
    base = ["g++", "-O1", "-g", "-std=c++17", "-DCDX_TEST_PURE_ONLY", "-o", exe, os.path.join(ROOT, "tests", "host_cpp", "test_host.cpp"),
            os.path.join(pkg_dir, "host", "proof_input.cpp"), "-L" + pkg_dir, "-lcodexcommit", "-lpthread", "-Wl,-rpath," + pkg_dir]
    res = subprocess.run(base[:4] + ["-fsanitize=undefined", "-fno-sanitize-recover=undefined"] + base[4:], capture_output=True, text=True)
    if res.returncode != 0:                                   # no UBSan runtime here: plain build
        subprocess.run(base, check=True)
    run = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert run.returncode == 0, run.stdout + run.stderr
    assert "pure logic): all checks passed" in run.stdout
