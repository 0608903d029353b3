"""C++ unit tests of the host mirror of reference/nim/proof_input, compiled against the in-tree library and run on the GPU."""
import os
import subprocess

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_host_mirror_cpp_unit_tests(tmp_path):
    pkg = os.path.join(ROOT, "codex-storage-proofs-circuits_b200")
    exe = str(tmp_path / "test_host")
    subprocess.run(["g++", "-O1", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "host_cpp", "test_host.cpp"),
                    os.path.join(pkg, "host", "proof_input.cpp"), "-L" + pkg, "-lcodexcommit", "-lpthread", "-Wl,-rpath," + pkg], check=True)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "all checks passed" in res.stdout
