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


def test_plain_c_example_against_the_oracle(tmp_path, orc):
    """examples/commit_dataset.c on the GPU(s) of the box: its dataset root, slot root and cell indices are the oracle's"""
    import re
    from test_abi_surface import build_c_example
    exe = build_c_example(tmp_path)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    roots = [orc.commit_fake_slot(12345 + 72 + 1001 * k, 256)[0] for k in range(5)]
    dset = orc.merkle_root(roots)
    assert int(re.search(r"dataSetRoot = (0x[0-9a-f]+)", res.stdout).group(1), 16) == dset
    assert int(re.search(r"slotRoot    = (0x[0-9a-f]+)", res.stdout).group(1), 16) == roots[3]
    idx = [int(v) for v in re.search(r"cell indices:((?: \d+)+)", res.stdout).group(1).split()]
    assert idx == [orc.cell_index(1234567, roots[3], 256, c) for c in range(1, 6)]
    assert "the dataset root: yes" in res.stdout
