"""Regression guard on the compiled hot loops (no GPU needed: nvcc cross-compiles, cuobjdump disassembles).
The cell kernel's time follows the number of FMA-heavy-pipe slots per S-box; this test pins the
structure the measurements in profiles/ were taken with: every 32x32->64 product is a fused IMAD.WIDE/IMAD.HI, 328 of
them per S-box (36 + 36 + 64 operand products, 3 x 64 reduction products), and almost nothing else on that pipe; and the
number of non-multiply instructions that the table-driven reduction brought down."""
import os
import re
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("nvcc") is None or shutil.which("cuobjdump") is None, reason="needs the CUDA toolchain")
def test_heavy_pipe_slots_per_sbox():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_stats.py"), "k_hash_cells_tma", "--max-wide", "1200"],
                         capture_output=True, text=True, check=True).stdout
    loops = [tuple(int(v) for v in m) for m in re.findall(r"wide products (\d+), other FMA-pipe (\d+), H = (\d+), A = (\d+)", out)]
    assert loops, out
    bodies = {wide // 328: (wide, other, h, a) for wide, other, h, a in loops if wide % 328 == 0}
    assert 1 in bodies and 3 in bodies, out                  # the internal-round loop and the three-way external-round body
    for n_sbox, (wide, other, h, a) in bodies.items():
        assert wide == 328 * n_sbox                          # no unfused mad.lo/mad.hi pairs, no extra products
        assert h <= 700 * n_sbox, out                        # 680 is the floor for an 8 x 32-bit CIOS S-box
        # everything that is not a multiply: 239 per internal round / 174 per S-box of the external body with the table-driven
        # reduction of round 2 (320 / 221 with round 1's chained conditional subtractions)
        assert a <= 260 * n_sbox, out
    sel = [int(m) for m in re.findall(r"\bSEL (\d+)", out)]
    assert sel and max(sel) <= 45, out                       # no select chains left: the remaining SELs materialise carries
