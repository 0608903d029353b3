#!/usr/bin/env python3
"""Workload for the `ncu --set full` capture: ONE launch of the cell-sponge kernel over a 1 GiB synthetic slot (after a
warm-up launch, which ncu is told to skip with --launch-skip).  Exits 0 without ncu as well."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
pkg = importlib.import_module("codex-storage-proofs-circuits_b200")
ctx = pkg.Context(0)
n_bytes = 1 << 30
d = torch.empty(n_bytes, dtype=torch.uint8, device="cuda")
out = torch.empty(n_bytes // 2048 * 32, dtype=torch.uint8, device="cuda")
ctx.fill_synthetic_dev(0xC0DE, 0, n_bytes, d.data_ptr())
for _ in range(2):
    ctx.hash_cells_dev(d.data_ptr(), n_bytes // 2048, 2048, out.data_ptr())
torch.cuda.synchronize()
print("ok", int(out[:8].cpu().numpy().view("uint64")[0]))
