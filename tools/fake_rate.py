#!/usr/bin/env python3
"""Rate of the reference's fake-data generator on the device (k_fake_cells) and of a commit that generates its bytes tile by
tile (cdx_slot_commit_fake) against the same slot committed resident; 1 GiB.  Not part of bench.py."""
import importlib, sys, time, os
sys.path.insert(0, os.getcwd())
import torch
pkg = importlib.import_module("codex-storage-proofs-circuits_b200")
ctx = pkg.Context(0)
n_cells = 1 << 19   # 1 GiB
d = torch.empty(n_cells * 2048, dtype=torch.uint8, device="cuda")
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ctx.fake_cells_dev(123, 0, n_cells, 2048, d.data_ptr()); torch.cuda.synchronize()
    print("k_fake_cells 1 GiB: %.1f ms" % (1e3 * (time.perf_counter() - t0)))
for rep in range(2):
    t0 = time.perf_counter()
    with ctx.slot_commit_fake(123, n_cells) as s: r = s.root
    print("slot_commit_fake 1 GiB: %.1f ms" % (1e3 * (time.perf_counter() - t0)))
    t0 = time.perf_counter()
    with ctx.slot_commit_dev(d.data_ptr(), n_cells * 2048) as s: assert s.root == r
    print("slot_commit_dev 1 GiB: %.1f ms" % (1e3 * (time.perf_counter() - t0)))
