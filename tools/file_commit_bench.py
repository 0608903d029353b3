#!/usr/bin/env python3
"""cdx_slot_commit_file throughput from the page cache; not part of bench.py"""
import importlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
pkg = importlib.import_module("codex-storage-proofs-circuits_b200")
ctx = pkg.Context(0)
gib = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
n = int(gib * (1 << 30)) // 65536 * 65536
d = torch.empty(n, dtype=torch.uint8, device="cuda")
ctx.fill_synthetic_dev(0xC0DE, 0, n, d.data_ptr()); torch.cuda.synchronize()
path = (sys.argv[2] if len(sys.argv) > 2 else "/dev/shm") + "/cdx_slot.dat"
d.cpu().numpy().tofile(path)
with ctx.slot_commit_dev(d.data_ptr(), n) as s:
    root = s.root
out = []
for rep in range(3):
    t0 = time.perf_counter()
    with ctx.slot_commit_file(path, n) as s:
        dt = time.perf_counter() - t0
        assert s.root == root
    out.append(n / dt / 1e9)
os.remove(path)
print(json.dumps({"file_gib": gib, "where": path, "GB_per_s": out}))
