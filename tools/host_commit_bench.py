#!/usr/bin/env python3
"""cdx_slot_commit_host from pinned vs pageable host memory (what a Nim seq[byte] is); not part of bench.py
usage: host_commit_bench.py [GiB = 4]"""
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
with ctx.slot_commit_dev(d.data_ptr(), n) as s:
    root = s.root
pageable = d.cpu()
pinned = torch.empty(n, dtype=torch.uint8, pin_memory=True)
pinned.copy_(pageable)
del d
res = {}
for name, buf in (("pinned", pinned), ("pageable", pageable)):
    rates = []
    for rep in range(4):
        t0 = time.perf_counter()
        with ctx.slot_commit_host(buf.data_ptr(), n_bytes=n) as s:
            dt = time.perf_counter() - t0
            assert s.root == root
        rates.append(n / dt / 1e9)
    res[name] = rates
# raw copy rates for context
dd = torch.empty(n, dtype=torch.uint8, device="cuda")
for name, buf in (("pinned", pinned), ("pageable", pageable)):
    torch.cuda.synchronize(); t0 = time.perf_counter(); dd.copy_(buf); torch.cuda.synchronize()
    res["h2d_copy_only_" + name] = n / (time.perf_counter() - t0) / 1e9
print(json.dumps({"gib": gib, "GB_per_s": res}))
