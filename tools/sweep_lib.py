#!/usr/bin/env python3
"""time k_hash_cells alone for alternative builds of the library (occupancy / variant sweeps); not part of the bench.
Prints a checksum over ALL cell hashes so that variants can be compared bit for bit."""
import ctypes as C, importlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
pkg = importlib.import_module("codex-storage-proofs-circuits_b200")
capi = pkg.capi
n_bytes = int(os.environ.get("SWEEP_GIB", "2")) << 30
n_cells = n_bytes // 2048
reps = int(os.environ.get("SWEEP_REPS", "3"))
for path in sys.argv[1:]:
    capi._lib = capi.load_library(os.path.join(ROOT, path))
    ctx = pkg.Context(0)
    st = torch.cuda.ExternalStream(ctx.stream)
    d = torch.empty(n_bytes, dtype=torch.uint8, device="cuda")
    out = torch.empty(n_cells * 32, dtype=torch.uint8, device="cuda")
    ctx.fill_synthetic_dev(0xC0DE, 0, n_bytes, d.data_ptr())
    ctx.hash_cells_dev(d.data_ptr(), n_cells, 2048, out.data_ptr())
    torch.cuda.synchronize()
    best = None
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st); ctx.hash_cells_dev(d.data_ptr(), n_cells, 2048, out.data_ptr()); e1.record(st); e1.synchronize()
        ms = e0.elapsed_time(e1); best = ms if best is None or ms < best else best
    w = out.view(torch.int64)
    chk = int(w.sum().item()) & (2**64 - 1)
    chk2 = int((w * torch.arange(1, w.numel() + 1, device="cuda", dtype=torch.int64)).sum().item()) & (2**64 - 1)
    print(json.dumps({"lib": path, "ms": best, "GB_per_s": n_bytes / best / 1e6, "Mperm_per_s": n_cells * 34 / best / 1e3,
                      "chk": hex(chk), "chk_weighted": hex(chk2)}), flush=True)
    ctx.close(); del d, out
