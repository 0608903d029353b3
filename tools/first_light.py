#!/usr/bin/env python3
"""First-light check on a B200: smoke parity, integer-multiply probe, and rough kernel rates (not the bench)."""
import importlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as g

pkg = importlib.import_module(g.PKG)
g.smoke()
ctx = pkg.Context(0)
res = {}
for kind, name in ((0, "imad_wide"), (1, "imad_wide_x_chain"), (2, "imad32")):
    ops, ms = ctx.probe_imad_rate(kind)
    res[name] = {"Tops": ops / 1e12, "ms": ms}
print(json.dumps(res))

def timed(fn, stream, reps=3):
    best = None
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); fn(); e1.record(stream); e1.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None or ms < best else best
    return best

st = torch.cuda.Stream()
with torch.cuda.stream(st):
    n = 1 << 20
    a = torch.empty(96 * n, dtype=torch.uint8, device="cuda")
    b = torch.empty_like(a)
    ctx.fill_synthetic_dev(1, 0, 96 * n, a.data_ptr(), st.cuda_stream)
    # mask top bits so inputs are < 2^254 (still possibly >= r: taken mod r)
    ms = timed(lambda: ctx.permutation_batch_dev(a.data_ptr(), b.data_ptr(), n, st.cuda_stream), st)
    print(json.dumps({"perm_batch_2^20_ms": ms, "Mperm_per_s": n / ms / 1e3}))
    for gib in (0.25, 1, 4):
        nbytes = int(gib * (1 << 30))
        d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        ctx.fill_synthetic_dev(0xC0DE, 0, nbytes, d.data_ptr(), st.cuda_stream)
        slot = [None]
        def run():
            if slot[0] is not None: slot[0].free()
            slot[0] = ctx.slot_commit_dev(d.data_ptr(), nbytes, 2048, 65536, st.cuda_stream)
        ms = timed(run, st, reps=2)
        perms = (nbytes // 65536) * 1120
        print(json.dumps({"slot_GiB": gib, "ms": ms, "GB_per_s": nbytes / ms / 1e6, "Mperm_per_s": perms / ms / 1e3, "root": hex(slot[0].root)[:18]}))
        slot[0].free(); del d
