#!/usr/bin/env python3
"""Challenge-answering rate against one retained commitment; not part of bench.py.

Commits a power-of-two synthetic slot once, then answers K challenges of S samples each
  (a) one challenge at a time: cdx_cell_indices, then cdx_slot_cell_paths    (what a per-request server does)
  (b) all at once: cdx_slot_prove_batch
and verifies every answer of (b) with the batched two-stage verifier (cdx_reconstruct_roots_host).
usage: proof_server_bench.py [slot GiB = 8] [K = 1000] [S = 100]
"""
import ctypes as C
import importlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

pkg = importlib.import_module("codex-storage-proofs-circuits_b200")
capi = pkg.capi
ctx = pkg.Context(0)
lib = ctx.lib
gib = float(sys.argv[1]) if len(sys.argv) > 1 else 8.0
K = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
S = int(sys.argv[3]) if len(sys.argv) > 3 else 100
DEPTH = 32
n = int(gib * (1 << 30))
assert n & (n - 1) == 0, "slot size must be a power of two so that sampling can run (sample/bn254.nim:19-20)"
d = torch.empty(n, dtype=torch.uint8, device="cuda")
ctx.fill_synthetic_dev(0xC0DE, 0, n, d.data_ptr())
torch.cuda.synchronize()
slot = ctx.slot_commit_dev(d.data_ptr(), n)
root = slot.root
n_cells, n_blocks, bdepth, sdepth = slot.shape
del d
torch.cuda.empty_cache()

ent = b"".join(capi.f2b((0x9e3779b97f4a7c15 * (k + 1)) % (1 << 250)) for k in range(K))
rootb = capi.f2b(root)
idx = (C.c_uint64 * (K * S))()
paths = C.create_string_buffer(32 * K * S * DEPTH + 32 * DEPTH)   # slack: the slot stage passes the array shifted by bdepth elements
leaves = C.create_string_buffer(32 * K * S)


chk = ctx._chk


# (a) one challenge at a time
one_idx = (C.c_uint64 * S)()
one_paths = C.create_string_buffer(32 * S * DEPTH)
one_leaves = C.create_string_buffer(32 * S)
ent_addr = capi._addr(ent)
n_serial = min(K, 200)
for warm in range(2):
    t0 = time.perf_counter()
    for k in range(n_serial):
        chk(lib.cdx_cell_indices(ctx.h, ent_addr + 32 * k, capi._addr(rootb), n_cells, S, C.addressof(one_idx)))
        chk(lib.cdx_slot_cell_paths(slot.h, C.addressof(one_idx), S, DEPTH, C.addressof(one_paths), C.addressof(one_leaves)))
    t_serial = (time.perf_counter() - t0) / n_serial
last_serial = (list(one_idx), one_paths.raw, one_leaves.raw)

# (b) batched
t_batch = []
for rep in range(4):
    t0 = time.perf_counter()
    chk(lib.cdx_slot_prove_batch(slot.h, ent_addr, K, S, DEPTH, C.addressof(idx), C.addressof(paths), C.addressof(leaves)))
    t_batch.append(time.perf_counter() - t0)
k = n_serial - 1
assert list(idx)[k * S:(k + 1) * S] == last_serial[0]
assert paths.raw[32 * DEPTH * S * k:32 * DEPTH * S * (k + 1)] == last_serial[1][:32 * DEPTH * S]
assert leaves.raw[32 * S * k:32 * S * (k + 1)] == last_serial[2]

# verify every answer: block level (5 steps), then slot level (sdepth steps) -- Slot.hs:189-217
total = K * S
cpb = n_cells // n_blocks
in_block = (C.c_uint64 * total)(*[i % cpb for i in idx])
in_slot = (C.c_uint64 * total)(*[i // cpb for i in idx])
block_roots = C.create_string_buffer(32 * total)
slot_roots = C.create_string_buffer(32 * total)
t0 = time.perf_counter()
chk(lib.cdx_reconstruct_roots_host(ctx.h, C.addressof(leaves), C.addressof(in_block), cpb, C.addressof(paths), DEPTH, bdepth, total,
                                   C.addressof(block_roots)))
chk(lib.cdx_reconstruct_roots_host(ctx.h, C.addressof(block_roots), C.addressof(in_slot), n_blocks, C.addressof(paths) + 32 * bdepth, DEPTH, sdepth,
                                   total, C.addressof(slot_roots)))
t_verify = time.perf_counter() - t0
assert slot_roots.raw == rootb * total, "a batched answer does not reconstruct the slot root"
slot.free()

best = min(t_batch[1:])
print(json.dumps({
    "slot_gib": gib, "n_cells": n_cells, "path_len": bdepth + sdepth, "challenges": K, "samples_per_challenge": S,
    "one_at_a_time": {"ms_per_challenge": 1e3 * t_serial, "challenges_per_s": 1.0 / t_serial, "calls": "cdx_cell_indices + cdx_slot_cell_paths"},
    "batched": {"ms_total": 1e3 * best, "challenges_per_s": K / best, "paths_per_s": total / best, "calls": "cdx_slot_prove_batch",
                "all_ms": [1e3 * t for t in t_batch]},
    "verify": {"ms_total": 1e3 * t_verify, "proofs_per_s": total / t_verify, "perms": total * (bdepth + sdepth),
               "calls": "2 x cdx_reconstruct_roots_host (block stage, slot stage)"},
    "all_answers_reconstruct_the_root": True,
}))
