#!/usr/bin/env python3
"""Host-buffer (e2e) commit rate per rank for different staging settings, with every rank streaming at once; run under
torchrun with N ranks (or alone).  Also measures the raw pinned H2D rate per rank under the same contention.
   CODEX_COMMIT_STAGE_TILES (2..4) x CODEX_COMMIT_TILE_MIB x CODEX_COMMIT_RAMP (SWEEP_SETTINGS=3x256x1,3x256x0,...) are read when a context is created, so one process can try
   several.  Prints one JSON line per setting on rank 0: min / mean / max over ranks."""
import importlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
pkg = importlib.import_module("codex-storage-proofs-circuits_b200")
rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
gib = float(os.environ.get("SWEEP_GIB", "10"))
n_bytes = int(gib * (1 << 30)) // 65536 * 65536
h = torch.empty(n_bytes, dtype=torch.uint8, pin_memory=True)
d = torch.empty(n_bytes, dtype=torch.uint8, device="cuda")
c0 = pkg.Context(local)
c0.fill_synthetic_dev(0xC0DE, 0, n_bytes, d.data_ptr())
torch.cuda.synchronize()
h.copy_(d)
torch.cuda.synchronize()
hn = h.numpy()

def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()

def reduce3(v):
    t = torch.tensor([v, -v, v], dtype=torch.float64, device="cuda")
    if world > 1:
        a = t.clone(); dist.all_reduce(a, op=dist.ReduceOp.MAX)
        s = t.clone(); dist.all_reduce(s, op=dist.ReduceOp.SUM)
        return -float(a[1]), float(s[2]) / world, float(a[0])
    return v, v, v

# raw H2D under contention
barrier()
t0 = time.perf_counter()
for _ in range(2):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
mn, mean, mx = reduce3(2 * n_bytes / dt / 1e9)
if rank == 0:
    print(json.dumps({"what": "raw pinned H2D, all ranks at once", "ranks": world, "GB_per_s_min": mn, "mean": mean, "max": mx}), flush=True)
# resident rate for reference
with c0.slot_commit_dev(d.data_ptr(), n_bytes) as s:
    root = s.root
barrier()
t0 = time.perf_counter()
for _ in range(2):
    with c0.slot_commit_dev(d.data_ptr(), n_bytes) as s:
        assert s.root == root
dt = time.perf_counter() - t0
mn, mean, mx = reduce3(2 * n_bytes / dt / 1e9)
if rank == 0:
    print(json.dumps({"what": "resident commit", "ranks": world, "GB_per_s_min": mn, "mean": mean, "max": mx}), flush=True)
del d
c0.close()
settings = [s.split("x") for s in os.environ.get("SWEEP_SETTINGS", "2x256,3x256,4x256,3x128,3x512,2x512").split(",")]
for st in settings:
    tiles, mib = st[0], st[1]
    ramp = st[2] if len(st) > 2 else "1"
    os.environ["CODEX_COMMIT_STAGE_TILES"], os.environ["CODEX_COMMIT_TILE_MIB"], os.environ["CODEX_COMMIT_RAMP"] = tiles, mib, ramp
    ctx = pkg.Context(local)
    with ctx.slot_commit_host(hn) as s:
        assert s.root == root
    barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        with ctx.slot_commit_host(hn) as s:
            r = s.root
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    barrier()
    assert r == root
    mn, mean, mx = reduce3(3 * n_bytes / dt / 1e9)
    if rank == 0:
        print(json.dumps({"what": "e2e commit from pinned host memory", "stage_tiles": int(tiles), "tile_mib": int(mib), "ramp_mode": int(ramp), "ranks": world,
                          "GB_per_s_min": mn, "mean": mean, "max": mx, "aggregate_at_min": mn * world}), flush=True)
    ctx.close()
if world > 1:
    dist.destroy_process_group()
