#!/usr/bin/env python3
"""BASELINE config 5: a dataset of 256 (or 250: odd nodes in the dataset tree) synthetic slots of mixed size, committed
across the GPUs of a box, dataset root + 100 sampled Merkle paths of one slot.  One JSON line on stdout.

  python tools/dataset_commit.py --nslots 256 --scale 0.01                       (1 GPU, sizes 10 MiB .. 1 GiB)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/dataset_commit.py --nslots 256
"""
import argparse, importlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "codex-storage-proofs-circuits_b200"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nslots", type=int, default=256)
    ap.add_argument("--min-gib", type=float, default=1.0)
    ap.add_argument("--max-gib", type=float, default=100.0)
    ap.add_argument("--scale", type=float, default=1.0, help="multiply both size bounds (use < 1 for a quick run)")
    ap.add_argument("--seed", type=int, default=12345)
    ap.add_argument("--entropy", type=int, default=1234567)
    ap.add_argument("--nsamples", type=int, default=100)
    ap.add_argument("--sampled-slot", type=int, default=3)
    ap.add_argument("--sampled-log2-cells", type=int, default=22, help="the sampled slot has 2^k cells (8 GiB at k = 22), scaled with --scale")
    args = ap.parse_args()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    pkg = importlib.import_module(PKG)
    dataset = importlib.import_module(PKG + ".dataset")
    rank, world, local_rank = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = pkg.Context(local_rank)
    gib = float(1 << 30)
    k = args.sampled_log2_cells
    while args.scale < 1.0 and (1 << k) * 2048 > args.max_gib * args.scale * gib and k > 5:
        k -= 1
    blocks = dataset.draw_slot_blocks(args.nslots, args.min_gib * args.scale * gib, args.max_gib * args.scale * gib, args.seed,
                                      pow2_slot=args.sampled_slot, pow2_blocks=(1 << k) // 32)
    sharded = importlib.import_module(PKG + ".sharded")
    comm = sharded.comm_from_torch(ctx) if world > 1 else None       # the library's NCCL communicator; torch only carries the id
    descs = dataset.synthetic_descs(blocks, args.seed)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ds = ctx.dataset_commit(comm, descs, keep_slot=args.sampled_slot)                      # cdx_dataset_commit: everything inside the library
    t_commit = time.perf_counter() - t0
    t1 = time.perf_counter()
    idx, paths, leaves = ds.prove(args.entropy, args.nsamples, 32)                           # cdx_dataset_prove (collective)
    t_prove = time.perf_counter() - t1
    wall = time.perf_counter() - t0
    stats = ds.stats
    tt = torch.tensor([t_commit, t_prove, float(stats["bytes_local"])], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    # self-check on rank 0 (GPU verifier kernel): every sampled path reconstructs to the slot root in two stages, and
    # the slot proof reconstructs to the dataset root (merkle.nim:51-74)
    if rank == 0:
        nb = blocks[args.sampled_slot]
        depth = max(1, (nb - 1).bit_length())
        roots = ds.slot_roots
        sroot = roots[args.sampled_slot]
        blk = ctx.reconstruct_roots(leaves, [ci % 32 for ci in idx], 32, paths, depth=5)
        top = ctx.reconstruct_roots(blk, [ci // 32 for ci in idx], nb, [p[5:] for p in paths], depth=depth)
        ok = all(t == sroot for t in top)
        dd = max(1, (args.nslots - 1).bit_length())
        sp = ds.slot_proof(args.sampled_slot, 8)
        ok &= ctx.reconstruct_roots([sroot], [args.sampled_slot], args.nslots, [sp], depth=dd)[0] == ds.root
        total = sum(blocks) * 65536
        perms = sum(b * 1119 for b in blocks)
        line = {"workload": f"dataset of {args.nslots} synthetic slots, sizes log-uniform {args.min_gib * args.scale:g}-{args.max_gib * args.scale:g} GiB, "
                            f"{world} GPU(s); cdx_dataset_commit (LPT packing, batching, sharding inside the library) + cdx_dataset_prove: "
                            f"dataset root + {args.nsamples} sampled paths of slot {args.sampled_slot}",
                "n_gpus": world, "bytes": total, "commit_s": float(tt[0]), "prove_s": float(tt[1]),
                "wall_s": wall, "GB_per_s": total / float(tt[0]) / 1e9, "perms_per_s": perms / float(tt[0]),
                "balance": float(tt[2]) / (total / world), "dataset_root": hex(ds.root), "rank0_stats": stats,
                "sampled_slot_cells": blocks[args.sampled_slot] * 32, "first_indices": idx[:5], "paths_self_check": bool(ok)}
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    ds.free()
    if comm is not None:
        comm.destroy()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
