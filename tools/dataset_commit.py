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
    t0 = time.perf_counter()
    res = dataset.commit_dataset(ctx, blocks, args.seed, args.sampled_slot, args.entropy, args.nsamples, rank=rank, world=world)
    wall = time.perf_counter() - t0
    # self-check on rank 0 (GPU compression only): every sampled path reconstructs to the slot root in two stages, and
    # the slot proof reconstructs to the dataset root (merkle.nim:51-74)
    ok = None
    if rank == 0:
        def reconstruct(leaf, j, m, path):
            h, bottom = leaf, 1
            for p in path:
                if j & 1: h = ctx.compress(p, h, bottom)
                elif j == m - 1: h = ctx.compress(h, p, bottom + 2)
                else: h = ctx.compress(h, p, bottom)
                bottom, j, m = 0, j >> 1, (m + 1) >> 1
            return h
        nb = blocks[args.sampled_slot]
        depth = max(1, (nb - 1).bit_length())
        sroot = res.slot_roots[args.sampled_slot]
        ok = True
        for ci, path, leaf in list(zip(res.cell_indices, res.merkle_paths, res.cell_hashes))[:10]:
            blk = reconstruct(leaf, ci % 32, 32, path[:5])
            ok &= reconstruct(blk, ci // 32, nb, path[5:5 + depth]) == sroot
        dd = len(res.dataset_layers) - 1
        ok &= reconstruct(sroot, args.sampled_slot, args.nslots, res.slot_proof[:dd]) == res.dataset_root
        total = res.bytes_committed
        perms = sum(b * 1119 for b in blocks)
        line = {"workload": f"dataset of {args.nslots} synthetic slots, sizes log-uniform {args.min_gib * args.scale:g}-{args.max_gib * args.scale:g} GiB, "
                            f"{world} GPU(s), LPT bin packing; dataset root + {args.nsamples} sampled paths of slot {args.sampled_slot}",
                "n_gpus": world, "bytes": total, "commit_s": res.timings["commit_s"], "roots_tree_paths_s": res.timings["roots_tree_paths_s"],
                "wall_s": wall, "GB_per_s": total / res.timings["commit_s"] / 1e9, "perms_per_s": perms / res.timings["commit_s"],
                "balance": max(res.per_rank_bytes) / (total / world), "dataset_root": hex(res.dataset_root),
                "dataset_tree_layers": [len(l) for l in res.dataset_layers], "sampled_slot_cells": blocks[args.sampled_slot] * 32,
                "first_indices": res.cell_indices[:5], "paths_self_check": bool(ok)}
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
