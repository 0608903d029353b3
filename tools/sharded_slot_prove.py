#!/usr/bin/env python3
"""One slot range-sharded over the GPUs of a box, then challenged (power-of-two variant; 8e).

Every rank generates and commits its own 2^T-aligned block range, ONE all-gather exchanges the level-T nodes, every
rank builds the replicated top tree; the sampled indices are derived from the root on every rank, each cell's owner
produces its Merkle path and a byte-wise SUM inside the library delivers them (cdx_slot_prove_batch_sharded).  Rank 0 then verifies
every path in two stages on its GPU (block tree, slot tree -- Slot.hs:189-217) and, with --check-whole, re-commits the
whole slot alone and compares root, paths and leaves with the sharded answer.

launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
            tools/sharded_slot_prove.py --total-gib 128 --samples 100
"""
import argparse
import importlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "codex-storage-proofs-circuits_b200"
CELL, BLOCK, SEED = 2048, 65536, 0xC0DE


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--total-gib", type=float, default=8.0)
    ap.add_argument("--samples", type=int, default=100)
    ap.add_argument("--entropy", type=int, default=1234567)
    ap.add_argument("--check-whole", action="store_true", help="rank 0 also commits the whole slot alone (needs it to fit one GPU)")
    args = ap.parse_args()

    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    pkg = importlib.import_module(PKG)
    sharded = importlib.import_module(PKG + ".sharded")
    rank, world, local_rank = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = pkg.Context(local_rank)

    n_total_blocks = int(args.total_gib * (1 << 30)) // BLOCK
    n_cells = n_total_blocks * (BLOCK // CELL)
    assert n_cells & (n_cells - 1) == 0, "sampling needs a power-of-two cell count (sample/bn254.nim:19-20)"
    top_level, ranges = sharded.plan_block_ranges(n_total_blocks, world)
    first_block, my_blocks = ranges[rank]
    n_bytes = my_blocks * BLOCK
    d = torch.empty(n_bytes, dtype=torch.uint8, device="cuda")
    ctx.fill_synthetic_dev(SEED, first_block * (BLOCK // 8), n_bytes, d.data_ptr())
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    comm = sharded.comm_from_torch(ctx)               # the library's NCCL communicator; torch only carries the 128-byte id

    def commit():                                     # range commit + exchange + top tree: one C-ABI call
        return ctx.slot_commit_sharded_dev(comm, d.data_ptr(), n_bytes, CELL, BLOCK, first_block, n_total_blocks, top_level)

    commit().free()                                   # warm-up: pool allocations, NCCL channels
    barrier()
    t0 = time.perf_counter()
    sh = commit()
    root = sh.root
    barrier()
    t_commit = time.perf_counter() - t0

    t0 = time.perf_counter()
    (indices,), (paths,), (leaves,) = sh.prove_batch_sharded(comm, [args.entropy], args.samples, 32)   # cdx_slot_prove_batch_sharded
    barrier()
    t_prove = time.perf_counter() - t0

    out = None
    if rank == 0:
        bdepth, sdepth = 5, (n_total_blocks - 1).bit_length()
        t0 = time.perf_counter()
        blocks = ctx.reconstruct_roots(leaves, [i % 32 for i in indices], 32, paths, depth=bdepth)
        roots = ctx.reconstruct_roots(blocks, [i // 32 for i in indices], n_total_blocks, [p[bdepth:] for p in paths], depth=sdepth)
        t_verify = time.perf_counter() - t0
        ok = roots == [root] * len(indices) and all(v == 0 for p in paths for v in p[bdepth + sdepth:])
        owners = [next(r for r, (b0, nb) in enumerate(ranges) if b0 <= i // 32 < b0 + nb) for i in indices]
        out = {"workload": f"{args.total_gib:g} GiB slot ({n_total_blocks} blocks, {n_cells} cells) range-sharded over {world} GPUs",
               "n_gpus": world, "exchange_level": top_level, "subtree_roots_exchanged": sharded.level_width(n_total_blocks, top_level),
               "blocks_per_rank": [nb for _, nb in ranges], "slot_root": hex(root),
               "commit_s": t_commit, "commit_GB_per_s": n_total_blocks * BLOCK / t_commit / 1e9,
               "samples": args.samples, "prove_ms": 1e3 * t_prove, "verify_ms": 1e3 * t_verify,
               "ranks_owning_a_sample": sorted(set(owners)), "all_paths_reconstruct_the_root": bool(ok)}
    sh.free()
    if args.check_whole:
        del d
        torch.cuda.empty_cache()
        if rank == 0:
            whole_bytes = n_total_blocks * BLOCK
            dw = torch.empty(whole_bytes, dtype=torch.uint8, device="cuda")
            ctx.fill_synthetic_dev(SEED, 0, whole_bytes, dw.data_ptr())
            with ctx.slot_commit_dev(dw.data_ptr(), whole_bytes, CELL, BLOCK) as whole:
                same_root = whole.root == root
                wp, wl = whole.cell_paths(indices, 32)
            out["whole_slot_on_one_gpu"] = {"same_root": bool(same_root), "same_paths": wp == paths, "same_leaves": wl == leaves}
            ok = ok and same_root and wp == paths and wl == leaves
    barrier()
    if rank == 0:
        os.dup2(real_stdout, 1)
        print(json.dumps(out), flush=True)
        if not ok:
            raise SystemExit("sharded answer is wrong")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
