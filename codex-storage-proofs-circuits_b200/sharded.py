"""Block-range sharding of one slot over the GPUs of a box (one process per GPU) -- SURVEY.md section 8(e).

The reference is single-process; what shards is its own structure: cells and blocks are independent, and level-l
node i of the slot tree covers blocks [i*2^l, (i+1)*2^l), so any 2^T-aligned block range is a set of complete
sub-trees (only the globally last range can contain odd nodes: reference/nim/proof_input/src/merkle/bn254.nim:38-53).
Each rank commits its range up to level T on its own GPU; the only exchange is the level-T nodes (32 bytes each, a
handful per rank); the small top tree is then built redundantly on every rank.

The whole data plane lives behind the C ABI (include/codex_commit.h): `cdx_slot_commit_sharded_*` commits the range,
combines the level-T nodes with one NCCL collective on the slot's stream and builds the top tree;
`cdx_slot_cell_paths_sharded` / `cdx_slot_prove_batch_sharded` answer challenges collectively.  This module only
 * creates the library's communicator for a torch.distributed job (torch broadcasts the 128-byte id, nothing else), and
 * keeps pure-Python twins of the two range planners (`cdx_plan_block_ranges`, `cdx_block_ranges_top_level`) so that the
   plan can be inspected and tested without the shared library.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple


def ceil_div(a: int, b: int) -> int:
    return -(-a // b)


def level_width(n_blocks: int, level: int) -> int:
    """global width of slot-tree level `level` (n, ceil(n/2), ...)"""
    w = n_blocks
    for _ in range(level):
        w = (w + 1) // 2
    return w


def plan_block_ranges(n_total_blocks: int, world_size: int, max_imbalance: float = 0.01,
                      min_chunks_per_rank: int = 1) -> Tuple[int, List[Tuple[int, int]]]:
    """Python twin of cdx_plan_block_ranges: the exchange level T and a contiguous, 2^T-aligned block range per rank.

    T is the largest level for which splitting the 2^T-block chunks evenly leaves the most loaded rank within
    `max_imbalance` of the ideal share (100 GiB over 8 GPUs: T = 13, 25 chunks of 8192 blocks per rank) -- the
    larger T, the fewer sub-tree roots are exchanged and the smaller the replicated top tree.  Falls back to
    T = 0 (exchange raw block hashes) for slots too small to balance.  Ranks with no blocks get (first, 0): an empty
    shard, which the library accepts (it contributes nothing to the exchange and still receives the top tree)."""
    assert n_total_blocks >= 1 and world_size >= 1
    best_t = 0
    t = max(0, (n_total_blocks - 1).bit_length())
    while t > 0:
        n_chunks = ceil_div(n_total_blocks, 1 << t)
        heaviest = min(ceil_div(n_chunks, world_size) << t, n_total_blocks)
        if n_chunks >= min_chunks_per_rank * world_size and heaviest <= (1.0 + max_imbalance) * n_total_blocks / world_size:
            best_t = t
            break
        t -= 1
    t = best_t
    n_chunks = ceil_div(n_total_blocks, 1 << t)
    ranges = []
    for r in range(world_size):
        c0, c1 = n_chunks * r // world_size, n_chunks * (r + 1) // world_size
        b0, b1 = min(c0 << t, n_total_blocks), min(c1 << t, n_total_blocks)
        ranges.append((b0, b1 - b0))
    return t, ranges


def fixed_ranges(blocks_per_rank: int, world_size: int) -> Tuple[int, List[Tuple[int, int]]]:
    """weak-scaling layout: every rank holds exactly blocks_per_rank blocks; T = the alignment those ranges allow
    (Python twin of cdx_block_ranges_top_level for this layout)"""
    t = 0
    while world_size > 1 and blocks_per_rank % (1 << (t + 1)) == 0 and (1 << (t + 1)) <= blocks_per_rank:
        t += 1
    return t, [(r * blocks_per_rank, blocks_per_rank) for r in range(world_size)]


def comm_from_torch(ctx, group=None, device: Optional[str] = None):
    """The library's communicator for this torch.distributed job: rank 0 asks the library for the NCCL id
    (cdx_comm_unique_id), torch broadcasts those 128 bytes, every rank joins (cdx_comm_init_rank).  After this call torch
    takes no further part in the data path."""
    import torch
    import torch.distributed as dist
    from . import capi

    if not (dist.is_available() and dist.is_initialized()):
        return ctx.comm_init(1, 0, None)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if world == 1:
        return ctx.comm_init(1, 0, None)
    if device is None:
        device = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.zeros(capi.COMM_ID_BYTES, dtype=torch.uint8)
    if rank == 0:
        t = torch.frombuffer(bytearray(capi.comm_unique_id()), dtype=torch.uint8).clone()
    t = t.to(device)
    dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    return ctx.comm_init(world, rank, bytes(t.cpu().numpy()))
