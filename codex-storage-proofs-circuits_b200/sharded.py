"""Block-range sharding of one slot over the GPUs of a box (one process per GPU) -- SURVEY.md section 8(e).

The reference is single-process; what shards is its own structure: cells and blocks are independent, and level-l
node i of the slot tree covers blocks [i*2^l, (i+1)*2^l), so any 2^T-aligned block range is a set of complete
sub-trees (only the globally last range can contain odd nodes: reference/nim/proof_input/src/merkle/bn254.nim:38-53).
Each rank commits its range up to level T on its own GPU; the only exchange is ONE all-gather of the level-T nodes
(32 bytes each, a handful per rank); the small top tree is then built redundantly on every rank.

The orchestration below is written against two small interfaces so the same code runs over NCCL on GPUs
(bench.py) and over gloo on CPU in the world_size-2 tests:
  slot-like:  .subtree_roots_tensor(device) -> uint8 tensor [n_local_nodes*32]
              .set_top_tensor(uint8 tensor [n_level_nodes*32])
              .root, .cell_paths(indices, max_depth)
  torch.distributed process group (or None for a single rank).
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
    """Choose the exchange level T and a contiguous, 2^T-aligned block range per rank.

    T is the largest level for which splitting the 2^T-block chunks evenly leaves the most loaded rank within
    `max_imbalance` of the ideal share (100 GiB over 8 GPUs: T = 13, 25 chunks of 8192 blocks per rank) -- the
    larger T, the fewer sub-tree roots are exchanged and the smaller the replicated top tree.  Falls back to
    T = 0 (exchange raw block hashes) for slots too small to balance.  Ranks with no blocks get (first, 0)."""
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
    """weak-scaling layout: every rank holds exactly blocks_per_rank blocks; T = the alignment those ranges allow"""
    t = 0
    while world_size > 1 and blocks_per_rank % (1 << (t + 1)) == 0 and (1 << (t + 1)) <= blocks_per_rank:
        t += 1
    return t, [(r * blocks_per_rank, blocks_per_rank) for r in range(world_size)]


def exchange_subtree_roots(slot, n_total_blocks: int, top_level: int, ranges: Sequence[Tuple[int, int]], group=None, device="cuda"):
    """all-gather the level-T nodes of every rank and install them as the complete level T of `slot`.

    Counts differ per rank (the last range is ragged), torch all_gather wants equal shapes: pad to the largest count,
    gather once, compact.  Payload: 32 B x (nodes per rank) -- a few hundred bytes."""
    import torch
    import torch.distributed as dist

    counts = []
    for (b0, nb) in ranges:
        if nb == 0:
            counts.append(0)
        else:
            counts.append(ceil_div(b0 + nb, 1 << top_level) - (b0 >> top_level))
    total = sum(counts)
    assert total == level_width(n_total_blocks, top_level), (counts, n_total_blocks, top_level)
    local = slot.subtree_roots_tensor(device)
    world = len(ranges)
    if world == 1 or group is None and not (dist.is_available() and dist.is_initialized()):
        gathered = local
    else:
        mx = max(counts)
        padded = torch.zeros(mx * 32, dtype=torch.uint8, device=local.device)
        padded[: local.numel()] = local
        out = torch.empty(world * mx * 32, dtype=torch.uint8, device=local.device)
        dist.all_gather_into_tensor(out, padded, group=group)
        gathered = torch.cat([out[r * mx * 32: r * mx * 32 + counts[r] * 32] for r in range(world)]).contiguous()
    assert gathered.numel() == total * 32
    slot.set_top_tensor(gathered)
    return gathered


def gather_cell_paths(slot, indices: Sequence[int], max_depth: int, group=None, device="cuda"):
    """Merkle paths of sampled cells of a sharded slot: the owner rank of each cell produces the whole path (its
    siblings below level T are local, those above are replicated), every other rank produces zeros; a SUM
    all-reduce over uint8 is therefore a copy from the owner."""
    import torch
    import torch.distributed as dist

    paths, leaves = slot.cell_paths(list(indices), max_depth)
    if group is None and not (dist.is_available() and dist.is_initialized()):
        return paths, leaves
    flat = b"".join(int(v).to_bytes(32, "little") for p in paths for v in p) + b"".join(int(v).to_bytes(32, "little") for v in leaves)
    t = torch.frombuffer(bytearray(flat), dtype=torch.uint8).to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    raw = bytes(t.cpu().numpy())
    n = len(indices)
    vals = [int.from_bytes(raw[i:i + 32], "little") for i in range(0, len(raw), 32)]
    return [vals[i * max_depth:(i + 1) * max_depth] for i in range(n)], vals[n * max_depth:]


class GpuShard:
    """adapter: a capi.Slot committed with cdx_slot_commit_range_* exposed through the slot-like interface"""

    def __init__(self, slot):
        self.slot = slot

    def subtree_roots_tensor(self, device="cuda"):
        import torch
        _, cnt, _ = self.slot.subtree_roots()
        t = torch.empty(cnt * 32, dtype=torch.uint8, device=device)
        if cnt:
            self.slot.subtree_roots_copy_dev(t.data_ptr())
            torch.cuda.synchronize()
        return t

    def set_top_tensor(self, t):
        import torch
        torch.cuda.current_stream().synchronize()   # the gather/compaction ran on torch's stream, the library copies on its own
        self.slot.set_top_dev(t.data_ptr(), t.numel() // 32)
        self._keep = t          # the library copies on its stream; keep the tensor until the next sync

    @property
    def root(self) -> int:
        return self.slot.root

    def cell_paths(self, indices, max_depth):
        return self.slot.cell_paths(indices, max_depth)

    def free(self):
        self.slot.free()
