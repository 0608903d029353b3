"""Dataset-level commitment over the GPUs of a box (BASELINE config 5): many slots of different sizes -> slot roots
-> dataset tree -> proof input for one sampled slot.

The reference builds every slot of a dataset one after the other in one process
(reference/nim/proof_input/src/gen_input/bn254.nim:41-51).  Slots are independent, so here they are dealt to the ranks
by longest-processing-time bin packing; each rank commits its slots on its own GPU (synthetic bytes generated on the
device, only the hashes kept), ONE all-reduce collects the 32-byte slot roots, and every rank builds the small
dataset tree (bottom key 1 again, odd nodes for a non-power-of-two slot count: merkle/bn254.nim:29-60).  The rank that
owns the sampled slot keeps its handle, samples the cell indices from the slot root (sample/bn254.nim:16-27) and
extracts all Merkle paths in one call.
"""
from __future__ import annotations

import math
import time
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

CELL, BLOCK = 2048, 65536


def splitmix64(x: int) -> int:
    x = (x + 0x9e3779b97f4a7c15) & (2**64 - 1)
    z = x
    z = ((z ^ (z >> 30)) * 0xbf58476d1ce4e5b9) & (2**64 - 1)
    z = ((z ^ (z >> 27)) * 0x94d049bb133111eb) & (2**64 - 1)
    return z ^ (z >> 31)


def draw_slot_blocks(n_slots: int, min_bytes: float, max_bytes: float, seed: int, pow2_slot: Optional[int] = None,
                     pow2_blocks: Optional[int] = None) -> List[int]:
    """slot sizes in blocks, log-uniform in [min_bytes, max_bytes], deterministic in `seed`;
    slot `pow2_slot` is forced to `pow2_blocks` (a power of two) so that sampling -- which needs a power-of-two cell
    count, sample/bn254.nim:19-20 -- can run on it."""
    out = []
    for k in range(n_slots):
        u = splitmix64(seed * 1000003 + k) / 2.0**64
        size = math.exp(math.log(min_bytes) + u * (math.log(max_bytes) - math.log(min_bytes)))
        out.append(max(1, int(size // BLOCK)))
    if pow2_slot is not None:
        assert pow2_blocks and pow2_blocks & (pow2_blocks - 1) == 0
        out[pow2_slot] = pow2_blocks
    return out


def lpt_assign(sizes: Sequence[int], world: int) -> List[List[int]]:
    """longest-processing-time bin packing: biggest slot first onto the least loaded rank"""
    load = [0] * world
    bins: List[List[int]] = [[] for _ in range(world)]
    for k in sorted(range(len(sizes)), key=lambda i: (-sizes[i], i)):
        r = min(range(world), key=lambda j: (load[j], j))
        bins[r].append(k)
        load[r] += sizes[k]
    return bins


def slot_seed(base_seed: int, k: int) -> int:
    """per-slot seed, the reference's rule (dataset.nim:32): seed + 72 + 1001*k"""
    return (base_seed + 72 + 1001 * k) & (2**64 - 1)


@dataclass
class DatasetCommitment:
    slot_roots: List[int]
    dataset_layers: List[List[int]]
    dataset_root: int
    sampled_slot: int
    slot_proof: List[int]                     # padded to max_log2_nslots
    cell_indices: List[int] = field(default_factory=list)
    merkle_paths: List[List[int]] = field(default_factory=list)   # padded to max_depth
    cell_hashes: List[int] = field(default_factory=list)
    timings: Dict[str, float] = field(default_factory=dict)
    bytes_committed: int = 0
    per_rank_bytes: List[int] = field(default_factory=list)


def commit_dataset(ctx, slot_blocks: Sequence[int], base_seed: int, sampled_slot: int, entropy: int, n_samples: int,
                   max_depth: int = 32, max_log2_nslots: int = 8, rank: int = 0, world: int = 1, group=None,
                   device: str = "cuda") -> DatasetCommitment:
    """Commit every slot of a synthetic dataset and produce the proof-input pieces for `sampled_slot`."""
    import torch
    import torch.distributed as dist

    n_slots = len(slot_blocks)
    assert 1 <= n_slots <= (1 << max_log2_nslots) and 0 <= sampled_slot < n_slots
    bins = lpt_assign(slot_blocks, world)
    mine = bins[rank]
    owner = next(r for r in range(world) if sampled_slot in bins[r])
    distributed = world > 1 and dist.is_available() and dist.is_initialized()

    buf_blocks = max([slot_blocks[k] for k in mine], default=1)
    buf = torch.empty(buf_blocks * BLOCK, dtype=torch.uint8, device=device)
    roots = torch.zeros(n_slots * 32, dtype=torch.uint8)
    kept = None
    torch.cuda.synchronize()
    if distributed:
        dist.barrier(group=group)
    t0 = time.perf_counter()
    for k in mine:
        nbytes = slot_blocks[k] * BLOCK
        ctx.fill_synthetic_dev(slot_seed(base_seed, k), 0, nbytes, buf.data_ptr())
        slot = ctx.slot_commit_dev(buf.data_ptr(), nbytes, CELL, BLOCK)
        r = slot.root                                 # synchronises: buf may be refilled for the next slot
        roots[32 * k:32 * k + 32] = torch.frombuffer(bytearray(int(r).to_bytes(32, "little")), dtype=torch.uint8)
        if k == sampled_slot:
            kept = slot
        else:
            slot.free()
    torch.cuda.synchronize()
    t_commit = time.perf_counter() - t0
    if distributed:                                   # every slot has exactly one owner: SUM over uint8 is a gather
        rt = roots.to(device)
        dist.all_reduce(rt, op=dist.ReduceOp.SUM, group=group)
        roots = rt.cpu()
    t1 = time.perf_counter()
    raw = bytes(roots.numpy())
    slot_roots = [int.from_bytes(raw[32 * k:32 * k + 32], "little") for k in range(n_slots)]
    layers = ctx.merkle_layers(slot_roots, bottom=True)                       # gen_input/bn254.nim:49
    dset_root = layers[-1][0]
    proof, kk, m = [], sampled_slot, n_slots                                  # merkleProof: merkle.nim:21-42
    for i in range(len(layers) - 1):
        j = kk ^ 1
        proof.append(layers[i][j] if j < m else 0)
        kk >>= 1
        m = (m + 1) >> 1
    assert len(proof) <= max_log2_nslots
    proof += [0] * (max_log2_nslots - len(proof))                             # padMerkleProof: types.nim:27-37
    out = DatasetCommitment(slot_roots, layers, dset_root, sampled_slot, proof)
    n_cells = slot_blocks[sampled_slot] * (BLOCK // CELL)
    payload = torch.zeros(n_samples * (8 + 32 * max_depth + 32), dtype=torch.uint8)
    if rank == owner:
        idx = ctx.cell_indices(entropy, slot_roots[sampled_slot], n_cells, n_samples)      # sample/bn254.nim:26-27
        paths, leaves = kept.cell_paths(idx, max_depth)
        blob = b"".join(int(i).to_bytes(8, "little") for i in idx) + \
            b"".join(int(v).to_bytes(32, "little") for p in paths for v in p) + b"".join(int(v).to_bytes(32, "little") for v in leaves)
        payload = torch.frombuffer(bytearray(blob), dtype=torch.uint8).clone()
        kept.free()
    if distributed:
        pt = payload.to(device)
        dist.broadcast(pt, src=owner, group=group)
        payload = pt.cpu()
    blob = bytes(payload.numpy())
    out.cell_indices = [int.from_bytes(blob[8 * i:8 * i + 8], "little") for i in range(n_samples)]
    off = 8 * n_samples
    vals = [int.from_bytes(blob[off + 32 * i:off + 32 * i + 32], "little") for i in range(n_samples * max_depth + n_samples)]
    out.merkle_paths = [vals[i * max_depth:(i + 1) * max_depth] for i in range(n_samples)]
    out.cell_hashes = vals[n_samples * max_depth:]
    t_tail = time.perf_counter() - t1
    per_rank = [sum(slot_blocks[k] for k in b) * BLOCK for b in bins]
    tm = torch.tensor([t_commit, t_tail], dtype=torch.float64)
    if distributed:
        tmd = tm.to(device)
        dist.all_reduce(tmd, op=dist.ReduceOp.MAX, group=group)
        tm = tmd.cpu()
    out.timings = {"commit_s": float(tm[0]), "roots_tree_paths_s": float(tm[1])}
    out.bytes_committed = sum(slot_blocks) * BLOCK
    out.per_rank_bytes = per_rank
    return out
