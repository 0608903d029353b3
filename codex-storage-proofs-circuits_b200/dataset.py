"""Dataset-level commitment over the GPUs of a box (BASELINE config 5): many slots of different sizes -> slot roots
-> dataset tree -> proof input for one sampled slot.

The reference builds every slot of a dataset one after the other in one process
(reference/nim/proof_input/src/gen_input/bn254.nim:41-51).  The whole of that loop now sits behind ONE C-ABI call,
`cdx_dataset_commit` (include/codex_commit.h): longest-processing-time packing of the slots onto the ranks, block-range
sharding of slots that would unbalance them, batched commitment of small slots, one NCCL collective for the 32-byte
roots, the dataset tree on every rank; `cdx_dataset_prove` answers a challenge against the retained slot.  What is left
here is workload description: the synthetic size distribution of the benchmark dataset, the reference's per-slot seed
rule, and a Python twin of the packing so that the plan can be inspected without the library.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

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
    """longest-processing-time bin packing: biggest slot first onto the least loaded rank (Python twin of the packing
    inside cdx_dataset_commit, without its sharding rule)"""
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


def synthetic_descs(slot_blocks: Sequence[int], base_seed: int, kind: Optional[int] = None):
    """slot descriptors for cdx_dataset_commit: synthetic (or fake) bytes with the reference's per-slot seeds"""
    from . import capi
    kind = capi.SRC_SYNTHETIC if kind is None else kind
    return [(kind, slot_seed(base_seed, k), nb * BLOCK) for k, nb in enumerate(slot_blocks)]
