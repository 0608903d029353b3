"""Circuit-semantics verifier for input.json.  TEST INFRASTRUCTURE ONLY (same rule as pyoracle.py).

This is the *second*, independent statement of the hot path's arithmetic: it follows the circom circuit that
consumes input.json, not the Haskell/Nim generators.  Where the generators build trees bottom-up, the circuit
walks Merkle paths with bit masks; agreeing on every sample is the structural pin the reference itself relies on
(its CI only checks that the circuit accepts the Nim output, .github/workflows/generate.yml:149-156).

  Permutation / S-box / rounds   circuit/poseidon2/poseidon2_perm.circom:10-198
  PoseidonSponge                 circuit/poseidon2/poseidon2_sponge.circom:28-99
  KeyedCompression               circuit/poseidon2/poseidon2_compr.circom:30-41
  RootFromMerklePath             circuit/codex/merkle.circom:44-114
  ProveSingleCell                circuit/codex/single_cell.circom:30-73
  SampleAndProve                 circuit/codex/sample_cells.circom:23-48,58-148
  CeilingLog2 / Log2 masks       circuit/lib/log2.circom:61-130
"""
from __future__ import annotations

import json
from typing import List, Sequence

try:
    from .poseidon2_rc import RC_EXT, RC_INT
except ImportError:                     # pragma: no cover
    from poseidon2_rc import RC_EXT, RC_INT

P = 21888242871839275222246405745257275088548364400416034343698204186575808495617


def _sbox(x):                           # poseidon2_perm.circom:10-18
    x2 = x * x % P
    x4 = x2 * x2 % P
    return x * x4 % P


def circom_permutation(inp: Sequence[int]) -> List[int]:      # poseidon2_perm.circom:163-198
    a = [(2 * inp[0] + inp[1] + inp[2]) % P, (inp[0] + 2 * inp[1] + inp[2]) % P, (inp[0] + inp[1] + 2 * inp[2]) % P]

    def ext(i, s):                      # poseidon2_perm.circom:98-147
        sb = [_sbox((s[j] + RC_EXT[i][j]) % P) for j in range(3)]
        return [(2 * sb[0] + sb[1] + sb[2]) % P, (sb[0] + 2 * sb[1] + sb[2]) % P, (sb[0] + sb[1] + 2 * sb[2]) % P]

    def inr(i, s):                      # poseidon2_perm.circom:23-93
        sb = _sbox((s[0] + RC_INT[i]) % P)
        return [(2 * sb + s[1] + s[2]) % P, (sb + 2 * s[1] + s[2]) % P, (sb + s[1] + 3 * s[2]) % P]

    for k in range(4):
        a = ext(k, a)
    for k in range(56):
        a = inr(k, a)
    for k in range(4):
        a = ext(k + 4, a)
    return a


def circom_sponge(inp: Sequence[int], rate: int) -> int:      # poseidon2_sponge.circom:28-99 (t=3, output_len=1)
    t = 3
    nblocks = ((len(inp) + 1) + (rate - 1)) // rate
    padded = list(inp) + [1] + [0] * (nblocks * rate - len(inp) - 1)
    state = [0] * (t - 1) + [2 ** 64 + 256 * t + rate]
    for m in range(nblocks):
        sorbed = [(state[i] + padded[m * rate + i]) % P for i in range(rate)]
        state = circom_permutation(sorbed + state[rate:])
    return state[0]


def keyed_compression(key: int, x: int, y: int) -> int:       # poseidon2_compr.circom:30-41
    return circom_permutation([x, y, key])[0]


def to_bits(x: int, n: int) -> List[int]:
    assert 0 <= x < (1 << n), "ToBits: does not fit"
    return [(x >> i) & 1 for i in range(n)]


def root_from_merkle_path(max_depth, leaf, path_bits, last_bits, mask_bits, merkle_path) -> int:
    """circuit/codex/merkle.circom:44-114, statement by statement."""
    mask_c = [1] + list(mask_bits[1:max_depth + 1])
    aux = [leaf]
    is_last = [0] * (max_depth + 1)
    is_last[max_depth] = 1
    for i in range(max_depth - 1, -1, -1):
        is_last[i] = is_last[i + 1] * (1 if path_bits[i] == last_bits[i] else 0)
    for i in range(max_depth):
        bottom = 1 if i == 0 else 0
        odd = is_last[i] * (1 - path_bits[i])
        L, Rr = aux[i], merkle_path[i]
        sw = (Rr - L) * path_bits[i] % P
        aux.append(keyed_compression(bottom + 2 * odd, (L + sw) % P, (Rr - sw) % P))
    return sum((mask_c[i] - mask_c[i + 1]) * aux[i + 1] for i in range(max_depth)) % P


def _ceiling_log2_bits_mask(inp: int, n: int):                 # circuit/lib/log2.circom:111-130
    bits = to_bits(inp - 1, n)
    aux = [0] * (n + 1)
    aux[n] = 1
    mask = [0] * (n + 1)
    for i in range(n - 1, -1, -1):
        aux[i] = aux[i + 1] * (1 - bits[i])
        mask[i] = 1 - aux[i]
    return bits, mask


def _log2_mask(inp: int, n: int):                              # circuit/lib/log2.circom:54-100
    mask = [1 if (2 ** i) < inp else 0 for i in range(n + 1)]
    assert mask[0] == 1 and mask[n] == 0, "Log2: nCellsPerSlot out of range"
    assert inp == sum(2 ** (i + 1) * (mask[i] - mask[i + 1]) for i in range(n)), "Log2: not a power of two"
    return mask


def sample_and_prove(inp: dict, max_depth: int, max_log2_nslots: int, block_tree_depth: int,
                     n_felems_per_cell: int, n_samples: int) -> None:
    """circuit/codex/sample_cells.circom:58-148; raises AssertionError where the circuit's `===` would fail."""
    F = lambda s: int(s) % P
    entropy, dset_root, slot_root = F(inp["entropy"]), F(inp["dataSetRoot"]), F(inp["slotRoot"])
    slot_index, n_cells, n_slots = int(inp["slotIndex"]), int(inp["nCellsPerSlot"]), int(inp["nSlotsPerDataSet"])
    slot_proof = [F(v) for v in inp["slotProof"]]
    assert len(slot_proof) == max_log2_nslots
    assert len(inp["cellData"]) == n_samples and len(inp["merklePaths"]) == n_samples

    # dataset-level inclusion of the slot root              sample_cells.circom:95-109
    last_bits, mask = _ceiling_log2_bits_mask(n_slots, max_log2_nslots)
    top = root_from_merkle_path(max_log2_nslots, slot_root, to_bits(slot_index, max_log2_nslots), last_bits, mask, slot_proof)
    assert top == dset_root, "top root check failed"

    lg_mask = _log2_mask(n_cells, max_depth)                   # sample_cells.circom:114-123
    last = lg_mask[:max_depth]
    for cnt in range(n_samples):
        data = [F(v) for v in inp["cellData"][cnt]]
        path = [F(v) for v in inp["merklePaths"][cnt]]
        assert len(data) == n_felems_per_cell and len(path) == max_depth
        h = circom_sponge([entropy, slot_root, cnt + 1], 2)    # sample_cells.circom:23-48
        index_bits = [lg_mask[i] * ((h >> i) & 1) for i in range(max_depth)]
        # ProveSingleCell                                      single_cell.circom:30-73
        bd = block_tree_depth
        cell_hash = circom_sponge(data, 2)
        pbot = root_from_merkle_path(bd, cell_hash, index_bits[:bd], last[:bd], lg_mask[:bd] + [0], path[:bd])
        pmid = root_from_merkle_path(max_depth - bd, pbot, index_bits[bd:], last[bd:], lg_mask[bd:max_depth] + [0], path[bd:])
        assert pmid == slot_root, f"sample {cnt}: middle/bottom root check failed"


def verify_input_json(text: str, max_depth=32, max_log2_nslots=8, cell_size=2048, block_size=65536, n_samples=None) -> None:
    inp = json.loads(text)
    k = block_size // cell_size
    bd = (k - 1).bit_length()
    assert (1 << bd) == k
    n_samples = len(inp["cellData"]) if n_samples is None else n_samples
    sample_and_prove(inp, max_depth, max_log2_nslots, bd, (cell_size + 30) // 31, n_samples)
