"""Python big-integer restatement of the slot-commitment path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the product path
(codex-storage-proofs-circuits_b200) never does.  It is the slow, obviously-correct twin of oracle/codex_oracle.c
and is used to cross-check it and to generate the goldens under tests/golden/ (tools/gen_goldens.py).

Parity pinning: the permutation is pinned by the reference's one stored known-answer
(reference/haskell/src/Poseidon2/Example.hs:13-22).  The reference stores NO expected values above the
permutation (its test-vector programs only print), so sponge / Merkle / slot / input.json parity is pinned
structurally: this file follows the Haskell statement, oracle/circuit_verifier.py follows the independent
circom statement, and tests require them to agree.

All citations are file:line under /root/reference.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Sequence

try:                                    # imported as oracle.pyoracle or as a script sibling
    from .poseidon2_rc import RC_EXT, RC_INT
except ImportError:                     # pragma: no cover
    from poseidon2_rc import RC_EXT, RC_INT

# BN254 scalar field (README.md:76, test/Params.hs:12)
R = 21888242871839275222246405745257275088548364400416034343698204186575808495617

# --------------------------------------------------------------------------------------------------
# Poseidon2 permutation, t = 3            reference/haskell/src/Poseidon2/Permutation.hs:14-45


def sbox(x: int) -> int:                # Permutation.hs:14-17
    x2 = x * x % R
    x4 = x2 * x2 % R
    return x4 * x % R


def internal_round(c: int, s):          # Permutation.hs:19-26
    x, y, z = s
    x = sbox((x + c) % R)
    return ((2 * x + y + z) % R, (x + 2 * y + z) % R, (x + y + 3 * z) % R)


def external_round(c, s):               # Permutation.hs:28-33
    x = sbox((s[0] + c[0]) % R)
    y = sbox((s[1] + c[1]) % R)
    z = sbox((s[2] + c[2]) % R)
    t = x + y + z
    return ((x + t) % R, (y + t) % R, (z + t) % R)


def linear_layer(s):                    # Permutation.hs:35-36
    t = s[0] + s[1] + s[2]
    return ((s[0] + t) % R, (s[1] + t) % R, (s[2] + t) % R)


def permutation(s):                     # Permutation.hs:40-45
    s = linear_layer(tuple(v % R for v in s))
    for r in range(4):
        s = external_round(RC_EXT[r], s)
    for c in RC_INT:
        s = internal_round(c, s)
    for r in range(4, 8):
        s = external_round(RC_EXT[r], s)
    return s


# --------------------------------------------------------------------------------------------------
# Sponges                                 reference/haskell/src/Poseidon2/Sponge.hs:13-43

IV_RATE1 = (1 << 64) + 0x0301           # Sponge.hs:17
IV_RATE2 = (1 << 64) + 0x0302           # Sponge.hs:34


def sponge1(xs: Sequence[int]) -> int:  # Sponge.hs:13-25
    s = (0, 0, IV_RATE1)
    for a in list(xs) + [1]:
        s = permutation(((s[0] + a) % R, s[1], s[2]))
    return s[0]


def sponge2(xs: Sequence[int]) -> int:  # Sponge.hs:30-43
    xs = list(xs)
    xs = xs + ([1] if len(xs) % 2 == 1 else [1, 0])
    s = (0, 0, IV_RATE2)
    for i in range(0, len(xs), 2):
        s = permutation(((s[0] + xs[i]) % R, (s[1] + xs[i + 1]) % R, s[2]))
    return s[0]


# --------------------------------------------------------------------------------------------------
# bytes -> field elements                 reference/haskell/src/Slot.hs:243-270, README.md:86-99


def bytes_to_elements(data: bytes) -> List[int]:
    """Append 0x01, zero-pad to a multiple of 31, read each 31-byte chunk little-endian (Slot.hs:243-270)."""
    padded = bytes(data) + b"\x01"
    if len(padded) % 31:
        padded += b"\x00" * (31 - len(padded) % 31)
    return [int.from_bytes(padded[i:i + 31], "little") for i in range(0, len(padded), 31)]


def hash_bytes(data: bytes) -> int:     # Slot.hs:227-228 (hashCell_), nim blocks/bn254.nim:27
    return sponge2(bytes_to_elements(data))


def hash_cell(cell: bytes, cell_size: int) -> int:   # Slot.hs:222-225, blocks/bn254.nim:23-29
    if len(cell) != cell_size:
        raise ValueError("hashCell: invalid cell data size")
    return hash_bytes(cell)


# --------------------------------------------------------------------------------------------------
# Keyed compression and Merkle trees      reference/haskell/src/Poseidon2/Merkle.hs:69-83,156-208
#                                         reference/nim/proof_input/src/merkle/bn254.nim:18-63

KEY_NONE, KEY_BOTTOM, KEY_ODD, KEY_ODD_BOTTOM = 0, 1, 2, 3     # Merkle.hs:171-175, merkle/bn254.nim:24-27


def compress(x: int, y: int, key: int = 0) -> int:   # Merkle.hs:202-203
    return permutation((x, y, key))[0]


def merkle_layers(xs: Sequence[int], bottom: bool = True) -> List[List[int]]:
    """All layers, bottom first (merkle/bn254.nim:29-63; Merkle.hs:69-78)."""
    xs = [v % R for v in xs]
    if not xs:
        raise ValueError("merkle tree of empty input")
    layers = []
    while True:
        m = len(xs)
        if not bottom and m == 1:
            layers.append(xs)
            return layers
        kb = 1 if bottom else 0
        ys = [compress(xs[2 * i], xs[2 * i + 1], kb) for i in range(m // 2)]
        if m % 2:
            ys.append(compress(xs[m - 1], 0, kb + 2))
        layers.append(xs)
        xs, bottom = ys, False


def merkle_root(xs: Sequence[int]) -> int:           # Merkle.hs:180-189
    return merkle_layers(xs)[-1][0]


@dataclass
class MerkleProof:                      # nim types.nim:14-18
    leaf_index: int
    leaf_value: int
    merkle_path: List[int]
    number_of_leaves: int


def merkle_proof(layers: List[List[int]], index: int) -> MerkleProof:   # nim merkle.nim:21-42
    depth, nleaves = len(layers) - 1, len(layers[0])
    assert 0 <= index < nleaves
    path, k, m = [], index, nleaves
    for i in range(depth):
        j = k ^ 1
        path.append(layers[i][j] if j < m else 0)
        k >>= 1
        m = (m + 1) >> 1
    return MerkleProof(index, layers[0][index], path, nleaves)


def reconstruct_root(proof: MerkleProof) -> int:     # nim merkle.nim:51-74, Merkle.hs:114-123
    m, j, h, bottom = proof.number_of_leaves, proof.leaf_index, proof.leaf_value, 1
    for p in proof.merkle_path:
        if j & 1:
            h = compress(p, h, bottom)
        elif j == m - 1:
            h = compress(h, p, bottom + 2)
        else:
            h = compress(h, p, bottom)
        bottom = 0
        j >>= 1
        m = (m + 1) >> 1
    return h


def merge_merkle_proofs(bot: MerkleProof, top: MerkleProof) -> MerkleProof:   # nim merkle.nim:86-100
    assert reconstruct_root(bot) == top.leaf_value
    return MerkleProof(top.leaf_index * bot.number_of_leaves + bot.leaf_index, bot.leaf_value,
                       bot.merkle_path + top.merkle_path, bot.number_of_leaves * top.number_of_leaves)


def pad_merkle_proof(p: MerkleProof, newlen: int) -> MerkleProof:              # nim types.nim:27-37
    assert newlen >= len(p.merkle_path)
    return MerkleProof(p.leaf_index, p.leaf_value, p.merkle_path + [0] * (newlen - len(p.merkle_path)),
                       p.number_of_leaves)


# --------------------------------------------------------------------------------------------------
# Data source                             reference/nim/proof_input/src/slot.nim:23-32, dataset.nim:32

M64 = (1 << 64) - 1


def gen_fake_cell(seed: int, idx: int, cell_size: int) -> bytes:              # slot.nim:23-32, Slot.hs:87-96
    seed1 = (seed + 0xdeadcafe) & M64
    seed2 = (idx + 0x98765432) & M64
    out, s = bytearray(cell_size), 1
    for i in range(cell_size):
        s = (s * ((s + seed1) & M64) * ((s + seed2) & M64) + s * (s ^ 0x5a5a5a5a) + seed1 * s + (seed2 + 17)) & M64
        s %= 1698428844001831
        out[i] = s & 0xFF
    return bytes(out)


def parametric_slot_seed(seed: int, k: int) -> int:                           # dataset.nim:32
    return (seed + 72 + 1001 * k) & M64


# --------------------------------------------------------------------------------------------------
# Slot / dataset commitment               reference/nim/proof_input/src/gen_input/bn254.nim:21-74


@dataclass
class GlobalConfig:                     # nim types.nim:87-91, cli.nim:53-58
    max_depth: int = 32
    max_log2_nslots: int = 8
    cell_size: int = 2048
    block_size: int = 65536

    @property
    def cells_per_block(self) -> int:   # types.nim:120-123
        k = self.block_size // self.cell_size
        assert k * self.cell_size == self.block_size, "block size is not divisible by cell size"
        return k


@dataclass
class DataSetConfig:                    # nim types.nim:81-85, cli.nim:60-65
    n_slots: int = 11
    n_cells: int = 256
    n_samples: int = 5
    seed: int = 12345                   # FakeData(seed)


def block_tree_layers(glob: GlobalConfig, block: bytes) -> List[List[int]]:   # blocks/bn254.nim:60-67
    assert len(block) == glob.block_size
    cs = glob.cell_size
    leaves = [hash_cell(block[i * cs:(i + 1) * cs], cs) for i in range(glob.cells_per_block)]
    return merkle_layers(leaves)


def slot_block_data(glob: GlobalConfig, seed: int, block_idx: int) -> bytes:  # slot.nim:70-73
    k = glob.cells_per_block
    return b"".join(gen_fake_cell(seed, block_idx * k + i, glob.cell_size) for i in range(k))


def build_slot_tree_full(glob: GlobalConfig, n_cells: int, seed: int):        # gen_input/bn254.nim:21-30
    nblocks = n_cells // glob.cells_per_block
    assert nblocks * glob.cells_per_block == n_cells
    mini = [block_tree_layers(glob, slot_block_data(glob, seed, b)) for b in range(nblocks)]
    big = merkle_layers([t[-1][0] for t in mini])
    return mini, big


def ceiling_log2(x: int) -> int:        # nim misc.nim:19-23
    return -1 if x == 0 else (x - 1).bit_length()


def cell_index(entropy: int, slot_root: int, n_cells: int, counter: int) -> int:   # sample/bn254.nim:16-24
    lg = ceiling_log2(n_cells)
    assert (1 << lg) == n_cells, "numberOfCells is assumed to be a power of two"
    return sponge2([entropy % R, slot_root, counter]) & (n_cells - 1)


def cell_indices(entropy: int, slot_root: int, n_cells: int, n_samples: int) -> List[int]:   # sample/bn254.nim:26-27
    return [cell_index(entropy, slot_root, n_cells, c) for c in range(1, n_samples + 1)]


@dataclass
class SlotProofInput:                   # nim types.nim:52-60
    data_set_root: int
    entropy: int
    n_slots: int
    n_cells: int
    slot_root: int
    slot_index: int
    slot_proof: MerkleProof
    cell_data: List[bytes] = field(default_factory=list)
    merkle_proofs: List[MerkleProof] = field(default_factory=list)


def generate_proof_input(glob: GlobalConfig, dset: DataSetConfig, slot_idx: int, entropy: int) -> SlotProofInput:
    """gen_input/bn254.nim:35-74 (without the per-sample slot rebuild of line 57, which recomputes identical trees)."""
    k = glob.cells_per_block
    trees = [build_slot_tree_full(glob, dset.n_cells, parametric_slot_seed(dset.seed, s)) for s in range(dset.n_slots)]
    slot_roots = [big[-1][0] for (_, big) in trees]
    dset_layers = merkle_layers(slot_roots)
    mini, big = trees[slot_idx]
    our_root = slot_roots[slot_idx]
    out = SlotProofInput(dset_layers[-1][0], entropy % R, dset.n_slots, dset.n_cells, our_root, slot_idx,
                         pad_merkle_proof(merkle_proof(dset_layers, slot_idx), glob.max_log2_nslots))
    seed = parametric_slot_seed(dset.seed, slot_idx)
    for ci in cell_indices(entropy, our_root, dset.n_cells, dset.n_samples):
        bot = merkle_proof(mini[ci // k], ci % k)
        top = merkle_proof(big, ci // k)
        out.cell_data.append(gen_fake_cell(seed, ci, glob.cell_size))
        out.merkle_proofs.append(pad_merkle_proof(merge_merkle_proofs(bot, top), glob.max_depth))
    return out


# --------------------------------------------------------------------------------------------------
# input.json writer                       reference/nim/proof_input/src/json/bn254.nim:19-74, json/shared.nim:17-25


def _q(x: int) -> str:                  # types/bn254.nim:29-37
    return '"' + str(x) + '"'


def _write_list(lines: List[str], prefix: str, xs, write_fun) -> None:        # json/shared.nim:17-25
    indent = " " * len(prefix)
    for i, x in enumerate(xs):
        write_fun(lines, (prefix + "[ ") if i == 0 else (indent + ", "), x)
    lines.append(indent + "]")


def _write_felems(lines, prefix, xs) -> None:                                 # json/bn254.nim:19-20
    _write_list(lines, prefix, xs, lambda ls, p, x: ls.append(p + _q(x)))


def export_proof_input(inp: SlotProofInput) -> str:                           # json/bn254.nim:57-74
    ls: List[str] = ["{"]
    ls.append('  "dataSetRoot":      ' + _q(inp.data_set_root))
    ls.append(', "entropy":          ' + _q(inp.entropy))
    ls.append(', "nCellsPerSlot":    ' + str(inp.n_cells))
    ls.append(', "nSlotsPerDataSet": ' + str(inp.n_slots))
    ls.append(', "slotIndex":        ' + str(inp.slot_index))
    ls.append(', "slotRoot":         ' + _q(inp.slot_root))
    ls.append(', "slotProof":')
    _write_felems(ls, "    ", inp.slot_proof.merkle_path)
    ls.append(', "cellData":')
    _write_list(ls, "    ", inp.cell_data, lambda l, p, c: _write_felems(l, p, bytes_to_elements(c)))
    ls.append(', "merklePaths":')
    _write_list(ls, "    ", inp.merkle_proofs, lambda l, p, m: _write_felems(l, p, m.merkle_path))
    ls.append("}")
    return "\n".join(ls) + "\n"


# --------------------------------------------------------------------------------------------------
# Optional acceleration: swap the three hot primitives for the C oracle (same results, ~1000x faster).  The
# orchestration above (trees, proofs, sampling, JSON) stays in Python.  Used to produce the config-1 golden.

def use_c_primitives(enable: bool = True) -> None:
    g = globals()
    if enable:
        try:
            from . import coracle as _c
        except ImportError:                 # pragma: no cover
            import coracle as _c
        g.setdefault("_py_impl", {k: g[k] for k in ("hash_bytes", "compress", "gen_fake_cell", "sponge2", "sponge1")})
        g["hash_bytes"] = _c.hash_bytes
        g["compress"] = _c.compress
        g["gen_fake_cell"] = _c.gen_fake_cell
        g["sponge2"] = _c.sponge2
        g["sponge1"] = _c.sponge1
    elif "_py_impl" in g:
        g.update(g.pop("_py_impl"))
