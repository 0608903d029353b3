"""GPU parity: the CUDA path through the C ABI vs the oracle and the frozen goldens, bit-exact.
Sizes are what the oracle finishes in seconds; the full BASELINE sizes are covered by test_gpu_properties.py."""
import json
import os
import random
import subprocess

import pytest

from conftest import GOLDEN, ROOT

pytestmark = pytest.mark.gpu
R = 21888242871839275222246405745257275088548364400416034343698204186575808495617


def splitmix_felts(seed, n):
    from tools.gen_goldens import random_felts
    return random_felts(seed, n)


# ---- hash layer --------------------------------------------------------------------------------------------

def test_permutation_kat_and_goldens(ctx, vectors):
    kat = vectors["permutation_kat"]
    assert ctx.permutation([int(v) for v in kat["in"]]) == tuple(int(v) for v in kat["out"])
    for case in vectors["permutations"]:
        assert ctx.permutation([int(v) for v in case["in"]]) == tuple(int(v) for v in case["out"])


def test_permutation_batch_vs_oracle(ctx, orc, pkg):
    """BASELINE config 2 shape at a size the oracle does in seconds: states (j, j+1, j+2) then random canonical elements."""
    n = 4096
    states = [(j, j + 1, j + 2) for j in range(n // 2)]
    fl = splitmix_felts(1, 3 * (n // 2))
    states += [tuple(fl[3 * i:3 * i + 3]) for i in range(n // 2)]
    blob = b"".join(pkg.capi.pack(s) for s in states)
    assert ctx.permutation_batch_bytes(blob) == orc.permutation_batch_bytes(blob)


def test_permutation_noncanonical_inputs_are_taken_mod_r(ctx, orc):
    s = (R + 5, 2**256 - 1, R)
    assert ctx.permutation(s) == orc.permutation(tuple(v % R for v in s))


def test_testvector_suite(ctx, vectors, pkg):
    """reference/nim/testvectors/src/testvectors.nim:20-72, every printed case"""
    for n in range(9):
        xs = list(range(1, n + 1))
        assert str(ctx.sponge(xs, 1)) == vectors["sponge_rate1"][n]
        assert str(ctx.sponge(xs, 2)) == vectors["sponge_rate2"][n]
    for n in range(81):
        b = bytes(range(1, n + 1))
        assert str(ctx.hash_bytes(b)) == vectors["hash_bytes"][n], n
    for n in range(1, 41):
        xs = list(range(1, n + 1))
        assert str(ctx.merkle_root(xs)) == vectors["merkle_root_felts"][n - 1], n
    for case in vectors["compress"]:
        assert str(ctx.compress(int(case["x"]), int(case["y"]), case["key"])) == case["out"]


def test_byte_merkle_roots(ctx, orc, vectors):
    """Merkle.digest(openArray[byte]) through its own entry point, cdx_merkle_root_bytes_host: chunking and tree both on the
    GPU, every n = 0..80 of testvectors.nim:60-66, against the golden values and the oracle"""
    for n in range(0, 81):
        data = bytes(range(1, n + 1))
        got = ctx.merkle_root_bytes(data)
        assert str(got) == vectors["merkle_root_bytes"][n], n
        assert got == orc.merkle_root(orc.bytes_to_elements(data)), n
    rnd = random.Random(61)
    for n in (31, 62, 93, 2048, 5000):                            # chunk-boundary lengths and a multi-level tree
        data = bytes(rnd.randrange(256) for _ in range(n))
        assert ctx.merkle_root_bytes(data) == orc.merkle_root(orc.bytes_to_elements(data)), n


def test_sponge_batches_random(ctx, orc):
    rnd = random.Random(3)
    for ln in (0, 1, 2, 3, 7, 67):
        items = [[rnd.randrange(R) for _ in range(ln)] for _ in range(5)]
        for rate in (1, 2):
            assert ctx.sponge_batch(items, rate) == [orc.sponge(it, rate) for it in items]


def test_hash_bytes_ragged_lengths(ctx, orc):
    """aligned-word loader (len % 4 == 0) and byte loader, incl. lengths around the 31-byte chunk and pad boundaries"""
    rnd = random.Random(4)
    for ln in [0, 1, 3, 4, 30, 31, 32, 61, 62, 63, 64, 92, 93, 124, 128, 256, 2044, 2047, 2048, 31 * 66, 31 * 67 + 1, 4096]:
        n_items = 3
        data = bytes(rnd.randrange(256) for _ in range(ln * n_items))
        got = ctx.hash_bytes_batch(data, n_items, ln)
        assert got == [orc.hash_bytes(data[i * ln:(i + 1) * ln]) for i in range(n_items)], ln


def test_cell_hash_goldens(ctx, vectors, orc):
    cells = {"zeros": bytes(2048), "ones_ff": b"\xff" * 2048, "ramp": bytes(i & 255 for i in range(2048)),
             "fake_seed15420_cell0": ctx.fake_cells(15420, 0, 1, 2048)}
    for k, v in cells.items():
        assert str(ctx.hash_bytes(v)) == vectors["cell_hashes"][k], k


def test_compress_all_keys(ctx, orc):
    rnd = random.Random(5)
    xs = [rnd.randrange(R) for _ in range(64)] + [0, R - 1]
    ys = [rnd.randrange(R) for _ in range(64)] + [0, R - 1]
    keys = [i % 4 for i in range(66)]
    assert ctx.compress_batch(xs, ys, keys) == [orc.compress(x, y, k) for x, y, k in zip(xs, ys, keys)]


# ---- Merkle trees ------------------------------------------------------------------------------------------

@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 6, 7, 8, 31, 32, 33, 64, 100, 163])
def test_merkle_layers_all_widths(ctx, orc, n):
    rnd = random.Random(n)
    xs = [rnd.randrange(R) for _ in range(n)]
    assert ctx.merkle_layers(xs) == orc.merkle_layers(xs)
    assert ctx.merkle_layers(xs, bottom=False) == orc.merkle_layers(xs, False)
    assert ctx.merkle_root(xs) == orc.merkle_root(xs)


# ---- data source -------------------------------------------------------------------------------------------

def test_fake_data_matches_reference_generator(ctx, orc, vectors):
    fk = vectors["fake_cell_sha256"]
    assert ctx.fake_cells(fk["seed"], 0, 1, 2048)[:16].hex() == fk["first16"]
    for seed, first, n, cs in [(15420, 0, 40, 2048), (2**64 - 5, 123456789, 9, 128), (7, 5, 33, 256)]:
        got = ctx.fake_cells(seed, first, n, cs)
        assert got == b"".join(orc.gen_fake_cell(seed, first + i, cs) for i in range(n))


# ---- slot commitment ---------------------------------------------------------------------------------------

def test_config1_slot(ctx, orc):
    """BASELINE config 1, slot 3: every cell hash, block hash, slot-tree layer and the root."""
    meta = json.load(open(os.path.join(GOLDEN, "meta.json")))["config1"]
    seed = 12345 + 72 + 3 * 1001
    with ctx.slot_commit_fake(seed, 2048) as slot:
        assert slot.shape == (2048, 64, 5, 6)
        root, bh, ch = orc.commit_fake_slot(seed, 2048, n_threads=4, want_cells=True)
        assert slot.read_layer(0, 0, 0, 2048) == ch
        assert slot.read_layer(0, 5, 0, 64) == bh == slot.read_layer(1, 0, 0, 64)
        layers = orc.merkle_layers(bh)
        for lvl, layer in enumerate(layers):
            assert slot.read_layer(1, lvl, 0, len(layer)) == layer
        assert str(slot.root) == meta["slotRoot"] == str(root)
        assert ctx.cell_indices(1234567, slot.root, 2048, 5) == meta["indices"]


@pytest.mark.parametrize("n_blocks", [1, 2, 3, 5, 6, 7, 13, 64, 100])
def test_slot_roots_ragged_block_counts(ctx, orc, n_blocks):
    """odd nodes at several levels, the one-block (key 3) slot tree, host-resident data"""
    rnd = random.Random(n_blocks)
    data = bytes(rnd.getrandbits(8) for _ in range(4096)) * 16 * n_blocks          # 64 KiB blocks, cheap to build
    data = bytearray(data)
    for b in range(n_blocks):                                                        # make every block distinct
        data[b * 65536:b * 65536 + 8] = b.to_bytes(8, "little")
    data = bytes(data)
    with ctx.slot_commit_host(data) as slot:
        root, bh, _ = orc.commit_slot(data, n_threads=4)
        assert slot.read_layer(1, 0, 0, n_blocks) == bh
        assert slot.root == root


@pytest.mark.parametrize("cell_size,block_size,n_cells", [(128, 4096, 256), (256, 4096, 64), (2048, 2048, 8), (64, 128, 10), (4096, 65536, 48),
                                                          (96, 192, 10), (32, 64, 6), (2048, 65536, 32 * 35), (160, 320, 70)])
def test_other_cell_and_block_sizes(ctx, orc, cell_size, block_size, n_cells):
    """testMain.hs small config (128/4096), one-cell blocks, two-cell blocks, bigger cells; TMA-staged rows with a
    partial last warp (rows past the end are zero-filled by the TMA unit), and geometries that take the plain-load
    kernel (cell size below 64 bytes)"""
    seed = 999
    with ctx.slot_commit_fake(seed, n_cells, cell_size, block_size) as slot:
        root, bh, ch = orc.commit_fake_slot(seed, n_cells, cell_size, block_size, want_cells=True)
        assert slot.read_layer(0, 0, 0, n_cells) == ch
        assert slot.read_layer(1, 0, 0, len(bh)) == bh
        assert slot.root == root


def test_adversarial_bytes(ctx, orc):
    """all-0xFF cells (every chunk >= 2^247), all-zero cells"""
    for fill in (b"\xff", b"\x00"):
        data = fill * (3 * 65536)
        with ctx.slot_commit_host(data) as slot:
            root, bh, ch = orc.commit_slot(data, want_cells=True)
            assert slot.read_layer(0, 0, 0, 96) == ch and slot.root == root


def test_dev_and_host_entry_points_agree(ctx):
    import torch
    n = 70 * 65536
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    ctx.fill_synthetic_dev(0xC0DE, 0, n, d.data_ptr())
    torch.cuda.synchronize()
    with ctx.slot_commit_dev(d.data_ptr(), n) as a:
        ra = a.root
    host = d.cpu().numpy()
    with ctx.slot_commit_host(host) as b:
        assert b.root == ra


def test_synthetic_fill_is_splitmix64(ctx):
    import torch
    from tools.gen_goldens import splitmix64_stream
    d = torch.empty(64, dtype=torch.uint8, device="cuda")
    ctx.fill_synthetic_dev(41, 100, 64, d.data_ptr())
    torch.cuda.synchronize()
    words = [int.from_bytes(bytes(d.cpu().numpy()[8 * i:8 * i + 8]), "little") for i in range(8)]
    # word i = mix(seed + first_word + i): the i-th output of a splitmix64 stream started at seed + first_word + i - 1 ... checked directly
    def mix(x):
        z = (x + 0x9e3779b97f4a7c15) & (2**64 - 1)
        z = ((z ^ (z >> 30)) * 0xbf58476d1ce4e5b9) & (2**64 - 1)
        z = ((z ^ (z >> 27)) * 0x94d049bb133111eb) & (2**64 - 1)
        return z ^ (z >> 31)
    assert words == [mix(41 + 100 + i) for i in range(8)]


# ---- paths -------------------------------------------------------------------------------------------------

def test_every_cell_path_of_a_ragged_slot(ctx, orc, pyorc):
    """5 blocks (odd nodes at two levels): every cell's merged path equals merkleProof(block tree) ++ merkleProof(slot tree),
    zero-padded (merkle.nim:21-42,86-100, types.nim:27-37), and reconstructs in two stages (Slot.hs:189-217)."""
    seed, n_cells = 4242, 5 * 32
    with ctx.slot_commit_fake(seed, n_cells) as slot:
        root, bh, ch = orc.commit_fake_slot(seed, n_cells, want_cells=True)
        big = orc.merkle_layers(bh)
        idx = list(range(n_cells))
        paths, leaves = slot.cell_paths(idx, 12)
        assert leaves == ch
        for i in idx:
            mini = orc.merkle_layers(ch[(i // 32) * 32:(i // 32 + 1) * 32])
            exp = pyorc.merkle_proof(mini, i % 32).merkle_path + pyorc.merkle_proof(big, i // 32).merkle_path
            assert paths[i] == exp + [0] * (12 - len(exp))
        i = 159                                                    # last cell: last block is an odd node at level 0 and 2
        assert paths[i][5] == 0
        blk = orc.reconstruct_root(leaves[i], i % 32, 32, paths[i][:5])
        assert orc.reconstruct_root(blk, i // 32, 5, paths[i][5:8]) == root


def test_batched_root_reconstruction(ctx, orc, pyorc):
    """cdx_reconstruct_roots_host == reconstructRoot (merkle.nim:51-74) for every leaf of trees with 1..20 leaves, and the
    two-stage check of padded cell paths (Slot.hs:189-217)"""
    for n in (1, 2, 3, 5, 8, 13, 20):
        layers = orc.merkle_layers([1000 + i for i in range(1, n + 1)])
        proofs = [pyorc.merkle_proof(layers, j) for j in range(n)]
        got = ctx.reconstruct_roots([p.leaf_value for p in proofs], list(range(n)), n, [p.merkle_path for p in proofs])
        assert got == [layers[-1][0]] * n
    seed, n_cells = 4242, 5 * 32
    with ctx.slot_commit_fake(seed, n_cells) as slot:
        idx = [0, 31, 32, 77, 159]
        paths, leaves = slot.cell_paths(idx, 32)
        block_roots = ctx.reconstruct_roots(leaves, [i % 32 for i in idx], 32, paths, depth=5)
        assert block_roots == [slot.read_layer(1, 0, i // 32, 1)[0] for i in idx]
        slot_roots = ctx.reconstruct_roots(block_roots, [i // 32 for i in idx], 5, [p[5:] for p in paths], depth=3)
        assert slot_roots == [slot.root] * len(idx)
        bad = [list(p) for p in paths]
        bad[2][1] ^= 1
        assert ctx.reconstruct_roots(leaves, [i % 32 for i in idx], 32, bad, depth=5)[2] != block_roots[2]


@pytest.mark.parametrize("n_cells,cell,block", [(5 * 32, 2048, 65536), (32, 2048, 65536), (8, 2048, 2048), (256, 128, 4096)])
def test_export_import_roundtrip(ctx, pkg, n_cells, cell, block):
    """a persisted commitment answers challenges exactly like the live one"""
    with ctx.slot_commit_fake(31337, n_cells, cell, block) as live:
        image = live.export()
        root, shape = live.root, live.shape
        idx = list(range(0, n_cells, max(1, n_cells // 7))) + [n_cells - 1]
        paths, leaves = live.cell_paths(idx, 24)
    with ctx.slot_import(image) as back:
        assert back.root == root and back.shape == shape
        assert back.cell_paths(idx, 24) == (paths, leaves)
        assert back.export() == image
    with pytest.raises(pkg.CodexCommitError):
        ctx.slot_import(image[:-32])
    with pytest.raises(pkg.CodexCommitError):
        ctx.slot_import(b"NOTASLOT" + image[8:])


def test_prove_batch_many_challenges(ctx, orc, pyorc, pkg):
    """cdx_slot_prove_batch == per challenge cellIndices (sample/bn254.nim:16-27) then merkleProof x2 / merge / pad
    (gen_input/bn254.nim:53-74), against the oracle; and the two-stage verifier accepts every answer"""
    seed, n_cells, n_samples, depth = 777, 8 * 32, 7, 32
    entropies = [1234567, 0, 1, R - 1, 0xdeadbeef << 200] + [random.Random(5).getrandbits(253) for _ in range(20)]
    with ctx.slot_commit_fake(seed, n_cells) as slot:
        root, bh, ch = orc.commit_fake_slot(seed, n_cells, want_cells=True)
        big = orc.merkle_layers(bh)
        assert slot.root == root
        idx, paths, leaves = slot.prove_batch(entropies, n_samples, depth)
        for k, e in enumerate(entropies):
            exp_idx = pyorc.cell_indices(e, root, n_cells, n_samples)
            assert idx[k] == exp_idx
            assert idx[k] == ctx.cell_indices(e, root, n_cells, n_samples)
            for c, i in enumerate(exp_idx):
                mini = orc.merkle_layers(ch[(i // 32) * 32:(i // 32 + 1) * 32])
                exp = pyorc.merkle_proof(mini, i % 32).merkle_path + pyorc.merkle_proof(big, i // 32).merkle_path
                assert paths[k][c] == exp + [0] * (depth - len(exp))
                assert leaves[k][c] == ch[i]
        flat_i = [i for row in idx for i in row]
        flat_p = [p for row in paths for p in row]
        flat_l = [v for row in leaves for v in row]
        blocks = ctx.reconstruct_roots(flat_l, [i % 32 for i in flat_i], 32, flat_p, depth=5)
        assert ctx.reconstruct_roots(blocks, [i // 32 for i in flat_i], 8, [p[5:] for p in flat_p], depth=3) == [root] * len(flat_i)
        assert slot.prove_batch([], n_samples, depth) == ([], [], [])
        with pytest.raises(pkg.CodexCommitError) as e:
            slot.prove_batch([1], 1, 7)                            # padMerkleProof: depth too small (types.nim:29)
        assert e.value.status == pkg.capi.CDX_ERR_RANGE
    with ctx.slot_commit_fake(seed, 5 * 32) as ragged:             # sampling needs a power-of-two cell count (sample/bn254.nim:19-20)
        with pytest.raises(pkg.CodexCommitError) as e:
            ragged.prove_batch([1], 1, depth)
        assert e.value.status == pkg.capi.CDX_ERR_NOT_POW2


def test_paths_errors(ctx, pkg):
    with ctx.slot_commit_fake(1, 64) as slot:
        with pytest.raises(pkg.CodexCommitError) as e:
            slot.cell_paths([64], 32)                              # index out of range (merkle.nim:27)
        assert e.value.status == pkg.capi.CDX_ERR_RANGE
        with pytest.raises(pkg.CodexCommitError) as e:
            slot.cell_paths([0], 5)                                # padMerkleProof: depth too small (types.nim:29)
        assert e.value.status == pkg.capi.CDX_ERR_RANGE


# ---- error behaviour (the reference asserts; the ABI returns status codes) -----------------------------------

def test_multi_gpu_and_dataset_entry_point_errors(ctx, pkg):
    """status codes of the round-2 entry points: every misuse is an error return with a message, never a crash"""
    E, capi = pkg.CodexCommitError, pkg.capi
    small = 2 * 65536
    with pytest.raises(E) as e:
        ctx.dataset_commit(None, [(capi.SRC_SYNTHETIC, 1, small)], keep_slot=1)        # keep_slot outside the dataset
    assert e.value.status == capi.CDX_ERR_RANGE
    with pytest.raises(E) as e:
        ctx.dataset_commit(None, [(capi.SRC_SYNTHETIC, 1, small + 100)])                 # not a whole number of blocks
    assert e.value.status == capi.CDX_ERR_SIZE
    with pytest.raises(E) as e:
        ctx.dataset_commit(None, [(7, 1, small)])                                        # unknown source kind
    assert e.value.status == capi.CDX_ERR_ARG
    with ctx.dataset_commit(None, [(capi.SRC_SYNTHETIC, 1, small), (capi.SRC_FAKE, 2, 3 * 65536)], keep_slot=1) as ds:
        with pytest.raises(E) as e:
            ds.prove(5, 4, 32)                                                           # 96 cells: not a power of two (sample/bn254.nim:19-20)
        assert e.value.status == capi.CDX_ERR_NOT_POW2
        with pytest.raises(E) as e:
            ds.slot_proof(2, 8)                                                          # slot index out of range
        assert e.value.status == capi.CDX_ERR_RANGE
        with pytest.raises(E) as e:
            ds.slot_proof(0, 0)                                                          # padMerkleProof: depth too small (types.nim:29)
        assert e.value.status == capi.CDX_ERR_RANGE
    with ctx.dataset_commit(None, [(capi.SRC_SYNTHETIC, 1, small)]) as ds:               # nothing kept
        with pytest.raises(E) as e:
            ds.prove(5, 4, 32)
        assert e.value.status == capi.CDX_ERR_STATE
    with pytest.raises(E) as e:
        ctx.comm_init(2, 2, b"\0" * 128)                                                 # rank outside 0..n-1
    assert e.value.status == capi.CDX_ERR_ARG
    with pytest.raises(E) as e:
        ctx.comm_init(2, 0, None)                                                        # several ranks need the id
    assert e.value.status == capi.CDX_ERR_STATE
    comm = ctx.comm_init(1, 0, None)
    d = bytes(8 * 65536)
    with pytest.raises(E) as e:
        ctx.slot_commit_sharded_host(comm, d, len(d), 2048, 65536, 3, 64, 2)            # first block not aligned to 2^top_level
    assert e.value.status == capi.CDX_ERR_SIZE
    with pytest.raises(E) as e:
        ctx.slot_commit_sharded_host(comm, d, len(d), 2048, 65536, 60, 64, 0)           # range runs past the end of the slot
    assert e.value.status == capi.CDX_ERR_RANGE
    with pytest.raises(E) as e:
        ctx.slots_commit_batch_host(d, [65536, 0])                                       # an empty slot in a batch
    assert e.value.status == capi.CDX_ERR_SIZE
    with pytest.raises(E):
        capi.plan_block_ranges(0, 4)
    comm.destroy()


def test_size_and_power_of_two_errors(ctx, pkg):
    E = pkg.CodexCommitError
    with pytest.raises(E) as e:
        ctx.slot_commit_host(bytes(65536 + 2048))                  # not a multiple of the block size
    assert e.value.status == pkg.capi.CDX_ERR_SIZE
    with pytest.raises(E) as e:
        ctx.slot_commit_host(bytes(3 * 2048), 2048, 3 * 2048)      # 3 cells per block: not a power of two
    assert e.value.status == pkg.capi.CDX_ERR_SIZE
    with pytest.raises(E) as e:
        ctx.slot_commit_host(bytes(4096), 2048, 3000)              # block not divisible by cell (types.nim:122)
    assert e.value.status == pkg.capi.CDX_ERR_SIZE
    with pytest.raises(E) as e:
        ctx.cell_indices(1, 2, 1000, 5)                            # sample/bn254.nim:19-20
    assert e.value.status == pkg.capi.CDX_ERR_NOT_POW2
    with pytest.raises(E) as e:
        ctx.merkle_root([])
    assert e.value.status == pkg.capi.CDX_ERR_ARG
    with pytest.raises(E) as e:
        ctx.sponge([1, 2], 3)
    assert e.value.status == pkg.capi.CDX_ERR_ARG


# ---- the host mirror of proof_input: cli -> input.json --------------------------------------------------------

def run_cli(args, tmp_path):
    cli = os.path.join(ROOT, "codex-storage-proofs-circuits_b200", "cli")
    assert os.path.exists(cli), "build the host cli first (__graft_entry__.build())"
    out = str(tmp_path / "input.json")
    res = subprocess.run([cli] + args + ["--output=" + out], capture_output=True, text=True)
    return res, out


def test_cli_config1_input_json_byte_exact(tmp_path):
    """BASELINE config 1 through the reference's flag surface (workflow/cli_args.sh:7-18)"""
    from oracle import circuit_verifier as cv
    args = "--field=bn254 --hash=poseidon2 --cellsize=2048 --blocksize=65536 --ncells=2048 --nslots=11 --index=3 --nsamples=5 " \
           "--seed=12345 --entropy=1234567 --depth=32 --maxslots=256".split()
    res, out = run_cli(args, tmp_path)
    assert res.returncode == 0, res.stderr
    txt = open(out).read()
    assert txt == open(os.path.join(GOLDEN, "input_config1.json")).read()
    cv.verify_input_json(txt, 32, 8, 2048, 65536)


def test_cli_small_config_and_short_flags(tmp_path):
    from oracle import circuit_verifier as cv
    args = "-F=bn254 -H:poseidon2 -c=128 -b=4096 -K=256 -s=5 -i=3 -n=10 -S=12345 -e=1234567 -d=16 -N=32".split()
    res, out = run_cli(args, tmp_path)
    assert res.returncode == 0, res.stderr
    txt = open(out).read()
    assert txt == open(os.path.join(GOLDEN, "input_small.json")).read()
    cv.verify_input_json(txt, 16, 5, 128, 4096)


def test_cli_slot_file_source(tmp_path, ctx):
    """SlotFile data source (slot.nim:57-68, dataset.nim:34): <base><k>.dat per slot; same result as the fake source it was dumped from"""
    from oracle import circuit_verifier as cv
    n_slots, n_cells = 3, 64
    for k in range(n_slots):
        with open(tmp_path / f"slotdata{k}.dat", "wb") as f:
            f.write(ctx.fake_cells(12345 + 72 + 1001 * k, 0, n_cells, 2048))
    common = f"--field=bn254 --ncells={n_cells} --nslots={n_slots} --index=1 --nsamples=4 --entropy=99".split()
    res_a, out_a = run_cli(common + [f"--file={tmp_path}/slotdata"], tmp_path)
    assert res_a.returncode == 0, res_a.stderr
    a = open(out_a).read()
    res_b, out_b = run_cli(common + ["--seed=12345"], tmp_path)
    assert res_b.returncode == 0, res_b.stderr
    assert a == open(out_b).read()
    cv.verify_input_json(a, 32, 8, 2048, 65536)


def test_cli_selfcheck(tmp_path):
    res, out = run_cli("--field=bn254 --ncells=512 --nslots=13 --index=7 --nsamples=20 --seed=666 --selfcheck".split(), tmp_path)
    assert res.returncode == 0, res.stderr
    assert "selfcheck: the proof input satisfies every constraint" in res.stdout
    from oracle import circuit_verifier as cv
    cv.verify_input_json(open(out).read(), 32, 8, 2048, 65536)


def test_cli_error_behaviour(tmp_path):
    res, _ = run_cli(["--field=bn254", "--ncells=1000"], tmp_path)           # checkPowerOfTwo (cli.nim:143, misc.nim:29-32)
    assert res.returncode != 0 and "expected to be a power of 2" in res.stderr
    res, _ = run_cli([], tmp_path)                                           # default field is goldilocks (cli.nim:47-51)
    assert res.returncode != 0 and "bn254" in res.stderr
    res, _ = run_cli(["--field=bn254", "--nslots=4", "--index=9"], tmp_path)
    assert res.returncode != 0 and "out of range" in res.stderr
