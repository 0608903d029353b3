"""GPU tests at sizes the oracle cannot re-hash in seconds: size-independent properties of the commitment
(sharded == unsharded, tiled host path == resident path, tree of read-back block hashes == root, every sampled path
reconstructs in two stages), plus the single-GPU emulation of the multi-rank exchange through the C ABI."""
import importlib
import random

import pytest

from conftest import PKG

pytestmark = pytest.mark.gpu
R = 21888242871839275222246405745257275088548364400416034343698204186575808495617


@pytest.fixture(scope="module")
def torch_mod():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


def synthetic(ctx, torch, n_bytes, seed=0xC0DE, first_word=0):
    d = torch.empty(n_bytes, dtype=torch.uint8, device="cuda")
    ctx.fill_synthetic_dev(seed, first_word, n_bytes, d.data_ptr())
    torch.cuda.synchronize()
    return d


def commit_in_ranges(ctx, torch, sharded, d, n_blocks, ranges, top_level):
    """what N ranks do, run back to back on one GPU: per-range commit, gather of the level-T nodes, replicated top"""
    shards = []
    for first, count in ranges:
        if count == 0:
            shards.append(None)
            continue
        ptr = d.data_ptr() + first * 65536
        shards.append(sharded.GpuShard(ctx.slot_commit_range_dev(ptr, count * 65536, 2048, 65536, first, n_blocks, top_level)))
    gathered = torch.cat([s.subtree_roots_tensor() for s in shards if s is not None]).contiguous()
    assert gathered.numel() == 32 * sharded.level_width(n_blocks, top_level)
    for s in shards:
        if s is not None:
            s.set_top_tensor(gathered)
    return shards


@pytest.mark.parametrize("n_blocks,world", [(16384, 4), (1000, 3), (163, 8), (25 * 64, 8)])
def test_sharded_commit_equals_whole_commit(ctx, torch_mod, n_blocks, world):
    torch = torch_mod
    sharded = importlib.import_module(PKG + ".sharded")
    d = synthetic(ctx, torch, n_blocks * 65536)
    with ctx.slot_commit_dev(d.data_ptr(), n_blocks * 65536) as whole:
        root = whole.root
        top_level, ranges = sharded.plan_block_ranges(n_blocks, world, max_imbalance=0.05)
        shards = commit_in_ranges(ctx, torch, sharded, d, n_blocks, ranges, top_level)
        live = [s for s in shards if s is not None]
        assert all(s.root == root for s in live)
        # paths: the owner produces the whole path, everyone else zeros -> the SUM over ranks is the path
        rnd = random.Random(n_blocks)
        cells = [rnd.randrange(32 * n_blocks) for _ in range(20)] + [0, 32 * n_blocks - 1]
        ref_paths, ref_leaves = whole.cell_paths(cells, 32)
        acc = [[0] * 32 for _ in cells]
        acc_leaf = [0] * len(cells)
        for s in live:
            p, l = s.cell_paths(cells, 32)
            for i in range(len(cells)):
                acc[i] = [a + b for a, b in zip(acc[i], p[i])]
                acc_leaf[i] += l[i]
        assert acc == ref_paths and acc_leaf == ref_leaves
        for s in live:
            s.free()


def test_one_gib_slot_properties(ctx, orc, torch_mod):
    torch = torch_mod
    n_blocks = 16384
    d = synthetic(ctx, torch, n_blocks * 65536)
    with ctx.slot_commit_dev(d.data_ptr(), n_blocks * 65536) as slot:
        n_cells, nb, bd, sd = slot.shape
        assert (n_cells, nb, bd, sd) == (524288, 16384, 5, 14)
        root = slot.root
        # a checksum of checksums: the tree over the read-back block hashes is the root (GPU tree API and oracle)
        bh = slot.read_layer(1, 0, 0, n_blocks)
        assert ctx.merkle_root(bh) == root == orc.merkle_root(bh)
        # spot-check cell hashes and one block tree against the oracle
        host = d[: 2 * 65536].cpu().numpy().tobytes()
        assert slot.read_layer(0, 0, 0, 64) == [orc.hash_bytes(host[i * 2048:(i + 1) * 2048]) for i in range(64)]
        tail = d[(n_blocks - 1) * 65536:].cpu().numpy().tobytes()
        assert slot.read_layer(1, 0, n_blocks - 1, 1)[0] == orc.merkle_root([orc.hash_bytes(tail[i * 2048:(i + 1) * 2048]) for i in range(32)])
        # 100 sampled paths (BASELINE config 5 shape) all reconstruct to the root in two stages (Slot.hs:189-217)
        idx = ctx.cell_indices(987654321, root, n_cells, 100)
        assert idx == [orc.cell_index(987654321, root, n_cells, c) for c in range(1, 101)]
        paths, leaves = slot.cell_paths(idx, 32)
        for i, p, leaf in zip(idx, paths, leaves):
            blk = orc.reconstruct_root(leaf, i % 32, 32, p[:5])
            assert orc.reconstruct_root(blk, i // 32, n_blocks, p[5:19]) == root
            assert all(v == 0 for v in p[19:])
    # the tiled host-buffer path (4 x 256 MiB tiles, copy overlapped with the sponge) gives the same root
    host_all = d.cpu().numpy()
    with ctx.slot_commit_host(host_all) as slot2:
        assert slot2.root == root
    # the same bytes from pinned memory (direct async copies instead of the pinned-chunk pipeline pageable memory takes)
    pinned = torch.empty(n_blocks * 65536, dtype=torch.uint8, pin_memory=True)
    pinned.copy_(d)
    torch.cuda.synchronize()
    with ctx.slot_commit_host(pinned.data_ptr(), n_bytes=n_blocks * 65536) as slot3:
        assert slot3.root == root
    # two ranks' worth of pageable host ranges (512 MiB each: the chunk pipeline with a non-zero first block), exchanged
    sharded = importlib.import_module(PKG + ".sharded")
    half = n_blocks // 2
    shards = [sharded.GpuShard(ctx.slot_commit_range_host(host_all[k * half * 65536:(k + 1) * half * 65536], 2048, 65536, k * half, n_blocks, 13))
              for k in range(2)]
    gathered = torch.cat([s.subtree_roots_tensor() for s in shards]).contiguous()
    for s in shards:
        s.set_top_tensor(gathered)
    assert [s.root for s in shards] == [root, root]
    for s in shards:
        s.free()


def test_idempotence_and_sensitivity(ctx, torch_mod):
    torch = torch_mod
    n = 300 * 65536
    d = synthetic(ctx, torch, n)
    with ctx.slot_commit_dev(d.data_ptr(), n) as a, ctx.slot_commit_dev(d.data_ptr(), n) as b:
        assert a.root == b.root
        ra = a.root
    d[n - 1] ^= 1                                                    # flip one bit of the very last byte
    torch.cuda.synchronize()
    with ctx.slot_commit_dev(d.data_ptr(), n) as c:
        assert c.root != ra


def test_permutation_batch_2_17_vs_oracle(ctx, orc, torch_mod):
    """BASELINE config 2 at 2^17 states fully checked (2^20 is timed by tools/first_light.py)"""
    torch = torch_mod
    n = 1 << 17
    a = synthetic(ctx, torch, 96 * n, seed=1)
    b = torch.empty_like(a)
    ctx.permutation_batch_dev(a.data_ptr(), b.data_ptr(), n)
    torch.cuda.synchronize()
    inp = a.cpu().numpy().tobytes()
    assert b.cpu().numpy().tobytes() == orc.permutation_batch_bytes(inp)     # inputs are arbitrary 256-bit values: taken mod r


def test_dataset_commit_small_vs_oracle(ctx, orc, torch_mod):
    """BASELINE config 5 in miniature: 13 slots of mixed size (odd nodes in the dataset tree), slot 3 sampled"""
    dataset = importlib.import_module(PKG + ".dataset")
    blocks = dataset.draw_slot_blocks(13, 3 * 65536, 40 * 65536, seed=99, pow2_slot=3, pow2_blocks=16)
    assert blocks[3] == 16 and len(set(blocks)) > 5
    res = dataset.commit_dataset(ctx, blocks, 99, 3, 1234567, 20, max_depth=32, max_log2_nslots=8)
    # every slot root against the oracle over the same synthetic bytes (bench.py's numpy twin of the device generator)
    import bench
    for k in (0, 3, 7, 12):
        data = bench.synthetic_bytes_host(dataset.slot_seed(99, k), 0, blocks[k] * 65536)
        root, _, _ = orc.commit_slot((data.ctypes.data, blocks[k] * 65536), n_threads=4)
        assert res.slot_roots[k] == root
    layers = orc.merkle_layers(res.slot_roots)
    assert res.dataset_layers == layers and res.dataset_root == layers[-1][0]
    assert orc.reconstruct_root(res.slot_roots[3], 3, 13, res.slot_proof[:len(layers) - 1]) == res.dataset_root
    assert res.slot_proof[len(layers) - 1:] == [0] * (8 - (len(layers) - 1))
    n_cells = 16 * 32
    assert res.cell_indices == [orc.cell_index(1234567, res.slot_roots[3], n_cells, c) for c in range(1, 21)]
    for ci, path, leaf in zip(res.cell_indices, res.merkle_paths, res.cell_hashes):
        blk = orc.reconstruct_root(leaf, ci % 32, 32, path[:5])
        assert orc.reconstruct_root(blk, ci // 32, 16, path[5:9]) == res.slot_roots[3]
    bins = dataset.lpt_assign(blocks, 4)
    assert sorted(k for b in bins for k in b) == list(range(13))
    loads = [sum(blocks[k] for k in b) for b in bins]
    assert max(loads) <= 1.34 * sum(blocks) / 4


def test_commit_from_file(ctx, torch_mod, tmp_path):
    """SlotFile at speed: pread -> pinned -> H2D -> sponge; equals the resident commitment; a short file
    reads as zeros past its end (slot.nim:64-65); offsets select a window"""
    import time
    torch = torch_mod
    n = 700 * 65536                                                      # 43.75 MiB ... plus a multi-tile case below
    d = synthetic(ctx, torch, n)
    host = d.cpu().numpy()
    path = str(tmp_path / "slot.dat")
    host.tofile(path)
    with ctx.slot_commit_dev(d.data_ptr(), n) as a, ctx.slot_commit_file(path, n) as b:
        assert a.root == b.root
    short = 600 * 65536 + 1234
    host[:short].tofile(path)
    padded = host.copy()
    padded[short:] = 0
    with ctx.slot_commit_host(padded) as a, ctx.slot_commit_file(path, n) as b:
        assert a.root == b.root
    host.tofile(path)
    with ctx.slot_commit_dev(d.data_ptr() + 100 * 65536, 64 * 65536) as a, ctx.slot_commit_file(path, 64 * 65536, offset=100 * 65536) as b:
        assert a.root == b.root
    big = 9 * 1024 * 65536 + 3 * 65536                                  # 579 MiB: three device tiles, ten pinned chunks, ragged tail
    d2 = synthetic(ctx, torch, big, seed=7)
    path2 = str(tmp_path / "big.dat")
    d2.cpu().numpy().tofile(path2)
    with ctx.slot_commit_dev(d2.data_ptr(), big) as a:
        ra = a.root
    t0 = time.perf_counter()
    with ctx.slot_commit_file(path2, big) as b:
        dt = time.perf_counter() - t0
        assert b.root == ra
    print(f"commit_file {big / 2**20:.0f} MiB from page cache: {big / dt / 1e9:.2f} GB/s")
