"""The N > 1 path on CPU: world_size-2 gloo run of the block-range sharding PROTOCOL the library implements over NCCL
(csrc/capi_multi.cuh: cdx_slot_exchange_top, prove_core), with an oracle-backed stand-in for the per-rank GPU slot and
torch.distributed/gloo standing in for NCCL.  What is under test is everything that does not depend on the device: range
planning (Python twin == the C planner), the global odd-node rule at the ragged tail, the exchange -- every rank places its
level-T nodes at their global positions in a zeroed level and ONE byte-wise SUM combines them --, the replicated top tree,
owner-only path assembly combined by the same byte-wise SUM, empty shards, and the communicator bootstrap
(sharded.comm_from_torch) up to the point where it needs a GPU."""
import importlib
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG, ROOT

R = 21888242871839275222246405745257275088548364400416034343698204186575808495617


class OracleShard:
    """slot-like stand-in: commits blocks [first, first+count) of a fake-data slot with the C oracle, builds levels
    0..T of the slot tree for that range with the GLOBAL odd-node rule (merkle/bn254.nim:38-53)."""

    def __init__(self, orc, seed, first_block, n_blocks, n_total_blocks, top_level):
        self.orc, self.first, self.n, self.total, self.T = orc, first_block, n_blocks, n_total_blocks, top_level
        self.cell_hashes, self.block_trees = [], []
        for b in range(first_block, first_block + n_blocks):
            ch = [orc.hash_bytes(orc.gen_fake_cell(seed, 32 * b + c, 2048)) for c in range(32)]
            self.cell_hashes += ch
            self.block_trees.append(orc.merkle_layers(ch))
        self.low = [[t[-1][0] for t in self.block_trees]]
        for l in range(top_level):
            cur, nxt = self.low[l], []
            for i in range((len(cur) + 1) // 2):
                key = 1 if l == 0 else 0
                if 2 * i + 1 < len(cur):
                    nxt.append(orc.compress(cur[2 * i], cur[2 * i + 1], key))
                else:
                    nxt.append(orc.compress(cur[2 * i], 0, key + 2))
            self.low.append(nxt)
        self.top = None

    def subtree_roots_tensor(self, device="cpu"):
        raw = b"".join(int(v).to_bytes(32, "little") for v in self.low[self.T])
        if not raw:                                      # empty shard
            return torch.zeros(0, dtype=torch.uint8)
        return torch.frombuffer(bytearray(raw), dtype=torch.uint8).clone()

    def set_top_tensor(self, t):
        raw = bytes(t.numpy())
        nodes = [int.from_bytes(raw[i:i + 32], "little") for i in range(0, len(raw), 32)]
        self.top = self.orc.merkle_layers(nodes, bottom=(self.T == 0))

    @property
    def root(self):
        return self.top[-1][0]

    def cell_paths(self, indices, max_depth):
        paths, leaves = [], []
        for ci in indices:
            b = ci // 32
            if not (self.first <= b < self.first + self.n):
                paths.append([0] * max_depth)
                leaves.append(0)
                continue
            lb = b - self.first
            path, k = [], ci % 32
            for l in range(5):
                path.append(self.block_trees[lb][l][k ^ 1])
                k >>= 1
            node, width = b, self.total
            n_levels = (len(self.top) - 1) + self.T
            for l in range(n_levels):
                sib = node ^ 1
                if sib >= width:
                    path.append(0)
                elif l < self.T:
                    path.append(self.low[l][sib - (self.first >> l)])
                else:
                    path.append(self.top[l - self.T][sib])
                node >>= 1
                width = (width + 1) // 2
            paths.append(path + [0] * (max_depth - len(path)))
            leaves.append(self.cell_hashes[32 * lb + ci % 32])
        return paths, leaves


def exchange_top_model(shard, n_total_blocks, top_level):
    """cdx_slot_exchange_top, line by line: zeroed level T, own nodes in place, byte-wise SUM over the ranks, top tree"""
    sharded = importlib.import_module(PKG + ".sharded")
    width = sharded.level_width(n_total_blocks, top_level)
    lvl = torch.zeros(32 * width, dtype=torch.uint8)
    own = shard.subtree_roots_tensor()
    first_node = shard.first >> top_level
    lvl[32 * first_node:32 * first_node + own.numel()] = own
    dist.all_reduce(lvl, op=dist.ReduceOp.SUM)
    shard.set_top_tensor(lvl)


def gather_paths_model(shard, indices, max_depth):
    """prove_core's combine: the owner fills a cell's path and leaf, everyone else zeros, byte-wise SUM"""
    paths, leaves = shard.cell_paths(list(indices), max_depth)
    flat = b"".join(int(v).to_bytes(32, "little") for p in paths for v in p) + b"".join(int(v).to_bytes(32, "little") for v in leaves)
    t = torch.frombuffer(bytearray(flat), dtype=torch.uint8).clone()
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    raw = bytes(t.numpy())
    n = len(indices)
    vals = [int.from_bytes(raw[i:i + 32], "little") for i in range(0, len(raw), 32)]
    return [vals[i * max_depth:(i + 1) * max_depth] for i in range(n)], vals[n * max_depth:]


def _worker(rank, world, port, n_total_blocks, seed, out_q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sharded = importlib.import_module(PKG + ".sharded")
        from oracle import coracle as orc
        top_level, ranges = sharded.plan_block_ranges(n_total_blocks, world, max_imbalance=0.35)
        first, count = ranges[rank]
        shard = OracleShard(orc, seed, first, count, n_total_blocks, top_level)       # count == 0: an empty shard
        exchange_top_model(shard, n_total_blocks, top_level)
        indices = sorted({0, 31, 32 * (n_total_blocks // 2), 32 * n_total_blocks - 1, 32 * (n_total_blocks - 1)})
        paths, leaves = gather_paths_model(shard, indices, 16)
        out_q.put((rank, top_level, ranges, shard.root, indices, paths, leaves))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_total_blocks,world", [(13, 2), (16, 2), (21, 2), (1, 2), (21, 3), (2, 3)])
def test_two_rank_sharded_commit_matches_single_process(orc, n_total_blocks, world):
    seed = 777
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + n_total_blocks + 40 * world
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total_blocks, seed, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    root, bh, ch = orc.commit_fake_slot(seed, 32 * n_total_blocks, want_cells=True)
    big = orc.merkle_layers(bh)
    depth = len(big) - 1
    for rank, top_level, ranges, r_root, indices, paths, leaves in results:
        assert sum(c for _, c in ranges) == n_total_blocks and ranges[0][0] == 0
        assert all(f % (1 << top_level) == 0 for f, c in ranges if c)
        assert r_root == root, f"rank {rank}: sharded root differs from the single-process root"
        for ci, path, leaf in zip(indices, paths, leaves):
            assert leaf == ch[ci]
            blk = orc.reconstruct_root(leaf, ci % 32, 32, path[:5])
            assert blk == bh[ci // 32]
            assert orc.reconstruct_root(blk, ci // 32, n_total_blocks, path[5:5 + depth]) == root
            assert all(v == 0 for v in path[5 + depth:])


def test_range_planning_properties(pkg):
    sharded = importlib.import_module(PKG + ".sharded")
    for n_blocks, world in [(1638400, 8), (163840, 1), (163840, 3), (5, 8), (1, 2), (2097152, 8), (1000, 7), (1, 1), (3, 2), (655360, 4)]:
        t, ranges = sharded.plan_block_ranges(n_blocks, world)
        assert (t, ranges) == pkg.capi.plan_block_ranges(n_blocks, world)      # the C planner (cdx_plan_block_ranges) is the one that ships
        assert pkg.capi.block_ranges_top_level(n_blocks, ranges) >= t
        assert len(ranges) == world and sum(c for _, c in ranges) == n_blocks
        pos = 0
        for f, c in ranges:
            assert f == pos or c == 0
            if c:
                assert f % (1 << t) == 0 and (c % (1 << t) == 0 or f + c == n_blocks)
            pos += c
        counts = [c for _, c in ranges]
        if n_blocks >= 64 * world:
            assert max(counts) <= 1.01 * n_blocks / world + 1          # the most loaded rank is within 1 % of ideal
    t, ranges = sharded.plan_block_ranges(1638400, 8)              # BASELINE config 4: 100 GiB over 8 GPUs
    assert t == 13 and all(c == 204800 for _, c in ranges)         # 25 chunks of 8192 blocks per GPU
    t, ranges = sharded.fixed_ranges(163840, 8)                    # bench.py weak-scaling layout
    assert t == 15 and ranges[3] == (3 * 163840, 163840)
    assert pkg.capi.block_ranges_top_level(8 * 163840, ranges) == 15
    with pytest.raises(pkg.CodexCommitError):
        pkg.capi.block_ranges_top_level(100, [(0, 40), (50, 50)])     # a gap
    assert sharded.level_width(163840, 15) == 5 and sharded.level_width(5, 3) == 1


def test_dataset_planning_helpers(pkg):
    """host logic of the dataset path (BASELINE config 5): slot sizes, per-slot seeds, LPT packing -- no GPU"""
    import importlib
    dataset = importlib.import_module(pkg.__name__ + ".dataset")
    blocks = dataset.draw_slot_blocks(256, 1 << 30, 100 << 30, seed=12345, pow2_slot=3, pow2_blocks=1 << 17)
    assert len(blocks) == 256 and blocks[3] == 1 << 17
    assert all((1 << 30) // 65536 - 1 <= b <= (100 << 30) // 65536 for k, b in enumerate(blocks) if k != 3)
    assert blocks == dataset.draw_slot_blocks(256, 1 << 30, 100 << 30, seed=12345, pow2_slot=3, pow2_blocks=1 << 17)   # deterministic
    assert blocks != dataset.draw_slot_blocks(256, 1 << 30, 100 << 30, seed=12346, pow2_slot=3, pow2_blocks=1 << 17)
    # the reference's per-slot seed rule (dataset.nim:32)
    assert dataset.slot_seed(12345, 3) == 12345 + 72 + 3003
    for world in (1, 2, 4, 8):
        bins = dataset.lpt_assign(blocks, world)
        assert sorted(k for b in bins for k in b) == list(range(256))          # every slot exactly once
        loads = [sum(blocks[k] for k in b) for b in bins]
        assert max(loads) <= 1.02 * sum(blocks) / world                        # LPT on 256 items: within 2 % of even
    assert dataset.lpt_assign([5], 4) == [[0], [], [], []]
    # the plan the library follows (cdx_dataset_plan): without sharded members it is the LPT packing above
    for world in (1, 2, 4, 8):
        owner = pkg.capi.dataset_plan([b * 65536 for b in blocks], world)
        assert all(o >= 0 for o in owner)                                       # 256 slots of <= 100 GiB: nothing exceeds a quarter of a share
        bins = dataset.lpt_assign(blocks, world)
        assert owner == [next(r for r in range(world) if k in bins[r]) for k in range(256)]
    # sharding rule: more than a quarter of the ideal share AND at least 256 MiB per rank
    gib = (1 << 30) // 65536
    assert pkg.capi.dataset_plan([100 * gib * 65536] * 4, 8) == [-1, -1, -1, -1]            # four 100 GiB slots on 8 GPUs: all sharded
    assert pkg.capi.dataset_plan([64 * 65536] * 11, 8) == [0, 1, 2, 3, 4, 5, 6, 7, 0, 1, 2]  # config 1: 4 MiB slots are never sharded
    plan = pkg.capi.dataset_plan([100 * gib * 65536] + [gib * 65536] * 20, 8)                # one dominant slot among small ones
    assert plan[0] == -1 and sorted(set(plan[1:])) == list(range(8))
    assert pkg.capi.dataset_plan([100 * gib * 65536], 1) == [0]                              # one rank: nothing to shard over
    with pytest.raises(pkg.CodexCommitError):
        pkg.capi.dataset_plan([65536 + 1], 2)


def _id_worker(rank, world, port, out_q):
    """the first half of sharded.comm_from_torch, up to the point where a GPU is needed: the library creates the NCCL id on
    rank 0 (cdx_comm_unique_id, dlopen of libnccl.so.2 -- no device involved), torch.distributed broadcasts the 128 bytes"""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        capi = importlib.import_module(PKG).capi
        t = torch.zeros(capi.COMM_ID_BYTES, dtype=torch.uint8)
        if rank == 0:
            t = torch.frombuffer(bytearray(capi.comm_unique_id()), dtype=torch.uint8).clone()
        dist.broadcast(t, src=0)
        out_q.put((rank, bytes(t.numpy())))
    finally:
        dist.destroy_process_group()


def test_communicator_id_bootstrap_over_gloo(pkg):
    try:
        pkg.capi.comm_unique_id()
    except pkg.CodexCommitError:
        pytest.skip("libnccl.so.2 is not loadable on this machine")
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_id_worker, args=(r, world, 29611, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert len(got[0]) == 128 and got[0] == got[1] and any(got[0])
