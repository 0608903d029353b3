"""The oracle against every stored vector the reference has for this path (one: the permutation KAT,
reference/haskell/src/Poseidon2/Example.hs:13-22), against the frozen goldens, and against its own independent
twins (pure-Python restatement; circom-semantics verifier).  CPU only."""
import json
import os
import random

import pytest

from conftest import GOLDEN

R = 21888242871839275222246405745257275088548364400416034343698204186575808495617


def test_permutation_kat_all_three_statements(orc, pyorc, vectors):
    from oracle import circuit_verifier as cv
    kat = vectors["permutation_kat"]
    exp = tuple(int(v) for v in kat["out"])
    assert exp[0] == 0x30610a447b7dec194697fb50786aa7421494bd64c221ba4d3b1af25fb07bd103   # literal from Example.hs
    assert orc.permutation((0, 1, 2)) == exp
    assert pyorc.permutation((0, 1, 2)) == exp
    assert tuple(cv.circom_permutation([0, 1, 2])) == exp


def test_round_constants_tables_consistent(pyorc):
    from oracle import poseidon2_rc as rc
    assert len(rc.RC_EXT) == 8 and all(len(t) == 3 for t in rc.RC_EXT) and len(rc.RC_INT) == 56
    assert all(0 < c < R for t in rc.RC_EXT for c in t) and all(0 < c < R for c in rc.RC_INT)
    # first/last constants as printed in RoundConsts.hs:32 and poseidon2_perm.circom:28
    assert rc.RC_EXT[0][0] == 0x2c4c51fd1bb9567c27e99f5712b49e0574178b41b6f0a476cddc41d242cf2b43
    assert rc.RC_INT[0] == 0x15ce7e5ae220e8623a40b3a3b22d441eff0c9be1ae1d32f1b777af84eea7e38c


def test_golden_permutations(orc, vectors):
    for case in vectors["permutations"]:
        assert orc.permutation([int(v) for v in case["in"]]) == tuple(int(v) for v in case["out"])


def test_golden_testvector_suite(orc, vectors):
    """shape of reference/nim/testvectors/src/testvectors.nim:20-72"""
    for n in range(9):
        xs = list(range(1, n + 1))
        assert str(orc.sponge1(xs)) == vectors["sponge_rate1"][n]
        assert str(orc.sponge2(xs)) == vectors["sponge_rate2"][n]
    for n in range(81):
        b = bytes(range(1, n + 1))
        assert str(orc.hash_bytes(b)) == vectors["hash_bytes"][n]
        assert str(orc.merkle_root(orc.bytes_to_elements(b))) == vectors["merkle_root_bytes"][n]
    for n in range(1, 41):
        assert str(orc.merkle_root(list(range(1, n + 1)))) == vectors["merkle_root_felts"][n - 1]
    for case in vectors["compress"]:
        assert str(orc.compress(int(case["x"]), int(case["y"]), case["key"])) == case["out"]


def test_python_twin_agrees_with_c_oracle(orc, pyorc):
    rnd = random.Random(11)
    for _ in range(10):
        s = tuple(rnd.randrange(R) for _ in range(3))
        assert orc.permutation(s) == pyorc.permutation(s)
    for n in (0, 1, 30, 31, 32, 61, 62, 63, 2048):
        d = bytes(rnd.randrange(256) for _ in range(n))
        assert orc.bytes_to_elements(d) == pyorc.bytes_to_elements(d)
    d = bytes(rnd.randrange(256) for _ in range(200))
    assert orc.hash_bytes(d) == pyorc.hash_bytes(d)
    for n in (1, 2, 3, 5, 8):
        xs = [rnd.randrange(R) for _ in range(n)]
        assert orc.merkle_layers(xs) == pyorc.merkle_layers(xs)
        assert orc.merkle_layers(xs, False) == pyorc.merkle_layers(xs, False)
    assert orc.gen_fake_cell(15420, 7, 256) == pyorc.gen_fake_cell(15420, 7, 256)


def test_bytes_to_elements_edge_cases(orc):
    """10* byte padding is unconditional: a 31k-byte input gives k+1 elements (Slot.hs:243-250)."""
    assert orc.bytes_to_elements(b"") == [1]
    assert orc.bytes_to_elements(b"\x00" * 31) == [0, 1]
    assert orc.bytes_to_elements(b"\xff" * 30) == [int.from_bytes(b"\xff" * 30 + b"\x01", "little")]
    e = orc.bytes_to_elements(b"\xff" * 2048)
    assert len(e) == 67 and e[0] == (1 << 248) - 1 and e[66] == 0xffff | (1 << 16)


def test_merkle_conventions(orc):
    """singleton = one key-3 compression; odd node = compress(x, 0, key|2); three bottom flags (Merkle.hs:69-83,171-189)."""
    assert orc.merkle_root([5]) == orc.compress(5, 0, 3)
    a = orc.compress(1, 2, 1)
    b = orc.compress(3, 0, 3)
    assert orc.merkle_root([1, 2, 3]) == orc.compress(a, b, 0)
    l5 = orc.merkle_layers([1, 2, 3, 4, 5])
    assert [len(l) for l in l5] == [5, 3, 2, 1]
    assert l5[2][1] == orc.compress(l5[1][2], 0, 2)
    assert orc.merkle_layers([7], False) == [[7]]          # non-bottom singleton is its own root (merkle/bn254.nim:34-36)


def test_all_leaves_proof_roundtrip(orc, pyorc):
    """testAllMerkleProofs (Merkle.hs:136-152): every leaf of trees with 1..24 leaves, inputs 1001.."""
    for n in range(1, 25):
        layers = orc.merkle_layers([1000 + i for i in range(1, n + 1)])
        root = layers[-1][0]
        for j in range(n):
            p = pyorc.merkle_proof(layers, j)
            assert orc.reconstruct_root(p.leaf_value, j, n, p.merkle_path) == root
            if j ^ 1 >= n:
                assert p.merkle_path[0] == 0                # out-of-range sibling is zero (merkle.nim:34)


def test_fake_data_golden(orc, vectors):
    import hashlib
    fk = vectors["fake_cell_sha256"]
    cell = orc.gen_fake_cell(fk["seed"], fk["idx"], fk["cell_size"])
    assert cell[:16].hex() == fk["first16"] == "83e5a6518375c7e551ac4ab6ec268822"
    assert hashlib.sha256(cell).hexdigest() == fk["sha256"]


def test_cell_hash_goldens(orc, vectors):
    cells = {"zeros": bytes(2048), "ones_ff": b"\xff" * 2048, "ramp": bytes(i & 255 for i in range(2048)),
             "fake_seed15420_cell0": orc.gen_fake_cell(15420, 0, 2048)}
    for k, v in cells.items():
        assert str(orc.hash_bytes(v)) == vectors["cell_hashes"][k]


def test_config1_slot_and_input_json(orc, pyorc):
    """BASELINE config 1: slot 3 of the 11-slot dataset; root, indices, and circuit acceptance of the golden JSON."""
    from oracle import circuit_verifier as cv
    meta = json.load(open(os.path.join(GOLDEN, "meta.json")))["config1"]
    root, bh, _ = orc.commit_fake_slot(pyorc.parametric_slot_seed(12345, 3), 2048, n_threads=4)
    assert str(root) == meta["slotRoot"]
    assert len(bh) == 64 and orc.merkle_root(bh) == root
    assert [orc.cell_index(1234567, root, 2048, c) for c in range(1, 6)] == meta["indices"] == [1839, 1819, 1754, 1592, 834]
    txt = open(os.path.join(GOLDEN, "input_config1.json")).read()
    js = json.loads(txt)
    assert js["slotRoot"] == meta["slotRoot"] and js["dataSetRoot"] == meta["dataSetRoot"]
    assert list(js.keys()) == ["dataSetRoot", "entropy", "nCellsPerSlot", "nSlotsPerDataSet", "slotIndex", "slotRoot",
                               "slotProof", "cellData", "merklePaths"]          # json/bn254.nim:61-73 order
    assert len(js["slotProof"]) == 8 and all(len(c) == 67 for c in js["cellData"]) and all(len(p) == 32 for p in js["merklePaths"])
    cv.verify_input_json(txt, 32, 8, 2048, 65536)


def test_small_config_end_to_end_pure_python(pyorc):
    """reference/haskell/cli/testMain.hs:12-24 shape, pure Python, compared with the frozen file byte for byte."""
    from oracle import circuit_verifier as cv
    g = pyorc.GlobalConfig(16, 5, 128, 4096)
    d = pyorc.DataSetConfig(5, 256, 10, 12345)
    inp = pyorc.generate_proof_input(g, d, 3, 1234567)
    txt = pyorc.export_proof_input(inp)
    assert txt == open(os.path.join(GOLDEN, "input_small.json")).read()
    assert [p.leaf_index for p in inp.merkle_proofs] == [56, 171, 70, 132, 14, 56, 74, 187, 249, 117]
    cv.verify_input_json(txt, 16, 5, 128, 4096)


def test_verifier_rejects_tampering():
    from oracle import circuit_verifier as cv
    txt = open(os.path.join(GOLDEN, "input_small.json")).read()
    js = json.loads(txt)
    js["merklePaths"][0][2] = str(int(js["merklePaths"][0][2]) + 1)
    with pytest.raises(AssertionError):
        cv.verify_input_json(json.dumps(js), 16, 5, 128, 4096)
    js = json.loads(txt)
    js["cellData"][1][0] = "5"
    with pytest.raises(AssertionError):
        cv.verify_input_json(json.dumps(js), 16, 5, 128, 4096)
