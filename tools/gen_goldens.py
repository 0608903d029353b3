#!/usr/bin/env python3
"""Generate tests/golden/*.json from the oracle (run in the build container; outputs are committed).

The reference stores exactly one expected value on this path -- the permutation KAT
(reference/haskell/src/Poseidon2/Example.hs:13-22) -- and its test-vector programs only print
(reference/nim/testvectors/src/testvectors.nim:20-72, reference/haskell/src/TestVectors.hs:28-75).  So the
goldens below are the oracle's own outputs, frozen: the KAT is copied from the reference, everything else is
computed by the pure-Python twin (oracle/pyoracle.py) and cross-checked against the C oracle before it is written.
The config-1 input.json (11 slots x 2048 cells) is computed with the C primitives (pure Python needs ~5 min) and
must pass the independent circom-semantics verifier (oracle/circuit_verifier.py) before it is written.
"""
import hashlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle as py, coracle as cc, circuit_verifier as cv

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def splitmix64_stream(seed):
    x = seed & (2**64 - 1)
    while True:
        x = (x + 0x9e3779b97f4a7c15) & (2**64 - 1)
        z = x
        z = ((z ^ (z >> 30)) * 0xbf58476d1ce4e5b9) & (2**64 - 1)
        z = ((z ^ (z >> 27)) * 0x94d049bb133111eb) & (2**64 - 1)
        yield z ^ (z >> 31)


def random_felts(seed, n):
    """n canonical field elements: four splitmix64 words, top word masked to 254 bits, rejection-sampled < r
    (SURVEY.md section 8d config 2)."""
    g, out = splitmix64_stream(seed), []
    while len(out) < n:
        w = [next(g) for _ in range(4)]
        v = w[0] | (w[1] << 64) | (w[2] << 128) | ((w[3] & ((1 << 62) - 1)) << 192)
        if v < py.R:
            out.append(v)
    return out


def main():
    S = str
    vec = {}
    # 1. the reference's stored KAT (Example.hs:13-22) -- copied, not computed
    vec["permutation_kat"] = {
        "source": "reference/haskell/src/Poseidon2/Example.hs:13-22",
        "in": ["0", "1", "2"],
        "out": [S(0x30610a447b7dec194697fb50786aa7421494bd64c221ba4d3b1af25fb07bd103),
                S(0x13f731d6ffbad391be22d2ac364151849e19fa38eced4e761bcd21dbdc600288),
                S(0x1433e2c8f68382c447c5c14b8b3df7cbfd9273dd655fe52f1357c27150da786f)],
    }
    assert [S(v) for v in py.permutation((0, 1, 2))] == vec["permutation_kat"]["out"]
    # 2. permutations of (j, j+1, j+2) and of random states
    perm_in = [(j, j + 1, j + 2) for j in range(16)] + [tuple(random_felts(1, 3 * 16)[3 * i:3 * i + 3]) for i in range(16)]
    perm_in += [(py.R - 1, py.R - 1, py.R - 1), (0, 0, 0)]
    vec["permutations"] = [{"in": [S(v) for v in s], "out": [S(v) for v in py.permutation(s)]} for s in perm_in]
    for s in perm_in:
        assert cc.permutation(s) == py.permutation(s)
    # 3. the testvectors suite (shape of testvectors.nim:20-72)
    vec["sponge_rate1"] = [S(py.sponge1(list(range(1, n + 1)))) for n in range(0, 9)]
    vec["sponge_rate2"] = [S(py.sponge2(list(range(1, n + 1)))) for n in range(0, 9)]
    vec["hash_bytes"] = [S(py.hash_bytes(bytes(range(1, n + 1)))) for n in range(0, 81)]
    vec["merkle_root_felts"] = [S(py.merkle_root(list(range(1, n + 1)))) for n in range(1, 41)]
    vec["merkle_root_bytes"] = [S(py.merkle_root(py.bytes_to_elements(bytes(range(1, n + 1))))) for n in range(0, 81)]
    for n in range(0, 9):
        assert S(cc.sponge1(list(range(1, n + 1)))) == vec["sponge_rate1"][n]
        assert S(cc.sponge2(list(range(1, n + 1)))) == vec["sponge_rate2"][n]
    for n in range(0, 81):
        assert S(cc.hash_bytes(bytes(range(1, n + 1)))) == vec["hash_bytes"][n]
        assert S(cc.merkle_root(cc.bytes_to_elements(bytes(range(1, n + 1))))) == vec["merkle_root_bytes"][n]
    for n in range(1, 41):
        assert S(cc.merkle_root(list(range(1, n + 1)))) == vec["merkle_root_felts"][n - 1]
    # 4. keyed compression, all four keys
    vec["compress"] = [{"x": "1", "y": "2", "key": k, "out": S(py.compress(1, 2, k))} for k in range(4)]
    # 5. cell hashes of adversarial 2048-byte cells
    cells = {"zeros": bytes(2048), "ones_ff": b"\xff" * 2048, "ramp": bytes(i & 255 for i in range(2048)),
             "fake_seed15420_cell0": py.gen_fake_cell(12345 + 72 + 3003, 0, 2048)}
    vec["cell_hashes"] = {k: S(py.hash_bytes(v)) for k, v in cells.items()}
    for k, v in cells.items():
        assert S(cc.hash_bytes(v)) == vec["cell_hashes"][k]
    vec["fake_cell_sha256"] = {"seed": 12345 + 72 + 3003, "idx": 0, "cell_size": 2048,
                               "sha256": hashlib.sha256(cells["fake_seed15420_cell0"]).hexdigest(),
                               "first16": cells["fake_seed15420_cell0"][:16].hex()}
    with open(os.path.join(OUT, "vectors.json"), "w") as f:
        json.dump(vec, f, indent=1)

    # 6. the small reference demo config (testMain.hs:12-24): pure Python end to end
    g = py.GlobalConfig(16, 5, 128, 4096)
    d = py.DataSetConfig(5, 256, 10, 12345)
    inp = py.generate_proof_input(g, d, 3, 1234567)
    txt = py.export_proof_input(inp)
    cv.verify_input_json(txt, 16, 5, 128, 4096)
    open(os.path.join(OUT, "input_small.json"), "w").write(txt)

    # 7. BASELINE config 1 (workflow/cli_args.sh shape): C primitives + circom-semantics verification
    py.use_c_primitives(True)
    g = py.GlobalConfig(32, 8, 2048, 65536)
    d = py.DataSetConfig(11, 2048, 5, 12345)
    inp = py.generate_proof_input(g, d, 3, 1234567)
    txt = py.export_proof_input(inp)
    py.use_c_primitives(False)
    cv.verify_input_json(txt, 32, 8, 2048, 65536)
    open(os.path.join(OUT, "input_config1.json"), "w").write(txt)
    meta = {"config1": {"cli": "--field=bn254 --hash=poseidon2 --cellsize=2048 --blocksize=65536 --ncells=2048 --nslots=11 "
                               "--index=3 --nsamples=5 --seed=12345 --entropy=1234567 --depth=32 --maxslots=256",
                        "slotRoot": S(inp.slot_root), "dataSetRoot": S(inp.data_set_root),
                        "indices": [p.leaf_index for p in inp.merkle_proofs]},
            "small": {"cfg": "cell 128, block 4096, nCells 256, nSlots 5, slot 3, 10 samples, depth 16, maxLog2NSlots 5"}}
    json.dump(meta, open(os.path.join(OUT, "meta.json"), "w"), indent=1)
    print("goldens written to", OUT)


if __name__ == "__main__":
    main()
