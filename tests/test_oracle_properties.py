"""Property tests of the oracle (hypothesis): the invariants the GPU tests lean on must hold for arbitrary inputs."""
from hypothesis import given, settings, strategies as st

R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
felts = st.integers(min_value=0, max_value=R - 1)


@settings(max_examples=40, deadline=None)
@given(st.binary(min_size=0, max_size=300))
def test_chunking_is_injective_and_bounded(orc_mod, data):
    elems = orc_mod.bytes_to_elements(data)
    assert len(elems) == len(data) // 31 + 1 and all(e < (1 << 248) for e in elems)
    raw = b"".join(e.to_bytes(31, "little") for e in elems)
    assert raw[:len(data)] == data and raw[len(data)] == 1 and not any(raw[len(data) + 1:])      # 10* padding, invertible


@settings(max_examples=25, deadline=None)
@given(st.lists(felts, min_size=1, max_size=40), st.data())
def test_every_proof_reconstructs(orc_mod, pyorc_mod, leaves, data):
    layers = orc_mod.merkle_layers(leaves)
    assert [len(l) for l in layers][0] == len(leaves) and len(layers[-1]) == 1
    j = data.draw(st.integers(min_value=0, max_value=len(leaves) - 1))
    p = pyorc_mod.merkle_proof(layers, j)
    assert orc_mod.reconstruct_root(p.leaf_value, j, len(leaves), p.merkle_path) == layers[-1][0]
    if len(leaves) > 1:                                            # a different leaf value must not verify
        assert orc_mod.reconstruct_root((p.leaf_value + 1) % R, j, len(leaves), p.merkle_path) != layers[-1][0]


@settings(max_examples=20, deadline=None)
@given(st.lists(felts, min_size=2, max_size=33))
def test_subtree_composition(orc_mod, leaves):
    """what the multi-GPU split relies on: the tree over 2^k-aligned chunk roots (non-bottom keys) equals the whole tree"""
    k = 1
    while (1 << (k + 1)) < len(leaves):
        k += 1
    chunk = 1 << k
    whole = orc_mod.merkle_layers(leaves)
    if chunk >= len(leaves):
        return
    level_k = whole[k]
    roots = []
    for c in range(0, len(leaves), chunk):
        part = leaves[c:c + chunk]
        if len(part) == chunk:
            roots.append(orc_mod.merkle_layers(part)[k][0])        # a complete aligned chunk is a complete sub-tree
    assert roots == level_k[:len(roots)]
    top = orc_mod.merkle_layers(level_k, bottom=False)
    assert top[-1][0] == whole[-1][0]


@settings(max_examples=30, deadline=None)
@given(felts, felts, st.integers(min_value=0, max_value=3))
def test_compress_is_the_first_permutation_word(orc_mod, x, y, key):
    assert orc_mod.compress(x, y, key) == orc_mod.permutation((x, y, key))[0]


def test_oracle_dedicated_squaring(orc):
    """the oracle's 10-product squaring (CPU baseline honesty) equals its general product and a^2 mod r"""
    import random
    R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
    rnd = random.Random(2)
    cases = [0, 1, 2, R - 1, R - 2, (1 << 64) - 1, (1 << 128) - 1, (1 << 192) - 1, (1 << 253), (1 << 254) % R, 0xffffffffffffffff << 64]
    cases += [rnd.randrange(R) for _ in range(5000)]
    cases += [int("ffffffffffffffff" * k + "0000000000000000" * (3 - k) + "00000000ffffffff", 16) % R for k in range(4)]
    for a in cases:
        s, m = orc.fr_sqr_check(a)
        assert s == m == a * a % R, hex(a)


def test_oracle_adx_product_equals_portable_product(orc):
    """the MULX/ADCX/ADOX Montgomery product (compiled in where the target has BMI2+ADX) against the portable C product
    and a*b mod r"""
    import random
    R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
    rnd = random.Random(4)
    edge = [0, 1, 2, R - 1, R - 2, (1 << 64) - 1, (1 << 128) - 1, (1 << 192) - 1, 1 << 253, (1 << 254) % R, 0xffffffffffffffff << 64,
            0xffffffffffffffff << 128, 0xffffffffffffffff << 192 & (R - 1)]
    cases = [(a, b) for a in edge for b in edge] + [(rnd.randrange(R), rnd.randrange(R)) for _ in range(5000)]
    for a, b in cases:
        fast, c = orc.fr_mul_check(a, b)
        assert fast == c == a * b % R, (hex(a), hex(b))
