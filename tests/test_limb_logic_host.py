"""The device headers' limb-level logic (lazy-reduction bounds, Poseidon2 schedule, 31-byte chunk reader, sponge
padding) compiled with g++ against a C emulation of the inline-PTX primitives (tests/host_emul/), compared with
the oracle.  This is a unit test of shared header code, not a CPU backend: nothing here ships."""
import ctypes as C
import os
import random
import subprocess

import pytest

from conftest import ROOT

R = 21888242871839275222246405745257275088548364400416034343698204186575808495617


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("emul") / "liblimb.so")
    src = os.path.join(ROOT, "tests", "host_emul", "limb_logic.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I" + os.path.join(ROOT, "tests", "host_emul"), "-o", out, src], check=True)
    return C.CDLL(out)


@pytest.fixture(scope="module")
def emul_ubsan(tmp_path_factory):
    """the same library built with -fsanitize=undefined -fno-sanitize-recover: any undefined behaviour in the shared header
    logic (over-wide shifts, misaligned or out-of-range accesses the compiler can see, signed overflow) aborts the test run.
    compute-sanitizer is closed on the GPU pool, so this is the sanitizer coverage the limb logic gets."""
    out = str(tmp_path_factory.mktemp("emul_ubsan") / "liblimb_ubsan.so")
    src = os.path.join(ROOT, "tests", "host_emul", "limb_logic.cpp")
    res = subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=undefined", "-fno-sanitize-recover=undefined", "-shared", "-fPIC",
                          "-I" + os.path.join(ROOT, "tests", "host_emul"), "-o", out, src], capture_output=True, text=True)
    if res.returncode != 0:
        pytest.skip("this g++ cannot build with -fsanitize=undefined: " + res.stderr[-200:])
    try:
        return C.CDLL(out)
    except OSError as e:
        pytest.skip("the UBSan runtime is not loadable: " + str(e))


def f2b(x): return int(x).to_bytes(32, "little")
def b2f(b): return int.from_bytes(b, "little")


def test_mont_mul_lazy_bounds(emul):
    rnd = random.Random(7)
    rinv = pow(1 << 256, -1, R)
    out = C.create_string_buffer(32)
    cases = [(rnd.randrange(2 * R), rnd.randrange(2 * R)) for _ in range(2000)]
    cases += [(2 * R - 1, 2 * R - 1), (0, 0), (R, R), (R - 1, (1 << 256) - 1), (4 * R - 1, 4 * R - 1), (4 * R, (1 << 256) - 1),
              ((1 << 256) - R - 1, 5 * R)]
    for a, b in cases:
        emul.emul_mont_mul_raw(f2b(a), f2b(b), out)
        v = b2f(out.raw)
        assert v % R == a * b * rinv % R
        assert v <= (a * b >> 256) + R            # documented bound: < a*b/2^256 + r
        if a < 2 * R and b < 2 * R:
            assert v < 2 * R


def test_mont_sqr_dedicated(emul):
    """36-product streaming squaring (variable-length product rows, two reduction rows per product row, upper rows added
    to the final window): same contract as mont_mul(a, a), bit-identical result"""
    rnd = random.Random(17)
    rinv = pow(1 << 256, -1, R)
    out, out2 = C.create_string_buffer(32), C.create_string_buffer(32)
    cases = [rnd.randrange(2 * R) for _ in range(3000)] + [0, 1, R, R - 1, 2 * R - 1, (1 << 254) - 1, 0xffffffff, (1 << 254) - (1 << 32)]
    cases += [int("ffffffff" * k + "00000000" * (7 - k) + "ffffffff", 16) % (2 * R) for k in range(7)]
    for a in cases:
        emul.emul_mont_sqr_raw(f2b(a), out)
        v = b2f(out.raw)
        assert v % R == a * a * rinv % R, hex(a)
        assert v <= (a * a >> 256) + R and v < 2 * R
        emul.emul_mont_mul_raw(f2b(a), f2b(a), out2)
        assert b2f(out2.raw) == v
    # the window bound of the streaming squaring is 2a + r < 2^256, i.e. a < 2.14 r: the emulation traps on any carry out
    # of the window, so running right up to that bound checks the bound itself (the S-box never goes beyond 2r)
    limit = ((1 << 256) - R) // 2
    for a in [limit - 1, limit - (1 << 200), 2 * R, 2 * R + 12345] + [rnd.randrange(2 * R, limit) for _ in range(500)]:
        emul.emul_mont_sqr_raw(f2b(a), out)
        v = b2f(out.raw)
        assert v % R == a * a * rinv % R and v <= (a * a >> 256) + R


def test_table_reduction_and_mixes_at_their_bounds(emul):
    """reduce_tab maps any v < 2^257 to the congruent value below B = r + 2^250; the S-box and both mixes keep their
    documented ranges when driven at the edges (state words B - 1, S-box results 1.64 r): the emulation traps on any carry
    out of a lazy add and on any reduction result >= B, so finishing is the check; values are compared mod r"""
    rnd = random.Random(23)
    B = R + (1 << 250)
    out = C.create_string_buffer(32)
    vals = [0, 1, R - 1, R, R + 1, B - 1, B, 2 * R, 5 * R, (1 << 256) - 1] + [rnd.randrange(1 << 256) for _ in range(3000)]
    vals += [k << 250 for k in range(64)] + [(k << 250) - 1 for k in range(1, 65)]
    for v in vals:
        emul.emul_reduce_tab(f2b(v), 0, out)
        w = b2f(out.raw)
        assert w % R == v % R and w < B, hex(v)
    for v in vals + [(1 << 256) - 1]:                      # with the carry: the value is v + 2^256
        emul.emul_reduce_tab(f2b(v), 1, out)
        w = b2f(out.raw)
        assert w % R == (v + (1 << 256)) % R and w < B, hex(v)
    for a, b in [(B - 1, (1 << 256) - 1), ((1 << 256) - 1, (1 << 256) - 1), (0, 0), (R, B)] + [(rnd.randrange(1 << 256), rnd.randrange(1 << 256)) for _ in range(500)]:
        emul.emul_add_reduce(f2b(a), f2b(b), out)
        w = b2f(out.raw)
        assert w % R == (a + b) % R and w < B
    rinv5 = pow(1 << 256, -4, R)                               # sbox on raw values: x^5 R^-4
    sbox_in_max = ((1 << 256) - R) // 2 - 1                    # mont_sqr's window bound, 2.14 r
    for x in [0, 1, B - 1, B + R - 1, 2 * R, sbox_in_max] + [rnd.randrange(B + R) for _ in range(300)]:
        emul.emul_sbox_raw(f2b(x), out)
        w = b2f(out.raw)
        assert w % R == pow(x, 5, R) * rinv5 % R
        if x < B + R:
            assert w < 164 * R // 100                          # the documented 1.64 r
    sb_max = 164 * R // 100
    bx, by, bz = (C.create_string_buffer(32) for _ in range(3))
    cases = [(sb_max, B - 1, B - 1), (sb_max, 0, 0), (0, B - 1, B - 1), (0, 0, 0)] + [(rnd.randrange(sb_max), rnd.randrange(B), rnd.randrange(B)) for _ in range(500)]
    for x, y, z in cases:
        for buf, v in ((bx, x), (by, y), (bz, z)):
            buf.raw = f2b(v)
        emul.emul_mix_internal_raw(bx, by, bz)
        s = x + y + z
        got = [b2f(b.raw) for b in (bx, by, bz)]
        assert [g % R for g in got] == [(x + s) % R, (y + s) % R, (2 * z + s) % R] and all(g < B for g in got)
    for x, y, z in [(sb_max, sb_max, sb_max), (0, 0, 0), (B - 1, B - 1, B - 1)] + [tuple(rnd.randrange(sb_max) for _ in range(3)) for _ in range(500)]:
        for buf, v in ((bx, x), (by, y), (bz, z)):
            buf.raw = f2b(v)
        emul.emul_mix_external_raw(bx, by, bz)
        s = x + y + z
        got = [b2f(b.raw) for b in (bx, by, bz)]
        assert [g % R for g in got] == [(x + s) % R, (y + s) % R, (z + s) % R] and all(g < B for g in got)


def test_permutation_and_compress(emul, orc):
    rnd = random.Random(8)
    out = C.create_string_buffer(96)
    for s in [(0, 1, 2), (R - 1, R - 1, R - 1), (0, 0, 0)] + [tuple(rnd.randrange(R) for _ in range(3)) for _ in range(50)]:
        emul.emul_permutation(b"".join(f2b(x) for x in s), out)
        assert tuple(b2f(out.raw[i:i + 32]) for i in (0, 32, 64)) == orc.permutation(s)
    o32 = C.create_string_buffer(32)
    for k in range(4):
        x, y = rnd.randrange(R), rnd.randrange(R)
        emul.emul_compress(f2b(x), f2b(y), k, o32)
        assert b2f(o32.raw) == orc.compress(x, y, k)


def test_chunk_reader_and_sponge_padding(emul, orc):
    rnd = random.Random(9)
    out = C.create_string_buffer(32)
    for n in list(range(0, 100)) + [128, 256, 2047, 2048, 31 * 66, 31 * 67, 4096]:
        d = bytes(rnd.randrange(256) for _ in range(n))
        emul.emul_hash_bytes(d, n, out)
        assert b2f(out.raw) == orc.hash_bytes(d), n
        if n % 4 == 0 and n > 0:
            buf = C.create_string_buffer(d, n)
            emul.emul_hash_cell_aligned(buf, n, out)
            assert b2f(out.raw) == orc.hash_bytes(d), n
    d = b"\xff" * 2048                              # every chunk >= 2^247
    buf = C.create_string_buffer(d, 2048)
    emul.emul_hash_cell_aligned(buf, 2048, out)
    assert b2f(out.raw) == orc.hash_bytes(d)
    for n in range(0, 9):
        xs = [rnd.randrange(R) for _ in range(n)]
        for rate in (1, 2):
            emul.emul_sponge(b"".join(f2b(x) for x in xs), n, rate, out)
            assert b2f(out.raw) == orc.sponge(xs, rate)


def test_limb_logic_under_ubsan(emul_ubsan, orc):
    """permutation, byte hashing on both loaders, sponge and compression through the UBSan build: results equal the oracle
    and nothing undefined is executed on the way"""
    rnd = random.Random(31)
    out = C.create_string_buffer(96)
    for s in [(0, 1, 2), (R - 1, R - 1, R - 1)] + [tuple(rnd.randrange(R) for _ in range(3)) for _ in range(10)]:
        emul_ubsan.emul_permutation(b"".join(f2b(x) for x in s), out)
        assert tuple(b2f(out.raw[i:i + 32]) for i in (0, 32, 64)) == orc.permutation(s)
    o32 = C.create_string_buffer(32)
    for n in list(range(0, 70)) + [2047, 2048, 4096]:
        d = bytes(rnd.randrange(256) for _ in range(n))
        emul_ubsan.emul_hash_bytes(d, n, o32)
        assert b2f(o32.raw) == orc.hash_bytes(d), n
        if n % 4 == 0 and n:
            buf = C.create_string_buffer(d, n)
            emul_ubsan.emul_hash_cell_aligned(buf, n, o32)
            assert b2f(o32.raw) == orc.hash_bytes(d), n
    for k in range(4):
        x, y = rnd.randrange(R), rnd.randrange(R)
        emul_ubsan.emul_compress(f2b(x), f2b(y), k, o32)
        assert b2f(o32.raw) == orc.compress(x, y, k)
