// poseidon2.cuh -- Poseidon2 (BN254, t = 3) permutation, rate-1/2 sponge steps, keyed compression and the
// 31-byte chunk reader, on Montgomery-form state held in registers (layer L1 of SURVEY.md section 1).
//
//   permutation            reference/haskell/src/Poseidon2/Permutation.hs:14-45, circuit/poseidon2/poseidon2_perm.circom:163-198
//   round constants        reference/haskell/src/Poseidon2/RoundConsts.hs:30-128 (Montgomery form in poseidon2_rc.cuh)
//   sponge, IV, 10* pad    reference/haskell/src/Poseidon2/Sponge.hs:13-43, circuit/poseidon2/poseidon2_sponge.circom:28-99
//   keyed compression      reference/haskell/src/Poseidon2/Merkle.hs:202-203, circuit/poseidon2/poseidon2_compr.circom:30-41
//   bytes -> elements      reference/haskell/src/Slot.hs:243-270, reference README.md:86-99
//
// Range discipline: the three state words are < r at every round boundary.  S-box input is x + c < 2r (lazy
// add), the three products stay < 1.76r / 1.59r / 1.6r, one conditional subtraction brings the result back
// below r.  All lanes of a warp are in the same round, so round constants come from the constant bank as
// uniform loads.
#pragma once
#include "fr.cuh"
#include "poseidon2_rc.cuh"

namespace cdx {

#if defined(__CUDACC__)
// [0..23] external rounds (round-major, 3 per round), [24..79] internal rounds; Montgomery form
static __constant__ Fr c_rc[80] = P2_RC_MONT_INIT;
#define CDX_RC(i) c_rc[i]
#else
static const Fr h_rc[80] = P2_RC_MONT_INIT;
#define CDX_RC(i) h_rc[i]
#endif


// capacity IV = 2^64 + 256*t + rate (Sponge.hs:17,34) in Montgomery form: computed once per thread (2 modmuls)
CDX_D Fr sponge_iv(int rate) {
  Fr a = fr_zero();
  a.l[0] = 0x0300u + (uint32_t)rate;
  a.l[2] = 1u;
  return to_mont(a);
}

#ifndef CDX_TABRED
#define CDX_TABRED 1
#endif

#if CDX_TABRED
// Range discipline of the permutation (units of r; B = 1 + 2^250/r = 1.0827 is what reduce_tab returns):
//   state words are < B at every round boundary;
//   S-box input t = w + c < B + 1 = 2.083 (mont_sqr wants < 2.14);  t^2 < 1.82,  t^4 < 1.63,  t^5 < 1.64 -- left unreduced;
//   the mixes add up to 2x' + y + z < 6.6 resp. x' + y + 3z < 6.0, i.e. at most 257 bits (2^257 = 10.6 r), and every new
//   state word goes through ONE table reduction (reduce_tab, fr.cuh) instead of a chain of conditional subtractions.

// x^5 with x < 2.14 r; result < 1.64 r                               Permutation.hs:14-17
CDX_D Fr sbox(const Fr& x) {
  Fr x2 = mont_sqr(x);
  Fr x4 = mont_sqr(x2);
  return mont_mul(x4, x);
}

// (x,y,z) <- (x+s, y+s, z+s), s = x+y+z; inputs < 1.64 r, outputs < B     Permutation.hs:35-36 and the external-round mix :28-33
CDX_D void mix_external(Fr& x, Fr& y, Fr& z) {
  const Fr s = add_lazy(add_lazy(x, y), z);       // < 4.92 r: fits 256 bits
  x = add_reduce(x, s);                           // < 6.56 r: 257 bits
  y = add_reduce(y, s);
  z = add_reduce(z, s);
}

// internal-round mix with x already through its S-box (x < 1.64 r; y, z < B)   Permutation.hs:19-26, matrix [[2,1,1],[1,2,1],[1,1,3]]
CDX_D void mix_internal(Fr& x, Fr& y, Fr& z) {
  const Fr s = add_lazy(x, add_lazy(y, z));       // < 3.81 r
  x = add_reduce(x, s);                           // 2x + y + z < 5.45 r
  y = reduce_tab(add_lazy(y, s), 0u);             // x + 2y + z < 4.89 r: no carry
  z = add_reduce(z, add_lazy(z, s));              // x + y + 3z < 5.97 r
}

// absorb one element into a state word: w + m*R mod r.  m standard form, any value < 2^256; w < B; result < B
CDX_D Fr absorb(const Fr& w, const Fr& m) {
  const Fr r2 = {CDX_R2_INIT};
  return add_reduce(w, mont_mul(r2, m));          // R2 < r is the row operand: product < 2r whatever m
}
CDX_D Fr absorb_one(const Fr& w) {
  const Fr one = {CDX_ONE_INIT};
  return add_reduce(w, one);
}
#else
// x^5 with x < 2r; result < r                                       Permutation.hs:14-17
CDX_D Fr sbox(const Fr& x) {
  Fr x2 = mont_sqr(x);
  Fr x4 = mont_sqr(x2);
  return reduce_once(mont_mul(x4, x));
}

// (x,y,z) <- (x+s, y+s, z+s), s = x+y+z                             Permutation.hs:35-36 and the external-round mix :28-33
CDX_D void mix_external(Fr& x, Fr& y, Fr& z) {
  Fr s = add_mod(add_mod(x, y), z);
  x = add_mod(x, s);
  y = add_mod(y, s);
  z = add_mod(z, s);
}

// internal-round mix with x already through its S-box            Permutation.hs:19-26, matrix [[2,1,1],[1,2,1],[1,1,3]]
CDX_D void mix_internal(Fr& x, Fr& y, Fr& z) {
  Fr s = add_mod(add_mod(x, y), z);
  x = add_mod(x, s);
  y = add_mod(y, s);
  z = add_mod(dbl_mod(z), s);
}

CDX_D Fr absorb(const Fr& w, const Fr& m) { return add_mod(w, to_mont(m)); }
CDX_D Fr absorb_one(const Fr& w) {
  const Fr one = {CDX_ONE_INIT};
  return add_mod(w, one);
}
#endif

// state words < B (1.083 r) in, < B out                             Permutation.hs:40-45
// Four S-box instances in the instruction stream: one in the internal-round loop (56 of the 80 S-box evaluations,
// nothing else in its body) and the three independent ones of an external round, in one loop body shared by all
// eight external rounds (the `half` loop runs it before and after the internal rounds).  With the first, larger
// squaring that body was 35 KB of SASS and the kernel stalled on instruction fetch, so one S-box was shared and the
// state rotated through it; at 27 KB the three-way body is 1 % faster than the rotation (same-box A/B).
CDX_D void permute(Fr& x, Fr& y, Fr& z) {
  mix_external(x, y, z);
#pragma unroll 1
  for (int half = 0; half < 2; ++half) {
#pragma unroll 1
    for (int i = 0; i < 4; ++i) {        // 4 external rounds; the three S-boxes of a round are independent
      x = sbox(add_lazy(x, CDX_RC(12 * half + 3 * i)));
      y = sbox(add_lazy(y, CDX_RC(12 * half + 3 * i + 1)));
      z = sbox(add_lazy(z, CDX_RC(12 * half + 3 * i + 2)));
      mix_external(x, y, z);
    }
    if (half == 0) {
#pragma unroll 1
      for (int r = 0; r < 56; ++r) {
        x = sbox(add_lazy(x, CDX_RC(24 + r)));
        mix_internal(x, y, z);
      }
    }
  }
}

// perm(x, y, key)[0] with key in {0,1,2,3}                          Merkle.hs:202-203
CDX_D Fr compress_keyed(const Fr& x, const Fr& y, uint32_t key) {
  Fr a = x, b = y, k = mont_from_u32(key);
  permute(a, b, k);
  return a;
}

// ---------------------------------------------------------------------------------------------------------
// 31-byte chunk reader.  `ld(i)` returns little-endian 32-bit word i of the PADDED stream
// data ++ 0x01 ++ 0x00...  (the loader owns the padding rule and where the bytes live: global memory, shared
// memory, registers).  Chunk k covers stream bytes [31k, 31k+31) and is returned as a standard-form integer
// < 2^248 (Slot.hs:243-270, README.md:86-99).
template <class LoadWord>
CDX_D Fr read_chunk(const LoadWord& ld, uint32_t k) {
  const uint32_t byte_off = 31u * k;
  const uint32_t w0 = byte_off >> 2;
  const uint32_t sh = (byte_off & 3u) * 8u;
  uint32_t w[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) w[i] = ld(w0 + i);
  Fr r;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
#if defined(__CUDA_ARCH__)
    r.l[i] = __funnelshift_r(w[i], w[i + 1], sh);
#else
    r.l[i] = sh ? ((w[i] >> sh) | (w[i + 1] << (32 - sh))) : w[i];
#endif
  }
  r.l[7] &= 0x00ffffffu;
  return r;
}

// padded-stream loader for a cell whose start is 4-byte aligned and whose length is a multiple of 4
struct AlignedWords {
  const uint32_t* words;
  uint32_t n_words;
  CDX_D uint32_t operator()(uint32_t i) const { return i < n_words ? words[i] : (i == n_words ? 1u : 0u); }
  CDX_D void begin_step(uint32_t) const {}   // loaders that stage data (RowRing in kernels.cuh) advance their pipeline here
};

// padded-stream loader for an arbitrary byte string (any alignment, any length)
struct AnyBytes {
  const uint8_t* p;
  uint32_t len;
  CDX_D uint32_t byte_at(uint32_t j) const { return j < len ? (uint32_t)p[j] : (j == len ? 1u : 0u); }
  CDX_D uint32_t operator()(uint32_t i) const {
    const uint32_t j = 4u * i;
    return byte_at(j) | (byte_at(j + 1) << 8) | (byte_at(j + 2) << 16) | (byte_at(j + 3) << 24);
  }
  CDX_D void begin_step(uint32_t) const {}
};

// number of field elements of a len-byte string: floor(len/31) + 1   (Slot.hs:243-250)
CDX_HD uint32_t n_chunks(uint32_t len) { return len / 31u + 1u; }

// sponge2 over the chunks of one byte string; returns the digest in Montgomery form, < r.
// One loop, one inlined permutation: step j absorbs elements (2j, 2j+1) of  chunks ++ pad,
// pad = [1] if the chunk count is odd, [1,0] if even.              Sponge.hs:30-43, blocks/bn254.nim:23-29
template <class LoadWord>
CDX_D Fr sponge2_bytes(const LoadWord& ld, uint32_t len_bytes) {
  const uint32_t n = n_chunks(len_bytes);
  const uint32_t n_perm = n / 2u + 1u;
  Fr s0 = fr_zero(), s1 = fr_zero(), s2 = sponge_iv(2);
#pragma unroll 1
  for (uint32_t j = 0; j < n_perm; ++j) {
    const uint32_t k = 2u * j;
    ld.begin_step(j);                 // step j reads padded-stream bytes [62 j, 62 j + 68)
    if (k < n) s0 = absorb(s0, read_chunk(ld, k));
    else s0 = absorb_one(s0);
    if (k + 1 < n) s1 = absorb(s1, read_chunk(ld, k + 1));
    else if (k + 1 == n) s1 = absorb_one(s1);
    permute(s0, s1, s2);
  }
  return s0;
}

// rate-1 / rate-2 sponge over n canonical field elements fetched by `get(i)` (standard form, any value < 2^256)
//                                                                   Sponge.hs:13-43
template <class GetElem>
CDX_D Fr sponge_elems(GetElem get, uint32_t n, int rate) {
  Fr s0 = fr_zero(), s1 = fr_zero(), s2 = sponge_iv(rate);
  const uint32_t n_perm = rate == 1 ? n + 1u : n / 2u + 1u;
#pragma unroll 1
  for (uint32_t j = 0; j < n_perm; ++j) {
    const uint32_t k = rate == 1 ? j : 2u * j;
    if (k < n) s0 = absorb(s0, get(k));
    else s0 = absorb_one(s0);
    if (rate == 2) {
      if (k + 1 < n) s1 = absorb(s1, get(k + 1));
      else if (k + 1 == n) s1 = absorb_one(s1);
    }
    permute(s0, s1, s2);
  }
  return s0;
}

}  // namespace cdx
