// fr29.cuh -- carry-free Montgomery multiplication and squaring for BN254 Fr on sm_100a.
//
// Why a second representation.  Integer multiplies issue only on the FMA-heavy pipe (64 lanes/clk/SM).  Measured on
// B200 (tools/probes/probe_pipes.cu): IMAD.WIDE.U32 with a plain 64-bit accumulate runs at the full 63.4
// lanes/clk/SM, but every form that touches the carry predicate (IMAD.WIDE.U32.X, carry-out, IMAD.HI) runs at half
// that.  The 8x32-bit CIOS product of fr.cuh needs a carry on ~128 of its 136 multiplies, so it saturates the pipe at
// ~0.49 of the IMAD.WIDE peak.  Here the operands are 9 limbs of 29 bits: every 29x29-bit product is < 2^58, eighteen
// of them fit a 64-bit column accumulator, so a product is 81 (square: 45) + a reduction of 81 carry-FREE
// IMAD.WIDE.U32, and the carries move between columns with shifts/adds on the ALU pipe, which runs alongside.
//
//   value = sum l[i] * 2^(29 i),   Montgomery radix R' = 2^261  (2^261 / r = 169.3: lots of headroom)
//   mul29 / sqr29: input limbs < 2^30 (one un-normalised addition is fine), values with a*b < 2^261 * 168 r;
//                  output limbs < 2^29 (normalised), value < a*b / 2^261 + r.
//
// Plain C++ on purpose (uint64_t accumulators compile to IMAD.WIDE.U32 Rd, Ra, Rb, Rc64): the same code is compiled
// for the host in tests/host_emul and compared with the oracle there.
#pragma once
#include "fr.cuh"

namespace cdx {

struct F29 {
  uint32_t l[9];
};

#define CDX29_MASK 0x1fffffffu
#define CDX29_NP 0x0fffffffu  // -r^-1 mod 2^29
// r in radix 2^29
#define CDX29_N0 0x10000001u
#define CDX29_N1 0x1f0fac9fu
#define CDX29_N2 0x0e5c2450u
#define CDX29_N3 0x07d090f3u
#define CDX29_N4 0x1585d283u
#define CDX29_N5 0x02db40c0u
#define CDX29_N6 0x00a6e141u
#define CDX29_N7 0x0e5c2634u
#define CDX29_N8 0x0030644eu
// 2^522 mod r (enters the R' = 2^261 Montgomery form) and 2^261 mod r (Montgomery one), 8x32-bit limbs
#define CDX29_R2_INIT {0x45b69bd4u, 0x38c2e14bu, 0x85883377u, 0x0ffedb18u, 0xabc6e54du, 0x7840f9f0u, 0x848b0f05u, 0x0a054a3eu}
#define CDX29_ONE_INIT {0x8fffff57u, 0x2fd4e156u, 0xa494b01au, 0x75bba827u, 0x819caa80u, 0x5301fa84u, 0x563d4475u, 0x0dc83629u}

// acc += a * b as ONE IMAD.WIDE.U32 with a 64-bit accumulate.  Written as PTX on the device because the C++ form
// `acc += (uint64_t)a * CONSTANT` is lowered as a 32x64-bit multiply (an extra add of zero to the high word per
// product); ptxas still folds constant operands into immediates.
CDX_D void mad_wide(uint64_t& acc, uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(a), "r"(b));
#else
  acc += (uint64_t)a * b;
#endif
}
CDX_D uint64_t mul_wide(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  uint64_t r;
  asm("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(a), "r"(b));
  return r;
#else
  return (uint64_t)a * b;
#endif
}

// 8x32 -> 9x29 (any 256-bit value; limb 8 gets the top 24 bits)
CDX_D F29 to29(const Fr& a) {
  F29 r;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const int bit = 29 * k, w = bit >> 5, sh = bit & 31;
    const uint32_t lo = a.l[w];
    const uint32_t hi = (w + 1 < 8) ? a.l[w + 1] : 0u;
#if defined(__CUDA_ARCH__)
    r.l[k] = __funnelshift_r(lo, hi, sh) & CDX29_MASK;
#else
    r.l[k] = (uint32_t)((((uint64_t)hi << 32) | lo) >> sh) & CDX29_MASK;
#endif
  }
  return r;
}

// 9x29 (limbs < 2^29, value < 2^256) -> 8x32
CDX_D Fr from29(const F29& x) {
  Fr r;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int bit = 32 * j, k = bit / 29, s = bit - 29 * k;   // word j = limb k bits s.. , then limb k+1
    r.l[j] = (x.l[k] >> s) + (x.l[k + 1] << (29 - s));
  }
  return r;
}

// one reduction step on column k: m = -c_k / r mod 2^29, add m*r to columns k..k+8, push the carry into column k+1
#define CDX29_REDC_STEP(c, k)                                   \
  do {                                                          \
    const uint32_t m = ((uint32_t)c[k] * CDX29_NP) & CDX29_MASK; \
    mad_wide(c[k + 0], m, CDX29_N0);                            \
    mad_wide(c[k + 1], m, CDX29_N1);                            \
    mad_wide(c[k + 2], m, CDX29_N2);                            \
    mad_wide(c[k + 3], m, CDX29_N3);                            \
    mad_wide(c[k + 4], m, CDX29_N4);                            \
    mad_wide(c[k + 5], m, CDX29_N5);                            \
    mad_wide(c[k + 6], m, CDX29_N6);                            \
    mad_wide(c[k + 7], m, CDX29_N7);                            \
    mad_wide(c[k + 8], m, CDX29_N8);                            \
    c[k + 1] += c[k] >> 29;                                     \
  } while (0)

// columns 9..17 -> nine normalised limbs
CDX_D F29 normalize_top(uint64_t* c) {
  F29 r;
#pragma unroll
  for (int k = 9; k < 17; ++k) {
    c[k + 1] += c[k] >> 29;
    r.l[k - 9] = (uint32_t)c[k] & CDX29_MASK;
  }
  r.l[8] = (uint32_t)c[17];
  return r;
}

// a * b * 2^-261 mod r (lazily reduced: < a*b/2^261 + r)
CDX_D F29 mul29(const F29& a, const F29& b) {
  uint64_t c[18];
#pragma unroll
  for (int k = 0; k < 18; ++k) c[k] = 0;
#pragma unroll
  for (int i = 0; i < 9; ++i) {
#pragma unroll
    for (int j = 0; j < 9; ++j) mad_wide(c[i + j], a.l[j], b.l[i]);   // row i completes column i
    CDX29_REDC_STEP(c, i);
  }
  return normalize_top(c);
}

// a * a * 2^-261 mod r: 45 products instead of 81 (cross terms use the doubled operand)
CDX_D F29 sqr29(const F29& a) {
  uint64_t c[18];
  uint32_t d[9];
#pragma unroll
  for (int k = 0; k < 18; ++k) c[k] = 0;
#pragma unroll
  for (int j = 0; j < 9; ++j) d[j] = a.l[j] << 1;
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    mad_wide(c[2 * i], a.l[i], a.l[i]);
#pragma unroll
    for (int j = i + 1; j < 9; ++j) mad_wide(c[i + j], a.l[i], d[j]);
    // after row i the columns up to 2i+1 are complete: reduce the ones not reduced yet
    if (2 * i < 9) CDX29_REDC_STEP(c, 2 * i);
    if (2 * i + 1 < 9) CDX29_REDC_STEP(c, 2 * i + 1);
  }
  return normalize_top(c);
}

// x^5 on a lazily added 8x32 input (< 2^256); result 8x32, < r.  x^2, x^4 and x^5 stay below 1.03 r.
CDX_D Fr sbox29(const Fr& x) {
  const F29 t = to29(x);
  const F29 x2 = sqr29(t);
  const F29 x4 = sqr29(x2);
  return reduce_once(from29(mul29(x4, t)));
}

}  // namespace cdx
