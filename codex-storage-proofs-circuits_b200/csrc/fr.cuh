// fr.cuh -- BN254 scalar field Fr on sm_100a: 8 x 32-bit limbs, Montgomery form (R = 2^256).
//
// This is layer L0 of the hot path: the reference gets this arithmetic from third-party
// packages (constantine for Nim, `proof_input.nimble:11`; zikkurat-algebra for Haskell,
// `storage-proof-ref.cabal:34`) -- call sites `types/bn254.nim:27,30,58`, `merkle/bn254.nim:24-27`.
//
// Multiplication is word-serial Montgomery (CIOS) arranged so that every 32x32->64 product is ONE
// IMAD.WIDE.U32(.X) with the carry chain riding on the predicate carry: products of even limbs accumulate in
// an "even" 256-bit accumulator (64-bit slots at limb positions 0,2,4,6) and products of odd limbs in an "odd"
// one (slots at 1,3,5,7); dividing by 2^32 after each row swaps their roles, so no product ever has to be
// re-aligned.  Cost per product: 120 IMAD.WIDE.U32[.X] + 8 IMAD.HI + 8 IMAD (the 2n^2+n = 136 of BASELINE.md
// section 2) and ~30 IADD3 on the ALU pipe.  Squaring (2/3 of all multiplications on this path) has its own
// routine with 36 instead of 64 operand products.
//
// What bounds it (measured, tools/probes/probe_pipes.cu): integer multiplies issue only on the FMA-heavy pipe;
// a 32-bit IMAD runs at 61-63 lanes/clk/SM, every form that produces the high half of the product (IMAD.HI,
// IMAD.WIDE with or without carry) at 25-32: a 32x32->64 product costs two pipe slots whatever form it takes, so
// the only lever left is the number of products.
//
// Range discipline ("units of r", r < 2^254 so 2^256 > 5.29 r):
//   mont_mul(a,b) needs a < 4.29 r (= 2^256 - r: the running row sum is < a + r and must fit 256 bits),
//   a*b < 22 r^2, b anything < 2^256; it returns < (a*b/(5.29 r^2) + 1) r  -- no final subtraction inside.
//   In particular inputs < 2r give outputs < 2r; inputs < r give outputs < 1.19 r.
//   add_mod / dbl_mod take inputs < r and return < r.  reduce_once maps [0,2r) -> [0,r).
//
// The carry-chain primitives below are inline PTX.  Compiling this header with a host compiler is only
// possible with -DCDX_HOST_EMUL, which pulls C emulations of exactly those primitives from
// tests/host_emul/fr_rows_host.h so that the limb-level logic ABOVE them (reductions, Poseidon2 schedule, byte
// chunking) can be unit-tested without a GPU.  The product library is always compiled by nvcc for sm_100a and
// contains no host arithmetic.
#pragma once
#include <stdint.h>
#include "fr_reduce_tab.cuh"

#if defined(__CUDACC__)
#define CDX_HD __host__ __device__ __forceinline__
#define CDX_D __device__ __forceinline__
#ifdef CDX_MODMUL_NOINLINE
#define CDX_MM __device__ __noinline__
#else
#define CDX_MM __device__ __forceinline__
#endif
#else
#define CDX_HD inline
#define CDX_D inline
#define CDX_MM inline
#endif

namespace cdx {

struct Fr {
  uint32_t l[8];
};

// r = 21888242871839275222246405745257275088548364400416034343698204186575808495617 (reference README.md:76)
#define CDX_N0 0xf0000001u
#define CDX_N1 0x43e1f593u
#define CDX_N2 0x79b97091u
#define CDX_N3 0x2833e848u
#define CDX_N4 0x8181585du
#define CDX_N5 0xb85045b6u
#define CDX_N6 0xe131a029u
#define CDX_N7 0x30644e72u
#define CDX_NP 0xefffffffu  // -r^-1 mod 2^32

// R^2 mod r = 2^512 mod r : multiplying by it enters Montgomery form
#define CDX_R2_INIT {0xae216da7u, 0x1bb8e645u, 0xe35c59e3u, 0x53fe3ab1u, 0x53bb8085u, 0x8c49833du, 0x7f4e44a5u, 0x0216d0b1u}
// R mod r = Montgomery form of 1
#define CDX_ONE_INIT {0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u, 0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u}

// Table-driven reduction (reduce_tab below).  The 4 KB table lives in global memory and is copied into shared memory by
// every CTA that runs field arithmetic (load_reduce_tab(), first statement of each kernel): the row a thread needs
// depends on its own value, which constant memory would serialise and shared memory serves in two LDS.128.
#if defined(__CUDACC__)
static __device__ const uint32_t g_reduce_tab[8 * CDX_REDUCE_TAB_ENTRIES] = CDX_REDUCE_TAB_INIT;
__shared__ uint4 s_reduce_tab[2 * CDX_REDUCE_TAB_ENTRIES];
static __device__ __forceinline__ void load_reduce_tab() {
  for (uint32_t i = threadIdx.x; i < 2 * CDX_REDUCE_TAB_ENTRIES; i += blockDim.x)
    s_reduce_tab[i] = reinterpret_cast<const uint4*>(g_reduce_tab)[i];
  __syncthreads();
}
#endif

#if defined(__CUDA_ARCH__)
// ---------------------------------------------------------------------------------------------------------
// carry-chain primitives, sm_100a PTX.  Each asm statement is self-contained with respect to the carry flag.

// e/o <- products of the even/odd limbs of a with bi (first row: accumulators start empty)
CDX_D void mont_row_first(uint32_t* e, uint32_t* o, const uint32_t* a, uint32_t bi) {
  asm("{\n\t"
      ".reg .u64 t;\n\t"
      "mul.wide.u32 t, %16, %24; mov.b64 {%0, %1}, t;\n\t"
      "mul.wide.u32 t, %18, %24; mov.b64 {%2, %3}, t;\n\t"
      "mul.wide.u32 t, %20, %24; mov.b64 {%4, %5}, t;\n\t"
      "mul.wide.u32 t, %22, %24; mov.b64 {%6, %7}, t;\n\t"
      "mul.wide.u32 t, %17, %24; mov.b64 {%8, %9}, t;\n\t"
      "mul.wide.u32 t, %19, %24; mov.b64 {%10, %11}, t;\n\t"
      "mul.wide.u32 t, %21, %24; mov.b64 {%12, %13}, t;\n\t"
      "mul.wide.u32 t, %23, %24; mov.b64 {%14, %15}, t;\n\t"
      "}"
      : "=r"(e[0]), "=r"(e[1]), "=r"(e[2]), "=r"(e[3]), "=r"(e[4]), "=r"(e[5]), "=r"(e[6]), "=r"(e[7]),
        "=r"(o[0]), "=r"(o[1]), "=r"(o[2]), "=r"(o[3]), "=r"(o[4]), "=r"(o[5]), "=r"(o[6]), "=r"(o[7])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(bi));
}

// Next row.  On entry e is the even-aligned accumulator (limb positions 0..7) and o is the previous row's even
// accumulator after the divide by 2^32 (o[k] sits at position k-1, o[0] is dead).  On exit e holds positions
// 0..7 and o positions 1..8 of  T + a*bi.
CDX_D void mont_row_next(uint32_t* e, uint32_t* o, const uint32_t* a, uint32_t bi) {
  asm("{\n\t"
      "add.cc.u32 %0, %0, %9;\n\t"
      "madc.lo.cc.u32 %8, %17, %24, %10;  madc.hi.cc.u32 %9, %17, %24, %11;\n\t"
      "madc.lo.cc.u32 %10, %19, %24, %12; madc.hi.cc.u32 %11, %19, %24, %13;\n\t"
      "madc.lo.cc.u32 %12, %21, %24, %14; madc.hi.cc.u32 %13, %21, %24, %15;\n\t"
      "madc.lo.cc.u32 %14, %23, %24, 0;   madc.hi.u32 %15, %23, %24, 0;\n\t"
      "mad.lo.cc.u32 %0, %16, %24, %0;  madc.hi.cc.u32 %1, %16, %24, %1;\n\t"
      "madc.lo.cc.u32 %2, %18, %24, %2; madc.hi.cc.u32 %3, %18, %24, %3;\n\t"
      "madc.lo.cc.u32 %4, %20, %24, %4; madc.hi.cc.u32 %5, %20, %24, %5;\n\t"
      "madc.lo.cc.u32 %6, %22, %24, %6; madc.hi.cc.u32 %7, %22, %24, %7;\n\t"
      "addc.u32 %15, %15, 0;\n\t"
      "}"
      : "+r"(e[0]), "+r"(e[1]), "+r"(e[2]), "+r"(e[3]), "+r"(e[4]), "+r"(e[5]), "+r"(e[6]), "+r"(e[7]),
        "+r"(o[0]), "+r"(o[1]), "+r"(o[2]), "+r"(o[3]), "+r"(o[4]), "+r"(o[5]), "+r"(o[6]), "+r"(o[7])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(bi));
}

// Reduction row: m = e[0] * (-r^-1) mod 2^32; (e,o) += m*r, which clears e[0].  The modulus limbs are immediates.
// (Tried and rejected, see DESIGN.md section 4: r0 = 2^32 - 2^28 + 1 and -r^-1 = -(2^28 + 1) allow m and the r0 column
// to be computed with shifts/adds instead of two multiplies, but ptxas moves the extra adds onto the FMA pipe as
// IMAD.X / IMAD.IADD and the dependency chain gets longer: 16.0 vs 17.0 GB/s on the cell kernel.)
CDX_D void mont_row_redc(uint32_t* e, uint32_t* o) {
  uint32_t m = e[0] * CDX_NP;
  asm("{\n\t"
      "mad.lo.cc.u32 %8, %17, %16, %8;    madc.hi.cc.u32 %9, %17, %16, %9;\n\t"
      "madc.lo.cc.u32 %10, %19, %16, %10; madc.hi.cc.u32 %11, %19, %16, %11;\n\t"
      "madc.lo.cc.u32 %12, %21, %16, %12; madc.hi.cc.u32 %13, %21, %16, %13;\n\t"
      "madc.lo.cc.u32 %14, %23, %16, %14; madc.hi.u32 %15, %23, %16, %15;\n\t"
      "mad.lo.cc.u32 %0, %18, %16, %0;  madc.hi.cc.u32 %1, %18, %16, %1;\n\t"
      "madc.lo.cc.u32 %2, %20, %16, %2; madc.hi.cc.u32 %3, %20, %16, %3;\n\t"
      "madc.lo.cc.u32 %4, %22, %16, %4; madc.hi.cc.u32 %5, %22, %16, %5;\n\t"
      "madc.lo.cc.u32 %6, %24, %16, %6; madc.hi.cc.u32 %7, %24, %16, %7;\n\t"
      "addc.u32 %15, %15, 0;\n\t"
      "}"
      : "+r"(e[0]), "+r"(e[1]), "+r"(e[2]), "+r"(e[3]), "+r"(e[4]), "+r"(e[5]), "+r"(e[6]), "+r"(e[7]),
        "+r"(o[0]), "+r"(o[1]), "+r"(o[2]), "+r"(o[3]), "+r"(o[4]), "+r"(o[5]), "+r"(o[6]), "+r"(o[7])
      : "r"(m), "n"(CDX_N1), "n"(CDX_N0), "n"(CDX_N3), "n"(CDX_N2), "n"(CDX_N5), "n"(CDX_N4), "n"(CDX_N7),
        "n"(CDX_N6));
}

// Reduction-only iteration (used by the squaring, which reduces a finished 512-bit product): the same as
// mont_row_next followed by mont_row_redc with no product row in between.  On entry e is the even-aligned window
// (positions 0..7), o the previous window after its divide by 2^32 (o[k] at position k-1, o[0] dead).  On exit o is
// the new odd-aligned accumulator (positions 1..8) and e the even one with its low limb cancelled.
CDX_D void mont_row_redc_shift(uint32_t* e, uint32_t* o) {
  // the Montgomery factor comes from the true low limb of the window, e[0] + o[1]: formed once, by the add that also
  // starts the carry chain (mul.lo does not touch the condition code)
  asm("{\n\t"
      ".reg .u32 m;\n\t"
      "add.cc.u32 %0, %0, %9;\n\t"
      "mul.lo.u32 m, %0, %16;\n\t"
      "madc.lo.cc.u32 %8, %17, m, %10;  madc.hi.cc.u32 %9, %17, m, %11;\n\t"
      "madc.lo.cc.u32 %10, %19, m, %12; madc.hi.cc.u32 %11, %19, m, %13;\n\t"
      "madc.lo.cc.u32 %12, %21, m, %14; madc.hi.cc.u32 %13, %21, m, %15;\n\t"
      "madc.lo.cc.u32 %14, %23, m, 0;   madc.hi.u32 %15, %23, m, 0;\n\t"
      "mad.lo.cc.u32 %0, %18, m, %0;  madc.hi.cc.u32 %1, %18, m, %1;\n\t"
      "madc.lo.cc.u32 %2, %20, m, %2; madc.hi.cc.u32 %3, %20, m, %3;\n\t"
      "madc.lo.cc.u32 %4, %22, m, %4; madc.hi.cc.u32 %5, %22, m, %5;\n\t"
      "madc.lo.cc.u32 %6, %24, m, %6; madc.hi.cc.u32 %7, %24, m, %7;\n\t"
      "addc.u32 %15, %15, 0;\n\t"
      "}"
      : "+r"(e[0]), "+r"(e[1]), "+r"(e[2]), "+r"(e[3]), "+r"(e[4]), "+r"(e[5]), "+r"(e[6]), "+r"(e[7]),
        "+r"(o[0]), "+r"(o[1]), "+r"(o[2]), "+r"(o[3]), "+r"(o[4]), "+r"(o[5]), "+r"(o[6]), "+r"(o[7])
      : "n"(CDX_NP), "n"(CDX_N1), "n"(CDX_N0), "n"(CDX_N3), "n"(CDX_N2), "n"(CDX_N5), "n"(CDX_N4), "n"(CDX_N7),
        "n"(CDX_N6));
}

// Product rows of the streaming squaring (mont_sqr): mont_row_next with a multiplicand of NPROD < 8 limbs, i.e. the
// same window shift, with plain carry propagation where the row has no product.  v = multiplicand limbs, b = multiplier.
CDX_D void sqr_row7(uint32_t* e, uint32_t* o, const uint32_t* v, uint32_t b) {
  asm("{\n\t"
      "add.cc.u32 %0, %0, %9;\n\t"
      "madc.lo.cc.u32 %8, %17, %23, %10;  madc.hi.cc.u32 %9, %17, %23, %11;\n\t"
      "madc.lo.cc.u32 %10, %19, %23, %12; madc.hi.cc.u32 %11, %19, %23, %13;\n\t"
      "madc.lo.cc.u32 %12, %21, %23, %14; madc.hi.cc.u32 %13, %21, %23, %15;\n\t"
      "addc.u32 %14, 0, 0;                mov.u32 %15, 0;\n\t"
      "mad.lo.cc.u32 %0, %16, %23, %0;  madc.hi.cc.u32 %1, %16, %23, %1;\n\t"
      "madc.lo.cc.u32 %2, %18, %23, %2; madc.hi.cc.u32 %3, %18, %23, %3;\n\t"
      "madc.lo.cc.u32 %4, %20, %23, %4; madc.hi.cc.u32 %5, %20, %23, %5;\n\t"
      "madc.lo.cc.u32 %6, %22, %23, %6; madc.hi.cc.u32 %7, %22, %23, %7;\n\t"
      "addc.u32 %15, %15, 0;\n\t"
      "}"
      : "+r"(e[0]), "+r"(e[1]), "+r"(e[2]), "+r"(e[3]), "+r"(e[4]), "+r"(e[5]), "+r"(e[6]), "+r"(e[7]),
        "+r"(o[0]), "+r"(o[1]), "+r"(o[2]), "+r"(o[3]), "+r"(o[4]), "+r"(o[5]), "+r"(o[6]), "+r"(o[7])
      : "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(b));
}
CDX_D void sqr_row6(uint32_t* e, uint32_t* o, const uint32_t* v, uint32_t b) {
  asm("{\n\t"
      "add.cc.u32 %0, %0, %9;\n\t"
      "madc.lo.cc.u32 %8, %17, %22, %10;  madc.hi.cc.u32 %9, %17, %22, %11;\n\t"
      "madc.lo.cc.u32 %10, %19, %22, %12; madc.hi.cc.u32 %11, %19, %22, %13;\n\t"
      "madc.lo.cc.u32 %12, %21, %22, %14; madc.hi.cc.u32 %13, %21, %22, %15;\n\t"
      "addc.u32 %14, 0, 0;                mov.u32 %15, 0;\n\t"
      "mad.lo.cc.u32 %0, %16, %22, %0;  madc.hi.cc.u32 %1, %16, %22, %1;\n\t"
      "madc.lo.cc.u32 %2, %18, %22, %2; madc.hi.cc.u32 %3, %18, %22, %3;\n\t"
      "madc.lo.cc.u32 %4, %20, %22, %4; madc.hi.cc.u32 %5, %20, %22, %5;\n\t"
      "addc.cc.u32 %6, %6, 0;           addc.cc.u32 %7, %7, 0;\n\t"
      "addc.u32 %15, %15, 0;\n\t"
      "}"
      : "+r"(e[0]), "+r"(e[1]), "+r"(e[2]), "+r"(e[3]), "+r"(e[4]), "+r"(e[5]), "+r"(e[6]), "+r"(e[7]),
        "+r"(o[0]), "+r"(o[1]), "+r"(o[2]), "+r"(o[3]), "+r"(o[4]), "+r"(o[5]), "+r"(o[6]), "+r"(o[7])
      : "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(b));
}
CDX_D void sqr_row5(uint32_t* e, uint32_t* o, const uint32_t* v, uint32_t b) {
  asm("{\n\t"
      "add.cc.u32 %0, %0, %9;\n\t"
      "madc.lo.cc.u32 %8, %17, %21, %10;  madc.hi.cc.u32 %9, %17, %21, %11;\n\t"
      "madc.lo.cc.u32 %10, %19, %21, %12; madc.hi.cc.u32 %11, %19, %21, %13;\n\t"
      "addc.cc.u32 %12, %14, 0;           addc.cc.u32 %13, %15, 0;\n\t"
      "addc.u32 %14, 0, 0;                mov.u32 %15, 0;\n\t"
      "mad.lo.cc.u32 %0, %16, %21, %0;  madc.hi.cc.u32 %1, %16, %21, %1;\n\t"
      "madc.lo.cc.u32 %2, %18, %21, %2; madc.hi.cc.u32 %3, %18, %21, %3;\n\t"
      "madc.lo.cc.u32 %4, %20, %21, %4; madc.hi.cc.u32 %5, %20, %21, %5;\n\t"
      "addc.cc.u32 %6, %6, 0;           addc.cc.u32 %7, %7, 0;\n\t"
      "addc.u32 %15, %15, 0;\n\t"
      "}"
      : "+r"(e[0]), "+r"(e[1]), "+r"(e[2]), "+r"(e[3]), "+r"(e[4]), "+r"(e[5]), "+r"(e[6]), "+r"(e[7]),
        "+r"(o[0]), "+r"(o[1]), "+r"(o[2]), "+r"(o[3]), "+r"(o[4]), "+r"(o[5]), "+r"(o[6]), "+r"(o[7])
      : "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(b));
}
// The ten operand products that lie entirely in the upper half of the square (rows a4..a7), added to the final
// window: e pairs sit at result positions (0,1),(2,3),(4,5),(6,7), o[2..7] at (1,2),(3,4),(5,6) -- o ends one limb below
// the top of the result, so a carry out of o[7] goes into e[7].
// in: a4 a5 a6 a7 = %14..%17, s5 s6 s7 (a_j << 1) = %18..%20, d6 d7 (doubled limbs) = %21 %22
CDX_D void sqr_upper(uint32_t* e, uint32_t* o, const uint32_t* a, const uint32_t* sh, const uint32_t* d) {
  asm("{\n\t"
      "mad.lo.cc.u32 %8, %18, %14, %8;    madc.hi.cc.u32 %9, %18, %14, %9;\n\t"
      "madc.lo.cc.u32 %10, %22, %14, %10; madc.hi.cc.u32 %11, %22, %14, %11;\n\t"
      "madc.lo.cc.u32 %12, %20, %16, %12; madc.hi.cc.u32 %13, %20, %16, %13;\n\t"
      "addc.u32 %7, %7, 0;\n\t"
      "mad.lo.cc.u32 %10, %19, %15, %10;  madc.hi.cc.u32 %11, %19, %15, %11;\n\t"
      "addc.cc.u32 %12, %12, 0;           addc.cc.u32 %13, %13, 0;\n\t"
      "addc.u32 %7, %7, 0;\n\t"
      "mad.lo.cc.u32 %0, %14, %14, %0;  madc.hi.cc.u32 %1, %14, %14, %1;\n\t"
      "madc.lo.cc.u32 %2, %21, %14, %2; madc.hi.cc.u32 %3, %21, %14, %3;\n\t"
      "madc.lo.cc.u32 %4, %22, %15, %4; madc.hi.cc.u32 %5, %22, %15, %5;\n\t"
      "madc.lo.cc.u32 %6, %17, %17, %6; madc.hi.u32 %7, %17, %17, %7;\n\t"
      "mad.lo.cc.u32 %2, %15, %15, %2;  madc.hi.cc.u32 %3, %15, %15, %3;\n\t"
      "madc.lo.cc.u32 %4, %16, %16, %4; madc.hi.cc.u32 %5, %16, %16, %5;\n\t"
      "addc.cc.u32 %6, %6, 0;           addc.u32 %7, %7, 0;\n\t"
      "}"
      : "+r"(e[0]), "+r"(e[1]), "+r"(e[2]), "+r"(e[3]), "+r"(e[4]), "+r"(e[5]), "+r"(e[6]), "+r"(e[7]),
        "+r"(o[2]), "+r"(o[3]), "+r"(o[4]), "+r"(o[5]), "+r"(o[6]), "+r"(o[7])
      : "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(sh[5]), "r"(sh[6]), "r"(sh[7]), "r"(d[6]), "r"(d[7]));
}

CDX_D uint32_t shl1_funnel(uint32_t lo, uint32_t hi) { return __funnelshift_l(lo, hi, 1); }              // (hi:lo << 1) >> 32

// r = e + (o >> 32 limbs aligned as after a row): r[k] = e[k] + o[k+1] with carry
CDX_D void mont_merge(uint32_t* r, const uint32_t* e, const uint32_t* o) {
  asm("add.cc.u32 %0, %8, %16; addc.cc.u32 %1, %9, %17; addc.cc.u32 %2, %10, %18; addc.cc.u32 %3, %11, %19;"
      "addc.cc.u32 %4, %12, %20; addc.cc.u32 %5, %13, %21; addc.cc.u32 %6, %14, %22; addc.u32 %7, %15, 0;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
      : "r"(e[0]), "r"(e[1]), "r"(e[2]), "r"(e[3]), "r"(e[4]), "r"(e[5]), "r"(e[6]), "r"(e[7]), "r"(o[1]),
        "r"(o[2]), "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]));
}

// r = a + b (256-bit, carry out discarded: callers keep sums below 2^256)
CDX_D void add256(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  asm("add.cc.u32 %0, %8, %16; addc.cc.u32 %1, %9, %17; addc.cc.u32 %2, %10, %18; addc.cc.u32 %3, %11, %19;"
      "addc.cc.u32 %4, %12, %20; addc.cc.u32 %5, %13, %21; addc.cc.u32 %6, %14, %22; addc.u32 %7, %15, %23;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(b[0]),
        "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
}

// d = a - r (the modulus) mod 2^256.  For a < 2r the sign bit of d tells whether that borrowed: a < r gives
// d > 2^256 - r > 2^255, a >= r gives d < r < 2^254 -- so the caller needs neither the borrow flag nor a compare chain.
CDX_D void sub_modulus(uint32_t* d, const uint32_t* a) {
  asm("sub.cc.u32 %0, %8, %16; subc.cc.u32 %1, %9, %17; subc.cc.u32 %2, %10, %18; subc.cc.u32 %3, %11, %19;"
      "subc.cc.u32 %4, %12, %20; subc.cc.u32 %5, %13, %21; subc.cc.u32 %6, %14, %22; subc.u32 %7, %15, %23;"
      : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "n"(CDX_N0),
        "n"(CDX_N1), "n"(CDX_N2), "n"(CDX_N3), "n"(CDX_N4), "n"(CDX_N5), "n"(CDX_N6), "n"(CDX_N7));
}

// r = a - b mod 2^256 (callers guarantee a >= b, or a + 2^256 >= b when the 257th bit is implied: see reduce_tab)
CDX_D void sub256(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  asm("sub.cc.u32 %0, %8, %16; subc.cc.u32 %1, %9, %17; subc.cc.u32 %2, %10, %18; subc.cc.u32 %3, %11, %19;"
      "subc.cc.u32 %4, %12, %20; subc.cc.u32 %5, %13, %21; subc.cc.u32 %6, %14, %22; subc.u32 %7, %15, %23;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(b[0]),
        "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
}

// r = a + b mod 2^256, returns the carry out (bit 256 of the sum)
CDX_D uint32_t add256c(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  uint32_t c;
  asm("add.cc.u32 %0, %9, %17; addc.cc.u32 %1, %10, %18; addc.cc.u32 %2, %11, %19; addc.cc.u32 %3, %12, %20;"
      "addc.cc.u32 %4, %13, %21; addc.cc.u32 %5, %14, %22; addc.cc.u32 %6, %15, %23; addc.cc.u32 %7, %16, %24;"
      "addc.u32 %8, 0, 0;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(c)
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(b[0]),
        "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
  return c;
}

// index of a 257-bit value (carry : limb 7) in the reduction table: its top 7 bits
CDX_D uint32_t reduce_tab_index(uint32_t top_limb, uint32_t carry) { return __funnelshift_l(top_limb, carry, 6); }
// row `idx` of the reduction table, from the CTA's shared-memory copy (two 16-byte loads on the otherwise idle LSU pipe)
CDX_D void reduce_tab_row(uint32_t* t, uint32_t idx) {
  const uint4 a = s_reduce_tab[2 * idx], b = s_reduce_tab[2 * idx + 1];
  t[0] = a.x; t[1] = a.y; t[2] = a.z; t[3] = a.w; t[4] = b.x; t[5] = b.y; t[6] = b.z; t[7] = b.w;
}
#elif defined(CDX_HOST_EMUL)
}  // namespace cdx
#include "fr_rows_host.h"  // tests/host_emul: C emulation of the primitives above (unit tests only)
namespace cdx {
#else
#if !defined(__CUDACC__)
#error "fr.cuh is sm_100a device code; host compilation is only for unit tests with -DCDX_HOST_EMUL"
#endif
// host pass of nvcc: declarations only, never called
void mont_row_first(uint32_t*, uint32_t*, const uint32_t*, uint32_t);
void mont_row_next(uint32_t*, uint32_t*, const uint32_t*, uint32_t);
void mont_row_redc(uint32_t*, uint32_t*);
void mont_row_redc_shift(uint32_t*, uint32_t*);
void sqr_row7(uint32_t*, uint32_t*, const uint32_t*, uint32_t);
void sqr_row6(uint32_t*, uint32_t*, const uint32_t*, uint32_t);
void sqr_row5(uint32_t*, uint32_t*, const uint32_t*, uint32_t);
void sqr_upper(uint32_t*, uint32_t*, const uint32_t*, const uint32_t*, const uint32_t*);
uint32_t shl1_funnel(uint32_t, uint32_t);
void mont_merge(uint32_t*, const uint32_t*, const uint32_t*);
void add256(uint32_t*, const uint32_t*, const uint32_t*);
void sub_modulus(uint32_t*, const uint32_t*);
void sub256(uint32_t*, const uint32_t*, const uint32_t*);
uint32_t add256c(uint32_t*, const uint32_t*, const uint32_t*);
uint32_t reduce_tab_index(uint32_t, uint32_t);
void reduce_tab_row(uint32_t*, uint32_t);
#endif

// ---------------------------------------------------------------------------------------------------------
// field operations built on the primitives (shared between device code and the host-emulated unit tests)

// a*b*2^-256 mod r, lazily reduced (see range discipline at the top)
CDX_D Fr mont_mul(const Fr& a, const Fr& b) {
  uint32_t e[8], o[8];
  mont_row_first(e, o, a.l, b.l[0]);
  mont_row_redc(e, o);
  mont_row_next(o, e, a.l, b.l[1]);
  mont_row_redc(o, e);
#pragma unroll
  for (int i = 2; i < 8; i += 2) {
    mont_row_next(e, o, a.l, b.l[i]);
    mont_row_redc(e, o);
    mont_row_next(o, e, a.l, b.l[i + 1]);
    mont_row_redc(o, e);
  }
  Fr r;
  mont_merge(r.l, e, o);
  return r;
}

// a*a*2^-256 mod r with 36 instead of 64 operand products (2/3 of all multiplications on this path are squarings:
// x^2 and x^4 of every S-box), in the same row pipeline as mont_mul.  a < 2r (so a < 2^255); result < a^2/2^256 + r < 2r.
//   a^2 = sum_i a_i * V_i * 2^(64 i),  V_i = a_i + 2^32 * 2*floor(a / 2^(32(i+1)))   (8-i limbs: a_i, a_{i+1} << 1, then
//   the limbs of 2a), so product row i starts at limb 2i and the window can drop two limbs (two reduction rows) per
//   product row.  Rows 4..7 lie entirely above the eight reduced limbs and are added to the final window.
// Window bound: after two reduction rows the window is < r + small, a product row adds < 2a, so it stays < 2^256
// as long as 2a + r < 2^256, i.e. a < 2.14 r.
CDX_D Fr mont_sqr(const Fr& a) {
  uint32_t d[8], sh[8];
#pragma unroll
  for (int j = 1; j < 8; ++j) {
    d[j] = shl1_funnel(a.l[j - 1], a.l[j]);     // limb j of 2a
    sh[j] = shl1_funnel(0u, a.l[j]);            // a_j << 1 = limb j of 2*(a with the limbs below j cleared)
  }
  d[0] = sh[0] = 0;
  const uint32_t v0[8] = {a.l[0], sh[1], d[2], d[3], d[4], d[5], d[6], d[7]};
  const uint32_t v1[7] = {a.l[1], sh[2], d[3], d[4], d[5], d[6], d[7]};
  const uint32_t v2[6] = {a.l[2], sh[3], d[4], d[5], d[6], d[7]};
  const uint32_t v3[5] = {a.l[3], sh[4], d[5], d[6], d[7]};
  uint32_t e[8], o[8];
  mont_row_first(e, o, v0, a.l[0]);
  mont_row_redc(e, o);
  mont_row_redc_shift(o, e);
  sqr_row7(e, o, v1, a.l[1]);
  mont_row_redc(e, o);
  mont_row_redc_shift(o, e);
  sqr_row6(e, o, v2, a.l[2]);
  mont_row_redc(e, o);
  mont_row_redc_shift(o, e);
  sqr_row5(e, o, v3, a.l[3]);
  mont_row_redc(e, o);
  mont_row_redc_shift(o, e);
  sqr_upper(e, o, a.l, sh, d);
  Fr r;
  mont_merge(r.l, e, o);
  return r;
}

// [0, 2r) -> [0, r)
CDX_D Fr reduce_once(const Fr& a) {
  Fr d;
  sub_modulus(d.l, a.l);
  const bool lt = (int32_t)d.l[7] < 0;   // a < r (see sub_modulus)
  Fr r;
#pragma unroll
  for (int i = 0; i < 8; ++i) r.l[i] = lt ? a.l[i] : d.l[i];
  return r;
}

// Table-driven reduction: any v < 2^257 (given as 256 bits + the carry out of the add that produced it) -> the
// congruent value v - q r in [0, r + 2^250) = [0, 1.0827 r), q = floor(floor(v / 2^250) * 2^250 / r) read from a table by the
// top 7 bits of v.  One shift, two shared-memory loads and ONE subtraction chain: no trial subtraction, no select.
// "B" below is this bound, 1.0827 r.
CDX_D Fr reduce_tab(const Fr& v, uint32_t carry) {
  uint32_t t[8];
  reduce_tab_row(t, reduce_tab_index(v.l[7], carry));
  Fr r;
  sub256(r.l, v.l, t);
  return r;
}

// a + b, 257-bit sum -> [0, B)
CDX_D Fr add_reduce(const Fr& a, const Fr& b) {
  Fr s;
  const uint32_t c = add256c(s.l, a.l, b.l);
  return reduce_tab(s, c);
}

// a + b without reduction (caller guarantees a + b < 2^256)
CDX_D Fr add_lazy(const Fr& a, const Fr& b) {
  Fr r;
  add256(r.l, a.l, b.l);
  return r;
}

// inputs < r, output < r
CDX_D Fr add_mod(const Fr& a, const Fr& b) { return reduce_once(add_lazy(a, b)); }
CDX_D Fr dbl_mod(const Fr& a) { return reduce_once(add_lazy(a, a)); }

CDX_D Fr fr_zero() {
  Fr r;
#pragma unroll
  for (int i = 0; i < 8; ++i) r.l[i] = 0;
  return r;
}

// standard form (< 2^256, any value) -> Montgomery form, result < r
CDX_D Fr to_mont(const Fr& a) {
  const Fr r2 = {CDX_R2_INIT};
  // R2 < r is the row operand, a < 2^256 = 5.29 r the per-row multiplier  =>  product < 2r
  return reduce_once(mont_mul(r2, a));
}

// Montgomery form (< 2r) -> canonical standard form (< r)
CDX_D Fr from_mont(const Fr& a) {
  Fr one = fr_zero();
  one.l[0] = 1;
  return reduce_once(mont_mul(a, one));  // a < 2r is the row operand
}

// small constant c (< 2^32) in Montgomery form
CDX_D Fr mont_from_u32(uint32_t c) {
  Fr a = fr_zero();
  a.l[0] = c;
  return to_mont(a);
}

}  // namespace cdx
