// fr_rows_host.h -- UNIT-TEST ONLY.  Plain C++ emulation of the inline-PTX carry-chain primitives of
// codex-storage-proofs-circuits_b200/csrc/fr.cuh, instruction by instruction (an explicit carry flag stands in
// for the PTX condition code), so the limb-level logic layered on top of them can be exercised by g++ on a
// machine without a GPU (tests/test_limb_logic_host.py).  Never compiled into the product library.
#pragma once
#include <stdint.h>

namespace cdx {
namespace emul {
struct CC {
  uint32_t cf = 0;
  uint32_t add_cc(uint32_t a, uint32_t b) { uint64_t s = (uint64_t)a + b; cf = (uint32_t)(s >> 32); return (uint32_t)s; }
  uint32_t addc_cc(uint32_t a, uint32_t b) { uint64_t s = (uint64_t)a + b + cf; cf = (uint32_t)(s >> 32); return (uint32_t)s; }
  uint32_t addc(uint32_t a, uint32_t b) { return a + b + cf; }
  uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return add_cc((uint32_t)((uint64_t)a * b), c); }
  uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return addc_cc((uint32_t)((uint64_t)a * b), c); }
  uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return addc_cc((uint32_t)(((uint64_t)a * b) >> 32), c); }
  uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { return addc((uint32_t)(((uint64_t)a * b) >> 32), c); }
  uint32_t sub_cc(uint32_t a, uint32_t b) { uint64_t d = (uint64_t)a - b; cf = (uint32_t)((d >> 32) & 1); return (uint32_t)d; }  // cf = borrow
  uint32_t subc_cc(uint32_t a, uint32_t b) { uint64_t d = (uint64_t)a - b - cf; cf = (uint32_t)((d >> 32) & 1); return (uint32_t)d; }
};
}  // namespace emul

static const uint32_t kModHost[8] = {CDX_N0, CDX_N1, CDX_N2, CDX_N3, CDX_N4, CDX_N5, CDX_N6, CDX_N7};

inline void mont_row_first(uint32_t* e, uint32_t* o, const uint32_t* a, uint32_t bi) {
  for (int j = 0; j < 8; j += 2) {
    uint64_t p = (uint64_t)a[j] * bi;      e[j] = (uint32_t)p; e[j + 1] = (uint32_t)(p >> 32);
    uint64_t q = (uint64_t)a[j + 1] * bi;  o[j] = (uint32_t)q; o[j + 1] = (uint32_t)(q >> 32);
  }
}

inline void mont_row_next(uint32_t* e, uint32_t* o, const uint32_t* a, uint32_t bi) {
  emul::CC c;
  e[0] = c.add_cc(e[0], o[1]);
  o[0] = c.madc_lo_cc(a[1], bi, o[2]); o[1] = c.madc_hi_cc(a[1], bi, o[3]);
  o[2] = c.madc_lo_cc(a[3], bi, o[4]); o[3] = c.madc_hi_cc(a[3], bi, o[5]);
  o[4] = c.madc_lo_cc(a[5], bi, o[6]); o[5] = c.madc_hi_cc(a[5], bi, o[7]);
  o[6] = c.madc_lo_cc(a[7], bi, 0);    o[7] = c.madc_hi(a[7], bi, 0);
  e[0] = c.mad_lo_cc(a[0], bi, e[0]);  e[1] = c.madc_hi_cc(a[0], bi, e[1]);
  e[2] = c.madc_lo_cc(a[2], bi, e[2]); e[3] = c.madc_hi_cc(a[2], bi, e[3]);
  e[4] = c.madc_lo_cc(a[4], bi, e[4]); e[5] = c.madc_hi_cc(a[4], bi, e[5]);
  e[6] = c.madc_lo_cc(a[6], bi, e[6]); e[7] = c.madc_hi_cc(a[6], bi, e[7]);
  o[7] = c.addc(o[7], 0);
}

inline void mont_row_redc(uint32_t* e, uint32_t* o) {
  const uint32_t* n = kModHost;
  uint32_t m = e[0] * CDX_NP;
  emul::CC c;
  o[0] = c.mad_lo_cc(n[1], m, o[0]);  o[1] = c.madc_hi_cc(n[1], m, o[1]);
  o[2] = c.madc_lo_cc(n[3], m, o[2]); o[3] = c.madc_hi_cc(n[3], m, o[3]);
  o[4] = c.madc_lo_cc(n[5], m, o[4]); o[5] = c.madc_hi_cc(n[5], m, o[5]);
  o[6] = c.madc_lo_cc(n[7], m, o[6]); o[7] = c.madc_hi(n[7], m, o[7]);
  e[0] = c.mad_lo_cc(n[0], m, e[0]);  e[1] = c.madc_hi_cc(n[0], m, e[1]);
  e[2] = c.madc_lo_cc(n[2], m, e[2]); e[3] = c.madc_hi_cc(n[2], m, e[3]);
  e[4] = c.madc_lo_cc(n[4], m, e[4]); e[5] = c.madc_hi_cc(n[4], m, e[5]);
  e[6] = c.madc_lo_cc(n[6], m, e[6]); e[7] = c.madc_hi_cc(n[6], m, e[7]);
  o[7] = c.addc(o[7], 0);
}

inline void mont_row_redc_shift(uint32_t* e, uint32_t* o) {
  const uint32_t* n = kModHost;
  const uint32_t m = (e[0] + o[1]) * CDX_NP;
  emul::CC c;
  e[0] = c.add_cc(e[0], o[1]);
  o[0] = c.madc_lo_cc(n[1], m, o[2]); o[1] = c.madc_hi_cc(n[1], m, o[3]);
  o[2] = c.madc_lo_cc(n[3], m, o[4]); o[3] = c.madc_hi_cc(n[3], m, o[5]);
  o[4] = c.madc_lo_cc(n[5], m, o[6]); o[5] = c.madc_hi_cc(n[5], m, o[7]);
  o[6] = c.madc_lo_cc(n[7], m, 0);    o[7] = c.madc_hi(n[7], m, 0);
  e[0] = c.mad_lo_cc(n[0], m, e[0]);  e[1] = c.madc_hi_cc(n[0], m, e[1]);
  e[2] = c.madc_lo_cc(n[2], m, e[2]); e[3] = c.madc_hi_cc(n[2], m, e[3]);
  e[4] = c.madc_lo_cc(n[4], m, e[4]); e[5] = c.madc_hi_cc(n[4], m, e[5]);
  e[6] = c.madc_lo_cc(n[6], m, e[6]); e[7] = c.madc_hi_cc(n[6], m, e[7]);
  o[7] = c.addc(o[7], 0);
}

template <int NPROD>
inline void sqr_row_emul(uint32_t* e, uint32_t* o, const uint32_t* v, uint32_t b) {
  emul::CC c;
  e[0] = c.add_cc(e[0], o[1]);
  for (int j = 0; j < 3; ++j) {
    const int k = 2 * j + 1;
    if (k < NPROD) { o[2 * j] = c.madc_lo_cc(v[k], b, o[2 * j + 2]); o[2 * j + 1] = c.madc_hi_cc(v[k], b, o[2 * j + 3]); }
    else { o[2 * j] = c.addc_cc(o[2 * j + 2], 0); o[2 * j + 1] = c.addc_cc(o[2 * j + 3], 0); }
  }
  o[6] = c.addc(0, 0); o[7] = 0;                    // k = 7 is never present (NPROD <= 7)
  c.cf = 0;
  for (int j = 0; j < 4; ++j) {
    const int k = 2 * j;
    if (k < NPROD) {
      e[2 * j] = j == 0 ? c.mad_lo_cc(v[k], b, e[0]) : c.madc_lo_cc(v[k], b, e[2 * j]);
      e[2 * j + 1] = c.madc_hi_cc(v[k], b, e[2 * j + 1]);
    } else { e[2 * j] = c.addc_cc(e[2 * j], 0); e[2 * j + 1] = c.addc_cc(e[2 * j + 1], 0); }
  }
  o[7] = c.addc(o[7], 0);
}
inline void sqr_row7(uint32_t* e, uint32_t* o, const uint32_t* v, uint32_t b) { sqr_row_emul<7>(e, o, v, b); }
inline void sqr_row6(uint32_t* e, uint32_t* o, const uint32_t* v, uint32_t b) { sqr_row_emul<6>(e, o, v, b); }
inline void sqr_row5(uint32_t* e, uint32_t* o, const uint32_t* v, uint32_t b) { sqr_row_emul<5>(e, o, v, b); }

inline void sqr_upper(uint32_t* e, uint32_t* o, const uint32_t* a, const uint32_t* sh, const uint32_t* d) {
  emul::CC c;
  o[2] = c.mad_lo_cc(sh[5], a[4], o[2]); o[3] = c.madc_hi_cc(sh[5], a[4], o[3]);
  o[4] = c.madc_lo_cc(d[7], a[4], o[4]); o[5] = c.madc_hi_cc(d[7], a[4], o[5]);
  o[6] = c.madc_lo_cc(sh[7], a[6], o[6]); o[7] = c.madc_hi_cc(sh[7], a[6], o[7]);
  e[7] = c.addc_cc(e[7], 0);
  if (c.cf) __builtin_trap();            // the result fits 256 bits: nothing may carry out of e[7]
  o[4] = c.mad_lo_cc(sh[6], a[5], o[4]); o[5] = c.madc_hi_cc(sh[6], a[5], o[5]);
  o[6] = c.addc_cc(o[6], 0);             o[7] = c.addc_cc(o[7], 0);
  e[7] = c.addc_cc(e[7], 0);
  if (c.cf) __builtin_trap();
  e[0] = c.mad_lo_cc(a[4], a[4], e[0]);  e[1] = c.madc_hi_cc(a[4], a[4], e[1]);
  e[2] = c.madc_lo_cc(d[6], a[4], e[2]); e[3] = c.madc_hi_cc(d[6], a[4], e[3]);
  e[4] = c.madc_lo_cc(d[7], a[5], e[4]); e[5] = c.madc_hi_cc(d[7], a[5], e[5]);
  e[6] = c.madc_lo_cc(a[7], a[7], e[6]); e[7] = c.madc_hi_cc(a[7], a[7], e[7]);
  if (c.cf) __builtin_trap();
  e[2] = c.mad_lo_cc(a[5], a[5], e[2]);  e[3] = c.madc_hi_cc(a[5], a[5], e[3]);
  e[4] = c.madc_lo_cc(a[6], a[6], e[4]); e[5] = c.madc_hi_cc(a[6], a[6], e[5]);
  e[6] = c.addc_cc(e[6], 0);             e[7] = c.addc_cc(e[7], 0);
  if (c.cf) __builtin_trap();
}

inline uint32_t shl1_funnel(uint32_t lo, uint32_t hi) { return (hi << 1) | (lo >> 31); }

inline void mont_merge(uint32_t* r, const uint32_t* e, const uint32_t* o) {
  emul::CC c;
  r[0] = c.add_cc(e[0], o[1]);
  for (int k = 1; k < 7; ++k) r[k] = c.addc_cc(e[k], o[k + 1]);
  r[7] = c.addc(e[7], 0);
}

inline void add256(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  emul::CC c;
  r[0] = c.add_cc(a[0], b[0]);
  for (int k = 1; k < 7; ++k) r[k] = c.addc_cc(a[k], b[k]);
  r[7] = c.addc_cc(a[7], b[7]);
  if (c.cf) __builtin_trap();          // add_lazy's contract: the sum fits 256 bits
}

inline void sub_modulus(uint32_t* d, const uint32_t* a) {
  emul::CC c;
  d[0] = c.sub_cc(a[0], kModHost[0]);
  for (int k = 1; k < 8; ++k) d[k] = c.subc_cc(a[k], kModHost[k]);
  if ((c.cf & 1u) != (d[7] >> 31)) __builtin_trap();   // the caller reads the borrow off the sign bit: only valid for a < 2r
}
static const uint32_t kReduceTabHost[8 * CDX_REDUCE_TAB_ENTRIES] = CDX_REDUCE_TAB_INIT;

inline void sub256(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  emul::CC c;
  r[0] = c.sub_cc(a[0], b[0]);
  for (int k = 1; k < 8; ++k) r[k] = c.subc_cc(a[k], b[k]);
  // reduce_tab's contract: the difference is < r + 2^250, whatever borrow the implied 257th bit absorbed
  static const uint32_t bound[8] = {CDX_N0, CDX_N1, CDX_N2, CDX_N3, CDX_N4, CDX_N5, CDX_N6, CDX_N7 + (1u << 26)};
  for (int k = 7; k >= 0; --k) {
    if (r[k] < bound[k]) break;
    if (r[k] > bound[k] || k == 0) __builtin_trap();
  }
}

// the add256 emulation above silently wraps; this one is for sums that are allowed to reach 257 bits
inline uint32_t add256c(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  emul::CC c;
  r[0] = c.add_cc(a[0], b[0]);
  for (int k = 1; k < 8; ++k) r[k] = c.addc_cc(a[k], b[k]);
  return c.cf;
}

inline uint32_t reduce_tab_index(uint32_t top_limb, uint32_t carry) {
  if (carry > 1) __builtin_trap();
  return (carry << 6) | (top_limb >> 26);
}
inline void reduce_tab_row(uint32_t* t, uint32_t idx) {
  if (idx >= CDX_REDUCE_TAB_ENTRIES) __builtin_trap();
  for (int k = 0; k < 8; ++k) t[k] = kReduceTabHost[8 * idx + k];
}
}  // namespace cdx
