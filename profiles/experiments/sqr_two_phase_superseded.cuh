// Superseded by the streaming squaring in csrc/fr.cuh (round 1, +5.0 % on the cell kernel: 626.8 -> 598.0 ms).
// First dedicated squaring: full 512-bit product (cross products into even/odd accumulators, doubling pass, diagonal
// pass), then eight reduction-only rows.  Kept for reference; not compiled.
// a*a*2^-256 mod r with 36 instead of 64 operand products (2/3 of all multiplications on this path are squarings:
// x^2 and x^4 of every S-box).  a < 2r; result < a^2/2^256 + r < 2r, same contract as mont_mul.
//   1. cross products a_i*a_j (i < j), 28 of them, row by row into an even- and an odd-aligned 512-bit accumulator
//      (a chain never has to ripple: the limb above its last pair has only ever received carries);
//   2. T = 2*(E + O) + sum a_i^2 * 2^(64 i): one add chain, one funnel-shift pass, one 8-product carry chain;
//   3. Montgomery-reduce the low half with 8 reduction-only rows, add the high half (< 0.76 r, no overflow).
CDX_D Fr mont_sqr(const Fr& a) {
  uint32_t E[16], O[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) E[k] = O[k] = 0;
#pragma unroll
  for (int i = 0; i < 7; ++i) {
    {  // products landing on even positions i+j
      bool first = true;
      int last = -1;
#pragma unroll
      for (int j = i + 1; j < 8; ++j) {
        if (((i + j) & 1) == 0) {
          if (first) cc_mad_first(E[i + j], E[i + j + 1], a.l[i], a.l[j]);
          else cc_mad_next(E[i + j], E[i + j + 1], a.l[i], a.l[j]);
          first = false;
          last = i + j;
        }
      }
      if (!first) cc_carry_into(E[last + 2]);
    }
    {  // products landing on odd positions
      bool first = true;
      int last = -1;
#pragma unroll
      for (int j = i + 1; j < 8; ++j) {
        if (((i + j) & 1) == 1) {
          if (first) cc_mad_first(O[i + j], O[i + j + 1], a.l[i], a.l[j]);
          else cc_mad_next(O[i + j], O[i + j + 1], a.l[i], a.l[j]);
          first = false;
          last = i + j;
        }
      }
      if (!first) cc_carry_into(O[last + 2]);
    }
  }
  // S = E + O (positions 1..15), then T = 2S
  uint32_t S[16], T[16];
  S[0] = 0;
  cc_add_first(S[1], E[1], O[1]);
#pragma unroll
  for (int k = 2; k < 15; ++k) cc_add_next(S[k], E[k], O[k]);
  cc_add_last(S[15], E[15], O[15]);
  T[0] = 0;
#pragma unroll
  for (int k = 1; k < 16; ++k) T[k] = shl1_funnel(S[k - 1], S[k]);
  // + diagonal squares, one carry chain over all 16 limbs
  cc_mad_first(T[0], T[1], a.l[0], a.l[0]);
#pragma unroll
  for (int i = 1; i < 8; ++i) cc_mad_next(T[2 * i], T[2 * i + 1], a.l[i], a.l[i]);
  // reduce the low half
  uint32_t e[8], o[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    e[k] = T[k];
    o[k] = 0;
  }
#pragma unroll
  for (int i = 0; i < 8; i += 2) {
    mont_row_redc_shift(e, o);
    mont_row_redc_shift(o, e);
  }
  Fr u, hi, r;
  mont_merge(u.l, e, o);
#pragma unroll
  for (int k = 0; k < 8; ++k) hi.l[k] = T[8 + k];
  add256(r.l, u.l, hi.l);
  return r;
}
