/* codex_oracle.c -- CPU restatement of the slot-commitment path (see codex_oracle.h: test infrastructure only).
 *
 * Arithmetic: BN254 scalar field Fr, 4x64-bit limbs, Montgomery form (R = 2^256), CIOS multiplication on
 * unsigned __int128.  The reference does this arithmetic in third-party libraries absent from /root/reference
 * (constantine @ bc3845aa for Nim, zikkurat-algebra 0.0.1 for Haskell); this file restates the published
 * algorithm (word-serial Montgomery multiplication) and everything above it follows the in-repo sources cited
 * at each function.
 */
#include "codex_oracle.h"
#include "poseidon2_rc.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef struct { uint64_t l[4]; } fr;

/* r = 21888242871839275222246405745257275088548364400416034343698204186575808495617  (README.md:76) */
static const uint64_t FR_MOD[4] = {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
static const uint64_t FR_NINV = 0xc2e1f593efffffffull;                 /* -r^-1 mod 2^64 */
static const fr FR_R2 = {{0x1bb8e645ae216da7ull, 0x53fe3ab1e35c59e3ull, 0x8c49833d53bb8085ull, 0x0216d0b17f4e44a5ull}};  /* 2^512 mod r */

static inline int fr_geq_mod(const uint64_t a[4]) {
  for (int i = 3; i >= 0; --i) {
    if (a[i] > FR_MOD[i]) return 1;
    if (a[i] < FR_MOD[i]) return 0;
  }
  return 1;
}

static inline void fr_sub_mod_inplace(uint64_t a[4]) {
  u128 b = 0;
  for (int i = 0; i < 4; ++i) {
    u128 d = (u128)a[i] - FR_MOD[i] - (uint64_t)b;
    a[i] = (uint64_t)d;
    b = (d >> 64) & 1;
  }
}

/* [0, 2r) -> [0, r) without a data-dependent branch (the comparison is a coin flip on this path: a mispredicted branch
 * per addition would cost more than the addition): trial subtraction, then select by the borrow */
static inline void fr_reduce_once(uint64_t a[4]) {
  uint64_t d[4];
  u128 b = 0;
  for (int i = 0; i < 4; ++i) {
    u128 t = (u128)a[i] - FR_MOD[i] - (uint64_t)b;
    d[i] = (uint64_t)t;
    b = (t >> 64) & 1;
  }
  const uint64_t keep = (uint64_t)0 - (uint64_t)b;      /* all ones if a < r */
  for (int i = 0; i < 4; ++i) a[i] = (a[i] & keep) | (d[i] & ~keep);
}

static inline fr fr_add(fr a, fr b) {
  fr c; u128 cy = 0;
  for (int i = 0; i < 4; ++i) { cy += (u128)a.l[i] + b.l[i]; c.l[i] = (uint64_t)cy; cy >>= 64; }
  /* r < 2^254 so a+b < 2^255: no carry out of limb 3 */
  fr_reduce_once(c.l);
  return c;
}

static inline fr fr_dbl(fr a) { return fr_add(a, a); }

/* Montgomery product a*b*2^-256 mod r, inputs/outputs fully reduced.  Word-serial (CIOS) form with the
 * multiply and reduce rows interleaved; because r < 2^254 the running value never needs a fifth limb
 * (the usual simplification for moduli with spare top bits). */
#define MAC(hi, lo, x, y, add1, add2) do { u128 _p = (u128)(x) * (y) + (add1) + (add2); lo = (uint64_t)_p; hi = (uint64_t)(_p >> 64); } while (0)
static inline fr fr_mul_c(fr a, fr b) {
  uint64_t t0 = 0, t1 = 0, t2 = 0, t3 = 0;
#define ROW(bi) do {                                         \
    uint64_t A, Cc, m, lo;                                   \
    MAC(A, t0, a.l[0], (bi), t0, 0);                         \
    m = t0 * FR_NINV;                                        \
    MAC(Cc, lo, m, FR_MOD[0], t0, 0); (void)lo;              \
    MAC(A, t1, a.l[1], (bi), t1, A);                         \
    MAC(Cc, t0, m, FR_MOD[1], t1, Cc);                       \
    MAC(A, t2, a.l[2], (bi), t2, A);                         \
    MAC(Cc, t1, m, FR_MOD[2], t2, Cc);                       \
    MAC(A, t3, a.l[3], (bi), t3, A);                         \
    MAC(Cc, t2, m, FR_MOD[3], t3, Cc);                       \
    t3 = Cc + A;                                             \
  } while (0)
  ROW(b.l[0]); ROW(b.l[1]); ROW(b.l[2]); ROW(b.l[3]);
#undef ROW
  fr r = {{t0, t1, t2, t3}};
  fr_reduce_once(r.l);
  return r;
}


#if defined(__x86_64__) && defined(__BMI2__) && defined(__ADX__) && !defined(ORC_NO_ASM)
/* The same word-serial product with the instructions tuned field libraries use on x86-64 (constantine, the reference's
 * arithmetic backend, among them): MULX for the 64x64->128 products and the two independent carry chains of ADCX (CF) and
 * ADOX (OF), so the low and the high halves of a row accumulate in parallel.  Per row: t += a*b_i (t4 = the word above),
 * m = t0 * (-r^-1), t = (t + m*r) / 2^64.  The five accumulator registers rotate instead of being shifted.  r < 2^254
 * keeps t4 within one word (the "no-carry" form).  Built only where the compiler targets BMI2+ADX; fr_mul_c is the
 * portable statement of the same function and the unit tests compare the two. */
#define ORC_ROW(T0, T1, T2, T3, T4, OFF)                                          \
  "xorl %%eax, %%eax\n\t"                                                        \
  "movq " #OFF "(%[b]), %%rdx\n\t"                                               \
  "mulx 0(%[a]), %%r13, %%r14\n\t"   "adox %%r13, " T0 "\n\t"                    \
  "mulx 8(%[a]), %%r13, %%r15\n\t"   "adcx %%r14, " T1 "\n\t" "adox %%r13, " T1 "\n\t" \
  "mulx 16(%[a]), %%r13, %%r14\n\t"  "adcx %%r15, " T2 "\n\t" "adox %%r13, " T2 "\n\t" \
  "mulx 24(%[a]), %%r13, %%r15\n\t"  "adcx %%r14, " T3 "\n\t" "adox %%r13, " T3 "\n\t" \
  "movl $0, %%r13d\n\t"              "movq %%r13, " T4 "\n\t"                    \
  "adcx %%r15, " T4 "\n\t"           "adox %%rax, " T4 "\n\t"                    \
  "movq " T0 ", %%rdx\n\t"           "imulq %[ninv], %%rdx\n\t"                  \
  "xorl %%eax, %%eax\n\t"                                                        \
  "mulx 0(%[q]), %%r13, %%r14\n\t"   "adox " T0 ", %%r13\n\t"                    \
  "mulx 8(%[q]), %%r13, %%r15\n\t"   "adcx %%r14, " T1 "\n\t" "adox %%r13, " T1 "\n\t" \
  "mulx 16(%[q]), %%r13, %%r14\n\t"  "adcx %%r15, " T2 "\n\t" "adox %%r13, " T2 "\n\t" \
  "mulx 24(%[q]), %%r13, %%r15\n\t"  "adcx %%r14, " T3 "\n\t" "adox %%r13, " T3 "\n\t" \
  "adcx %%r15, " T4 "\n\t"           "adox %%rax, " T4 "\n\t"
static inline fr fr_mul(fr a, fr b) {
  fr r;
  __asm__ volatile(
      "xorl %%r8d, %%r8d\n\t" "xorl %%r9d, %%r9d\n\t" "xorl %%r10d, %%r10d\n\t" "xorl %%r11d, %%r11d\n\t"
      ORC_ROW("%%r8", "%%r9", "%%r10", "%%r11", "%%r12", 0)
      ORC_ROW("%%r9", "%%r10", "%%r11", "%%r12", "%%r8", 8)
      ORC_ROW("%%r10", "%%r11", "%%r12", "%%r8", "%%r9", 16)
      ORC_ROW("%%r11", "%%r12", "%%r8", "%%r9", "%%r10", 24)
      "movq %%r12, 0(%[r])\n\t" "movq %%r8, 8(%[r])\n\t" "movq %%r9, 16(%[r])\n\t" "movq %%r10, 24(%[r])\n\t"
      :
      : [a] "r"(a.l), [b] "r"(b.l), [q] "r"(FR_MOD), [ninv] "m"(FR_NINV), [r] "r"(r.l)
      : "rax", "rdx", "r8", "r9", "r10", "r11", "r12", "r13", "r14", "r15", "cc", "memory");
  fr_reduce_once(r.l);
  return r;
}
#define ORC_HAVE_ASM 1
#else
static inline fr fr_mul(fr a, fr b) { return fr_mul_c(a, b); }
#define ORC_HAVE_ASM 0
#endif

/* Montgomery square a*a*2^-256 mod r: the 6 cross products once, doubled, plus the 4 squares (10 multiplications instead
 * of 16), then a separate 4-row Montgomery reduction of the 512-bit product -- what constantine and every tuned field
 * library do for x^2; two thirds of the multiplications of this path are squarings (x^2 and x^4 of every S-box).
 * a < r, so a^2 + (sum of m_i r 2^(64 i)) < 2^508 + 2^510 fits the eight limbs. */
static inline fr fr_sqr(fr a) {
  uint64_t t[8], c;
  u128 p;
  p = (u128)a.l[0] * a.l[1];                 t[1] = (uint64_t)p; c = (uint64_t)(p >> 64);
  p = (u128)a.l[0] * a.l[2] + c;             t[2] = (uint64_t)p; c = (uint64_t)(p >> 64);
  p = (u128)a.l[0] * a.l[3] + c;             t[3] = (uint64_t)p; t[4] = (uint64_t)(p >> 64);
  p = (u128)a.l[1] * a.l[2] + t[3];          t[3] = (uint64_t)p; c = (uint64_t)(p >> 64);
  p = (u128)a.l[1] * a.l[3] + t[4] + c;      t[4] = (uint64_t)p; t[5] = (uint64_t)(p >> 64);
  p = (u128)a.l[2] * a.l[3] + t[5];          t[5] = (uint64_t)p; t[6] = (uint64_t)(p >> 64);
  t[7] = t[6] >> 63;
  t[6] = (t[6] << 1) | (t[5] >> 63);
  t[5] = (t[5] << 1) | (t[4] >> 63);
  t[4] = (t[4] << 1) | (t[3] >> 63);
  t[3] = (t[3] << 1) | (t[2] >> 63);
  t[2] = (t[2] << 1) | (t[1] >> 63);
  t[1] = t[1] << 1;
  p = (u128)a.l[0] * a.l[0];                 t[0] = (uint64_t)p; c = (uint64_t)(p >> 64);
  p = (u128)t[1] + c;                        t[1] = (uint64_t)p; c = (uint64_t)(p >> 64);
  p = (u128)a.l[1] * a.l[1] + t[2] + c;      t[2] = (uint64_t)p; c = (uint64_t)(p >> 64);
  p = (u128)t[3] + c;                        t[3] = (uint64_t)p; c = (uint64_t)(p >> 64);
  p = (u128)a.l[2] * a.l[2] + t[4] + c;      t[4] = (uint64_t)p; c = (uint64_t)(p >> 64);
  p = (u128)t[5] + c;                        t[5] = (uint64_t)p; c = (uint64_t)(p >> 64);
  p = (u128)a.l[3] * a.l[3] + t[6] + c;      t[6] = (uint64_t)p; c = (uint64_t)(p >> 64);
  t[7] += c;
  uint64_t up = 0;                             /* carry into the limb above the current row */
  for (int i = 0; i < 4; ++i) {              /* row i clears limb i */
    const uint64_t m = t[i] * FR_NINV;
    p = (u128)m * FR_MOD[0] + t[i];          c = (uint64_t)(p >> 64);
    p = (u128)m * FR_MOD[1] + t[i + 1] + c;  t[i + 1] = (uint64_t)p; c = (uint64_t)(p >> 64);
    p = (u128)m * FR_MOD[2] + t[i + 2] + c;  t[i + 2] = (uint64_t)p; c = (uint64_t)(p >> 64);
    p = (u128)m * FR_MOD[3] + t[i + 3] + c;  t[i + 3] = (uint64_t)p; c = (uint64_t)(p >> 64);
    p = (u128)t[i + 4] + c + up;             t[i + 4] = (uint64_t)p; up = (uint64_t)(p >> 64);
  }
  /* up == 0 here: a^2 + sum m_i r 2^(64 i) < 2^512 */
  fr r = {{t[4], t[5], t[6], t[7]}};
  fr_reduce_once(r.l);
  return r;
}

static inline fr fr_from_std(const uint64_t a[4]) { fr x = {{a[0], a[1], a[2], a[3]}}; return fr_mul(x, FR_R2); }
static inline fr fr_to_std(fr a) { fr one = {{1, 0, 0, 0}}; return fr_mul(a, one); }
static inline fr fr_from_u64(uint64_t v) { uint64_t a[4] = {v, 0, 0, 0}; return fr_from_std(a); }

static inline void load_le(const uint8_t *p, uint64_t a[4]) { memcpy(a, p, 32); }          /* little-endian host */
static inline void store_le(uint8_t *p, const uint64_t a[4]) { memcpy(p, a, 32); }

static inline fr fr_from_bytes(const uint8_t p[32]) {   /* value reduced mod r if needed (inputs are < 2^256) */
  uint64_t a[4]; load_le(p, a);
  while (fr_geq_mod(a)) fr_sub_mod_inplace(a);
  return fr_from_std(a);
}
static inline void fr_to_bytes(fr x, uint8_t p[32]) { fr s = fr_to_std(x); store_le(p, s.l); }

/* ---------------------------------------------------------------------------------------------------------- */
/* round constants, converted to Montgomery form once */

static fr RC_EXT_M[8][3], RC_INT_M[56], FR_ZERO_M, FR_ONE_M, IV1_M, IV2_M, KEY_M[4];
static pthread_once_t g_once = PTHREAD_ONCE_INIT;

static void init_tables(void) {
  for (int r = 0; r < 8; ++r) for (int j = 0; j < 3; ++j) RC_EXT_M[r][j] = fr_from_std(P2_RC_EXT[r][j]);
  for (int r = 0; r < 56; ++r) RC_INT_M[r] = fr_from_std(P2_RC_INT[r]);
  memset(&FR_ZERO_M, 0, sizeof FR_ZERO_M);
  FR_ONE_M = fr_from_u64(1);
  /* capacity IV = 2^64 + 256*t + rate            reference/haskell/src/Poseidon2/Sponge.hs:17,34 */
  uint64_t iv1[4] = {0x0301, 1, 0, 0}, iv2[4] = {0x0302, 1, 0, 0};
  IV1_M = fr_from_std(iv1);
  IV2_M = fr_from_std(iv2);
  for (uint64_t k = 0; k < 4; ++k) KEY_M[k] = fr_from_u64(k);
}
static inline void ensure_init(void) { pthread_once(&g_once, init_tables); }

/* ---------------------------------------------------------------------------------------------------------- */
/* Poseidon2                                       reference/haskell/src/Poseidon2/Permutation.hs:14-45 */

static int g_use_sqr = 1;   /* bench.py's CPU legs time both forms on the host they run on and keep the faster one */
void orc_set_use_sqr(int on) { g_use_sqr = on; }

static inline fr sbox(fr x) {                       /* Permutation.hs:14-17 */
  if (g_use_sqr) {
    fr x2 = fr_sqr(x), x4 = fr_sqr(x2);
    return fr_mul(x4, x);
  }
  fr x2 = fr_mul(x, x), x4 = fr_mul(x2, x2);
  return fr_mul(x4, x);
}

static inline void linear_layer(fr s[3]) {          /* Permutation.hs:35-36; also the external-round mix */
  fr t = fr_add(fr_add(s[0], s[1]), s[2]);
  s[0] = fr_add(s[0], t); s[1] = fr_add(s[1], t); s[2] = fr_add(s[2], t);
}

static inline void external_round(const fr c[3], fr s[3]) {   /* Permutation.hs:28-33 */
  for (int j = 0; j < 3; ++j) s[j] = sbox(fr_add(s[j], c[j]));
  linear_layer(s);
}

static inline void internal_round(fr c, fr s[3]) {  /* Permutation.hs:19-26: [[2,1,1],[1,2,1],[1,1,3]] */
  fr x = sbox(fr_add(s[0], c));
  fr t = fr_add(fr_add(x, s[1]), s[2]);
  s[0] = fr_add(x, t);
  s[1] = fr_add(s[1], t);
  s[2] = fr_add(fr_dbl(s[2]), t);
}

static void permute(fr s[3]) {                      /* Permutation.hs:40-45 */
  linear_layer(s);
  for (int r = 0; r < 4; ++r) external_round(RC_EXT_M[r], s);
  for (int r = 0; r < 56; ++r) internal_round(RC_INT_M[r], s);
  for (int r = 4; r < 8; ++r) external_round(RC_EXT_M[r], s);
}

void orc_permutation(const uint8_t in[96], uint8_t out[96]) {
  ensure_init();
  fr s[3];
  for (int j = 0; j < 3; ++j) s[j] = fr_from_bytes(in + 32 * j);
  permute(s);
  for (int j = 0; j < 3; ++j) fr_to_bytes(s[j], out + 32 * j);
}

/* unit-test hooks */
int orc_have_asm(void) { return ORC_HAVE_ASM; }
void orc_fr_mul_check(const uint8_t a[32], const uint8_t b[32], uint8_t out_fast[32], uint8_t out_c[32]) {
  fr x = fr_from_bytes(a), y = fr_from_bytes(b);
  fr_to_bytes(fr_mul(x, y), out_fast);
  fr_to_bytes(fr_mul_c(x, y), out_c);
}
/* a^2 through the dedicated squaring and through the general product (both canonical) */
void orc_fr_sqr_check(const uint8_t a[32], uint8_t out_sqr[32], uint8_t out_mul[32]) {
  fr x = fr_from_bytes(a);
  fr_to_bytes(fr_sqr(x), out_sqr);
  fr_to_bytes(fr_mul(x, x), out_mul);
}

void orc_permutation_batch(const uint8_t *in, uint8_t *out, size_t n) {
  for (size_t i = 0; i < n; ++i) orc_permutation(in + 96 * i, out + 96 * i);
}

/* sponge over Montgomery-form elements supplied by a callback-free pull: elements are materialised by the caller */
static fr sponge2_m(const fr *xs, size_t n) {        /* Sponge.hs:30-43 */
  fr s[3] = {FR_ZERO_M, FR_ZERO_M, IV2_M};
  size_t i = 0;
  for (; i + 2 <= n; i += 2) { s[0] = fr_add(s[0], xs[i]); s[1] = fr_add(s[1], xs[i + 1]); permute(s); }
  if (i < n) { s[0] = fr_add(s[0], xs[i]); s[1] = fr_add(s[1], FR_ONE_M); }      /* pad [x]  = [x,1] */
  else       { s[0] = fr_add(s[0], FR_ONE_M); }                                  /* pad []   = [1,0] */
  permute(s);
  return s[0];
}

static fr sponge1_m(const fr *xs, size_t n) {        /* Sponge.hs:13-25 */
  fr s[3] = {FR_ZERO_M, FR_ZERO_M, IV1_M};
  for (size_t i = 0; i < n; ++i) { s[0] = fr_add(s[0], xs[i]); permute(s); }
  s[0] = fr_add(s[0], FR_ONE_M);
  permute(s);
  return s[0];
}

int orc_sponge(const uint8_t *elems, size_t n, int rate, uint8_t out[32]) {
  ensure_init();
  if (rate != 1 && rate != 2) return -1;
  fr *xs = (fr *)malloc((n ? n : 1) * sizeof(fr));
  if (!xs) return -2;
  for (size_t i = 0; i < n; ++i) {
    uint64_t a[4]; load_le(elems + 32 * i, a);
    if (fr_geq_mod(a)) { free(xs); return -1; }
    xs[i] = fr_from_std(a);
  }
  fr h = rate == 1 ? sponge1_m(xs, n) : sponge2_m(xs, n);
  free(xs);
  fr_to_bytes(h, out);
  return 0;
}

/* ---------------------------------------------------------------------------------------------------------- */
/* bytes -> elements                               reference/haskell/src/Slot.hs:243-270, README.md:86-99 */

static inline void chunk_std(const uint8_t *data, size_t len, size_t k, uint64_t a[4]) {
  /* k-th 31-byte chunk of (data ++ 0x01 ++ 0x00...), little-endian integer < 2^248 */
  uint8_t buf[32];
  memset(buf, 0, 32);
  size_t off = 31 * k;
  if (off + 31 <= len) memcpy(buf, data + off, 31);
  else { size_t m = len - off; memcpy(buf, data + off, m); buf[m] = 0x01; }
  load_le(buf, a);
}

size_t orc_bytes_to_elements(const uint8_t *data, size_t len, uint8_t *out) {
  size_t n = len / 31 + 1;
  for (size_t k = 0; k < n; ++k) { uint64_t a[4]; chunk_std(data, len, k, a); store_le(out + 32 * k, a); }
  return n;
}

static fr hash_bytes_m(const uint8_t *data, size_t len) {      /* Slot.hs:227-228 */
  size_t n = len / 31 + 1;
  fr s[3] = {FR_ZERO_M, FR_ZERO_M, IV2_M};
  size_t k = 0;
  uint64_t a[4];
  for (; k + 2 <= n; k += 2) {
    chunk_std(data, len, k, a);     s[0] = fr_add(s[0], fr_from_std(a));
    chunk_std(data, len, k + 1, a); s[1] = fr_add(s[1], fr_from_std(a));
    permute(s);
  }
  if (k < n) { chunk_std(data, len, k, a); s[0] = fr_add(s[0], fr_from_std(a)); s[1] = fr_add(s[1], FR_ONE_M); }
  else       { s[0] = fr_add(s[0], FR_ONE_M); }
  permute(s);
  return s[0];
}

void orc_hash_bytes(const uint8_t *data, size_t len, uint8_t out[32]) {
  ensure_init();
  fr_to_bytes(hash_bytes_m(data, len), out);
}

/* ---------------------------------------------------------------------------------------------------------- */
/* keyed compression + Merkle                      reference/haskell/src/Poseidon2/Merkle.hs:69-83,156-208
 *                                                 reference/nim/proof_input/src/merkle/bn254.nim:29-63 */

static inline fr compress_m(fr x, fr y, unsigned key) {        /* Merkle.hs:202-203 */
  fr s[3] = {x, y, KEY_M[key & 3]};
  permute(s);
  return s[0];
}

void orc_compress(const uint8_t x[32], const uint8_t y[32], uint32_t key, uint8_t out[32]) {
  ensure_init();
  fr k = key < 4 ? KEY_M[key] : fr_from_u64(key);
  fr s[3] = {fr_from_bytes(x), fr_from_bytes(y), k};
  permute(s);
  fr_to_bytes(s[0], out);
}

size_t orc_merkle_total_nodes(size_t n, int bottom_layer) {
  size_t total = 0;
  int bottom = bottom_layer;
  if (n == 0) return 0;
  for (;;) {
    total += n;
    if (!bottom && n == 1) return total;
    n = (n + 1) / 2;
    bottom = 0;
  }
}

/* one level: ys[i] = compress(xs[2i], xs[2i+1], key), odd tail compress(x, 0, key+2)   merkle/bn254.nim:38-53 */
static void merkle_level_m(const fr *xs, size_t m, int bottom, fr *ys) {
  unsigned kb = bottom ? 1u : 0u;
  for (size_t i = 0; i < m / 2; ++i) ys[i] = compress_m(xs[2 * i], xs[2 * i + 1], kb);
  if (m & 1) ys[m / 2] = compress_m(xs[m - 1], FR_ZERO_M, kb + 2);
}

static int merkle_layers_m(const fr *leaves, size_t n, int bottom, fr *out) {
  /* out holds all layers concatenated; returns number of layers */
  if (n == 0) return -1;
  memcpy(out, leaves, n * sizeof(fr));
  const fr *cur = out;
  fr *next = out + n;
  int layers = 1;
  size_t m = n;
  for (;;) {
    if (!bottom && m == 1) return layers;
    merkle_level_m(cur, m, bottom, next);
    cur = next; m = (m + 1) / 2; next += m; bottom = 0; ++layers;
  }
}

int orc_merkle_layers(const uint8_t *leaves, size_t n, int bottom_layer, uint8_t *layers_out) {
  ensure_init();
  if (n == 0) return -1;
  size_t total = orc_merkle_total_nodes(n, bottom_layer);
  fr *buf = (fr *)malloc(total * sizeof(fr));
  fr *lv = (fr *)malloc(n * sizeof(fr));
  if (!buf || !lv) { free(buf); free(lv); return -2; }
  for (size_t i = 0; i < n; ++i) lv[i] = fr_from_bytes(leaves + 32 * i);
  int layers = merkle_layers_m(lv, n, bottom_layer, buf);
  for (size_t i = 0; i < total; ++i) fr_to_bytes(buf[i], layers_out + 32 * i);
  free(buf); free(lv);
  return layers;
}

static fr merkle_root_m(fr *work, size_t n) {       /* in-place, destroys work; Merkle.hs:180-189 */
  int bottom = 1;
  for (;;) {
    if (!bottom && n == 1) return work[0];
    merkle_level_m(work, n, bottom, work);          /* ys[i] only reads xs[2i], xs[2i+1] >= i: safe in place */
    n = (n + 1) / 2; bottom = 0;
  }
}

int orc_merkle_root(const uint8_t *leaves, size_t n, uint8_t out[32]) {
  ensure_init();
  if (n == 0) return -1;
  fr *w = (fr *)malloc(n * sizeof(fr));
  if (!w) return -2;
  for (size_t i = 0; i < n; ++i) w[i] = fr_from_bytes(leaves + 32 * i);
  fr_to_bytes(merkle_root_m(w, n), out);
  free(w);
  return 0;
}

void orc_reconstruct_root(const uint8_t leaf[32], uint64_t j, uint64_t m, const uint8_t *path, size_t path_len,
                          uint8_t out[32]) {     /* merkle.nim:51-74 */
  ensure_init();
  fr h = fr_from_bytes(leaf);
  unsigned bottom = 1;
  for (size_t i = 0; i < path_len; ++i) {
    fr p = fr_from_bytes(path + 32 * i);
    if (j & 1)           h = compress_m(p, h, bottom);
    else if (j == m - 1) h = compress_m(h, p, bottom + 2);
    else                 h = compress_m(h, p, bottom);
    bottom = 0; j >>= 1; m = (m + 1) >> 1;
  }
  fr_to_bytes(h, out);
}

/* ---------------------------------------------------------------------------------------------------------- */
/* fake data                                       reference/nim/proof_input/src/slot.nim:23-32 */

void orc_gen_fake_cell(uint64_t seed, uint64_t idx, size_t cell_size, uint8_t *out) {
  uint64_t seed1 = seed + 0xdeadcafeull, seed2 = idx + 0x98765432ull, s = 1;
  for (size_t i = 0; i < cell_size; ++i) {
    s = s * (s + seed1) * (s + seed2) + s * (s ^ 0x5a5a5a5aull) + seed1 * s + (seed2 + 17);
    s %= 1698428844001831ull;
    out[i] = (uint8_t)s;
  }
}

/* ---------------------------------------------------------------------------------------------------------- */
/* slot commitment                                 reference/nim/proof_input/src/gen_input/bn254.nim:21-30,
 *                                                 reference/nim/proof_input/src/blocks/bn254.nim:60-67 */

typedef struct {
  const uint8_t *data;       /* NULL => fake data */
  uint64_t seed;
  size_t cell_size, cells_per_block, b0, b1;
  fr *cell_hashes;           /* may be NULL */
  fr *block_hashes;
} job_t;

static void *block_worker(void *arg) {
  job_t *j = (job_t *)arg;
  size_t k = j->cells_per_block, cs = j->cell_size;
  fr *leaves = (fr *)malloc(k * sizeof(fr));
  uint8_t *cell = j->data ? NULL : (uint8_t *)malloc(cs);
  for (size_t b = j->b0; b < j->b1; ++b) {
    for (size_t c = 0; c < k; ++c) {
      size_t ci = b * k + c;
      const uint8_t *p;
      if (j->data) p = j->data + ci * cs;
      else { orc_gen_fake_cell(j->seed, ci, cs, cell); p = cell; }
      leaves[c] = hash_bytes_m(p, cs);                         /* hashCell: blocks/bn254.nim:23-29 */
      if (j->cell_hashes) j->cell_hashes[ci] = leaves[c];
    }
    j->block_hashes[b] = merkle_root_m(leaves, k);             /* networkBlockTree root: blocks/bn254.nim:60-64 */
  }
  free(leaves); free(cell);
  return NULL;
}

static int commit_common(const uint8_t *data, uint64_t seed, size_t n_cells, size_t cell_size, size_t block_size,
                         int n_threads, uint8_t *cell_hashes_out, uint8_t *block_hashes_out, uint8_t root_out[32]) {
  ensure_init();
  if (cell_size == 0 || block_size == 0 || block_size % cell_size) return -1;
  size_t k = block_size / cell_size;
  if (n_cells == 0 || n_cells % k) return -1;
  size_t nb = n_cells / k;
  if (n_threads < 1) n_threads = 1;
  if ((size_t)n_threads > nb) n_threads = (int)nb;
  fr *bh = (fr *)malloc(nb * sizeof(fr));
  fr *ch = cell_hashes_out ? (fr *)malloc(n_cells * sizeof(fr)) : NULL;
  pthread_t *th = (pthread_t *)malloc(n_threads * sizeof(pthread_t));
  job_t *jobs = (job_t *)malloc(n_threads * sizeof(job_t));
  for (int t = 0; t < n_threads; ++t) {
    job_t jb = {data, seed, cell_size, k, nb * t / n_threads, nb * (t + 1) / n_threads, ch, bh};
    jobs[t] = jb;
    if (t + 1 < n_threads) pthread_create(&th[t], NULL, block_worker, &jobs[t]);
  }
  block_worker(&jobs[n_threads - 1]);
  for (int t = 0; t + 1 < n_threads; ++t) pthread_join(th[t], NULL);
  if (ch) for (size_t i = 0; i < n_cells; ++i) fr_to_bytes(ch[i], cell_hashes_out + 32 * i);
  if (block_hashes_out) for (size_t i = 0; i < nb; ++i) fr_to_bytes(bh[i], block_hashes_out + 32 * i);
  fr_to_bytes(merkle_root_m(bh, nb), root_out);                 /* big tree: gen_input/bn254.nim:28-29 */
  free(bh); free(ch); free(th); free(jobs);
  return 0;
}

int orc_commit_slot(const uint8_t *data, size_t n_bytes, size_t cell_size, size_t block_size, int n_threads,
                    uint8_t *cell_hashes_out, uint8_t *block_hashes_out, uint8_t root_out[32]) {
  if (!data || block_size == 0 || n_bytes == 0 || n_bytes % block_size) return -1;
  return commit_common(data, 0, n_bytes / cell_size, cell_size, block_size, n_threads, cell_hashes_out,
                       block_hashes_out, root_out);
}

int orc_commit_fake_slot(uint64_t seed, size_t n_cells, size_t cell_size, size_t block_size, int n_threads,
                         uint8_t *cell_hashes_out, uint8_t *block_hashes_out, uint8_t root_out[32]) {
  return commit_common(NULL, seed, n_cells, cell_size, block_size, n_threads, cell_hashes_out, block_hashes_out,
                       root_out);
}

/* ---------------------------------------------------------------------------------------------------------- */
/* sampling                                        reference/nim/proof_input/src/sample/bn254.nim:16-24 */

int64_t orc_cell_index(const uint8_t entropy[32], const uint8_t slot_root[32], uint64_t n_cells, uint64_t counter) {
  ensure_init();
  if (n_cells == 0 || (n_cells & (n_cells - 1))) return -1;
  fr xs[3] = {fr_from_bytes(entropy), fr_from_bytes(slot_root), fr_from_u64(counter)};
  fr h = fr_to_std(sponge2_m(xs, 3));
  return (int64_t)(h.l[0] & (n_cells - 1));                     /* extractLowBits: types/bn254.nim:47-59 */
}

/* ---------------------------------------------------------------------------------------------------------- */
/* decimal output                                  reference/nim/proof_input/src/types/bn254.nim:29-33 */

int orc_to_decimal(const uint8_t x[32], char *buf) {
  uint32_t w[8];
  memcpy(w, x, 32);
  char tmp[80];
  int n = 0;
  for (;;) {
    int nz = 0;
    uint64_t rem = 0;
    for (int i = 7; i >= 0; --i) {
      uint64_t cur = (rem << 32) | w[i];
      w[i] = (uint32_t)(cur / 1000000000u);
      rem = cur % 1000000000u;
      nz |= w[i] != 0;
    }
    for (int d = 0; d < 9; ++d) { tmp[n++] = (char)('0' + rem % 10); rem /= 10; }
    if (!nz) break;
  }
  while (n > 1 && tmp[n - 1] == '0') --n;
  for (int i = 0; i < n; ++i) buf[i] = tmp[n - 1 - i];
  buf[n] = 0;
  return n;
}
