// UNIT-TEST ONLY: builds the device headers with g++ against the C emulation of the PTX primitives
// (fr_rows_host.h) and exports a few entry points so pytest can compare the limb-level logic (lazy reduction
// bounds, Poseidon2 schedule, 31-byte chunk reader, sponge padding) with the oracle on a machine without a GPU.
#define CDX_HOST_EMUL 1
#include "../../codex-storage-proofs-circuits_b200/csrc/poseidon2.cuh"
#include <string.h>

using namespace cdx;

static Fr load(const uint8_t* p) { Fr a; memcpy(a.l, p, 32); return a; }
static void store(uint8_t* p, const Fr& a) { memcpy(p, a.l, 32); }

extern "C" {
// raw Montgomery product of two 256-bit values (no conversion): out = a*b*2^-256 mod r, lazily reduced
void emul_mont_mul_raw(const uint8_t* a, const uint8_t* b, uint8_t* out) { store(out, mont_mul(load(a), load(b))); }
void emul_mont_sqr_raw(const uint8_t* a, uint8_t* out) { store(out, mont_sqr(load(a))); }
void emul_to_mont(const uint8_t* a, uint8_t* out) { store(out, to_mont(load(a))); }
void emul_from_mont(const uint8_t* a, uint8_t* out) { store(out, from_mont(load(a))); }
void emul_add_mod(const uint8_t* a, const uint8_t* b, uint8_t* out) { store(out, add_mod(load(a), load(b))); }
// table-driven reduction and the mixes built on it, on RAW limb values (no conversion): the tests drive them at the
// edges of their documented input ranges; the emulated primitives trap on any violated bound
void emul_reduce_tab(const uint8_t* v, uint32_t carry, uint8_t* out) { store(out, reduce_tab(load(v), carry)); }
void emul_add_reduce(const uint8_t* a, const uint8_t* b, uint8_t* out) { store(out, add_reduce(load(a), load(b))); }
void emul_mix_internal_raw(uint8_t* x, uint8_t* y, uint8_t* z) {
  Fr a = load(x), b = load(y), c = load(z);
  mix_internal(a, b, c);
  store(x, a); store(y, b); store(z, c);
}
void emul_mix_external_raw(uint8_t* x, uint8_t* y, uint8_t* z) {
  Fr a = load(x), b = load(y), c = load(z);
  mix_external(a, b, c);
  store(x, a); store(y, b); store(z, c);
}
void emul_sbox_raw(const uint8_t* x, uint8_t* out) { store(out, sbox(load(x))); }
void emul_permutation(const uint8_t* in, uint8_t* out) {
  Fr x = to_mont(load(in)), y = to_mont(load(in + 32)), z = to_mont(load(in + 64));
  permute(x, y, z);
  store(out, from_mont(x)); store(out + 32, from_mont(y)); store(out + 64, from_mont(z));
}
void emul_hash_bytes(const uint8_t* data, uint32_t len, uint8_t* out) {
  AnyBytes ld{data, len};
  store(out, from_mont(sponge2_bytes(ld, len)));
}
// aligned-cell path: data must be 4-byte aligned and len % 4 == 0
void emul_hash_cell_aligned(const uint8_t* data, uint32_t len, uint8_t* out) {
  AlignedWords ld{(const uint32_t*)data, len / 4};
  store(out, from_mont(sponge2_bytes(ld, len)));
}
void emul_sponge(const uint8_t* elems, uint32_t n, int rate, uint8_t* out) {
  auto get = [&](uint32_t i) { return load(elems + 32 * i); };
  store(out, from_mont(sponge_elems(get, n, rate)));
}
void emul_compress(const uint8_t* x, const uint8_t* y, uint32_t key, uint8_t* out) {
  store(out, from_mont(compress_keyed(to_mont(load(x)), to_mont(load(y)), key)));
}
}
