/* codex_oracle.h -- CPU restatement (plain C) of the slot-commitment hot path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load this library.  The product (libcodexcommit.so) never links or calls it.
 *
 * Field elements cross this API as 32-byte little-endian canonical integers (value < r), the same convention
 * as include/codex_commit.h, so outputs can be compared with memcmp.
 *
 * Parity pinning: permutation pinned by the reference's stored KAT (reference/haskell/src/Poseidon2/Example.hs:13-22);
 * everything above it is pinned structurally only (the reference stores no expected values) -- see DESIGN.md.
 * Citations are file:line under /root/reference.
 */
#ifndef CODEX_ORACLE_H
#define CODEX_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Poseidon2 t=3 permutation                       reference/haskell/src/Poseidon2/Permutation.hs:40-45 */
void orc_permutation(const uint8_t in[96], uint8_t out[96]);
void orc_permutation_batch(const uint8_t *in, uint8_t *out, size_t n);
/* 1 if the MULX/ADCX/ADOX product was compiled in; a*b mod r (standard form in and out) through it and through the portable C product */
int orc_have_asm(void);
void orc_fr_mul_check(const uint8_t a[32], const uint8_t b[32], uint8_t out_fast[32], uint8_t out_c[32]);
/* x^2 and x^4 of the S-box through the dedicated squaring (1, default) or the general product (0): same results */
void orc_set_use_sqr(int on);
/* unit-test hook: a^2 mod r (standard form in and out) through fr_sqr and through fr_mul(a, a) */
void orc_fr_sqr_check(const uint8_t a[32], uint8_t out_sqr[32], uint8_t out_mul[32]);

/* rate-1 / rate-2 sponge over n field elements    reference/haskell/src/Poseidon2/Sponge.hs:13-43
 * returns 0, or -1 if rate is not 1 or 2 / an element is >= r */
int orc_sponge(const uint8_t *elems, size_t n, int rate, uint8_t out[32]);

/* bytes -> 31-byte LE chunks with 10* padding     reference/haskell/src/Slot.hs:243-270
 * writes floor(len/31)+1 elements, returns that count */
size_t orc_bytes_to_elements(const uint8_t *data, size_t len, uint8_t *out);

/* sponge2(bytes_to_elements(data))                reference/haskell/src/Slot.hs:227-228,
 *                                                 reference/nim/proof_input/src/blocks/bn254.nim:27 */
void orc_hash_bytes(const uint8_t *data, size_t len, uint8_t out[32]);

/* perm(x, y, key)[0]                              reference/haskell/src/Poseidon2/Merkle.hs:202-203 */
void orc_compress(const uint8_t x[32], const uint8_t y[32], uint32_t key, uint8_t out[32]);

/* number of nodes in all layers of a tree over n leaves (bottom_layer as in merkleTreeWorker) */
size_t orc_merkle_total_nodes(size_t n, int bottom_layer);

/* all layers, bottom first, concatenated          reference/nim/proof_input/src/merkle/bn254.nim:29-63
 * layers_out must hold orc_merkle_total_nodes(n, bottom_layer) elements; returns number of layers, <0 on error */
int orc_merkle_layers(const uint8_t *leaves, size_t n, int bottom_layer, uint8_t *layers_out);

/* root only                                       reference/haskell/src/Poseidon2/Merkle.hs:180-189 */
int orc_merkle_root(const uint8_t *leaves, size_t n, uint8_t out[32]);

/* verifier walk                                   reference/nim/proof_input/src/merkle.nim:51-74 */
void orc_reconstruct_root(const uint8_t leaf[32], uint64_t leaf_index, uint64_t n_leaves,
                          const uint8_t *path, size_t path_len, uint8_t out[32]);

/* fake data                                       reference/nim/proof_input/src/slot.nim:23-32 */
void orc_gen_fake_cell(uint64_t seed, uint64_t idx, size_t cell_size, uint8_t *out);

/* Whole-slot commitment over resident bytes: cell hashes -> block trees -> slot tree.
 *                                                 reference/nim/proof_input/src/gen_input/bn254.nim:21-30
 * n_bytes must be a multiple of block_size; block_size a multiple of cell_size.
 * cell_hashes_out (n_cells*32) and block_hashes_out (n_blocks*32) may be NULL.
 * n_threads >= 1: blocks are split over pthreads (the reference itself is single-threaded).
 * returns 0 / negative on bad sizes. */
int orc_commit_slot(const uint8_t *data, size_t n_bytes, size_t cell_size, size_t block_size, int n_threads,
                    uint8_t *cell_hashes_out, uint8_t *block_hashes_out, uint8_t root_out[32]);

/* same, over the reference's fake data (never materialises the slot) */
int orc_commit_fake_slot(uint64_t seed, size_t n_cells, size_t cell_size, size_t block_size, int n_threads,
                         uint8_t *cell_hashes_out, uint8_t *block_hashes_out, uint8_t root_out[32]);

/* low log2(n_cells) bits of sponge2([entropy, slot_root, counter])
 *                                                 reference/nim/proof_input/src/sample/bn254.nim:16-24
 * returns -1 if n_cells is not a power of two */
int64_t orc_cell_index(const uint8_t entropy[32], const uint8_t slot_root[32], uint64_t n_cells, uint64_t counter);

/* decimal string without leading zeros ("0" for zero); buf >= 80 bytes; returns length
 *                                                 reference/nim/proof_input/src/types/bn254.nim:29-33 */
int orc_to_decimal(const uint8_t x[32], char *buf);

#ifdef __cplusplus
}
#endif
#endif
