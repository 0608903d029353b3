/* codex_commit.h -- C ABI of libcodexcommit.so, the B200 (sm_100a) slot-commitment backend.
 *
 * The reference (codex-storage-proofs-circuits, reference/nim/proof_input) has no FFI seam of its own: its
 * arithmetic comes from the Nim packages `poseidon2` and `constantine`, called from a handful of places.  The
 * entry points below are exactly what a Nim `importc` binding would need to replace those call sites and the
 * loops around them; each one cites the reference interface it stands in for (file:line under /root/reference,
 * `nim/` = reference/nim/proof_input/src/).  INTEGRATION.md shows the Nim-side stub.
 *
 * Conventions
 *   - A field element (Nim `F`, `Hash`, `Root`, `Entropy`: nim/types/bn254.nim:20-23) crosses the ABI as 32 bytes,
 *     little-endian, canonical (value < r).  Inputs >= r are taken mod r.  Outputs are always canonical.
 *   - The caller owns every buffer.  `_host` entry points take host pointers and perform the transfers
 *     themselves; `_dev` entry points take CUDA device pointers and a cudaStream_t (as void*) and are
 *     asynchronous with respect to the host.
 *   - Every function returns 0 (CDX_OK) or a negative cdx_status; nothing aborts (the reference asserts:
 *     nim/blocks/bn254.nim:26,34, nim/sample/bn254.nim:20).  cdx_last_error(ctx) gives the message.
 *   - One cdx_ctx per (host thread, GPU).  Handles are not shared between threads.
 *   - There is NO CPU implementation behind this ABI: without a CUDA device cdx_ctx_create fails with
 *     CDX_ERR_CUDA and nothing else can be called.
 */
#ifndef CODEX_COMMIT_H
#define CODEX_COMMIT_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CDX_ABI_VERSION 1
#define CDX_FELT_BYTES 32

typedef enum cdx_status {
  CDX_OK = 0,
  CDX_ERR_ARG = -1,      /* null pointer, zero count, bad rate/key */
  CDX_ERR_SIZE = -2,     /* sizes not divisible as the reference asserts (nim/types.nim:120-123, nim/blocks/bn254.nim:26,34) */
  CDX_ERR_NOT_POW2 = -3, /* numberOfCells must be a power of two (nim/sample/bn254.nim:19-20) */
  CDX_ERR_RANGE = -4,    /* index out of range (nim/merkle.nim:27) or depth too small (nim/types.nim:29) */
  CDX_ERR_CUDA = -5,     /* CUDA runtime error, or no device */
  CDX_ERR_ALLOC = -6,
  CDX_ERR_STATE = -7     /* handle not in the state the call needs (e.g. sharded slot without its top tree) */
} cdx_status;

typedef struct cdx_ctx cdx_ctx;   /* per-GPU context: device, streams, staging buffers */
typedef struct cdx_slot cdx_slot; /* a committed slot: every Merkle layer resident in HBM */

/* ---- context ------------------------------------------------------------------------------------------ */
int cdx_abi_version(void);
int cdx_ctx_create(int device, cdx_ctx** out);
void cdx_ctx_destroy(cdx_ctx* ctx);
const char* cdx_last_error(const cdx_ctx* ctx);
const char* cdx_status_string(int status);
/* number of kernels this context has launched since creation (bench.py's gpu_launches) */
uint64_t cdx_launch_count(const cdx_ctx* ctx);
/* the CUDA stream (cudaStream_t) host-pointer entry points run on; lets callers time them with events */
void* cdx_ctx_stream(const cdx_ctx* ctx);

/* ---- hash layer (what nim-poseidon2 provides to the reference) ------------------------------------------ */

/* n independent Poseidon2 permutations, 3 elements in / 3 out each.
 * Replaces: permutation, reference/haskell/src/Poseidon2/Permutation.hs:40-45 (Nim: poseidon2 `perm`). */
int cdx_permutation_batch_host(cdx_ctx* ctx, const uint8_t* in, uint8_t* out, size_t n);
int cdx_permutation_batch_dev(cdx_ctx* ctx, const void* d_in, void* d_out, size_t n, void* stream);

/* n_items sponges, each over `len` field elements (item i at elems + i*len*32), rate 1 or 2.
 * Replaces: Sponge.digest(openArray[F], rate) -- nim/sample/bn254.nim:23, testvectors.nim:26,34;
 * reference/haskell/src/Poseidon2/Sponge.hs:13-43. */
int cdx_sponge_felts_batch_host(cdx_ctx* ctx, const uint8_t* elems, size_t n_items, size_t len, int rate, uint8_t* out);

/* n_items byte strings of equal length `len` (item i at data + i*len), each hashed with the rate-2 sponge after
 * 31-byte 10* chunking.  Replaces: Sponge.digest(openArray[byte], rate=2) -- nim/blocks/bn254.nim:27 (hashCell),
 * testvectors.nim:45; reference/haskell/src/Slot.hs:222-270. */
int cdx_hash_bytes_batch_host(cdx_ctx* ctx, const uint8_t* data, size_t n_items, size_t len, uint8_t* out);

/* Cell hashes only, device to device: n_cells cells of cell_size bytes (a multiple of 4) at d_data -> n_cells field
 * elements at d_out.  This is the dominant kernel of the path (34 permutations per 2048-byte cell); bench.py times
 * it alone for the roofline.  Replaces: the hashCell loop of networkBlockTree -- nim/blocks/bn254.nim:23-29,63. */
int cdx_hash_cells_dev(cdx_ctx* ctx, const void* d_data, size_t n_cells, size_t cell_size, void* d_out, void* stream);

/* out[i] = perm(x[i], y[i], keys[i])[0], keys in 0..3.
 * Replaces: compress(x, y, key=) via compressWithkey -- nim/merkle/bn254.nim:18,50,53;
 * reference/haskell/src/Poseidon2/Merkle.hs:202-203. */
int cdx_compress_batch_host(cdx_ctx* ctx, const uint8_t* x, const uint8_t* y, const uint32_t* keys, size_t n, uint8_t* out);

/* ---- Merkle trees --------------------------------------------------------------------------------------- */

/* Total node count over all layers of a tree with n leaves (layers n, ceil(n/2), ..., 1).  With
 * bottom_layer != 0 a single leaf still gets one key-3 compression (2 layers); otherwise a single node is its
 * own root.  Mirrors merkleTreeWorker, nim/merkle/bn254.nim:29-60. */
size_t cdx_merkle_total_nodes(size_t n, int bottom_layer);
int cdx_merkle_num_layers(size_t n, int bottom_layer);

/* All layers bottom-first, concatenated, into layers_out (cdx_merkle_total_nodes(n) * 32 bytes).
 * Replaces: merkleTreeBN254 / merkleTree -- nim/merkle/bn254.nim:62-63, nim/blocks/bn254.nim:17-19;
 * reference/haskell/src/Poseidon2/Merkle.hs:69-83.  Key rule: bit0 = first level only, bit1 = odd node
 * (single child paired with 0), single leaf = one key-3 compression. */
int cdx_merkle_layers_host(cdx_ctx* ctx, const uint8_t* leaves, size_t n, int bottom_layer, uint8_t* layers_out);

/* Root only.  Replaces: Merkle.digest / merkleDigestBN254 -- nim/merkle/bn254.nim:20, testvectors.nim:56. */
int cdx_merkle_root_host(cdx_ctx* ctx, const uint8_t* leaves, size_t n, uint8_t root_out[32]);

/* ---- slot commitment ------------------------------------------------------------------------------------ */

/* Commit a whole slot: hash every cell (34 permutations per 2048-byte cell), build every block tree
 * (cells_per_block leaves each), then the slot tree over the block hashes.  All layers stay in HBM inside the
 * returned handle so that paths can be extracted later without rehashing.
 * Replaces: buildSlotTreeFull -- nim/gen_input/bn254.nim:21-30 (hashCell, networkBlockTree, merkleTree);
 * reference/haskell/src/Slot.hs:151-179.
 * Requirements (else CDX_ERR_SIZE): cell_size a multiple of 4, block_size a multiple of cell_size with a
 * power-of-two quotient, n_bytes a non-zero multiple of block_size. */
int cdx_slot_commit_host(cdx_ctx* ctx, const uint8_t* data, size_t n_bytes, size_t cell_size, size_t block_size, cdx_slot** out);
int cdx_slot_commit_dev(cdx_ctx* ctx, const void* d_data, size_t n_bytes, size_t cell_size, size_t block_size, void* stream, cdx_slot** out);

/* Commit a slot straight from a file: bytes [offset, offset + n_bytes) of `path` stream through pinned host buffers
 * (parallel pread) -> H2D -> cell sponge, all three overlapped; only the hashes stay resident.  Bytes past the end of
 * the file read as zeros, like the reference's ignored short reads.
 * Replaces: slotLoadCellData(SlotFile) + buildSlotTreeFull -- nim/slot.nim:57-68 (one open/seek/read per 2 KiB cell),
 * nim/dataset.nim:34-41, nim/gen_input/bn254.nim:21-30.   */
int cdx_slot_commit_file(cdx_ctx* ctx, const char* path, uint64_t offset, size_t n_bytes, size_t cell_size, size_t block_size, cdx_slot** out);

/* Same over the reference's fake data source, generated on the device (never crosses PCIe).
 * Replaces: slotLoadBlockData + genFakeCell -- nim/slot.nim:23-32,70-73, seed as in nim/dataset.nim:32. */
int cdx_slot_commit_fake(cdx_ctx* ctx, uint64_t seed, size_t n_cells, size_t cell_size, size_t block_size, cdx_slot** out);

/* Sharded commitment (one process per GPU).  This rank holds blocks [first_block, first_block + n_local_blocks)
 * of a slot of n_total_blocks; first_block must be a multiple of 2^top_level, and so must n_local_blocks unless
 * the range ends at n_total_blocks.  Levels 0..top_level of the slot tree are built for the local range; the
 * level-top_level nodes are the sub-tree roots to exchange (all-gather).  cdx_slot_set_top then installs the
 * gathered level and builds the replicated upper levels.  (The reference is single-process; the tree
 * conventions are nim/merkle/bn254.nim:29-60 with the odd-node rule applied to the GLOBAL layer width.) */
int cdx_slot_commit_range_dev(cdx_ctx* ctx, const void* d_data, size_t n_local_bytes, size_t cell_size, size_t block_size,
                              uint64_t first_block, uint64_t n_total_blocks, int top_level, void* stream, cdx_slot** out);
int cdx_slot_commit_range_host(cdx_ctx* ctx, const uint8_t* data, size_t n_local_bytes, size_t cell_size, size_t block_size,
                               uint64_t first_block, uint64_t n_total_blocks, int top_level, cdx_slot** out);
int cdx_slot_subtree_root_count(const cdx_slot* slot, uint64_t* first_node, uint64_t* n_nodes);
/* copy this rank's level-top_level nodes into a caller buffer (e.g. a torch tensor handed to NCCL all-gather) */
int cdx_slot_subtree_roots_copy_dev(const cdx_slot* slot, void* d_dst, void* stream);
/* device pointer to this rank's level-top_level nodes (n_nodes * 32 bytes), for NCCL */
const void* cdx_slot_subtree_roots_dev(const cdx_slot* slot);
int cdx_slot_set_top_dev(cdx_slot* slot, const void* d_level_nodes, uint64_t n_level_nodes, void* stream);

/* Persist / restore a commitment (a proof server answers repeated challenges from a retained
 * commitment instead of re-hashing the slot).  The image is a small header followed by every retained layer (canonical
 * 32-byte elements), about 3.1 % of the slot size; it is self-describing and checked on import.  Only whole
 * (non-sharded, top tree present) slots can be exported.  The reference keeps nothing: it rebuilds the slot tree for
 * every sample (nim/gen_input/bn254.nim:57). */
size_t cdx_slot_export_size(const cdx_slot* slot);
int cdx_slot_export(const cdx_slot* slot, uint8_t* image, size_t image_bytes);
int cdx_slot_import(cdx_ctx* ctx, const uint8_t* image, size_t image_bytes, cdx_slot** out);

void cdx_slot_free(cdx_slot* slot);

/* treeRoot of the slot tree -- nim/merkle.nim:14-17. */
int cdx_slot_root(const cdx_slot* slot, uint8_t root_out[32]);
/* shape of the retained trees */
int cdx_slot_shape(const cdx_slot* slot, uint64_t* n_cells, uint64_t* n_blocks, uint32_t* block_tree_depth, uint32_t* slot_tree_depth);
/* Copy `count` nodes starting at `first` out of a retained layer.  tree 0 = the forest of block trees
 * (level 0 = cell hashes ... level block_tree_depth = block hashes), tree 1 = the slot tree (level 0 = block
 * hashes ... level slot_tree_depth = root).  For a sharded slot, indices are global; only locally held nodes
 * (or replicated upper levels) can be read.  Mirrors MerkleTree.layers, nim/types.nim:20-22. */
int cdx_slot_read_layer(const cdx_slot* slot, int tree, uint32_t level, uint64_t first, uint64_t count, uint8_t* out);

/* Batched Merkle-path extraction for sampled cells: for each cell index, the block-tree path
 * (block_tree_depth siblings) followed by the slot-tree path (slot_tree_depth siblings), out-of-range sibling =
 * 0, zero-padded to max_depth; out is n_samples * max_depth * 32 bytes; leaf_out (may be NULL) gets the cell
 * hashes.  Replaces, per sample: merkleProof x2, mergeMerkleProofs' path concat, padMerkleProof --
 * nim/merkle.nim:21-42,86-100, nim/types.nim:27-37, nim/gen_input/bn254.nim:56-63.
 * For a sharded slot only levels held by this rank are filled, the rest are left zero (sum over ranks = path). */
int cdx_slot_cell_paths(const cdx_slot* slot, const uint64_t* cell_indices, size_t n_samples, size_t max_depth, uint8_t* out, uint8_t* leaf_out);

/* Proof-server call: answer n_challenges challenges against one retained commitment in one pass.  For challenge k
 * (entropies + 32 k) it produces what cdx_cell_indices followed by cdx_slot_cell_paths would: indices_out[k * n_samples
 * + c-1], paths_out[(k * n_samples + c-1) * max_depth * 32 ...], leaves_out (may be NULL) likewise -- the indices are
 * derived from the slot root on the device and never visit the host in between.  The slot's cell count must be a
 * power of two (nim/sample/bn254.nim:19-20).  Replaces, per challenge: cellIndices (nim/sample/bn254.nim:26-27) and
 * the per-sample proof loop of generateProofInput (nim/gen_input/bn254.nim:53-74), which rebuilds the slot tree for
 * every sample (:57). */
int cdx_slot_prove_batch(const cdx_slot* slot, const uint8_t* entropies, size_t n_challenges, size_t n_samples, size_t max_depth,
                         uint64_t* indices_out, uint8_t* paths_out, uint8_t* leaves_out);

/* Batched verifier walk: roots_out[i] = the root reconstructed from leaf i at index indices[i] in a tree of n_leaves
 * leaves along the first `depth` elements of its path (paths are n x path_stride elements, so padded paths can be
 * passed as they are).  Replaces: reconstructRoot / checkMerkleProof -- nim/merkle.nim:51-77 (also the check inside
 * mergeMerkleProofs, :88-89); same semantics as RootFromMerklePath, circuit/codex/merkle.circom:44-114. */
int cdx_reconstruct_roots_host(cdx_ctx* ctx, const uint8_t* leaves, const uint64_t* indices, uint64_t n_leaves, const uint8_t* paths,
                               size_t path_stride, size_t depth, size_t n, uint8_t* roots_out);

/* ---- sampling and data source --------------------------------------------------------------------------- */

/* indices[c-1] = low log2(n_cells) bits of sponge2([entropy, slot_root, c]), c = 1..n_samples.
 * Replaces: cellIndices / cellIndex / extractLowBits -- nim/sample/bn254.nim:16-27, nim/types/bn254.nim:47-59. */
int cdx_cell_indices(cdx_ctx* ctx, const uint8_t entropy[32], const uint8_t slot_root[32], uint64_t n_cells, size_t n_samples, uint64_t* indices);

/* The reference's fake cells [first_cell, first_cell + n_cells) into a host buffer (n_cells * cell_size bytes).
 * Replaces: slotLoadCellData(FakeData) -- nim/slot.nim:23-32,51-55. */
int cdx_fake_cells_host(cdx_ctx* ctx, uint64_t seed, uint64_t first_cell, size_t n_cells, size_t cell_size, uint8_t* out);
int cdx_fake_cells_dev(cdx_ctx* ctx, uint64_t seed, uint64_t first_cell, size_t n_cells, size_t cell_size, void* d_out, void* stream);

/* Counter-based synthetic slot bytes for benchmarks: 64-bit word i of the slot = splitmix64(seed + first_word + i)
 * (SURVEY.md section 8d config 3).  n_bytes must be a multiple of 8. */
int cdx_fill_synthetic_dev(cdx_ctx* ctx, uint64_t seed, uint64_t first_word, size_t n_bytes, void* d_out, void* stream);

/* ---- measurement ---------------------------------------------------------------------------------------- */

/* Integer-multiply roofline probe: runs a multiply stream with ILP 8 on every SM and returns the achieved rate in
 * thread-level instructions per second.  kind 0 = IMAD.WIDE.U32 without carry (each product feeds the next
 * multiplicand), 1 = IMAD.WIDE.U32.X carry chains as in the Montgomery rows, 2 = 32-bit IMAD. */
int cdx_probe_imad_rate(cdx_ctx* ctx, int kind, double* ops_per_second, double* elapsed_ms);

#ifdef __cplusplus
}
#endif
#endif /* CODEX_COMMIT_H */
