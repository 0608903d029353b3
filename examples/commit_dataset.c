/* commit_dataset.c -- the C ABI from plain C99: commit a small dataset on every visible GPU, print the dataset root, answer
 * one challenge against the sampled slot and re-check one of its Merkle paths with the batched verifier.
 *
 *   gcc -std=c99 -O2 -Iinclude examples/commit_dataset.c -Lcodex-storage-proofs-circuits_b200 -lcodexcommit \
 *       -Wl,-rpath,$PWD/codex-storage-proofs-circuits_b200 -o commit_dataset && ./commit_dataset
 *
 * This is what the Nim host does through `importc` (codex-storage-proofs-circuits_b200/nim/): the same calls, in the same
 * order, as generateProofInput (reference/nim/proof_input/src/gen_input/bn254.nim:35-74). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "codex_commit.h"

#define N_SLOTS 5
#define N_SAMPLES 5
#define MAX_DEPTH 32
#define MAX_LOG2_NSLOTS 8

static void print_felt(const char *label, const uint8_t f[32]) {
  printf("%s0x", label);
  for (int i = 31; i >= 0; --i) printf("%02x", f[i]);
  printf("\n");
}

#define CHECK(call)                                                                  \
  do {                                                                               \
    int rc_ = (call);                                                                \
    if (rc_ != CDX_OK) {                                                             \
      fprintf(stderr, "%s failed: %s\n", #call, cdx_status_string(rc_));             \
      return 1;                                                                      \
    }                                                                                \
  } while (0)

int main(void) {
  if (cdx_device_count() < 1) {
    fprintf(stderr, "no CUDA device: this backend has no CPU path\n");
    return 2;
  }
  cdx_group *group = NULL;
  CHECK(cdx_group_create(NULL, 0, &group));                    /* every visible GPU; one GPU needs no NCCL */
  const int n_gpus = cdx_group_size(group);

  /* the reference's fake data source: slot k of a dataset with seed 12345 has seed 12345 + 72 + 1001 k (dataset.nim:32);
   * 256 cells of 2048 bytes per slot, as in reference/haskell/cli/testMain.hs but with the default cell size */
  cdx_slot_desc slots[N_SLOTS];
  memset(slots, 0, sizeof slots);
  for (int k = 0; k < N_SLOTS; ++k) {
    slots[k].kind = CDX_SRC_FAKE;
    slots[k].seed = 12345u + 72u + 1001u * (unsigned)k;
    slots[k].n_bytes = 256u * 2048u;
  }
  const int64_t sampled_slot = 3;
  cdx_dataset **ds = calloc((size_t)n_gpus, sizeof *ds);
  if (!ds) return 1;
  if (cdx_group_dataset_commit(group, slots, N_SLOTS, 2048, 65536, sampled_slot, ds) != CDX_OK) {
    fprintf(stderr, "dataset commit failed: %s\n", cdx_group_last_error(group));
    return 1;
  }
  uint8_t dataset_root[32], slot_roots[N_SLOTS][32], slot_proof[MAX_LOG2_NSLOTS][32];
  CHECK(cdx_dataset_root(ds[0], dataset_root));
  CHECK(cdx_dataset_slot_roots(ds[0], &slot_roots[0][0]));
  CHECK(cdx_dataset_slot_proof(ds[0], (uint64_t)sampled_slot, MAX_LOG2_NSLOTS, &slot_proof[0][0]));
  printf("GPUs: %d\n", n_gpus);
  print_felt("dataSetRoot = ", dataset_root);
  print_felt("slotRoot    = ", slot_roots[sampled_slot]);

  uint8_t entropy[32] = {0};
  entropy[0] = 0x87; entropy[1] = 0xd6; entropy[2] = 0x12;     /* 1234567, little-endian */
  uint64_t indices[N_SAMPLES];
  static uint8_t paths[N_SAMPLES][MAX_DEPTH][32], leaves[N_SAMPLES][32];
  if (cdx_group_dataset_prove(group, ds, entropy, N_SAMPLES, MAX_DEPTH, indices, &paths[0][0][0], &leaves[0][0]) != CDX_OK) {
    fprintf(stderr, "prove failed: %s\n", cdx_group_last_error(group));
    return 1;
  }
  printf("cell indices:");
  for (int i = 0; i < N_SAMPLES; ++i) printf(" %llu", (unsigned long long)indices[i]);
  printf("\n");

  /* verifier side, two stages like the circuit (single_cell.circom:41-71): cell -> block root (5 levels), block -> slot root */
  cdx_ctx *ctx = cdx_group_ctx(group, 0);
  uint64_t within[N_SAMPLES], block[N_SAMPLES];
  uint8_t block_roots[N_SAMPLES][32], rebuilt[N_SAMPLES][32];
  for (int i = 0; i < N_SAMPLES; ++i) {
    within[i] = indices[i] % 32;
    block[i] = indices[i] / 32;
  }
  CHECK(cdx_reconstruct_roots_host(ctx, &leaves[0][0], within, 32, &paths[0][0][0], MAX_DEPTH, 5, N_SAMPLES, &block_roots[0][0]));
  CHECK(cdx_reconstruct_roots_host(ctx, &block_roots[0][0], block, 8, &paths[0][5][0], MAX_DEPTH, 3, N_SAMPLES, &rebuilt[0][0]));
  int ok = 1;
  for (int i = 0; i < N_SAMPLES; ++i) ok &= memcmp(rebuilt[i], slot_roots[sampled_slot], 32) == 0;
  /* and the slot root under the dataset root */
  uint64_t slot_index = (uint64_t)sampled_slot;
  uint8_t top[32];
  CHECK(cdx_reconstruct_roots_host(ctx, slot_roots[sampled_slot], &slot_index, N_SLOTS, &slot_proof[0][0], MAX_LOG2_NSLOTS, 3, 1, top));
  ok &= memcmp(top, dataset_root, 32) == 0;
  printf("every sampled path reconstructs the slot root, and the slot proof the dataset root: %s\n", ok ? "yes" : "NO");

  cdx_group_datasets_free(group, ds);
  free(ds);
  cdx_group_destroy(group);
  return ok ? 0 : 1;
}
