// Unit tests of the C++ host mirror of reference/nim/proof_input (host/proof_input.hpp) over the CUDA backend.
// Built and run by tests/test_gpu_host_cpp.py on a GPU box; every hash below is computed by libcodexcommit.so.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../codex-storage-proofs-circuits_b200/host/proof_input.hpp"

using namespace codex;

static int failures = 0;
#define CHECK(cond)                                                         \
  do {                                                                      \
    if (!(cond)) { std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond); ++failures; } \
  } while (0)
template <class Fn> static bool throws(Fn fn) {
  try { fn(); } catch (const AssertionDefect&) { return true; }
  return false;
}
static F felt(uint64_t v) { return intToBN254((int64_t)v); }

int main() {
  // ---- pure host helpers (no hashing) ----
  CHECK(toDecimalF(felt(0)) == "0");
  CHECK(toDecimalF(felt(1234567)) == "1234567");
  CHECK(toQuotedDecimalF(felt(42)) == "\"42\"");
  CHECK(toDecimalF(intToBN254(-5)) == "21888242871839275222246405745257275088548364400416034343698204186575808495612");   // r - 5
  CHECK(ceilingLog2(0) == -1 && ceilingLog2(1) == 0 && ceilingLog2(256) == 8 && ceilingLog2(257) == 9 && floorLog2(255) == 7);
  CHECK(exactLog2(32) == 5 && throws([] { exactLog2(33); }) && throws([] { checkPowerOfTwo(1000, "nCells"); }));
  CHECK(parseField("BN254") == FieldSelect::BN254 && parseField("goldilocks") == FieldSelect::Goldilocks && throws([] { parseField("bls"); }));
  CHECK(throws([] { toFieldHashCombo(FieldSelect::BN254, HashSelect::Monolith); }));
  GlobalConfig g;
  CHECK(cellsPerBlock(g) == 32);
  g.blockSize = 3000;
  CHECK(throws([&] { cellsPerBlock(g); }));
  g = GlobalConfig();
  {
    std::vector<uint8_t> b62(62, 0xff);
    auto e = elements(b62);                       // 31k bytes -> k+1 elements (Slot.hs:243-250)
    CHECK(e.size() == 3 && e[2][0] == 1 && e[2][1] == 0 && e[0][30] == 0xff && e[0][31] == 0);
    CHECK(elements(std::vector<uint8_t>()).size() == 1 && elements(std::vector<uint8_t>(2048, 7)).size() == 67);
  }
  CHECK(extractLowBits(felt(0xabcdef), 8) == 0xef && extractLowBits(felt(0xabcdef), 64) == 0xabcdef);
  CHECK(parametricSlotSeed(12345, 3) == 12345 + 72 + 3003);

  // more pure host logic: decimal rendering (types/bn254.nim:29-33) at the edges, and the JSON layout on a hand-made input
  {
    F big{};
    for (int i = 0; i < 32; ++i) big[i] = 0xff;
    CHECK(toDecimalF(big) == "115792089237316195423570985008687907853269984665640564039457584007913129639935");   // 2^256 - 1
    F p10{};                                          // 10^18 = 0x0de0b6b3a7640000
    const uint64_t v = 1000000000000000000ull;
    std::memcpy(p10.data(), &v, 8);
    CHECK(toDecimalF(p10) == "1000000000000000000");
    CHECK(toDecimalF(intToBN254(-1)) == "21888242871839275222246405745257275088548364400416034343698204186575808495616");   // r - 1
    SlotProofInput prf;
    prf.dataSetRoot = felt(7);
    prf.entropy = felt(8);
    prf.nCells = 64;
    prf.nSlots = 3;
    prf.slotIndex = 2;
    prf.slotRoot = felt(9);
    prf.slotProof.merklePath = {felt(1), felt(2)};
    CellProofInput c;
    c.cellData = Cell(62, 0);
    c.merkleProof.merklePath = {felt(3)};
    prf.proofInputs = {c};
    const std::string js = proofInputToJson(prf);
    CHECK(js == "{\n  \"dataSetRoot\":      \"7\"\n, \"entropy\":          \"8\"\n, \"nCellsPerSlot\":    64\n, \"nSlotsPerDataSet\": 3\n, \"slotIndex\":        2\n"
                ", \"slotRoot\":         \"9\"\n, \"slotProof\":\n    [ \"1\"\n    , \"2\"\n    ]\n, \"cellData\":\n    [ [ \"0\"\n      , \"0\"\n      , \"1\"\n      ]\n    ]\n"
                ", \"merklePaths\":\n    [ [ \"3\"\n      ]\n    ]\n}\n");
    if (js.size() < 10) std::printf("%s", js.c_str());
  }
#ifdef CDX_TEST_PURE_ONLY
  std::printf(failures ? "%d FAILURES\n" : "host mirror (pure logic): all checks passed\n", failures);
  return failures ? 1 : 0;
#endif

  // ---- over the GPU backend ----
  Backend be(0);
  HashConfig h;
  h.field = FieldSelect::BN254;
  h.combo = toFieldHashCombo(h.field, h.hashFun);
  const CompressWithKey cwk = [&be](int key, const F& x, const F& y) { return compressWithKey(be, key, x, y); };
  // Merkle.hs:136-152 testAllMerkleProofs: every leaf of trees with 1..12 leaves
  for (int n = 1; n <= 12; ++n) {
    std::vector<F> leaves;
    for (int i = 1; i <= n; ++i) leaves.push_back(felt(1000 + i));
    const MerkleTree t = merkleTree(be, h, leaves);
    CHECK(treeNumberOfLeaves(t) == n && treeRoot(t) == merkleDigestBN254(be, leaves));
    CHECK(treeDepth(t) == (n == 1 ? 1 : ceilingLog2(n)));
    for (int j = 0; j < n; ++j) {
      const MerkleProof p = merkleProof(t, j);
      CHECK(checkMerkleProof(cwk, treeRoot(t), p));
      if ((j ^ 1) >= n) CHECK(p.merklePath[0] == F{});          // out-of-range sibling is zero (merkle.nim:34)
      const MerkleProof padded = padMerkleProof(p, 8);
      CHECK(padded.merklePath.size() == 8 && padded.merklePath[7] == F{});
    }
    CHECK(throws([&] { merkleProof(t, n); }));
  }
  CHECK(compressWithKey(be, 3, felt(5), F{}) == treeRoot(merkleTree(be, h, {felt(5)})));   // singleton = one key-3 compression
  CHECK(throws([&] { compressWithKey(be, 4, felt(1), felt(2)); }));
  // hashCell size assertion (blocks/bn254.nim:26) and block tree == tree over cell hashes
  CHECK(throws([&] { hashCell(be, h, g, Cell(100, 0)); }));
  {
    SlotConfig sc;
    sc.nCells = 64;
    sc.dataSrc.seed = 777;
    const Block blk = slotLoadBlockData(be, g, sc, 1);
    CHECK((int64_t)blk.size() == g.blockSize);
    std::vector<Hash> ch;
    for (int c = 0; c < 32; ++c) ch.push_back(hashCell(be, h, g, slotLoadCellData(be, g, sc, 32 + c)));
    const MerkleTree bt = networkBlockTree(be, h, g, blk);
    CHECK(bt.layers[0] == ch && treeRoot(bt) == hashNetworkBlock(be, h, g, blk) && treeDepth(bt) == 5);
    // merge of a bottom and a top proof, and the mismatch assertion (merkle.nim:86-100)
    const MerkleTree top = merkleTree(be, h, {felt(9), treeRoot(bt), felt(11)});
    const MerkleProof merged = mergeMerkleProofs(cwk, merkleProof(bt, 7), merkleProof(top, 1));
    CHECK(merged.leafIndex == 32 + 7 && merged.numberOfLeaves == 96 && merged.merklePath.size() == 5 + 2);
    CHECK(throws([&] { mergeMerkleProofs(cwk, merkleProof(bt, 7), merkleProof(top, 0)); }));
  }
  // sampling: counters run 1..n, power-of-two assertion (sample/bn254.nim:16-27)
  {
    const auto idx = cellIndices(be, h, felt(1234567), felt(99), 2048, 5);
    CHECK(idx.size() == 5 && cellIndex(be, h, felt(1234567), felt(99), 2048, 3) == idx[2]);
    for (auto i : idx) CHECK(i >= 0 && i < 2048);
    CHECK(throws([&] { cellIndices(be, h, felt(1), felt(2), 1000, 5); }));
  }
  // generateProofInputBN254 argument checks (gen_input/bn254.nim:38-39, dataset.nim:46)
  {
    DataSetConfig d;
    d.nCells = 64;
    d.nSlots = 3;
    CHECK(throws([&] { generateProofInputBN254(be, h, g, d, 5, felt(1)); }));
    d.nCells = 48;                                   // not a multiple of 32 cells per block
    CHECK(throws([&] { generateProofInputBN254(be, h, g, d, 0, felt(1)); }));
    d.nCells = 64;
    const SlotProofInput in = generateProofInputBN254(be, h, g, d, 2, felt(1234567));
    CHECK(in.proofInputs.size() == 5 && in.slotProof.merklePath.size() == 8 && in.nCells == 64 && in.nSlots == 3);
    for (const auto& p : in.proofInputs) CHECK(p.merkleProof.merklePath.size() == 32 && (int64_t)p.cellData.size() == g.cellSize);
    // verifier side: accepts the generated input, pinpoints tampering
    std::string why;
    CHECK(checkProofInputBN254(be, g, in, &why));
    {
      SlotProofInput bad = in;
      bad.proofInputs[1].merkleProof.merklePath[2][0] ^= 1;
      CHECK(!checkProofInputBN254(be, g, bad, &why) && why.find("sample 1") != std::string::npos);
      bad = in;
      bad.proofInputs[3].cellData[100] ^= 0x80;
      CHECK(!checkProofInputBN254(be, g, bad, &why) && why.find("sample 3: leaf value") != std::string::npos);
      bad = in;
      bad.dataSetRoot[5] ^= 1;
      CHECK(!checkProofInputBN254(be, g, bad, &why) && why == "top root check failed");
      bad = in;
      bad.entropy[0] ^= 1;
      CHECK(!checkProofInputBN254(be, g, bad, &why) && why.find("cell index") != std::string::npos);
      bad = in;
      bad.slotProof.merklePath[7][0] = 9;
      CHECK(!checkProofInputBN254(be, g, bad, &why) && why == "slotProof is not zero-padded");
    }
    const std::string js = proofInputToJson(in);
    CHECK(js.rfind("{\n  \"dataSetRoot\":      \"", 0) == 0 && js.find(", \"nSlotsPerDataSet\": 3\n") != std::string::npos);
  }
  // ---- every GPU of the box through the C ABI alone (on a one-GPU box the same calls run with one rank, without NCCL) ----
  {
    const int nGpus = be.visibleGpus();
    std::printf("visible GPUs: %d\n", nGpus);
    // a 40 MiB + 3 blocks slot of counter bytes in host memory; whole-slot commitment on GPU 0 is the expectation
    const size_t nBlocks = 643, nBytes = nBlocks * 65536;
    std::vector<uint8_t> slot(nBytes);
    uint64_t x = 88172645463325252ull;
    for (size_t i = 0; i < nBytes; i += 8) {
      x ^= x << 13; x ^= x >> 7; x ^= x << 17;
      std::memcpy(&slot[i], &x, 8);
    }
    cdx_slot* whole = nullptr;
    be.check(cdx_slot_commit_host(be.ctx(), slot.data(), nBytes, 2048, 65536, &whole), "whole");
    F wholeRoot{};
    be.check(cdx_slot_root(whole, wholeRoot.data()), "root");
    const std::vector<uint64_t> cells = {0, 31, 32 * 321 + 5, 32 * nBlocks - 1};
    std::vector<F> wantPaths(cells.size() * 32), wantLeaves(cells.size());
    be.check(cdx_slot_cell_paths(whole, cells.data(), cells.size(), 32, wantPaths[0].data(), wantLeaves[0].data()), "paths");

    // (a) one thread per GPU, each with its own context and communicator rank, cdx_slot_commit_sharded_host
    {
      uint8_t id[CDX_COMM_ID_BYTES] = {0};
      if (nGpus > 1) CHECK(cdx_comm_unique_id(id) == CDX_OK);
      std::vector<uint64_t> first(nGpus), count(nGpus);
      int T = 0;
      CHECK(cdx_plan_block_ranges(nBlocks, nGpus, &T, first.data(), count.data()) == CDX_OK);
      std::vector<F> roots(nGpus), gotPaths(cells.size() * 32), gotLeaves(cells.size());
      std::vector<int> rcs(nGpus, -100);
      std::vector<std::thread> thr;
      for (int r = 0; r < nGpus; ++r)
        thr.emplace_back([&, r]() {
          cdx_ctx* c = nullptr;
          cdx_comm* comm = nullptr;
          cdx_slot* sh = nullptr;
          int rc = cdx_ctx_create(r, &c);
          if (rc == CDX_OK) rc = cdx_comm_init_rank(c, nGpus, r, id, &comm);
          if (rc == CDX_OK)
            rc = cdx_slot_commit_sharded_host(c, comm, count[r] ? slot.data() + first[r] * 65536 : nullptr, count[r] * 65536, 2048, 65536, first[r],
                                              nBlocks, T, &sh);
          if (rc == CDX_OK) rc = cdx_slot_root(sh, roots[r].data());
          std::vector<F> p(cells.size() * 32), l(cells.size());
          if (rc == CDX_OK) rc = cdx_slot_cell_paths_sharded(sh, comm, cells.data(), cells.size(), 32, p[0].data(), l[0].data());
          if (rc == CDX_OK && r == nGpus - 1) { gotPaths = p; gotLeaves = l; }
          if (rc != CDX_OK) std::printf("rank %d: %s\n", r, cdx_last_error(c));
          cdx_slot_free(sh);
          cdx_comm_destroy(comm);
          cdx_ctx_destroy(c);
          rcs[r] = rc;
        });
      for (auto& t : thr) t.join();
      for (int r = 0; r < nGpus; ++r) CHECK(rcs[r] == CDX_OK && roots[r] == wholeRoot);
      CHECK(gotPaths == wantPaths && gotLeaves == wantLeaves);
    }
    // (b) the same through the group entry points
    {
      cdx_group* grp = be.group();
      CHECK(cdx_group_size(grp) == nGpus);
      std::vector<cdx_slot*> shards(nGpus, nullptr);
      const int rc = cdx_group_slot_commit_host(grp, slot.data(), nBytes, 2048, 65536, shards.data());
      if (rc != CDX_OK) std::printf("group: %s\n", cdx_group_last_error(grp));
      CHECK(rc == CDX_OK);
      for (int r = 0; r < nGpus && rc == CDX_OK; ++r) {
        F root{};
        CHECK(cdx_slot_root(shards[r], root.data()) == CDX_OK && root == wholeRoot);
      }
      std::vector<F> p(cells.size() * 32), l(cells.size());
      CHECK(cdx_group_slot_cell_paths(grp, shards.data(), cells.data(), cells.size(), 32, p[0].data(), l[0].data()) == CDX_OK);
      CHECK(p == wantPaths && l == wantLeaves);
      cdx_group_slots_free(grp, shards.data());
    }
    cdx_slot_free(whole);
    // (c) generateProofInputBN254 over a dataset large enough for the all-GPU path (20 slots x 64 MiB of the reference's fake
    //     data = 1.25 GiB) equals the same call pinned to one GPU, and the verifier accepts it
    {
      DataSetConfig d;
      d.nCells = 32768;
      d.nSlots = 20;
      d.nSamples = 7;
      setenv("CODEX_COMMIT_GROUP_MIN_GIB", "1", 1);          // the all-GPU path normally starts at 32 GiB
      const SlotProofInput all = generateProofInputBN254(be, h, g, d, 13, felt(424242));
      unsetenv("CODEX_COMMIT_GROUP_MIN_GIB");
      setenv("CODEX_COMMIT_GPUS", "1", 1);
      const SlotProofInput one = generateProofInputBN254(be, h, g, d, 13, felt(424242));
      unsetenv("CODEX_COMMIT_GPUS");
      CHECK(proofInputToJson(all) == proofInputToJson(one));
      std::string why;
      CHECK(checkProofInputBN254(be, g, all, &why));
    }
  }
  std::printf(failures ? "%d FAILURES\n" : "host mirror: all checks passed\n", failures);
  return failures ? 1 : 0;
}
