// proof_input.cpp -- see proof_input.hpp.  Orchestration, gathering and formatting only; all arithmetic is CUDA.
#include "proof_input.hpp"

#include <algorithm>
#include <cstdio>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>

namespace codex {

static void nim_assert(bool cond, const std::string& msg) {
  if (!cond) throw AssertionDefect(msg);
}

// ---- nim/types.nim, nim/misc.nim ----------------------------------------------------------------------------

int64_t cellsPerBlock(const GlobalConfig& glob) {   // types.nim:120-123
  nim_assert(glob.cellSize > 0, "cell size must be positive");
  const int64_t k = glob.blockSize / glob.cellSize;
  nim_assert(k * glob.cellSize == glob.blockSize, "block size is not divisible by cell size");
  return k;
}

static std::string lower(std::string s) {
  std::transform(s.begin(), s.end(), s.begin(), [](unsigned char c) { return (char)std::tolower(c); });
  return s;
}

FieldSelect parseField(const std::string& s) {      // types.nim:127-132
  const std::string l = lower(s);
  if (l == "bn254") return FieldSelect::BN254;
  if (l == "goldilocks") return FieldSelect::Goldilocks;
  throw AssertionDefect("parsefield: unrecognized field `" + s + "`");
}

HashSelect parseHashFun(const std::string& s) {     // types.nim:134-139
  const std::string l = lower(s);
  if (l == "poseidon2") return HashSelect::Poseidon2;
  if (l == "monolith") return HashSelect::Monolith;
  throw AssertionDefect("parsefield: unrecognized hash function `" + s + "`");
}

FieldHashCombo toFieldHashCombo(FieldSelect f, HashSelect h) {   // types.nim:144-157
  if (f == FieldSelect::BN254) {
    if (h == HashSelect::Poseidon2) return FieldHashCombo::BN254_Poseidon2;
    throw AssertionDefect("invalid hash function `Monolith` choice for field `BN254`");
  }
  return h == HashSelect::Poseidon2 ? FieldHashCombo::Goldilocks_Poseidon2 : FieldHashCombo::Goldilocks_Monolith;
}

int floorLog2(int64_t x) {     // misc.nim:10-16
  int k = -1;
  for (int64_t y = x; y > 0; y >>= 1) ++k;
  return k;
}
int ceilingLog2(int64_t x) { return x == 0 ? -1 : floorLog2(x - 1) + 1; }   // misc.nim:18-22
int exactLog2(int64_t x) {     // misc.nim:24-27
  const int k = ceilingLog2(x);
  nim_assert(k >= 0 && x == ((int64_t)1 << k), "exactLog2: not a power of two");
  return k;
}
int64_t checkPowerOfTwo(int64_t x, const std::string& what) {   // misc.nim:29-32
  const int k = ceilingLog2(x);
  nim_assert(k >= 0 && x == ((int64_t)1 << k), "`" + what + "` is expected to be a power of 2");
  return x;
}

MerkleProof padMerkleProof(const MerkleProof& old, int newlen) {   // types.nim:27-37
  const int pad = newlen - (int)old.merklePath.size();
  nim_assert(pad >= 0, "padMerkleProof: the path is longer than the requested length");
  MerkleProof p = old;
  p.merklePath.resize(newlen, F{});   // zero elements
  return p;
}

// ---- nim/types/bn254.nim ------------------------------------------------------------------------------------

static const uint8_t kModulusLE[32] = {0x01, 0x00, 0x00, 0xf0, 0x93, 0xf5, 0xe1, 0x43, 0x91, 0x70, 0xb9, 0x79, 0x48, 0xe8, 0x33, 0x28,
                                       0x5d, 0x58, 0x81, 0x81, 0xb6, 0x45, 0x50, 0xb8, 0x29, 0xa0, 0x31, 0xe1, 0x72, 0x4e, 0x64, 0x30};

F intToBN254(int64_t x) {
  F f{};
  uint64_t mag = x < 0 ? (uint64_t)(-(x + 1)) + 1 : (uint64_t)x;
  for (int i = 0; i < 8; ++i) f[i] = (uint8_t)(mag >> (8 * i));
  if (x < 0) {   // r - |x|
    int borrow = 0;
    for (int i = 0; i < 32; ++i) {
      int d = (int)kModulusLE[i] - (int)f[i] - borrow;
      borrow = d < 0;
      f[i] = (uint8_t)(d + (borrow ? 256 : 0));
    }
  }
  return f;
}

std::string toDecimalF(const F& a) {   // types/bn254.nim:29-33: leading zeros stripped, "0" for zero
  uint32_t w[8];
  std::memcpy(w, a.data(), 32);
  std::string digits;
  for (;;) {
    bool nz = false;
    uint64_t rem = 0;
    for (int i = 7; i >= 0; --i) {
      const uint64_t cur = (rem << 32) | w[i];
      w[i] = (uint32_t)(cur / 1000000000u);
      rem = cur % 1000000000u;
      nz |= w[i] != 0;
    }
    for (int d = 0; d < 9; ++d) {
      digits.push_back((char)('0' + rem % 10));
      rem /= 10;
    }
    if (!nz) break;
  }
  while (digits.size() > 1 && digits.back() == '0') digits.pop_back();
  std::reverse(digits.begin(), digits.end());
  return digits;
}

std::string toQuotedDecimalF(const F& a) { return "\"" + toDecimalF(a) + "\""; }

uint64_t extractLowBits(const F& fld, int k) {   // types/bn254.nim:47-59
  nim_assert(k > 0 && k <= 64, "extractLowBits: k out of range");
  uint64_t v = 0;
  std::memcpy(&v, fld.data(), 8);
  return k == 64 ? v : (v & (((uint64_t)1 << k) - 1));
}

std::vector<F> elements(const std::vector<uint8_t>& bytes) {   // 31-byte LE chunks of bytes ++ 0x01 ++ 0x00..; Slot.hs:243-270
  const size_t n = bytes.size() / 31 + 1;
  std::vector<F> out(n, F{});
  for (size_t k = 0; k < n; ++k) {
    const size_t off = 31 * k;
    const size_t m = std::min<size_t>(31, bytes.size() - off);
    if (m) std::memcpy(out[k].data(), bytes.data() + off, m);      // an empty input has no data pointer to copy from
    if (m < 31) out[k][m] = 0x01;
  }
  return out;
}

// ---- backend ------------------------------------------------------------------------------------------------

Backend::Backend(int device) : device_(device) {
  const int rc = cdx_ctx_create(device, &ctx_);
  if (rc != CDX_OK)
    throw AssertionDefect(std::string("cdx_ctx_create failed: ") + cdx_status_string(rc) + " (this backend needs a CUDA device; there is no CPU path)");
}
Backend::~Backend() {
  cdx_group_destroy(group_);
  cdx_ctx_destroy(ctx_);
}
int Backend::visibleGpus() const {
  int n = cdx_device_count() - device_;             // this backend's device and the ones after it
  if (n < 1) n = 1;
  if (const char* cap = std::getenv("CODEX_COMMIT_GPUS")) {
    const int c = std::atoi(cap);
    if (c >= 1 && c < n) n = c;
  }
  return n;
}
cdx_group* Backend::group() {
  if (!group_) {
    std::vector<int> devs;
    const int n = visibleGpus();
    for (int i = 0; i < n; ++i) devs.push_back(device_ + i);
    const int rc = cdx_group_create(devs.data(), (int)devs.size(), &group_);
    if (rc != CDX_OK) throw AssertionDefect(std::string("cdx_group_create failed: ") + cdx_status_string(rc));
  }
  return group_;
}
void Backend::check(int rc, const char* what) const {
  if (rc != CDX_OK) throw AssertionDefect(std::string(what) + ": " + cdx_last_error(ctx_) + " [" + cdx_status_string(rc) + "]");
}

// ---- nim/merkle/bn254.nim, nim/merkle.nim -------------------------------------------------------------------

F compressWithKey(Backend& be, int key, const F& x, const F& y) {
  nim_assert(key >= 0 && key <= 3, "compressWithKey: key must be 0..3");
  F out{};
  const uint32_t k = (uint32_t)key;
  be.check(cdx_compress_batch_host(be.ctx(), x.data(), y.data(), &k, 1, out.data()), "compress");
  return out;
}

F merkleDigestBN254(Backend& be, const std::vector<F>& xs) {
  nim_assert(!xs.empty(), "merkle digest of empty input");
  F out{};
  be.check(cdx_merkle_root_host(be.ctx(), xs[0].data(), xs.size(), out.data()), "Merkle.digest");
  return out;
}

MerkleTree merkleTreeBN254(Backend& be, const std::vector<F>& xs) {
  nim_assert(!xs.empty(), "merkle tree of empty input");
  const size_t total = cdx_merkle_total_nodes(xs.size(), 1);
  std::vector<F> flat(total);
  be.check(cdx_merkle_layers_host(be.ctx(), xs[0].data(), xs.size(), 1, flat[0].data()), "merkleTree");
  MerkleTree t;
  size_t off = 0, m = xs.size();
  const int nl = cdx_merkle_num_layers(xs.size(), 1);
  for (int l = 0; l < nl; ++l) {
    t.layers.emplace_back(flat.begin() + off, flat.begin() + off + m);
    off += m;
    m = (m + 1) / 2;
  }
  return t;
}

int treeDepth(const MerkleTree& t) { return (int)t.layers.size() - 1; }
int64_t treeNumberOfLeaves(const MerkleTree& t) { return (int64_t)t.layers.at(0).size(); }
Hash treeRoot(const MerkleTree& t) {
  nim_assert(!t.layers.empty() && t.layers.back().size() == 1, "treeRoot: topmost layer is not a singleton");
  return t.layers.back()[0];
}

MerkleProof merkleProof(const MerkleTree& tree, int64_t index) {   // merkle.nim:21-42
  const int depth = treeDepth(tree);
  const int64_t nleaves = treeNumberOfLeaves(tree);
  nim_assert(index >= 0 && index < nleaves, "merkleProof: index out of range");
  MerkleProof p;
  p.merklePath.resize(depth);
  int64_t k = index, m = nleaves;
  for (int i = 0; i < depth; ++i) {
    const int64_t j = k ^ 1;
    p.merklePath[i] = j < m ? tree.layers[i][j] : F{};
    k >>= 1;
    m = (m + 1) >> 1;
  }
  p.leafIndex = index;
  p.leafValue = tree.layers[0][index];
  p.numberOfLeaves = nleaves;
  return p;
}

Hash reconstructRoot(const CompressWithKey& c, const MerkleProof& proof) {   // merkle.nim:51-74
  int64_t m = proof.numberOfLeaves, j = proof.leafIndex;
  Hash h = proof.leafValue;
  int bottomFlag = 1;
  for (const Hash& p : proof.merklePath) {
    if (j & 1) h = c(bottomFlag, p, h);
    else if (j == m - 1) h = c(bottomFlag + 2, h, p);
    else h = c(bottomFlag, h, p);
    bottomFlag = 0;
    j >>= 1;
    m = (m + 1) >> 1;
  }
  return h;
}

bool checkMerkleProof(const CompressWithKey& c, const Root& root, const MerkleProof& p) { return root == reconstructRoot(c, p); }

MerkleProof mergeMerkleProofs(const CompressWithKey& c, const MerkleProof& bot, const MerkleProof& top) {   // merkle.nim:86-100
  const Hash botRoot = reconstructRoot(c, bot);
  nim_assert(botRoot == top.leafValue, "mergeMerkleProofs: bottom root does not match the top leaf");
  MerkleProof p;
  p.leafIndex = top.leafIndex * bot.numberOfLeaves + bot.leafIndex;
  p.leafValue = bot.leafValue;
  p.numberOfLeaves = bot.numberOfLeaves * top.numberOfLeaves;
  p.merklePath = bot.merklePath;
  p.merklePath.insert(p.merklePath.end(), top.merklePath.begin(), top.merklePath.end());
  return p;
}

// ---- nim/blocks/bn254.nim -----------------------------------------------------------------------------------

MerkleTree merkleTree(Backend& be, const HashConfig& h, const std::vector<Hash>& what) {
  nim_assert(h.combo == FieldHashCombo::BN254_Poseidon2, "merkleTree: only BN254/Poseidon2 is implemented by this backend");
  return merkleTreeBN254(be, what);
}

Hash hashCell(Backend& be, const HashConfig& h, const GlobalConfig& g, const Cell& cellData) {
  nim_assert(h.field == FieldSelect::BN254, "hashCell: field must be BN254");
  nim_assert(h.hashFun == HashSelect::Poseidon2, "hashCell: hash must be Poseidon2");
  nim_assert((int64_t)cellData.size() == g.cellSize, "cells are expected to be exactly " + std::to_string(g.cellSize) + " bytes");
  Hash out{};
  be.check(cdx_hash_bytes_batch_host(be.ctx(), cellData.data(), 1, cellData.size(), out.data()), "hashCell");
  return out;
}

static std::vector<Hash> hashBlockCells(Backend& be, const HashConfig& h, const GlobalConfig& g, const Block& blockData) {
  nim_assert(h.field == FieldSelect::BN254, "field must be BN254");
  nim_assert((int64_t)blockData.size() == g.blockSize, "network blocks are expected to be exactly" + std::to_string(g.blockSize) + " bytes");
  const int64_t k = cellsPerBlock(g);
  std::vector<Hash> leaves(k);
  be.check(cdx_hash_bytes_batch_host(be.ctx(), blockData.data(), (size_t)k, (size_t)g.cellSize, leaves[0].data()), "hashCell (block)");
  return leaves;
}

Hash hashNetworkBlock(Backend& be, const HashConfig& h, const GlobalConfig& g, const Block& blockData) {
  return merkleDigestBN254(be, hashBlockCells(be, h, g, blockData));
}

MerkleTree networkBlockTree(Backend& be, const HashConfig& h, const GlobalConfig& g, const Block& blockData) {
  return merkleTree(be, h, hashBlockCells(be, h, g, blockData));
}

// ---- nim/sample/bn254.nim -----------------------------------------------------------------------------------

std::vector<int64_t> cellIndices(Backend& be, const HashConfig& h, const Entropy& e, const Root& slotRoot, int64_t numberOfCells, int64_t nSamples) {
  nim_assert(h.field == FieldSelect::BN254, "cellIndex: field must be BN254");
  const int lg = ceilingLog2(numberOfCells);
  nim_assert(lg >= 0 && ((int64_t)1 << lg) == numberOfCells, "for this version, `numberOfCells` is assumed to be a power of two");
  std::vector<uint64_t> idx((size_t)std::max<int64_t>(nSamples, 0));
  if (nSamples > 0) be.check(cdx_cell_indices(be.ctx(), e.data(), slotRoot.data(), (uint64_t)numberOfCells, idx.size(), idx.data()), "cellIndices");
  return std::vector<int64_t>(idx.begin(), idx.end());
}

int64_t cellIndex(Backend& be, const HashConfig& h, const Entropy& e, const Root& slotRoot, int64_t numberOfCells, int counter) {
  nim_assert(counter >= 1, "cellIndex: counters start at 1");
  return cellIndices(be, h, e, slotRoot, numberOfCells, counter).back();
}

// ---- nim/slot.nim, nim/dataset.nim --------------------------------------------------------------------------

static Cell readFileRange(const std::string& fname, int64_t offset, int64_t len) {   // slot.nim:57-68 (short reads zero-filled)
  std::ifstream f(fname, std::ios::binary);
  nim_assert(f.good(), "cannot open slot data file `" + fname + "`");
  Cell cell((size_t)len, 0);
  f.seekg(offset);
  f.read(reinterpret_cast<char*>(cell.data()), len);
  return cell;
}

Cell slotLoadCellData(Backend& be, const GlobalConfig& g, const SlotConfig& cfg, CellIdx idx) {
  if (cfg.dataSrc.kind == DataSourceKind::FakeData) {
    Cell cell((size_t)g.cellSize);
    be.check(cdx_fake_cells_host(be.ctx(), cfg.dataSrc.seed, (uint64_t)idx, 1, (size_t)g.cellSize, cell.data()), "genFakeCell");
    return cell;
  }
  nim_assert(g.cellSize <= 16384, "cell size exceeds the reference's 16384-byte file buffer");   // slot.nim:61-62
  return readFileRange(cfg.dataSrc.filename, g.cellSize * idx, g.cellSize);
}

Block slotLoadBlockData(Backend& be, const GlobalConfig& g, const SlotConfig& cfg, BlockIdx idx) {
  const int64_t k = cellsPerBlock(g);
  if (cfg.dataSrc.kind == DataSourceKind::FakeData) {
    Block blk((size_t)g.blockSize);
    be.check(cdx_fake_cells_host(be.ctx(), cfg.dataSrc.seed, (uint64_t)(idx * k), (size_t)k, (size_t)g.cellSize, blk.data()), "genFakeCell (block)");
    return blk;
  }
  return readFileRange(cfg.dataSrc.filename, g.blockSize * idx, g.blockSize);
}

Seed parametricSlotSeed(Seed seed, SlotIdx k) { return seed + 72 + 1001 * (uint64_t)k; }   // wrap-around, dataset.nim:31-32

SlotConfig slotCfgFromDataSetCfg(const DataSetConfig& d, SlotIdx idx) {
  nim_assert(idx >= 0 && idx < d.nSlots, "slot index out of range");
  SlotConfig s;
  s.nCells = d.nCells;
  s.nSamples = d.nSamples;
  s.dataSrc = d.dataSrc;
  if (d.dataSrc.kind == DataSourceKind::FakeData) s.dataSrc.seed = parametricSlotSeed(d.dataSrc.seed, idx);
  else s.dataSrc.filename = d.dataSrc.filename + std::to_string(idx) + ".dat";   // dataset.nim:34
  return s;
}

Cell dataSetLoadCellData(Backend& be, const GlobalConfig& g, const DataSetConfig& d, SlotIdx s, CellIdx c) {
  return slotLoadCellData(be, g, slotCfgFromDataSetCfg(d, s), c);
}
Block dataSetLoadBlockData(Backend& be, const GlobalConfig& g, const DataSetConfig& d, SlotIdx s, BlockIdx b) {
  return slotLoadBlockData(be, g, slotCfgFromDataSetCfg(d, s), b);
}

// ---- nim/gen_input/bn254.nim --------------------------------------------------------------------------------

namespace {
// CODEX_HOST_TRACE=1: wall-clock milestones of generateProofInputBN254 on stderr
struct Trace {
  const bool on = std::getenv("CODEX_HOST_TRACE") != nullptr;
  const std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  void mark(const char* what) const {
    if (on) std::fprintf(stderr, "[trace] %8.3f s  %s\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(), what);
  }
};

// the dataset committed on one GPU (comm == NULL) or on every GPU of the backend's group: handles per rank
struct DatasetHandles {
  Backend& be;
  cdx_group* group = nullptr;
  std::vector<cdx_dataset*> ds;
  explicit DatasetHandles(Backend& b) : be(b) {}
  ~DatasetHandles() {
    if (group) cdx_group_datasets_free(group, ds.data());
    else
      for (cdx_dataset* d : ds) cdx_dataset_free(d);
  }
};
}  // namespace

SlotProofInput generateProofInputBN254(Backend& be, const HashConfig& hashCfg, const GlobalConfig& globCfg, const DataSetConfig& dsetCfg,
                                       SlotIdx slotIdx, const Entropy& entropy) {   // gen_input/bn254.nim:35-79
  nim_assert(hashCfg.combo == FieldHashCombo::BN254_Poseidon2, "only --field=bn254 --hash=poseidon2 is implemented by this backend");
  const int64_t nslots = dsetCfg.nSlots, ncells = dsetCfg.nCells;
  const int64_t cpb = cellsPerBlock(globCfg);
  const int64_t nblocks = ncells / cpb;
  nim_assert(nblocks * cpb == ncells && nblocks > 0, "slot size is not divisible by the block size");
  nim_assert(slotIdx >= 0 && slotIdx < nslots, "slot index out of range");

  // Every slot is committed exactly once (the reference rebuilds the sampled slot per sample, :57; same trees), by ONE
  // library call: cdx_dataset_commit deals the slots to the GPUs, batches the small ones, shards the huge ones, combines
  // the roots and builds the dataset tree (:41-51).  Slot sources as slotCfgFromDataSetCfg gives them (dataset.nim:45-51).
  std::vector<SlotConfig> slotCfgs;
  std::vector<cdx_slot_desc> descs((size_t)nslots);
  for (SlotIdx i = 0; i < nslots; ++i) slotCfgs.push_back(slotCfgFromDataSetCfg(dsetCfg, i));
  for (SlotIdx i = 0; i < nslots; ++i) {
    cdx_slot_desc& d = descs[(size_t)i];
    std::memset(&d, 0, sizeof d);
    d.n_bytes = (uint64_t)(globCfg.cellSize * ncells);
    if (slotCfgs[(size_t)i].dataSrc.kind == DataSourceKind::FakeData) {
      d.kind = CDX_SRC_FAKE;
      d.seed = slotCfgs[(size_t)i].dataSrc.seed;
    } else {
      d.kind = CDX_SRC_FILE;                 // pread -> pinned buffers -> H2D -> sponge, overlapped (short files read as zeros, slot.nim:64-65)
      d.path = slotCfgs[(size_t)i].dataSrc.filename.c_str();
    }
  }
  DatasetHandles h(be);
  const Trace trace;
  // All GPUs when it pays: creating the group costs one NCCL initialisation (a second or two for eight GPUs, once per
  // process), which a single B200 spends committing ~30 GB; a one-shot cli therefore takes the group from 32 GiB up
  // (CODEX_COMMIT_GROUP_MIN_GIB overrides the threshold; a long-lived host that keeps its Backend pays the start-up once).
  const uint64_t totalBytes = (uint64_t)nslots * descs[0].n_bytes;
  double minGib = 32.0;
  if (const char* v = std::getenv("CODEX_COMMIT_GROUP_MIN_GIB")) minGib = std::atof(v);
  const bool useGroup = (double)totalBytes >= minGib * (double)((uint64_t)1 << 30) && be.visibleGpus() > 1;
  if (useGroup) {
    h.group = be.group();
    trace.mark("group of all GPUs created");
    h.ds.assign((size_t)cdx_group_size(h.group), nullptr);
    const int rc = cdx_group_dataset_commit(h.group, descs.data(), descs.size(), (size_t)globCfg.cellSize, (size_t)globCfg.blockSize, slotIdx, h.ds.data());
    if (rc != CDX_OK) throw AssertionDefect(std::string("buildSlotTree (all GPUs): ") + cdx_group_last_error(h.group) + " [" + cdx_status_string(rc) + "]");
  } else {
    h.ds.assign(1, nullptr);
    be.check(cdx_dataset_commit(be.ctx(), nullptr, descs.data(), descs.size(), (size_t)globCfg.cellSize, (size_t)globCfg.blockSize, slotIdx, &h.ds[0]),
             "buildSlotTree");
  }
  trace.mark(useGroup ? "dataset committed on all GPUs" : "dataset committed on one GPU");
  cdx_dataset* ds0 = h.ds[0];
  std::vector<Root> slotRoots((size_t)nslots);
  be.check(cdx_dataset_slot_roots(ds0, slotRoots[0].data()), "treeRoot");
  const Root ourSlotRoot = slotRoots[(size_t)slotIdx];
  Hash dsetRoot{};
  be.check(cdx_dataset_root(ds0, dsetRoot.data()), "treeRoot (dataset)");                // :49-50
  nim_assert(nslots == 1 || ceilingLog2(nslots) <= globCfg.maxLog2NSlots, "padMerkleProof: the path is longer than the requested length");
  MerkleProof slotProof;                                                                 // :51, already padded (types.nim:27-37)
  slotProof.leafIndex = slotIdx;
  slotProof.leafValue = ourSlotRoot;
  slotProof.numberOfLeaves = nslots;
  slotProof.merklePath.assign((size_t)globCfg.maxLog2NSlots, F{});
  {
    const int rc = cdx_dataset_slot_proof(ds0, (uint64_t)slotIdx, (size_t)globCfg.maxLog2NSlots, slotProof.merklePath.empty() ? nullptr : slotProof.merklePath[0].data());
    if (rc == CDX_ERR_RANGE) throw AssertionDefect("padMerkleProof: the path is longer than the requested length");
    be.check(rc, "merkleProof (dataset)");
  }

  // sampled indices (:53), cell hashes and merged, padded paths (:56-63) in one call; the kept slot may live on any GPU
  const int lg = ceilingLog2(ncells);
  nim_assert(lg >= 0 && ((int64_t)1 << lg) == ncells, "for this version, `numberOfCells` is assumed to be a power of two");
  const uint32_t bd = cpb == 1 ? 1 : (uint32_t)exactLog2(cpb);
  const uint32_t sd = nblocks == 1 ? 1 : (uint32_t)ceilingLog2(nblocks);
  nim_assert((int)(bd + sd) <= globCfg.maxDepth, "padMerkleProof: the path is longer than the requested length");
  const size_t ns = (size_t)std::max<int64_t>(dsetCfg.nSamples, 0);
  std::vector<uint64_t> idx64(ns);
  std::vector<F> paths(ns * (size_t)globCfg.maxDepth), leaves(ns);
  if (ns) {
    if (useGroup) {
      const int rc = cdx_group_dataset_prove(h.group, h.ds.data(), entropy.data(), ns, (size_t)globCfg.maxDepth, idx64.data(), paths[0].data(), leaves[0].data());
      if (rc != CDX_OK) throw AssertionDefect(std::string("merkleProof (all GPUs): ") + cdx_group_last_error(h.group) + " [" + cdx_status_string(rc) + "]");
    } else {
      be.check(cdx_dataset_prove(ds0, entropy.data(), ns, (size_t)globCfg.maxDepth, idx64.data(), paths[0].data(), leaves[0].data()), "merkleProof (batched)");
    }
  }
  const std::vector<int64_t> indices(idx64.begin(), idx64.end());
  trace.mark("indices, cell hashes and paths of all samples");

  // mergeMerkleProofs (merkle.nim:86-100) re-hashes every bottom proof and asserts it lands on the top proof's leaf (the
  // block hash held in the slot tree); here that check runs for all samples in ONE batched verifier launch
  std::vector<F> botRoots(ns);
  std::vector<uint64_t> botIdx(ns);
  for (size_t s = 0; s < ns; ++s) botIdx[s] = (uint64_t)(indices[s] % cpb);
  if (ns) be.check(cdx_reconstruct_roots_host(be.ctx(), leaves[0].data(), botIdx.data(), (uint64_t)cpb, paths[0].data(), (size_t)globCfg.maxDepth,
                                              bd, ns, botRoots[0].data()), "reconstructRoot (batched)");
  auto blockHashOf = [&](int64_t blockIdx) {            // layers[0][blockIdx] of the kept slot tree, from whichever GPU holds it
    F out{};
    for (cdx_dataset* d : h.ds) {
      cdx_slot* kept = cdx_dataset_kept_slot(d);
      if (kept && cdx_slot_read_layer(kept, 1, 0, (uint64_t)blockIdx, 1, out.data()) == CDX_OK) return out;
    }
    throw AssertionDefect("block hash " + std::to_string(blockIdx) + " of the sampled slot is not held by any GPU");
  };
  SlotProofInput out;
  const SlotConfig& ourSlotCfg = slotCfgs[(size_t)slotIdx];
  for (size_t s = 0; s < ns; ++s) {
    const int64_t cellIdx = indices[s], blockIdx = cellIdx / cpb;
    const F* p = &paths[s * (size_t)globCfg.maxDepth];
    nim_assert(botRoots[s] == blockHashOf(blockIdx), "mergeMerkleProofs: bottom root does not match the top leaf");
    MerkleProof merged;                                                                  // merkle.nim:91-99
    merged.leafIndex = blockIdx * cpb + cellIdx % cpb;
    merged.leafValue = leaves[s];
    merged.numberOfLeaves = cpb * nblocks;
    merged.merklePath.assign(p, p + bd + sd);
    CellProofInput cpi;
    cpi.cellData = slotLoadCellData(be, globCfg, ourSlotCfg, cellIdx);                  // :60
    cpi.merkleProof = padMerkleProof(merged, globCfg.maxDepth);                         // :63
    out.proofInputs.push_back(std::move(cpi));
  }
  trace.mark("sampled cells loaded, bottom proofs checked");
  out.dataSetRoot = dsetRoot;
  out.entropy = entropy;
  out.nCells = ncells;
  out.nSlots = nslots;
  out.slotIndex = slotIdx;
  out.slotRoot = ourSlotRoot;
  out.slotProof = slotProof;
  return out;
}

// ---- verifier side ------------------------------------------------------------------------------------------

bool checkProofInputBN254(Backend& be, const GlobalConfig& g, const SlotProofInput& prf, std::string* why) {
  auto no = [&](const std::string& msg) {
    if (why) *why = msg;
    return false;
  };
  const int64_t cpb = cellsPerBlock(g);
  if (prf.nCells <= 0 || prf.nCells % cpb) return no("nCellsPerSlot is not a whole number of blocks");
  const int lg = ceilingLog2(prf.nCells);
  if (((int64_t)1 << lg) != prf.nCells) return no("nCellsPerSlot is not a power of two");              // sample_cells.circom:114-123
  const int64_t nblocks = prf.nCells / cpb;
  const size_t ns = prf.proofInputs.size();
  const int bd = cpb == 1 ? 1 : exactLog2(cpb);
  const int sd = nblocks == 1 ? 1 : ceilingLog2(nblocks);
  if ((int)prf.slotProof.merklePath.size() != g.maxLog2NSlots) return no("slotProof has the wrong length");
  if (prf.slotIndex < 0 || prf.slotIndex >= prf.nSlots) return no("slotIndex out of range");
  // 1. slot root -> dataset root
  {
    const int dd = prf.nSlots == 1 ? 1 : ceilingLog2(prf.nSlots);
    if (dd > g.maxLog2NSlots) return no("nSlotsPerDataSet exceeds maxLog2NSlots");
    const uint64_t idx = (uint64_t)prf.slotIndex;
    F top{};
    be.check(cdx_reconstruct_roots_host(be.ctx(), prf.slotRoot.data(), &idx, (uint64_t)prf.nSlots, prf.slotProof.merklePath[0].data(),
                                        (size_t)g.maxLog2NSlots, (size_t)dd, 1, top.data()), "dataset path");
    if (top != prf.dataSetRoot) return no("top root check failed");
    for (int i = dd; i < g.maxLog2NSlots; ++i)
      if (prf.slotProof.merklePath[(size_t)i] != F{}) return no("slotProof is not zero-padded");
  }
  if (ns == 0) return true;
  // 2. sampled indices
  HashConfig h;
  h.field = FieldSelect::BN254;
  h.combo = FieldHashCombo::BN254_Poseidon2;
  const std::vector<int64_t> idx = cellIndices(be, h, prf.entropy, prf.slotRoot, prf.nCells, (int64_t)ns);
  // 3. cell hashes from the cell data (as field elements, exactly what the circuit is given)
  std::vector<F> leaves(ns);
  {
    std::vector<uint8_t> cells;
    for (const auto& p : prf.proofInputs) {
      if ((int64_t)p.cellData.size() != g.cellSize) return no("cell data has the wrong size");
      cells.insert(cells.end(), p.cellData.begin(), p.cellData.end());
    }
    be.check(cdx_hash_bytes_batch_host(be.ctx(), cells.data(), ns, (size_t)g.cellSize, leaves[0].data()), "cell hashes");
  }
  // 4./5. block-level, then slot-level reconstruction
  std::vector<F> paths(ns * (size_t)g.maxDepth), blockRoots(ns), slotRoots(ns);
  std::vector<uint64_t> within(ns), block(ns);
  for (size_t s = 0; s < ns; ++s) {
    const MerkleProof& mp = prf.proofInputs[s].merkleProof;
    if ((int)mp.merklePath.size() != g.maxDepth || bd + sd > g.maxDepth) return no("merkle path has the wrong length");
    if (mp.leafIndex != idx[s]) return no("sample " + std::to_string(s) + ": cell index does not match H(entropy | slotRoot | counter)");
    if (mp.leafValue != leaves[s]) return no("sample " + std::to_string(s) + ": leaf value is not the hash of the cell data");
    std::copy(mp.merklePath.begin(), mp.merklePath.end(), paths.begin() + s * (size_t)g.maxDepth);
    within[s] = (uint64_t)(idx[s] % cpb);
    block[s] = (uint64_t)(idx[s] / cpb);
    for (int i = bd + sd; i < g.maxDepth; ++i)
      if (mp.merklePath[(size_t)i] != F{}) return no("sample " + std::to_string(s) + ": merkle path is not zero-padded");
  }
  be.check(cdx_reconstruct_roots_host(be.ctx(), leaves[0].data(), within.data(), (uint64_t)cpb, paths[0].data(), (size_t)g.maxDepth, (size_t)bd, ns,
                                      blockRoots[0].data()), "block-level reconstruction");
  be.check(cdx_reconstruct_roots_host(be.ctx(), blockRoots[0].data(), block.data(), (uint64_t)nblocks, paths[(size_t)bd].data(), (size_t)g.maxDepth,
                                      (size_t)sd, ns, slotRoots[0].data()), "slot-level reconstruction");
  for (size_t s = 0; s < ns; ++s)
    if (slotRoots[s] != prf.slotRoot) return no("sample " + std::to_string(s) + ": middle/bottom root check failed");
  return true;
}

// ---- nim/json/bn254.nim, nim/json/shared.nim ----------------------------------------------------------------

namespace {
void writeFieldElems(std::ostream& h, const std::string& prefix, const std::vector<F>& xs) {   // shared.nim:17-25 + bn254.nim:19-20
  const std::string indent(prefix.size(), ' ');
  for (size_t i = 0; i < xs.size(); ++i) h << (i == 0 ? prefix + "[ " : indent + ", ") << toQuotedDecimalF(xs[i]) << "\n";
  h << indent << "]\n";
}
template <class T, class Fn>
void writeList(std::ostream& h, const std::string& prefix, const std::vector<T>& xs, Fn fn) {
  const std::string indent(prefix.size(), ' ');
  for (size_t i = 0; i < xs.size(); ++i) fn(h, i == 0 ? prefix + "[ " : indent + ", ", xs[i]);
  h << indent << "]\n";
}
}  // namespace

std::string proofInputToJson(const SlotProofInput& prf) {   // json/bn254.nim:57-74
  std::ostringstream h;
  h << "{\n";
  h << "  \"dataSetRoot\":      " << toQuotedDecimalF(prf.dataSetRoot) << "\n";
  h << ", \"entropy\":          " << toQuotedDecimalF(prf.entropy) << "\n";
  h << ", \"nCellsPerSlot\":    " << prf.nCells << "\n";
  h << ", \"nSlotsPerDataSet\": " << prf.nSlots << "\n";
  h << ", \"slotIndex\":        " << prf.slotIndex << "\n";
  h << ", \"slotRoot\":         " << toQuotedDecimalF(prf.slotRoot) << "\n";
  h << ", \"slotProof\":\n";
  writeFieldElems(h, "    ", prf.slotProof.merklePath);
  h << ", \"cellData\":\n";
  writeList(h, "    ", prf.proofInputs, [](std::ostream& o, const std::string& p, const CellProofInput& c) { writeFieldElems(o, p, elements(c.cellData)); });
  h << ", \"merklePaths\":\n";
  writeList(h, "    ", prf.proofInputs, [](std::ostream& o, const std::string& p, const CellProofInput& c) { writeFieldElems(o, p, c.merkleProof.merklePath); });
  h << "}\n";
  return h.str();
}

void exportProofInputBN254(const HashConfig& h, const std::string& fname, const SlotProofInput& prf) {
  nim_assert(h.field == FieldSelect::BN254, "exportProofInputBN254: field must be BN254");
  std::ofstream f(fname, std::ios::binary);
  nim_assert(f.good(), "cannot open `" + fname + "` for writing");
  f << proofInputToJson(prf);
}

}  // namespace codex
