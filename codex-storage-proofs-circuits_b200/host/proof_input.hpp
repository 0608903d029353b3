// proof_input.hpp -- host-side mirror of reference/nim/proof_input (BN254 / Poseidon2 path) over the C ABI.
//
// The reference host is Nim; no Nim toolchain exists in the build image, so the host layer is C++ with the SAME
// proc names, argument meaning and error behaviour (Nim `assert`/`raiseAssert` -> codex::AssertionDefect with
// the reference's message; the CLI turns that into a non-zero exit like the Nim binary's abort).
// Every hash, compression, tree level and sampled index is computed by libcodexcommit.so (CUDA); this layer only
// orchestrates, gathers and formats.  `nim/` below = reference/nim/proof_input/src/.
#pragma once
#include <array>
#include <cstdint>
#include <functional>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/codex_commit.h"

namespace codex {

struct AssertionDefect : std::runtime_error { using std::runtime_error::runtime_error; };

// ---- nim/types.nim:7-8, nim/types/bn254.nim:20-23 -----------------------------------------------------------
using F = std::array<uint8_t, 32>;   // canonical little-endian
using Hash = F;
using Root = F;
using Entropy = F;
using Cell = std::vector<uint8_t>;
using Block = std::vector<uint8_t>;
using Seed = uint64_t;
using CellIdx = int64_t;
using BlockIdx = int64_t;
using SlotIdx = int64_t;

struct MerkleProof {                 // nim/types.nim:14-18
  int64_t leafIndex = 0;
  Hash leafValue{};
  std::vector<Hash> merklePath;      // bottom -> top
  int64_t numberOfLeaves = 0;
};
struct MerkleTree {                  // nim/types.nim:20-22
  std::vector<std::vector<Hash>> layers;   // layers[0] = leaves, layers.back() = {root}
};

enum class DataSourceKind { SlotFile, FakeData };                 // nim/types.nim:64-74
struct DataSource { DataSourceKind kind = DataSourceKind::FakeData; std::string filename; Seed seed = 12345; };
struct SlotConfig { int64_t nCells = 256; int64_t nSamples = 5; DataSource dataSrc; };                 // :76-79
struct DataSetConfig { int64_t nSlots = 11; int64_t nCells = 256; int64_t nSamples = 5; DataSource dataSrc; };   // :81-85
struct GlobalConfig { int maxDepth = 32; int maxLog2NSlots = 8; int64_t cellSize = 2048; int64_t blockSize = 65536; };   // :87-91
enum class FieldSelect { BN254, Goldilocks };                     // nim/types.nim:98-100
enum class HashSelect { Poseidon2, Monolith };
enum class FieldHashCombo { BN254_Poseidon2, Goldilocks_Poseidon2, Goldilocks_Monolith };
struct HashConfig { FieldSelect field = FieldSelect::Goldilocks; HashSelect hashFun = HashSelect::Poseidon2;
                    FieldHashCombo combo = FieldHashCombo::Goldilocks_Poseidon2; };                       // :93-96, cli.nim:47-51

struct CellProofInput { Cell cellData; MerkleProof merkleProof; };                                       // nim/types.nim:48-50
struct SlotProofInput {                                                                                   // nim/types.nim:52-60
  Root dataSetRoot{}; Entropy entropy{}; int64_t nSlots = 0; int64_t nCells = 0; Root slotRoot{}; SlotIdx slotIndex = 0;
  MerkleProof slotProof; std::vector<CellProofInput> proofInputs;
};

// ---- nim/types.nim:120-160, nim/misc.nim -------------------------------------------------------------------
int64_t cellsPerBlock(const GlobalConfig& glob);
FieldSelect parseField(const std::string& s);
HashSelect parseHashFun(const std::string& s);
FieldHashCombo toFieldHashCombo(FieldSelect f, HashSelect h);
int floorLog2(int64_t x);
int ceilingLog2(int64_t x);
int exactLog2(int64_t x);
int64_t checkPowerOfTwo(int64_t x, const std::string& what);
MerkleProof padMerkleProof(const MerkleProof& old, int newlen);                                          // nim/types.nim:27-37

// ---- nim/types/bn254.nim ------------------------------------------------------------------------------------
F intToBN254(int64_t x);                                   // :27 (negative x -> r - |x|, as toF does)
std::string toDecimalF(const F& a);                        // :29-33
std::string toQuotedDecimalF(const F& a);                  // :35-37
uint64_t extractLowBits(const F& fld, int k);              // :47-59
std::vector<F> elements(const std::vector<uint8_t>& bytes);   // nim-poseidon2 `bytes.elements(F)` used at nim/json/bn254.nim:25

// The backend: one GPU context.  All functions below that hash take it explicitly (the Nim procs are free
// functions over a CPU library; here the library is a device).
class Backend {
 public:
  explicit Backend(int device = 0);
  ~Backend();
  Backend(const Backend&) = delete;
  Backend& operator=(const Backend&) = delete;
  cdx_ctx* ctx() const { return ctx_; }
  void check(int rc, const char* what) const;              // throws AssertionDefect on rc != 0
  // Every visible GPU of this process as one group (cdx_group: a context, a communicator rank and a worker thread per
  // device), created on first use.  generateProofInputBN254 commits datasets of 32 GiB and more through it
  // (CODEX_COMMIT_GROUP_MIN_GIB changes that threshold); the environment variable CODEX_COMMIT_GPUS caps the number of
  // GPUs (1 = never use the group).
  cdx_group* group();
  int visibleGpus() const;
 private:
  cdx_ctx* ctx_ = nullptr;
  cdx_group* group_ = nullptr;
  int device_ = 0;
};

// ---- nim/merkle/bn254.nim, nim/merkle.nim -------------------------------------------------------------------
F compressWithKey(Backend& be, int key, const F& x, const F& y);                                        // merkle/bn254.nim:18
F merkleDigestBN254(Backend& be, const std::vector<F>& xs);                                             // :20
MerkleTree merkleTreeBN254(Backend& be, const std::vector<F>& xs);                                      // :62-63
int treeDepth(const MerkleTree& t);                                                                      // merkle.nim:8-9
int64_t treeNumberOfLeaves(const MerkleTree& t);                                                         // :11-12
Hash treeRoot(const MerkleTree& t);                                                                      // :14-17
MerkleProof merkleProof(const MerkleTree& t, int64_t index);                                             // :21-42
using CompressWithKey = std::function<F(int, const F&, const F&)>;                                       // :46
Hash reconstructRoot(const CompressWithKey& c, const MerkleProof& p);                                    // :51-74
bool checkMerkleProof(const CompressWithKey& c, const Root& root, const MerkleProof& p);                 // :76-77
MerkleProof mergeMerkleProofs(const CompressWithKey& c, const MerkleProof& bot, const MerkleProof& top);  // :86-100

// ---- nim/blocks/bn254.nim -----------------------------------------------------------------------------------
MerkleTree merkleTree(Backend& be, const HashConfig& h, const std::vector<Hash>& what);                  // :17-19
Hash hashCell(Backend& be, const HashConfig& h, const GlobalConfig& g, const Cell& cellData);            // :23-29
Hash hashNetworkBlock(Backend& be, const HashConfig& h, const GlobalConfig& g, const Block& blockData);  // :48-55
MerkleTree networkBlockTree(Backend& be, const HashConfig& h, const GlobalConfig& g, const Block& blockData);   // :60-67

// ---- nim/sample/bn254.nim -----------------------------------------------------------------------------------
int64_t cellIndex(Backend& be, const HashConfig& h, const Entropy& e, const Root& slotRoot, int64_t numberOfCells, int counter);   // :16-24
std::vector<int64_t> cellIndices(Backend& be, const HashConfig& h, const Entropy& e, const Root& slotRoot, int64_t numberOfCells, int64_t nSamples);   // :26-27

// ---- nim/slot.nim, nim/dataset.nim --------------------------------------------------------------------------
Cell slotLoadCellData(Backend& be, const GlobalConfig& g, const SlotConfig& cfg, CellIdx idx);           // slot.nim:51-68
Block slotLoadBlockData(Backend& be, const GlobalConfig& g, const SlotConfig& cfg, BlockIdx idx);        // slot.nim:70-73
Seed parametricSlotSeed(Seed seed, SlotIdx k);                                                           // dataset.nim:32
SlotConfig slotCfgFromDataSetCfg(const DataSetConfig& d, SlotIdx idx);                                   // dataset.nim:45-51
Cell dataSetLoadCellData(Backend& be, const GlobalConfig& g, const DataSetConfig& d, SlotIdx s, CellIdx c);    // :55-57
Block dataSetLoadBlockData(Backend& be, const GlobalConfig& g, const DataSetConfig& d, SlotIdx s, BlockIdx b); // :59-61

// ---- nim/gen_input/bn254.nim, nim/json/bn254.nim -------------------------------------------------------------
SlotProofInput generateProofInputBN254(Backend& be, const HashConfig& h, const GlobalConfig& g, const DataSetConfig& d,
                                       SlotIdx slotIdx, const Entropy& entropy);                         // gen_input/bn254.nim:78-79
void exportProofInputBN254(const HashConfig& h, const std::string& fname, const SlotProofInput& prf);    // json/bn254.nim:77-79
std::string proofInputToJson(const SlotProofInput& prf);                                                 // the text exportProofInput writes (:57-74)

// ---- verifier side (what the circuit re-computes) ---------------------------------------
// Re-derives everything the SampleAndProve circuit constrains, on the GPU, in five batched launches: the dataset path
// (circuit/codex/sample_cells.circom:95-109), the sampled indices (:23-48,125-147), every cell hash from its 67 field
// elements, the block-level and then the slot-level root of every sample (circuit/codex/single_cell.circom:41-71 --
// two stages, each restarting with the bottom-layer key).  Returns true if the input would satisfy the circuit;
// otherwise false with the first failed check in *why.
bool checkProofInputBN254(Backend& be, const GlobalConfig& g, const SlotProofInput& prf, std::string* why = nullptr);

}  // namespace codex
