// cli.cpp -- command line of the B200 proof-input generator; flag surface and defaults of
// reference/nim/proof_input/src/cli.nim:37-237 (std/parseopt syntax: -k=v, -k:v, --key=v, --key:v).
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>

#include "proof_input.hpp"

using namespace codex;

struct FullConfig {            // cli.nim:37-45, defaults :47-76
  HashConfig hashCfg;
  GlobalConfig globCfg;
  DataSetConfig dsetCfg;
  int64_t slotIndex = 0;
  int64_t entropy = 1234567;
  std::string outFile, circomFile;
  bool verbose = false;
  bool selfcheck = false;
  int device = 0;
};

static void printHelp() {      // cli.nim:80-105
  std::puts("usage:");
  std::puts("$ ./cli [options] --output=proof_input.json --circom=proof_main.circom");
  std::puts("");
  std::puts("available options:");
  std::puts(" -h, --help                         : print this help");
  std::puts(" -v, --verbose                      : verbose output (print the actual parameters)");
  std::puts(" -d, --depth      = <maxdepth>      : maximum depth of the slot tree (eg. 32)");
  std::puts(" -N, --maxslots   = <maxslots>      : maximum number of slots (eg. 256)");
  std::puts(" -c, --cellsize   = <cellSize>      : cell size in bytes (eg. 2048)");
  std::puts(" -b, --blocksize  = <blockSize>     : block size in bytes (eg. 65536)");
  std::puts(" -s, --nslots     = <nslots>        : number of slots in the dataset (eg. 13)");
  std::puts(" -n, --nsamples   = <nsamples>      : number of samples we prove (eg. 100)");
  std::puts(" -e, --entropy    = <entropy>       : external randomness (eg. 1234567)");
  std::puts(" -S, --seed       = <seed>          : seed to generate the fake data (eg. 12345)");
  std::puts(" -f, --file       = <datafile>      : slot data file, base name (eg. \"slotdata\" would mean \"slotdata5.dat\" for slot index = 5)");
  std::puts(" -i, --index      = <slotIndex>     : index of the slot (within the dataset) we prove");
  std::puts(" -k, --log2ncells = <log2(ncells)>  : log2 of the number of cells inside this slot (eg. 10)");
  std::puts(" -K, --ncells     = <ncells>        : number of cells inside this slot (eg. 1024; must be a power of two)");
  std::puts(" -o, --output     = <input.json>    : the JSON file into which we write the proof input");
  std::puts(" -C, --circom     = <main.circom>   : the circom main component to create with these parameters");
  std::puts(" -F, --field      = <field>         : the underlying field: \"bn254\" or \"goldilocks\"");
  std::puts(" -H, --hash       = <hash>          : the hash function to use: \"poseidon2\" or \"monolith\"");
  std::puts(" -G, --gpu        = <device>        : CUDA device ordinal (this backend only; default 0)");
  std::puts("     --selfcheck                    : re-derive on the GPU everything the circuit constrains before writing (this backend only)");
  std::puts("");
  std::exit(0);
}

static int64_t parseInt(const std::string& v) {
  size_t pos = 0;
  long long x = 0;
  try { x = std::stoll(v, &pos); } catch (...) { pos = 0; }
  if (pos == 0 || pos != v.size()) throw AssertionDefect("invalid integer: " + v);
  return x;
}

static FullConfig parseCliOptions(int argc, char** argv) {   // cli.nim:109-162
  FullConfig cfg;
  for (int a = 1; a < argc; ++a) {
    std::string arg = argv[a];
    if (arg.empty() || arg[0] != '-') continue;               // positional arguments are ignored (cli.nim:122-124)
    const bool isLong = arg.size() > 1 && arg[1] == '-';
    std::string body = arg.substr(isLong ? 2 : 1), key = body, value;
    const size_t sep = body.find_first_of("=:");
    if (sep != std::string::npos) { key = body.substr(0, sep); value = body.substr(sep + 1); }
    else if (!isLong && body.size() > 1) { key = body.substr(0, 1); value = body.substr(1); }
    auto is = [&](const char* s, const char* l) { return key == s || key == l; };
    if (is("h", "help")) printHelp();
    else if (is("v", "verbose")) cfg.verbose = true;
    else if (is("d", "depth")) cfg.globCfg.maxDepth = (int)parseInt(value);
    else if (is("N", "maxslots")) cfg.globCfg.maxLog2NSlots = ceilingLog2(parseInt(value));
    else if (is("c", "cellsize")) cfg.globCfg.cellSize = checkPowerOfTwo(parseInt(value), "cellSize");
    else if (is("b", "blocksize")) cfg.globCfg.blockSize = checkPowerOfTwo(parseInt(value), "blockSize");
    else if (is("s", "nslots")) cfg.dsetCfg.nSlots = parseInt(value);
    else if (is("n", "nsamples")) cfg.dsetCfg.nSamples = parseInt(value);
    else if (is("e", "entropy")) cfg.entropy = parseInt(value);
    else if (is("S", "seed")) { cfg.dsetCfg.dataSrc = DataSource(); cfg.dsetCfg.dataSrc.kind = DataSourceKind::FakeData; cfg.dsetCfg.dataSrc.seed = (uint64_t)parseInt(value); }
    else if (is("f", "file")) { cfg.dsetCfg.dataSrc = DataSource(); cfg.dsetCfg.dataSrc.kind = DataSourceKind::SlotFile; cfg.dsetCfg.dataSrc.filename = value; }
    else if (is("i", "index")) cfg.slotIndex = parseInt(value);
    else if (is("k", "log2ncells")) cfg.dsetCfg.nCells = (int64_t)1 << parseInt(value);
    else if (is("K", "ncells")) cfg.dsetCfg.nCells = checkPowerOfTwo(parseInt(value), "nCells");
    else if (is("o", "output")) cfg.outFile = value;
    else if (is("C", "circom")) cfg.circomFile = value;
    else if (is("F", "field")) cfg.hashCfg.field = parseField(value);
    else if (is("H", "hash")) cfg.hashCfg.hashFun = parseHashFun(value);
    else if (is("G", "gpu")) cfg.device = (int)parseInt(value);
    else if (key == "selfcheck") cfg.selfcheck = true;
    else {
      std::cout << "Unknown option: " << key << "\nuse --help to get a list of options\n";
      std::exit(0);
    }
  }
  cfg.hashCfg.combo = toFieldHashCombo(cfg.hashCfg.field, cfg.hashCfg.hashFun);
  return cfg;
}

static void printConfig(const FullConfig& c) {   // cli.nim:166-182
  std::cout << "field      = " << (c.hashCfg.field == FieldSelect::BN254 ? "BN254" : "Goldilocks") << "\n";
  std::cout << "hash func. = " << (c.hashCfg.hashFun == HashSelect::Poseidon2 ? "Poseidon2" : "Monolith") << "\n";
  std::cout << "maxDepth   = " << c.globCfg.maxDepth << "\n";
  std::cout << "maxSlots   = " << ((int64_t)1 << c.globCfg.maxLog2NSlots) << "\n";
  std::cout << "cellSize   = " << c.globCfg.cellSize << "\n";
  std::cout << "blockSize  = " << c.globCfg.blockSize << "\n";
  std::cout << "nSamples   = " << c.dsetCfg.nSamples << "\n";
  std::cout << "entropy    = " << c.entropy << "\n";
  std::cout << "slotIndex  = " << c.slotIndex << "\n";
  std::cout << "nCells     = " << c.dsetCfg.nCells << "\n";
  if (c.dsetCfg.dataSrc.kind == DataSourceKind::FakeData) std::cout << "dataSource = (kind: FakeData, seed: " << c.dsetCfg.dataSrc.seed << ")\n";
  else std::cout << "dataSource = (kind: SlotFile, filename: \"" << c.dsetCfg.dataSrc.filename << "\")\n";
}

static void writeCircomMainComponent(const FullConfig& c, const std::string& fname) {   // cli.nim:186-204
  const int blockTreeDepth = exactLog2(c.globCfg.blockSize / c.globCfg.cellSize);
  const int64_t nFieldElemsPerCell = (c.globCfg.cellSize + 30) / 31;
  std::ofstream f(fname);
  f << "pragma circom 2.0.0;\n";
  f << "include \"sample_cells.circom\";\n";
  f << "// SampleAndProven( maxDepth, maxLog2NSlots, blockTreeDepth, nFieldElemsPerCell, nSamples )\n";
  f << "component main {public [entropy,dataSetRoot,slotIndex]} = SampleAndProve(" << c.globCfg.maxDepth << ", " << c.globCfg.maxLog2NSlots << ", "
    << blockTreeDepth << ", " << nFieldElemsPerCell << ", " << c.dsetCfg.nSamples << ");\n";
}

int main(int argc, char** argv) {   // cli.nim:208-237
  try {
    const FullConfig cfg = parseCliOptions(argc, argv);
    if (cfg.verbose) printConfig(cfg);
    if (cfg.circomFile.empty() && cfg.outFile.empty()) {
      std::cout << "nothing to do!\nuse --help for getting a list of options\n";
      return 0;
    }
    if (!cfg.circomFile.empty()) {
      std::cout << "writing circom main component into `" << cfg.circomFile << "`\n";
      writeCircomMainComponent(cfg, cfg.circomFile);
    }
    if (!cfg.outFile.empty()) {
      std::cout << "writing proof input into `" << cfg.outFile << "`...\n";
      if (cfg.hashCfg.field != FieldSelect::BN254)
        throw AssertionDefect("this backend implements --field=bn254 --hash=poseidon2 only (the reference's default field is goldilocks: pass --field=bn254)");
      Backend be(cfg.device);
      const Entropy entropy = intToBN254(cfg.entropy);
      const SlotProofInput prf = generateProofInputBN254(be, cfg.hashCfg, cfg.globCfg, cfg.dsetCfg, cfg.slotIndex, entropy);
      if (cfg.selfcheck) {
        std::string why;
        if (!checkProofInputBN254(be, cfg.globCfg, prf, &why)) throw AssertionDefect("selfcheck failed: " + why);
        std::cout << "selfcheck: the proof input satisfies every constraint the circuit re-computes\n";
      }
      exportProofInputBN254(cfg.hashCfg, cfg.outFile, prf);
    }
    std::cout << "done\n";
    return 0;
  } catch (const AssertionDefect& e) {
    std::cerr << "Error: unhandled exception: " << e.what() << " [AssertionDefect]\n";
    return 1;
  }
}
