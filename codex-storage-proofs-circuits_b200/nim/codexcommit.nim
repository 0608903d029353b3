## codexcommit.nim -- the Nim side of the B200 backend: F <-> 32-byte marshalling and one wrapper per call site of the
## external `poseidon2` package in reference/nim/proof_input/src, over the generated binding codexcommit_abi.nim (every
## exported function of libcodexcommit.so, include/codex_commit.h).
##
## Delivered as source: the build image has no Nim toolchain, so this file and the patch series in nim/patches/ were never
## compiled here; the C++ host layer (host/proof_input.cpp) is the executed twin -- same calls, same arguments, same
## order -- and it is what the tests run.  To use it, copy codexcommit.nim and codexcommit_abi.nim into
## reference/nim/proof_input/src/, apply nim/patches/*.patch from the repository root of the reference and build with
##   nimble build -d:release --passL:"-L<dir> -lcodexcommit -Wl,-rpath,<dir>"
## (or leave the dynlib pragma to dlopen libcodexcommit.so from LD_LIBRARY_PATH).
##
## Purity: the reference declares the hashing procs as `func`.  A GPU call is not a side effect the caller can observe
## (same inputs, same output, no Nim-visible state), so the wrappers are `func`s that cast away the effect of the importc
## call, and the patched modules keep every reference signature unchanged.

import std/sequtils
import constantine/math/arithmetic, constantine/math/io/io_bigints, constantine/math/io/io_fields
import poseidon2/types          # F
import ./codexcommit_abi
export codexcommit_abi

type Felt* = array[32, byte]    ## canonical little-endian field element, the library's wire format

var gCtx {.threadvar.}: CdxCtx    ## one context per host thread (cdx_ctx is not shared between threads)
var gGroup: CdxGroup              ## every visible GPU, created on first use by the dataset commit

proc ctx*(): CdxCtx =
  if pointer(gCtx) == nil:
    doAssert cdx_ctx_create(0, addr gCtx) == CDX_OK, "no CUDA device: this backend has no CPU path"
  gCtx

proc group*(): CdxGroup =
  if pointer(gGroup) == nil:
    doAssert cdx_group_create(nil, 0, addr gGroup) == CDX_OK, "cdx_group_create failed"
  gGroup

func toFelt*(x: F): Felt =
  discard result.marshal(x.toBig(), littleEndian)

func toF*(b: Felt): F =
  var big: BigInt[254]
  big.unmarshal(b, littleEndian)
  result.fromBig(big)

func toFelts*(xs: openArray[F]): seq[Felt] = xs.mapIt(it.toFelt)
func toFs*(xs: openArray[Felt]): seq[F] = xs.mapIt(it.toF)

template chk*(rc: cint) =
  if rc != CDX_OK: raiseAssert($cdx_last_error(ctx()) & " [" & $cdx_status_string(rc) & "]")

template pure(body: untyped): untyped =
  {.cast(noSideEffect).}:
    body

# ---- blocks/bn254.nim:27        Sponge.digest(cellData, rate=2)
func gpuHashBytes*(data: openArray[byte]): F =
  var o: Felt
  pure: chk cdx_hash_bytes_batch_host(ctx(), (if data.len > 0: unsafeAddr data[0] else: nil), 1, csize_t(data.len), addr o[0])
  o.toF

# ---- blocks/bn254.nim:52,63     the hashCell loop over the cells of one block, as one launch
func gpuHashCells*(blockData: openArray[byte], cellSize: int): seq[F] =
  let n = blockData.len div cellSize
  var o = newSeq[Felt](n)
  pure: chk cdx_hash_bytes_batch_host(ctx(), unsafeAddr blockData[0], csize_t(n), csize_t(cellSize), addr o[0][0])
  o.toFs

# ---- sample/bn254.nim:23        Sponge.digest(@[entropy, slotRoot, toF(counter)], rate=2)
func gpuSpongeFelts*(xs: openArray[F], rate = 2): F =
  var inp = xs.toFelts
  var o: Felt
  pure: chk cdx_sponge_felts_batch_host(ctx(), (if inp.len > 0: addr inp[0][0] else: nil), 1, csize_t(inp.len), cint(rate), addr o[0])
  o.toF

# ---- sample/bn254.nim:26-27     all counters 1..nSamples in one launch
func gpuCellIndices*(entropy, slotRoot: F, numberOfCells, nSamples: int): seq[int] =
  var e = entropy.toFelt
  var r = slotRoot.toFelt
  var idx = newSeq[uint64](nSamples)
  if nSamples > 0:
    pure: chk cdx_cell_indices(ctx(), addr e[0], addr r[0], uint64(numberOfCells), csize_t(nSamples), addr idx[0])
  idx.mapIt(int(it))

# ---- merkle/bn254.nim:18        compress(x, y, key = toF(key))
func gpuCompress*(key: int, x, y: F): F =
  var fx = x.toFelt
  var fy = y.toFelt
  var k = uint32(key)
  var o: Felt
  pure: chk cdx_compress_batch_host(ctx(), addr fx[0], addr fy[0], addr k, 1, addr o[0])
  o.toF

# ---- merkle/bn254.nim:20        Merkle.digest(xs)
func gpuMerkleDigest*(xs: openArray[F]): F =
  var inp = xs.toFelts
  var o: Felt
  pure: chk cdx_merkle_root_host(ctx(), addr inp[0][0], csize_t(inp.len), addr o[0])
  o.toF

# ---- merkle/bn254.nim:29-63     merkleTreeWorker: every layer, bottom first
func gpuMerkleLayers*(xs: openArray[F]): seq[seq[F]] =
  var inp = xs.toFelts
  let n = inp.len
  var flat = newSeq[Felt](int(cdx_merkle_total_nodes(csize_t(n), 1)))
  pure: chk cdx_merkle_layers_host(ctx(), addr inp[0][0], csize_t(n), 1, addr flat[0][0])
  var off = 0
  var m = n
  for _ in 0 ..< int(cdx_merkle_num_layers(csize_t(n), 1)):
    result.add flat[off ..< off + m].toFs
    off += m
    m = (m + 1) div 2

# ---- testvectors.nim:60-66      Merkle.digest(openArray[byte])
func gpuMerkleDigestBytes*(data: openArray[byte]): F =
  var o: Felt
  pure: chk cdx_merkle_root_bytes_host(ctx(), (if data.len > 0: unsafeAddr data[0] else: nil), csize_t(data.len), addr o[0])
  o.toF

# ---- slot.nim:23-32,51-55       genFakeCell (optional: the CPU generator of the reference is kept by the patches)
proc gpuFakeCells*(seed: uint64, firstCell, nCells, cellSize: int): seq[byte] =
  result = newSeq[byte](nCells * cellSize)
  chk cdx_fake_cells_host(ctx(), seed, uint64(firstCell), csize_t(nCells), csize_t(cellSize), addr result[0])
