## codexcommit.nim -- Nim binding of libcodexcommit.so (include/codex_commit.h).
##
## Delivered as source: the build image has no Nim toolchain, so this file is exercised only through the
## equivalent C++ host layer (host/proof_input.cpp) and the ctypes binding (capi.py), which call the same symbols
## with the same arguments.  Drop it next to reference/nim/proof_input/src/ and link with
##   --passL:"-L<dir> -lcodexcommit -Wl,-rpath,<dir>"
{.push callconv: cdecl, dynlib: "libcodexcommit.so".}

type
  CdxCtx*  = distinct pointer
  CdxSlot* = distinct pointer
  Felt*    = array[32, byte]        ## canonical little-endian field element

proc cdx_ctx_create*(device: cint, ctx: ptr CdxCtx): cint {.importc.}
proc cdx_ctx_destroy*(ctx: CdxCtx) {.importc.}
proc cdx_last_error*(ctx: CdxCtx): cstring {.importc.}

proc cdx_sponge_felts_batch_host*(ctx: CdxCtx, elems: ptr byte, nItems, len: csize_t, rate: cint, outp: ptr byte): cint {.importc.}
proc cdx_hash_bytes_batch_host*(ctx: CdxCtx, data: ptr byte, nItems, len: csize_t, outp: ptr byte): cint {.importc.}
proc cdx_compress_batch_host*(ctx: CdxCtx, x, y: ptr byte, keys: ptr uint32, n: csize_t, outp: ptr byte): cint {.importc.}
proc cdx_merkle_total_nodes*(n: csize_t, bottomLayer: cint): csize_t {.importc.}
proc cdx_merkle_num_layers*(n: csize_t, bottomLayer: cint): cint {.importc.}
proc cdx_merkle_layers_host*(ctx: CdxCtx, leaves: ptr byte, n: csize_t, bottomLayer: cint, layersOut: ptr byte): cint {.importc.}
proc cdx_merkle_root_host*(ctx: CdxCtx, leaves: ptr byte, n: csize_t, rootOut: ptr byte): cint {.importc.}

proc cdx_slot_commit_host*(ctx: CdxCtx, data: ptr byte, nBytes, cellSize, blockSize: csize_t, slot: ptr CdxSlot): cint {.importc.}
proc cdx_slot_commit_fake*(ctx: CdxCtx, seed: uint64, nCells, cellSize, blockSize: csize_t, slot: ptr CdxSlot): cint {.importc.}
proc cdx_slot_free*(slot: CdxSlot) {.importc.}
proc cdx_slot_root*(slot: CdxSlot, rootOut: ptr byte): cint {.importc.}
proc cdx_slot_shape*(slot: CdxSlot, nCells, nBlocks: ptr uint64, blockDepth, slotDepth: ptr uint32): cint {.importc.}
proc cdx_slot_read_layer*(slot: CdxSlot, tree: cint, level: uint32, first, count: uint64, outp: ptr byte): cint {.importc.}
proc cdx_slot_cell_paths*(slot: CdxSlot, cellIndices: ptr uint64, nSamples, maxDepth: csize_t, outp, leafOut: ptr byte): cint {.importc.}
proc cdx_slot_prove_batch*(slot: CdxSlot, entropies: ptr byte, nChallenges, nSamples, maxDepth: csize_t, indicesOut: ptr uint64, pathsOut, leavesOut: ptr byte): cint {.importc.}
proc cdx_cell_indices*(ctx: CdxCtx, entropy, slotRoot: ptr byte, nCells: uint64, nSamples: csize_t, indices: ptr uint64): cint {.importc.}
proc cdx_fake_cells_host*(ctx: CdxCtx, seed, firstCell: uint64, nCells, cellSize: csize_t, outp: ptr byte): cint {.importc.}
{.pop.}

# ---------------------------------------------------------------------------------------------------------------
# Drop-in replacements for the call sites of the external `poseidon2` package in reference/nim/proof_input/src.
# F <-> Felt marshalling uses constantine's little-endian (un)marshal, the inverse of what the library emits.

import constantine/math/arithmetic, constantine/math/io/io_fields, constantine/serialization/codecs
import poseidon2/types          # only for the type F

var gCtx: CdxCtx
proc ctx(): CdxCtx =
  if pointer(gCtx) == nil:
    doAssert cdx_ctx_create(0, addr gCtx) == 0, "no CUDA device: this backend has no CPU path"
  gCtx

proc toFelt*(x: F): Felt = discard result.marshal(x.toBig(), littleEndian)
proc toF*(b: Felt): F = (var big: BigInt[254]; big.unmarshal(b, littleEndian); result.fromBig(big))

template chk(rc: cint) = doAssert rc == 0, $cdx_last_error(ctx())

# blocks/bn254.nim:27      Sponge.digest(cellData, rate=2)
proc spongeDigestBytes*(data: openArray[byte]): F =
  var o: Felt
  chk cdx_hash_bytes_batch_host(ctx(), unsafeAddr data[0], 1, csize_t(data.len), addr o[0])
  o.toF

# sample/bn254.nim:23      Sponge.digest(@[entropy, slotRoot, toF(counter)], rate=2)
proc spongeDigestFelts*(xs: openArray[F], rate = 2): F =
  var inp = newSeq[Felt](xs.len)
  for i, x in xs: inp[i] = x.toFelt
  var o: Felt
  chk cdx_sponge_felts_batch_host(ctx(), cast[ptr byte](addr inp[0]), 1, csize_t(xs.len), cint(rate), addr o[0])
  o.toF

# merkle/bn254.nim:18      compress(x, y, key = toF(key))
proc compressWithkey*(key: int, x, y: F): F =
  var fx = x.toFelt; var fy = y.toFelt; var k = uint32(key); var o: Felt
  chk cdx_compress_batch_host(ctx(), addr fx[0], addr fy[0], addr k, 1, addr o[0])
  o.toF

# merkle/bn254.nim:20,62   Merkle.digest(xs) / merkleTreeBN254(xs)
proc merkleDigestBN254*(xs: openArray[F]): F =
  var inp = newSeq[Felt](xs.len)
  for i, x in xs: inp[i] = x.toFelt
  var o: Felt
  chk cdx_merkle_root_host(ctx(), cast[ptr byte](addr inp[0]), csize_t(xs.len), addr o[0])
  o.toF

# gen_input/bn254.nim:21-33  buildSlotTreeFull as one device commitment; the handle replaces (miniTrees, bigTree):
#   treeRoot(bigTree)                         -> cdx_slot_root
#   merkleProof(blockTree, i) & merkleProof(bigTree, b), merged and padded (merkle.nim:21-100, types.nim:27-37)
#                                             -> cdx_slot_cell_paths (all samples in one call)
#   cellIndices (sample/bn254.nim:26)         -> cdx_cell_indices
