// REJECTED EXPERIMENT (round 2), kept as evidence; not compiled into the product.
// Software-pipelined internal rounds of the Poseidon2 permutation: the y/z update of round r is moved to the top of the
// loop body of round r+1, where it is independent of the S-box that follows, so that ptxas can interleave its ~45 ALU
// instructions with the multiply stream; only x' + (y + z + c) -> + x' -> one table reduction stay between two S-boxes.
// Bit-exact (same checksums over 2^20 cell hashes, host emulation green), same 80 registers, 596 instead of 591
// instructions per round -- and 0.5 % SLOWER on a B200 (profiles/r2_sweep_swpipe.txt: 114.62 ms vs 113.99 ms per 2 GiB):
// with six warps per scheduler the ALU tail of one warp is already covered by the multiplies of the others, and the
// rotated body only lengthens the live ranges.  Needs c_rc[81] with a zero last entry and neg_2r() (2r - a) in fr.cuh.
#if 0
#if CDX_TABRED && CDX_SWPIPE
      // The 56 internal rounds, software-pipelined by one round.  Only x goes through the S-box, and the next S-box
      // needs nothing but x: the y and z updates of round r (two table reductions, ~45 ALU instructions) do not sit
      // between S-box r and S-box r+1 on the dependency chain.  In round order they would still be scheduled there --
      // they need s = x' + y + z, known only when S-box r ends, and ptxas does not move code across the loop edge -- so
      // the loop body is rotated: it starts with the PENDING y/z update of the previous round, which is independent of
      // the S-box that follows it in the same basic block and fills the issue slots of the multiply stream.  After the
      // S-box only x' + (y + z + c) -> + x' -> one table reduction remain before the next S-box can start; the state
      // carried around the loop is t = x + c (the next S-box input), y, z and s.
      //   bounds: t < B, S-box result < 1.27 r (first round: t < 2.083, < 1.64 r); y + s < 4.9 r; 2z + s and 2x' + y + z + c
      //   < 6.5 r use the 257-bit reduction.
      // First round: "pending" must be neutral: s = 2r - z and y := y + z give y + s = y + 2r and z + (z + s) = z + 2r.
      Fr t = add_lazy(x, CDX_RC(24));
      Fr s = neg_2r(z);
      y = add_lazy(y, z);
#pragma unroll 1
      for (int r = 0; r < 56; ++r) {
        y = reduce_tab(add_lazy(y, s), 0u);            // pending update of the previous round: y + s
        z = add_reduce(z, add_lazy(z, s));             //                                       2z + s
        const Fr u = add_lazy(y, z);
        const Fr uc = add_lazy(u, CDX_RC(25 + r));     // + the NEXT round's constant (entry 80 = 0 after the last round)
        const Fr xs = sbox(t);
        s = add_lazy(xs, u);                           // s = x' + y + z for the next pending update
        t = add_reduce(xs, add_lazy(xs, uc));          // next S-box input: 2x' + y + z + c
      }
      x = t;
      y = reduce_tab(add_lazy(y, s), 0u);
      z = add_reduce(z, add_lazy(z, s));
#endif
#endif
