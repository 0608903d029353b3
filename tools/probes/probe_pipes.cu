// probe_pipes.cu -- issue rates of the instruction forms a 256-bit modular multiply can be built from (B200, sm_100a).
// Every kernel's SASS was checked with cuobjdump: ptxas hoists or re-associates anything it can (an earlier version of
// this probe "measured" IMAD.WIDE at full rate while the loop had been rewritten into IADD3 chains), so the multiply
// probes feed each product back into the next multiplicand, and the carry-chain probes use the PTX carry flag.
// Result on B200 (profiles/r1_probe_pipes.txt): 32-bit IMAD 61-63 /clk/SM; IMAD.HI, IMAD.WIDE (with or without a
// carry) 25-32 /clk/SM -- a 32x32->64 product costs two FMA-heavy slots whatever form it takes; IADD3(.X) 62 /clk/SM
// on the ALU pipe; DFMA 59 /clk/SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 2048
#define UNROLL 8

// (lo,hi) = a*b ; a = lo ^ hi       -> IMAD.WIDE.U32 Rd, Ra, Rb, RZ  + LOP3
__global__ void k_mulwide_xor(uint32_t seed, uint32_t* sink) {
  uint32_t b = seed * 2654435761u + threadIdx.x * 40503u + blockIdx.x;
  uint32_t a0 = b ^ 1, a1 = b ^ 2, a2 = b ^ 3, a3 = b ^ 4, a4 = b ^ 5, a5 = b ^ 6, a6 = b ^ 7, a7 = b ^ 8;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      uint64_t p0 = (uint64_t)a0 * b, p1 = (uint64_t)a1 * b, p2 = (uint64_t)a2 * b, p3 = (uint64_t)a3 * b;
      uint64_t p4 = (uint64_t)a4 * b, p5 = (uint64_t)a5 * b, p6 = (uint64_t)a6 * b, p7 = (uint64_t)a7 * b;
      a0 = (uint32_t)p0 ^ (uint32_t)(p0 >> 32); a1 = (uint32_t)p1 ^ (uint32_t)(p1 >> 32); a2 = (uint32_t)p2 ^ (uint32_t)(p2 >> 32); a3 = (uint32_t)p3 ^ (uint32_t)(p3 >> 32);
      a4 = (uint32_t)p4 ^ (uint32_t)(p4 >> 32); a5 = (uint32_t)p5 ^ (uint32_t)(p5 >> 32); a6 = (uint32_t)p6 ^ (uint32_t)(p6 >> 32); a7 = (uint32_t)p7 ^ (uint32_t)(p7 >> 32);
    }
  }
  uint32_t acc = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
  if (acc == 0x12345678u) sink[0] = acc;
}
// lo = a*b (32-bit) ; a = lo ^ c     -> IMAD + LOP3   (control: same shape with the 1-slot multiply)
__global__ void k_mullo_xor(uint32_t seed, uint32_t* sink) {
  uint32_t b = seed * 2654435761u + threadIdx.x * 40503u + blockIdx.x;
  uint32_t a0 = b ^ 1, a1 = b ^ 2, a2 = b ^ 3, a3 = b ^ 4, a4 = b ^ 5, a5 = b ^ 6, a6 = b ^ 7, a7 = b ^ 8;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      a0 = (a0 * b) ^ seed; a1 = (a1 * b) ^ seed; a2 = (a2 * b) ^ seed; a3 = (a3 * b) ^ seed;
      a4 = (a4 * b) ^ seed; a5 = (a5 * b) ^ seed; a6 = (a6 * b) ^ seed; a7 = (a7 * b) ^ seed;
    }
  }
  uint32_t acc = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
  if (acc == 0x12345678u) sink[0] = acc;
}
// hi = mulhi(a,b) ; a = hi ^ c       -> IMAD.HI + LOP3
__global__ void k_mulhi_xor(uint32_t seed, uint32_t* sink) {
  uint32_t b = seed * 2654435761u + threadIdx.x * 40503u + blockIdx.x;
  uint32_t a0 = b ^ 1, a1 = b ^ 2, a2 = b ^ 3, a3 = b ^ 4, a4 = b ^ 5, a5 = b ^ 6, a6 = b ^ 7, a7 = b ^ 8;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      a0 = __umulhi(a0, b) ^ seed; a1 = __umulhi(a1, b) ^ seed; a2 = __umulhi(a2, b) ^ seed; a3 = __umulhi(a3, b) ^ seed;
      a4 = __umulhi(a4, b) ^ seed; a5 = __umulhi(a5, b) ^ seed; a6 = __umulhi(a6, b) ^ seed; a7 = __umulhi(a7, b) ^ seed;
    }
  }
  uint32_t acc = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
  if (acc == 0x12345678u) sink[0] = acc;
}
// 64-bit accumulate kept alive: acc64 += a*b with a = low word of acc (ptxas splits this into IMAD.WIDE RZ + IADD3 pair)
__global__ void k_madwide_acc(uint32_t seed, uint32_t* sink) {
  uint32_t b = seed * 2654435761u + threadIdx.x * 40503u + blockIdx.x;
  uint64_t c0 = b ^ 1, c1 = b ^ 2, c2 = b ^ 3, c3 = b ^ 4, c4 = b ^ 5, c5 = b ^ 6, c6 = b ^ 7, c7 = b ^ 8;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      c0 += (uint64_t)(uint32_t)c0 * b; c1 += (uint64_t)(uint32_t)c1 * b; c2 += (uint64_t)(uint32_t)c2 * b; c3 += (uint64_t)(uint32_t)c3 * b;
      c4 += (uint64_t)(uint32_t)c4 * b; c5 += (uint64_t)(uint32_t)c5 * b; c6 += (uint64_t)(uint32_t)c6 * b; c7 += (uint64_t)(uint32_t)c7 * b;
    }
  }
  uint64_t acc = c0 ^ c1 ^ c2 ^ c3 ^ c4 ^ c5 ^ c6 ^ c7;
  if ((uint32_t)acc == 0x12345678u && (acc >> 32) == 1) sink[0] = 1;
}

#define DECL uint32_t a = seed + threadIdx.x, b = seed * 3u + blockIdx.x; \
  uint32_t r0 = a, r1 = b, r2 = a ^ b, r3 = a + b, r4 = a * 3, r5 = b * 5, r6 = a * 7, r7 = b * 9; \
  uint32_t s0 = 1, s1 = 2, s2 = 3, s3 = 4, s4 = 5, s5 = 6, s6 = 7, s7 = 8;
#define SINK uint32_t acc = r0 ^ r1 ^ r2 ^ r3 ^ r4 ^ r5 ^ r6 ^ r7 ^ s0 ^ s1 ^ s2 ^ s3 ^ s4 ^ s5 ^ s6 ^ s7; if (acc == 0x12345678u) sink[0] = acc;
#define REGS16 "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(r6), "+r"(r7), "+r"(s0), "+r"(s1), "+r"(s2), "+r"(s3), "+r"(s4), "+r"(s5), "+r"(s6), "+r"(s7)

// 1: IMAD.WIDE.U32.X carry chains of 4 (2 chains per asm = 8 wide ops)
__global__ void k_wide_x_chain4(uint32_t seed, uint32_t* sink) { DECL
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int u = 0; u < UNROLL; ++u)
      asm volatile(
        "mad.lo.cc.u32 %0,%16,%17,%0; madc.hi.cc.u32 %8,%16,%17,%8; madc.lo.cc.u32 %1,%16,%17,%1; madc.hi.cc.u32 %9,%16,%17,%9;\n\t"
        "madc.lo.cc.u32 %2,%16,%17,%2; madc.hi.cc.u32 %10,%16,%17,%10; madc.lo.cc.u32 %3,%16,%17,%3; madc.hi.u32 %11,%16,%17,%11;\n\t"
        "mad.lo.cc.u32 %4,%16,%17,%4; madc.hi.cc.u32 %12,%16,%17,%12; madc.lo.cc.u32 %5,%16,%17,%5; madc.hi.cc.u32 %13,%16,%17,%13;\n\t"
        "madc.lo.cc.u32 %6,%16,%17,%6; madc.hi.cc.u32 %14,%16,%17,%14; madc.lo.cc.u32 %7,%16,%17,%7; madc.hi.u32 %15,%16,%17,%15;"
        : REGS16 : "r"(a), "r"(b));
  } SINK }

// 7: pure ALU: IADD3 with carry chains (add.cc/addc), 16 per asm
__global__ void k_alu_carry(uint32_t seed, uint32_t* sink) { DECL
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int u = 0; u < UNROLL; ++u)
      asm volatile("add.cc.u32 %0,%0,%16; addc.cc.u32 %1,%1,%17; addc.cc.u32 %2,%2,%16; addc.cc.u32 %3,%3,%17; addc.cc.u32 %4,%4,%16; addc.cc.u32 %5,%5,%17; addc.cc.u32 %6,%6,%16; addc.u32 %7,%7,%17;\n\t"
                   "add.cc.u32 %8,%8,%16; addc.cc.u32 %9,%9,%17; addc.cc.u32 %10,%10,%16; addc.cc.u32 %11,%11,%17; addc.cc.u32 %12,%12,%16; addc.cc.u32 %13,%13,%17; addc.cc.u32 %14,%14,%16; addc.u32 %15,%15,%17;"
        : REGS16 : "r"(a), "r"(b));
  } SINK }

// 8: DFMA
__global__ void k_dfma(uint32_t seed, uint32_t* sink) {
  double x = 1.0 + 1e-9 * (seed + threadIdx.x), y = 1.0 - 1e-9 * blockIdx.x;
  double d0 = x, d1 = y, d2 = x + 1, d3 = y + 1, d4 = x + 2, d5 = y + 2, d6 = x + 3, d7 = y + 3;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int u = 0; u < UNROLL; ++u)
      asm volatile("fma.rn.f64 %0,%0,%8,%9; fma.rn.f64 %1,%1,%8,%9; fma.rn.f64 %2,%2,%8,%9; fma.rn.f64 %3,%3,%8,%9;\n\t"
                   "fma.rn.f64 %4,%4,%8,%9; fma.rn.f64 %5,%5,%8,%9; fma.rn.f64 %6,%6,%8,%9; fma.rn.f64 %7,%7,%8,%9;"
        : "+d"(d0), "+d"(d1), "+d"(d2), "+d"(d3), "+d"(d4), "+d"(d5), "+d"(d6), "+d"(d7) : "d"(x), "d"(y));
  }
  double acc = d0 + d1 + d2 + d3 + d4 + d5 + d6 + d7;
  if (acc == 0.123) sink[0] = 1;
}


// 9: co-issue, same thread: 16 IMAD.WIDE-class ops (two X chains of 4 wide pairs) + n_dfma DFMAs per unroll step
template <int NDFMA>
__global__ void k_mix_thread(uint32_t seed, uint32_t* sink) { DECL
  double x = 1.0 + 1e-9 * (seed + threadIdx.x), y = 1.0 - 1e-9 * blockIdx.x;
  double d0 = x, d1 = y, d2 = x + 1, d3 = y + 1, d4 = x + 2, d5 = y + 2, d6 = x + 3, d7 = y + 3;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      asm volatile(
        "mad.lo.cc.u32 %0,%16,%17,%0; madc.hi.cc.u32 %8,%16,%17,%8; madc.lo.cc.u32 %1,%16,%17,%1; madc.hi.cc.u32 %9,%16,%17,%9;\n\t"
        "madc.lo.cc.u32 %2,%16,%17,%2; madc.hi.cc.u32 %10,%16,%17,%10; madc.lo.cc.u32 %3,%16,%17,%3; madc.hi.u32 %11,%16,%17,%11;\n\t"
        "mad.lo.cc.u32 %4,%16,%17,%4; madc.hi.cc.u32 %12,%16,%17,%12; madc.lo.cc.u32 %5,%16,%17,%5; madc.hi.cc.u32 %13,%16,%17,%13;\n\t"
        "madc.lo.cc.u32 %6,%16,%17,%6; madc.hi.cc.u32 %14,%16,%17,%14; madc.lo.cc.u32 %7,%16,%17,%7; madc.hi.u32 %15,%16,%17,%15;"
        : REGS16 : "r"(a), "r"(b));
#pragma unroll
      for (int k = 0; k < NDFMA / 8; ++k)
        asm volatile("fma.rn.f64 %0,%0,%8,%9; fma.rn.f64 %1,%1,%8,%9; fma.rn.f64 %2,%2,%8,%9; fma.rn.f64 %3,%3,%8,%9;\n\t"
                     "fma.rn.f64 %4,%4,%8,%9; fma.rn.f64 %5,%5,%8,%9; fma.rn.f64 %6,%6,%8,%9; fma.rn.f64 %7,%7,%8,%9;"
          : "+d"(d0), "+d"(d1), "+d"(d2), "+d"(d3), "+d"(d4), "+d"(d5), "+d"(d6), "+d"(d7) : "d"(x), "d"(y));
    }
  }
  double dacc = d0 + d1 + d2 + d3 + d4 + d5 + d6 + d7;
  if (dacc == 0.123) sink[1] = 1;
  SINK }

// 10: co-issue, different warps of one block: even warps run the IMAD.WIDE.X chains, odd warps run DFMA chains
__global__ void k_mix_warps(uint32_t seed, uint32_t* sink, int imad_iters, int dfma_iters) {
  if (((threadIdx.x >> 5) & 1) == 0) { DECL
    for (int it = 0; it < imad_iters; ++it) {
#pragma unroll
      for (int u = 0; u < UNROLL; ++u)
        asm volatile(
          "mad.lo.cc.u32 %0,%16,%17,%0; madc.hi.cc.u32 %8,%16,%17,%8; madc.lo.cc.u32 %1,%16,%17,%1; madc.hi.cc.u32 %9,%16,%17,%9;\n\t"
          "madc.lo.cc.u32 %2,%16,%17,%2; madc.hi.cc.u32 %10,%16,%17,%10; madc.lo.cc.u32 %3,%16,%17,%3; madc.hi.u32 %11,%16,%17,%11;\n\t"
          "mad.lo.cc.u32 %4,%16,%17,%4; madc.hi.cc.u32 %12,%16,%17,%12; madc.lo.cc.u32 %5,%16,%17,%5; madc.hi.cc.u32 %13,%16,%17,%13;\n\t"
          "madc.lo.cc.u32 %6,%16,%17,%6; madc.hi.cc.u32 %14,%16,%17,%14; madc.lo.cc.u32 %7,%16,%17,%7; madc.hi.u32 %15,%16,%17,%15;"
          : REGS16 : "r"(a), "r"(b));
    } SINK
  } else {
    double x = 1.0 + 1e-9 * (seed + threadIdx.x), y = 1.0 - 1e-9 * blockIdx.x;
    double d0 = x, d1 = y, d2 = x + 1, d3 = y + 1, d4 = x + 2, d5 = y + 2, d6 = x + 3, d7 = y + 3;
    for (int it = 0; it < dfma_iters; ++it) {
#pragma unroll
      for (int u = 0; u < UNROLL; ++u)
        asm volatile("fma.rn.f64 %0,%0,%8,%9; fma.rn.f64 %1,%1,%8,%9; fma.rn.f64 %2,%2,%8,%9; fma.rn.f64 %3,%3,%8,%9;\n\t"
                     "fma.rn.f64 %4,%4,%8,%9; fma.rn.f64 %5,%5,%8,%9; fma.rn.f64 %6,%6,%8,%9; fma.rn.f64 %7,%7,%8,%9;"
          : "+d"(d0), "+d"(d1), "+d"(d2), "+d"(d3), "+d"(d4), "+d"(d5), "+d"(d6), "+d"(d7) : "d"(x), "d"(y));
    }
    double dacc = d0 + d1 + d2 + d3 + d4 + d5 + d6 + d7;
    if (dacc == 0.123) sink[1] = 1;
  }
}

static void run_mix_warps(int sm, int imad_iters, int dfma_iters) {
  uint32_t* sink; cudaMalloc(&sink, 8);
  const int blocks = sm * 8, threads = 256;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0); k_mix_warps<<<blocks, threads>>>(12345u + rep, sink, imad_iters, dfma_iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep > 0 && ms < best) best = ms;
  }
  double half = (double)blocks * threads / 2 * UNROLL;
  double wide = half * imad_iters * 8, dfma = half * dfma_iters * 8;   // 8 wide products (16 IMAD.WIDE-class ops) / 8 DFMA per unroll step
  printf("warps split: imad_iters %5d dfma_iters %5d  %8.3f ms   wide products %6.2f /clk/SM   DFMA %6.2f /clk/SM\n", imad_iters, dfma_iters, best,
         wide / (best * 1e-3) / (sm * 1.965e9), dfma / (best * 1e-3) / (sm * 1.965e9));
  cudaFree(sink);
}
template <class K> static void run_mix_thread(const char* name, K kernel, int sm, int ndfma) {
  uint32_t* sink; cudaMalloc(&sink, 8);
  const int blocks = sm * 8, threads = 256;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0); kernel<<<blocks, threads>>>(12345u + rep, sink); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep > 0 && ms < best) best = ms;
  }
  double steps = (double)blocks * threads * ITERS * UNROLL;
  printf("%-40s %8.3f ms   wide products %6.2f /clk/SM   DFMA %6.2f /clk/SM\n", name, best, steps * 8 / (best * 1e-3) / (sm * 1.965e9), steps * ndfma / (best * 1e-3) / (sm * 1.965e9));
  cudaFree(sink);
}

template <class K> static void run(const char* name, K kernel, int sm, double per_unroll = 8) {
  uint32_t* sink; cudaMalloc(&sink, 4);
  const int blocks = sm * 8, threads = 256;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0); kernel<<<blocks, threads>>>(12345u + rep, sink); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep > 0 && ms < best) best = ms;
  }
  double mults = (double)blocks * threads * ITERS * UNROLL * per_unroll;
  printf("%-46s %8.3f ms  %7.3f T mult/s  (%5.1f multiplies/clk/SM at 1.965 GHz)\n", name, best, mults / (best * 1e-3) / 1e12, mults / (best * 1e-3) / (sm * 1.965e9));
  cudaFree(sink);
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); int sm = p.multiProcessorCount;
  printf("%s, %d SMs\n", p.name, sm);
  run("IMAD (lo)         + LOP3", k_mullo_xor, sm);
  run("IMAD.HI           + LOP3", k_mulhi_xor, sm);
  run("IMAD.WIDE Rd,a,b,RZ + LOP3", k_mulwide_xor, sm);
  run("acc64 += a*b  (IMAD.WIDE RZ + IADD3 + IADD3.X)", k_madwide_acc, sm);
  run("IMAD.WIDE.U32.X carry chains of 4", k_wide_x_chain4, sm);
  run("IADD3 / IADD3.X carry chains (ALU pipe)", k_alu_carry, sm, 16);
  run("DFMA (fp64 pipe)", k_dfma, sm);
  printf("-- co-issue of the FMA-heavy integer pipe and the fp64 pipe --\n");
  run_mix_thread("same thread: 8 wide.X pairs + 0 DFMA", k_mix_thread<0>, sm, 0);
  run_mix_thread("same thread: 8 wide.X pairs + 8 DFMA", k_mix_thread<8>, sm, 8);
  run_mix_thread("same thread: 8 wide.X pairs + 16 DFMA", k_mix_thread<16>, sm, 16);
  run_mix_thread("same thread: 8 wide.X pairs + 24 DFMA", k_mix_thread<24>, sm, 24);
  run_mix_warps(sm, ITERS, 0);
  run_mix_warps(sm, 0, ITERS);
  run_mix_warps(sm, 0, 2 * ITERS);
  run_mix_warps(sm, ITERS, ITERS);
  run_mix_warps(sm, ITERS, 2 * ITERS);
  run_mix_warps(sm, ITERS, 3 * ITERS);
  return 0;
}
