// Micro-benchmark of the on-chip ceilings that bound the packed LDPC decoder ("an integer-ALU / shared-memory ceiling measured
// by a micro-benchmark on the same GPU").  One CTA per SM, 768 threads (the decoder's geometry: 6 warps per scheduler), every thread runs a
// long unrolled body of independent dependency chains built from inline PTX so that ptxas cannot fold or reorder the mix away:
//   lop3     only LOP3 (three register operands)                          -> the ALU pipe's issue rate
//   alu      LOP3 / PRMT / IADD3 / SHF / VABSDIFF4 in the decoder's ratio  -> same pipe, mixed opcodes
//   mix      the decoder's measured opcode mix (profiles/r01n_*): ALU ops + IMAD + IDP.4A + LDS + STS in its proportions, no barriers, no
//            branches: what the SM sustains on this instruction blend when nothing but the pipes limits it
// Prints one JSON object: warp instructions per cycle per scheduler for each body and the ALU-pipe share of the mix.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/_bin/alu_ceiling tools/ubench/alu_ceiling.cu   (see __graft_entry__.build)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define LOP3(d, a, b, c) asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(a), "r"(b), "r"(c))
#define LOPS(d, a, b, c) asm volatile("lop3.b32 %0, %1, %2, %3, 0xCA;" : "=r"(d) : "r"(a), "r"(b), "r"(c))
#define PRMT(d, a) asm volatile("prmt.b32 %0, %1, %2, 0xba98;" : "=r"(d) : "r"(a), "r"(0u))
#define ADD3(d, a, b) asm volatile("{ .reg .u32 t; add.u32 t, %1, 0x80808080; sub.u32 %0, t, %2; }" : "=r"(d) : "r"(a), "r"(b))
#define SHF(d, a, b, c) asm volatile("shf.r.wrap.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c))
#define VABS(d, a, b) asm volatile("vabsdiff4.u32.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(0u))
#define IMAD(d, a, b, c) asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c))
#define IDP(d, a, b, c) asm volatile("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c))
#define LDS(d, addr) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(d) : "r"(addr))
#define STS(addr, v) asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v))

constexpr int kChains = 8;

template <int MODE, int UNROLL>
__global__ void __launch_bounds__(768, 1) body(uint32_t *out, int iters, uint32_t one)
{
  extern __shared__ uint32_t sm[];
  uint32_t x[kChains], y[kChains];
  for (int c = 0; c < kChains; c++) { x[c] = threadIdx.x * 2654435761u + c; y[c] = x[c] ^ 0x5bd1e995u; }
  for (int i = threadIdx.x; i < 12288; i += blockDim.x) sm[i] = i * 747796405u;
  __syncthreads();
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm) + 4u * threadIdx.x;
#pragma unroll UNROLL
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int c = 0; c < kChains; c++) {
      if (MODE == 0) {          // 8 LOP3
        LOP3(x[c], x[c], y[c], one); LOP3(y[c], y[c], x[c], one); LOP3(x[c], x[c], y[c], one); LOP3(y[c], y[c], x[c], one);
        LOP3(x[c], x[c], y[c], one); LOP3(y[c], y[c], x[c], one); LOP3(x[c], x[c], y[c], one); LOP3(y[c], y[c], x[c], one);
      } else if (MODE == 1) {   // ALU only, decoder ratio: 13 LOP3 : 6 PRMT : 2.5 IADD3 : 2.4 SHF : 0.9 VABSDIFF4  ->  per 16: 9 LOP3, 4 PRMT, 1 IADD, 1 SHF, 1 VABS
        uint32_t t, u;
        VABS(t, x[c], y[c]); PRMT(u, t); LOPS(t, u, t, one); LOP3(u, x[c], y[c], t); LOPS(x[c], u, x[c], y[c]); ADD3(t, x[c], y[c]);
        PRMT(u, t); LOPS(y[c], u, y[c], x[c]); LOPS(t, u, x[c], y[c]); SHF(u, t, y[c], one); PRMT(t, u); LOP3(x[c], t, u, x[c]);
        LOPS(y[c], t, x[c], y[c]); PRMT(u, y[c]); LOP3(x[c], u, x[c], one); LOP3(y[c], y[c], x[c], u);
      } else {                  // decoder mix per 30: 13 ALU-logic/permute (7 LOP3 3 PRMT 1 IADD 1 SHF 1 VABS) + 2 more LOP3 = 15 ALU, 5 IMAD, 2 IDP, 4 LDS, 1 STS, (3 left to ISETP/BRA: not modelled)
        uint32_t t, u, v, w;
        const uint32_t a0 = base + ((x[c] & 0x7Fu) << 7);
        LDS(t, a0); LDS(u, a0 + 4); SHF(v, t, u, one); LDS(w, a0 + 512);
        VABS(t, v, w); PRMT(u, t); LOPS(t, u, t, one); IMAD(u, w, one, v); LOP3(u, v, w, u); LOPS(x[c], u, x[c], t);
        ADD3(t, x[c], y[c]); PRMT(u, t); LOPS(y[c], u, y[c], x[c]); LOPS(t, u, x[c], y[c]); IMAD(u, t, one, y[c]);
        PRMT(w, u); LOP3(x[c], w, u, x[c]); IMAD(t, x[c], one, w); LOPS(y[c], w, t, y[c]); IMAD(u, y[c], one, t);
        LDS(w, a0 + 1024); IDP(x[c], w, one, x[c]); IDP(y[c], w, u, y[c]); IMAD(t, u, one, x[c]); LOP3(x[c], t, y[c], one); LOP3(y[c], y[c], x[c], t);
        STS(a0 + 2048, x[c]);
      }
    }
  }
  uint32_t r = 0;
  for (int c = 0; c < kChains; c++) r ^= x[c] ^ y[c];
  if (r == 0x12345u) out[threadIdx.x] = r;
}

template <int MODE, int UNROLL>
static double run(int per_iter, int sms, int clock_khz)
{
  uint32_t *d;
  cudaMalloc(&d, 4096);
  cudaFuncSetAttribute(body<MODE, UNROLL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152);
  const int iters = 4000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  body<MODE, UNROLL><<<sms, 768, 49152>>>(d, 200, 1u);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(e0);
    body<MODE, UNROLL><<<sms, 768, 49152>>>(d, iters, 1u);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  cudaFree(d);
  const double warp_inst = (double)iters * per_iter * 24.0;                 // per SM: 24 warps
  const double cycles = best * 1e-3 * clock_khz * 1e3;
  return warp_inst / cycles / 4.0;                                          // per scheduler
}

int main()
{
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, 0) != cudaSuccess) { printf("{\"error\": \"no device\"}\n"); return 1; }
  int clk = 0;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  // SASS instructions per loop iteration, counted with cuobjdump (sm_100a, nvcc 12.9): 67 = 64 LOP3 + 3 loop; 131 = 72 LOP3 + 32 PRMT + 8 IADD3 + 8 SHF +
  // 8 VABSDIFF4 + 3 loop; 243 = 128 ALU (80 LOP3, 24 PRMT, 8 IADD3, 8 SHF, 8 VABSDIFF4) + 56 IMAD + 16 IDP.4A + 32 LDS + 8 STS + 3 loop
  // (ALU share 52.7 %, the decoder's is 52.6 %; LSU 16.5 % vs 17.2 %; FMA pipe 29.6 % vs 21 %)
  const double lop3 = run<0, 1>(67, p.multiProcessorCount, clk), alu = run<1, 1>(131, p.multiProcessorCount, clk), mix = run<2, 1>(243, p.multiProcessorCount, clk);
  // the same blend as straight-line code: the loop unrolled 40 times = ~9 600 instructions = 150 KB of SASS per trip, as large as the decoder's
  // unrolled row code -- what instruction fetch costs when nothing is re-used from the L0 instruction cache
  const double mix_big = run<2, 40>(240, p.multiProcessorCount, clk);
  // footprint sweep of the same blend: instruction-cache knees (3.9 KB of SASS per unrolled iteration)
  const double f2 = run<2, 2>(240, p.multiProcessorCount, clk), f4 = run<2, 4>(240, p.multiProcessorCount, clk), f8 = run<2, 8>(240, p.multiProcessorCount, clk),
               f16 = run<2, 16>(240, p.multiProcessorCount, clk), f25 = run<2, 25>(240, p.multiProcessorCount, clk), f32 = run<2, 32>(240, p.multiProcessorCount, clk),
               f80 = run<2, 80>(240, p.multiProcessorCount, clk);
  printf("{\"sm_count\": %d, \"clock_khz\": %d, \"lop3_ipc_per_scheduler\": %.4f, \"alu_mix_ipc_per_scheduler\": %.4f, \"decoder_mix_ipc_per_scheduler\": %.4f, "
         "\"decoder_mix_straight_line_ipc_per_scheduler\": %.4f, \"decoder_mix_alu_share\": %.4f, "
         "\"ipc_vs_code_kb\": {\"3.9\": %.4f, \"7.7\": %.4f, \"15\": %.4f, \"31\": %.4f, \"61\": %.4f, \"96\": %.4f, \"123\": %.4f, \"154\": %.4f, \"307\": %.4f}}\n",
         p.multiProcessorCount, clk, lop3, alu, mix, mix_big, 128.0 / 243.0, mix, f2, f4, f8, f16, f25, f32, mix_big, f80);
  return 0;
}
