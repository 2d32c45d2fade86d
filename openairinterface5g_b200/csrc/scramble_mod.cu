// Gold-sequence scrambling / LLR unscrambling and the QAM mapper.
//   nr_codeword_scrambling, nr_codeword_unscrambling   reference openair1/PHY/NR_TRANSPORT/nr_scrambling.c:30-96
//   lte_gold_generic                                    reference openair1/PHY/LTE_TRANSPORT/transport_proto.h:633-680
//   nr_modulation + nr_generate_modulation_table        reference MODULATION/nr_modulation.c:115-244, NR_REFSIG/nr_gen_mod_table.c:33-98
// The reference produces the Gold sequence 32 bits at a time with a serial word recurrence (state' = L(state), GF(2)-linear on the
// 32-bit word).  Here every thread jumps straight to its own run of words with precomputed powers L^(2^k) (binary 32x32 matrices),
// then walks 16 words with the same recurrence -- identical bits, no serial dependence across the code word.
#include "nrb200_ctx.h"
#include <vector>

namespace nrb200 {

constexpr int kGoldPow = 22;     // jump distances up to 2^22 words = 2^27 bits
constexpr int kGoldRun = 16;     // words per thread

struct GoldTables { uint32_t m[2][kGoldPow][32]; };   // m[g][k][i] = image of bit i under 2^k word steps of generator g
static GoldTables *d_gold = nullptr;
static uint32_t *d_modtab = nullptr;                   // per Qm: 2^Qm symbols {re | im << 16}; offsets 0, 4, 20, 84

static inline uint32_t step1(uint32_t x) { x = (x >> 1) ^ (x >> 4); return x ^ (x << 31) ^ (x << 28); }
static inline uint32_t step2(uint32_t x) { x = (x >> 1) ^ (x >> 2) ^ (x >> 3) ^ (x >> 4); return x ^ (x << 31) ^ (x << 30) ^ (x << 29) ^ (x << 28); }

int scramble_mod_init()
{
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  if (d_gold) return 0;
  std::vector<GoldTables> h(1);
  for (int g = 0; g < 2; g++) {
    for (int i = 0; i < 32; i++) h[0].m[g][0][i] = g == 0 ? step1(1u << i) : step2(1u << i);
    for (int k = 1; k < kGoldPow; k++)
      for (int i = 0; i < 32; i++) {          // M^(2^k) e_i = M^(2^(k-1)) (M^(2^(k-1)) e_i)
        const uint32_t v = h[0].m[g][k - 1][i];
        uint32_t y = 0;
        for (int b = 0; b < 32; b++) if ((v >> b) & 1) y ^= h[0].m[g][k - 1][b];
        h[0].m[g][k][i] = y;
      }
  }
  if (cudaMalloc(&d_gold, sizeof(GoldTables)) != cudaSuccess) return -1;
  cudaMemcpy(d_gold, h.data(), sizeof(GoldTables), cudaMemcpyHostToDevice);
  // modulation tables: float32 arithmetic in the order of nr_gen_mod_table.c
  std::vector<uint32_t> t(4 + 16 + 64 + 256);
  const float val = 32768.0f, s2 = 0.70711f, s10 = 0.31623f, s42 = 0.15430f, s170 = 0.076696f;
  auto sym = [&](short lr, short li, float sc) { const short re = (short)(lr * val * sc * s2), im = (short)(li * val * sc * s2); return ((uint32_t)(uint16_t)re) | ((uint32_t)(uint16_t)im << 16); };
  auto sg = [](int idx, int b) { return 1 - 2 * ((idx >> b) & 1); };
  for (int i = 0; i < 4; i++) t[i] = sym((short)sg(i, 0), (short)sg(i, 1), s2);
  for (int i = 0; i < 16; i++) t[4 + i] = sym((short)(sg(i, 0) * (2 - sg(i, 2))), (short)(sg(i, 1) * (2 - sg(i, 3))), s10);
  for (int i = 0; i < 64; i++) t[20 + i] = sym((short)(sg(i, 0) * (4 - sg(i, 2) * (2 - sg(i, 4)))), (short)(sg(i, 1) * (4 - sg(i, 3) * (2 - sg(i, 5)))), s42);
  for (int i = 0; i < 256; i++)
    t[84 + i] = sym((short)(sg(i, 0) * (8 - sg(i, 2) * (4 - sg(i, 4) * (2 - sg(i, 6))))), (short)(sg(i, 1) * (8 - sg(i, 3) * (4 - sg(i, 5) * (2 - sg(i, 7))))), s170);
  if (cudaMalloc(&d_modtab, t.size() * 4) != cudaSuccess) return -1;
  cudaMemcpy(d_modtab, t.data(), t.size() * 4, cudaMemcpyHostToDevice);
  return 0;
}

__device__ __forceinline__ uint32_t dstep1(uint32_t x) { x = (x >> 1) ^ (x >> 4); return x ^ (x << 31) ^ (x << 28); }
__device__ __forceinline__ uint32_t dstep2(uint32_t x) { x = (x >> 1) ^ (x >> 2) ^ (x >> 3) ^ (x >> 4); return x ^ (x << 31) ^ (x << 30) ^ (x << 29) ^ (x << 28); }
__device__ __forceinline__ uint32_t matvec(const uint32_t *__restrict__ col, uint32_t x)
{
  uint32_t y = 0;
#pragma unroll 8
  for (int b = 0; b < 32; b++) y ^= ((x >> b) & 1u) ? __ldg(col + b) : 0u;
  return y;
}
// generator states after `steps` word steps from the reset values (x1 = 1 + 2^31, x2 = c_init with bit 31 completed)
__device__ __forceinline__ void gold_jump(const GoldTables *__restrict__ T, uint32_t c_init, uint32_t steps, uint32_t &x1, uint32_t &x2)
{
  x1 = 1u + (1u << 31);
  x2 = c_init ^ ((c_init ^ (c_init >> 1) ^ (c_init >> 2) ^ (c_init >> 3)) << 31);
  for (int k = 0; steps; k++, steps >>= 1)
    if (steps & 1u) { x1 = matvec(T->m[0][k], x1); x2 = matvec(T->m[1][k], x2); }
}

// mode 0: scramble (in = one bit per byte, out = packed words)   mode 1: unscramble int16 LLRs in place
__global__ void __launch_bounds__(256) gold_kernel(const GoldTables *__restrict__ T, int mode, uint32_t c_init, uint32_t size,
                                                   const uint8_t *__restrict__ in, uint32_t *__restrict__ out, int16_t *__restrict__ llr)
{
  const uint32_t nw = (size + 31) >> 5;
  const uint32_t w0 = (blockIdx.x * blockDim.x + threadIdx.x) * kGoldRun;
  if (w0 >= nw) return;
  uint32_t x1, x2;
  gold_jump(T, c_init, 49u + w0, x1, x2);                 // the reset loop performs 49 steps, every call one more (transport_proto.h:655-676)
  for (uint32_t w = w0; w < w0 + kGoldRun && w < nw; w++) {
    x1 = dstep1(x1); x2 = dstep2(x2);
    const uint32_t s = x1 ^ x2;
    if (mode == 0) {
      uint32_t v = 0;
#pragma unroll 8
      for (int i = 0; i < 32; i++) if (32 * w + i < size) v |= (uint32_t)(in[32 * w + i] & 1u) << i;
      out[w] = v ^ s;
    } else {
#pragma unroll 8
      for (int i = 0; i < 32; i++) {
        const uint32_t n = 32 * w + i;
        if (n < size && ((s >> i) & 1u)) llr[n] = (int16_t)(uint16_t)(0u - (uint16_t)llr[n]);   // mullo_epi16 by -1 wraps
      }
    }
  }
}

__global__ void __launch_bounds__(256) modulate_kernel(const uint32_t *__restrict__ tab, int Qm, uint32_t nsym, const uint8_t *__restrict__ bits,
                                                       uint32_t *__restrict__ out)
{
  const uint32_t *t = tab + (Qm == 2 ? 0 : Qm == 4 ? 4 : Qm == 6 ? 20 : 84);
  const uint32_t mask = (1u << Qm) - 1u;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nsym; i += gridDim.x * blockDim.x) {
    const uint32_t n = i * Qm, byte = n >> 3, sh = n & 7;
    const uint32_t two = (uint32_t)bits[byte] | ((uint32_t)bits[byte + 1] << 8);   // Qm <= 8 bits starting anywhere in a byte: 2 bytes suffice
    out[i] = __ldg(t + ((two >> sh) & mask));
  }
}

int launch_gold(int mode, uint32_t c_init, uint32_t size, const uint8_t *in, uint32_t *out, int16_t *llr, cudaStream_t st)
{
  if (scramble_mod_init() != 0) return -5;
  if (size == 0) return 0;
  const uint32_t nw = (size + 31) >> 5, nthreads = (nw + kGoldRun - 1) / kGoldRun;
  gold_kernel<<<(nthreads + 255) / 256, 256, 0, st>>>(d_gold, mode, c_init, size, in, out, llr);
  ctx().launches++;
  NRB200_CUDA_OK(cudaGetLastError(), "gold launch");
  return 0;
}

int launch_modulate(int Qm, uint32_t length_bits, const uint8_t *bits, int16_t *out, cudaStream_t st)
{
  if (scramble_mod_init() != 0) return -5;
  if (Qm != 2 && Qm != 4 && Qm != 6 && Qm != 8) return -4;
  const uint32_t nsym = length_bits / Qm;
  if (nsym == 0) return 0;
  modulate_kernel<<<std::min<unsigned>((nsym + 255) / 256, 148 * 8), 256, 0, st>>>(d_modtab, Qm, nsym, bits, (uint32_t *)out);
  ctx().launches++;
  NRB200_CUDA_OK(cudaGetLastError(), "modulate launch");
  return 0;
}

}  // namespace nrb200
