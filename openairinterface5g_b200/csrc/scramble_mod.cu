// Gold-sequence scrambling / LLR unscrambling and the QAM mapper.
//   nr_codeword_scrambling, nr_codeword_unscrambling   reference openair1/PHY/NR_TRANSPORT/nr_scrambling.c:30-96
//   lte_gold_generic                                    reference openair1/PHY/LTE_TRANSPORT/transport_proto.h:633-680
//   nr_modulation + nr_generate_modulation_table        reference MODULATION/nr_modulation.c:115-244, NR_REFSIG/nr_gen_mod_table.c:33-98
// The reference produces the Gold sequence 32 bits at a time with a serial word recurrence (state' = L(state), GF(2)-linear on the
// 32-bit word).  Here every thread jumps straight to its own word with precomputed powers L^(2^k) (binary 32x32 matrices),
// (nibble-indexed tables: 8 loads per product) -- identical bits, no serial dependence across the code word.
#include "nrb200_ctx.h"
#include "gold_seq.cuh"
#include <cstring>
#include <vector>

namespace nrb200 {

constexpr int kGoldBlk = 256;    // words (= threads) per CTA

static GoldTables *d_gold_dev[kMaxDevices] = {nullptr};  // per device
static uint32_t *d_modtab_dev[kMaxDevices] = {nullptr};                  // per Qm: 2^Qm symbols {re | im << 16}; offsets 0, 4, 20, 84

static inline uint32_t step1(uint32_t x) { x = (x >> 1) ^ (x >> 4); return x ^ (x << 31) ^ (x << 28); }
static inline uint32_t step2(uint32_t x) { x = (x >> 1) ^ (x >> 2) ^ (x >> 3) ^ (x >> 4); return x ^ (x << 31) ^ (x << 30) ^ (x << 29) ^ (x << 28); }

#define d_gold (d_gold_dev[ctx().dev])
#define d_modtab (d_modtab_dev[ctx().dev])
const GoldTables *gold_tables_dev() { return d_gold; }
const uint32_t *mod_tables_dev() { return d_modtab; }

int scramble_mod_init()
{
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  if (d_gold) return 0;
  std::vector<GoldTables> h(1);
  for (int g = 0; g < 2; g++) {
    uint32_t col[32], nxt[32];
    for (int i = 0; i < 32; i++) col[i] = g == 0 ? step1(1u << i) : step2(1u << i);
    for (int k = 0; k < kGoldPow; k++) {
      for (int j = 0; j < 8; j++)
        for (int v = 0; v < 16; v++) {
          uint32_t y = 0;
          for (int b = 0; b < 4; b++) if ((v >> b) & 1) y ^= col[4 * j + b];
          h[0].t[g][k][j][v] = y;
        }
      for (int i = 0; i < 32; i++) {            // M^(2^(k+1)) e_i = M^(2^k) (M^(2^k) e_i)
        uint32_t y = 0;
        for (int b = 0; b < 32; b++) if ((col[i] >> b) & 1) y ^= col[b];
        nxt[i] = y;
      }
      std::memcpy(col, nxt, sizeof(col));
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
// One CTA = 256 words of the sequence (8192 bits).  Phase 1: thread t produces word t.  Phase 2: the CTA sweeps its 8192 elements
// with coalesced accesses.   mode 0: scramble (in = one bit per byte, out = packed words)   mode 1: unscramble int16 LLRs in place
__global__ void __launch_bounds__(kGoldBlk) gold_kernel(const GoldTables *__restrict__ T, int mode, int aligned, uint32_t c_init, uint32_t size,
                                                        const uint8_t *__restrict__ in, uint32_t *__restrict__ out, int16_t *__restrict__ llr)
{
  __shared__ uint32_t s_gold[kGoldBlk];
  const uint32_t nw = (size + 31) >> 5, wb = blockIdx.x * kGoldBlk, t = threadIdx.x;
  if (wb + t < nw) s_gold[t] = gold_word(T, c_init, wb + t);
  __syncthreads();
  const uint32_t e0 = wb * 32u;
  if (mode == 0) {
#pragma unroll 2
    for (int k = 0; k < 8; k++) {                         // 4 input bytes per thread per pass, 8 lanes make one word
      const uint32_t e = e0 + 4u * t + 1024u * k;
      uint32_t nib = 0;
      if (aligned && e + 4 <= size) {
        const uint32_t v = *reinterpret_cast<const uint32_t *>(in + e) & 0x01010101u;
        nib = (v | (v >> 7) | (v >> 14) | (v >> 21)) & 15u;
      } else {
        for (int i = 0; i < 4; i++) if (e + i < size) nib |= (uint32_t)(in[e + i] & 1u) << i;
      }
      uint32_t v = nib << (4u * (t & 7u));
      v |= __shfl_xor_sync(0xffffffffu, v, 1);
      v |= __shfl_xor_sync(0xffffffffu, v, 2);
      v |= __shfl_xor_sync(0xffffffffu, v, 4);
      if ((t & 7u) == 0 && e < size) out[e >> 5] = v ^ s_gold[(e - e0) >> 5];
    }
  } else {
#pragma unroll 4
    for (int k = 0; k < 16; k++) {                        // 2 LLRs per thread per pass
      const uint32_t e = e0 + 2u * t + 512u * k;
      if (e >= size) break;
      const uint32_t g = s_gold[(e - e0) >> 5] >> (e & 31u);
      if (aligned && e + 2 <= size) {
        uint32_t v = *reinterpret_cast<uint32_t *>(llr + e);
        // 16-bit wrapping negation of the halves selected by the two sequence bits (mullo_epi16 by -1: -32768 stays)
        const uint32_t m = ((g & 1u) ? 0xffffu : 0u) | ((g & 2u) ? 0xffff0000u : 0u);
        v = __vsub2(v ^ m, m);                            // (x ^ -1) - (-1) = -x per halfword, wrapping
        *reinterpret_cast<uint32_t *>(llr + e) = v;
      } else {
        for (int i = 0; i < 2; i++) if (e + i < size && ((g >> i) & 1u)) llr[e + i] = (int16_t)(uint16_t)(0u - (uint16_t)llr[e + i]);
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
  const uint32_t nw = (size + 31) >> 5;
  const int aligned = mode == 0 ? ((uintptr_t)in & 3u) == 0 : ((uintptr_t)llr & 3u) == 0;
  gold_kernel<<<(nw + kGoldBlk - 1) / kGoldBlk, kGoldBlk, 0, st>>>(d_gold, mode, aligned, c_init, size, in, out, llr);
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
