// Gold-sequence jump-ahead shared by the scrambling kernels and the PUSCH inner receiver (see scramble_mod.cu for the derivation).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace nrb200 {

constexpr int kGoldPow = 22;     // jump distances up to 2^22 words = 2^27 bits

// L^(2^k) as 8 nibble tables: t[g][k][j][v] = image of (v << 4j).  One matrix-vector product = 8 loads + 8 XORs.
struct GoldTables { uint32_t t[2][kGoldPow][8][16]; };

int scramble_mod_init();                    // builds the tables on first use; 0 ok
const GoldTables *gold_tables_dev();        // device pointer (valid after scramble_mod_init)
const uint32_t *mod_tables_dev();           // QAM tables {re | im << 16}: QPSK at 0, 16QAM at 4, 64QAM at 20, 256QAM at 84

__device__ __forceinline__ uint32_t gold_matvec(const uint32_t (*__restrict__ t)[16], uint32_t x)
{
  uint32_t y = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) y ^= __ldg(&t[j][(x >> (4 * j)) & 15u]);
  return y;
}
// Gold word number w of the sequence started from c_init: the reset loop performs 49 word steps, every call one more
// (transport_proto.h:655-676), so word w is the XOR of both generator states after 50 + w steps.
__device__ __forceinline__ uint32_t gold_word(const GoldTables *__restrict__ T, uint32_t c_init, uint32_t w)
{
  uint32_t x1 = 1u + (1u << 31);
  uint32_t x2 = c_init ^ ((c_init ^ (c_init >> 1) ^ (c_init >> 2) ^ (c_init >> 3)) << 31);
  uint32_t steps = 50u + w;
  for (int k = 0; steps; k++, steps >>= 1)
    if (steps & 1u) { x1 = gold_matvec(T->t[0][k], x1); x2 = gold_matvec(T->t[1][k], x2); }
  return x1 ^ x2;
}

}  // namespace nrb200
