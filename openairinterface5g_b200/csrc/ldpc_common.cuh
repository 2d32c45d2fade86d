// Shared device-side declarations for the LDPC kernels.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "nrb200_graph.h"

namespace nrb200 {

constexpr int kCrcTableLen = 8448 + 32;   // longest code block (bits) the CRC-stop mode can see
constexpr int kCrcChunk = 8192;           // long messages are folded in chunks of this many bits, counted from the END of the message
constexpr int kCrcMaxChunks = 256;        // => up to 2 097 152 bits (a transport block is <= ~1.3 Mbit)

// Launch arguments of the decode kernels (POD, passed by value).
struct DecodeArgs {
  const int8_t *llr;        // n_cb x llr_stride
  uint8_t *out;             // n_cb x out_stride
  int32_t *iters;           // n_cb
  const uint8_t *abort_flags;  // optional n_cb: non-zero = decode_abort_t already set when the call was made
  const uint32_t *crc_tab;  // x^j mod g, j < kCrcTableLen, for the selected crc_type (CRC-stop mode only)
  uint32_t n_cb, llr_stride, out_stride;
  uint32_t crc_len_bits;    // the reference's p_decParams->E handed to check_crc (nrLDPC_decoder.c:858)
  uint8_t numMaxIter, outMode, use_crc, quirks;
};

// Hard-decision output (reference nrLDPC_bnProc.h:1321-1380). hd holds one 0/1 byte per LLR position in shared memory.
// BIT mode packs MSB first; BITINT8 writes one bit per byte; LLRINT8 is what the reference actually produces for that
// mode: hard bits again, because llr2bit runs in place over the LLR output (nrLDPC_decoder.c:866-877).
__device__ __forceinline__ void write_output(const DecodeArgs &a, int cb, const uint8_t *hd, int numLLR)
{
  uint8_t *o = a.out + (size_t)cb * a.out_stride;
  if (a.outMode == 0) {
    const int nbytes = (numLLR + 7) >> 3;
    for (int j = threadIdx.x; j < nbytes; j += blockDim.x) {
      unsigned b = 0;
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const int i = j * 8 + k;
        b |= (i < numLLR ? (unsigned)hd[i] : 0u) << (7 - k);
      }
      o[j] = (uint8_t)b;
    }
  } else {
    for (int i = threadIdx.x; i < numLLR; i += blockDim.x) o[i] = hd[i];
  }
}

// check_crc (reference crc_byte.c:314-379) on the first crc_len_bits hard bits: payload CRC == trailing CRC bits
// <=> the whole string (payload || crc) leaves remainder 0.  Remainder = XOR over set bits i of x^(n-1-i) mod g.
// Every thread returns the same verdict.  scratch: one int of shared memory.
__device__ __forceinline__ int crc_check_block(const DecodeArgs &a, const uint8_t *hd, int *scratch)
{
  const int n = (int)a.crc_len_bits;
  unsigned rem = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x)
    if (hd[i]) rem ^= __ldg(a.crc_tab + (n - 1 - i));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) rem ^= __shfl_xor_sync(0xffffffffu, rem, o);
  if (threadIdx.x == 0) *scratch = 0;
  __syncthreads();
  if ((threadIdx.x & 31) == 0 && rem) atomicXor(scratch, (int)rem);
  __syncthreads();
  const int r = *scratch;
  __syncthreads();
  return r == 0;
}

}  // namespace nrb200
