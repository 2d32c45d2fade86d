// Shared device-side declarations for the LDPC kernels.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "nrb200_graph.h"

namespace nrb200 {

constexpr int kCrcTableLen = 8448 + 32;   // longest code block (bits) the CRC-stop mode can see
constexpr int kCrcChunk = 8192;           // long messages are folded in chunks of this many bits, counted from the END of the message
constexpr int kCrcMaxChunks = 256;        // => up to 2 097 152 bits (a transport block is <= ~1.3 Mbit)

constexpr int kLlMaxBatch = 32;           // code blocks one low-latency launch can carry (their row numbers travel in the launch arguments)

// Control record of one staging row of the low-latency path (mapped pinned host memory, written by the kernel, polled by the host thread
// that waits for the block: no stream synchronisation, no copy engine on the way).  32 bytes.
struct LlCtrl {
  uint8_t done[8];     // done[rank] = the launch's sequence byte once CTA `rank` of the block's cluster has stored its share of the output
  int32_t iters;       // what LDPCdecoder returns
  uint8_t abort;       // host -> device: decode_abort_t::failed as the waiting host thread last saw it (polled every iteration, nrLDPC_decoder.c:557)
  uint8_t pad[3];
  uint64_t t_begin, t_end;   // %globaltimer (ns) when CTA 0 of the block started and when it stored its completion byte: device-side time of the block
};

// Launch arguments of the decode kernels (POD, passed by value).
struct DecodeArgs {
  const int8_t *llr;        // n_cb x llr_stride
  uint8_t *out;             // n_cb x out_stride
  int32_t *iters;           // n_cb
  const uint8_t *abort_flags;  // optional n_cb: non-zero = decode_abort_t set (read once per iteration)
  const uint32_t *crc_tab;  // x^j mod g, j < kCrcTableLen, for the selected crc_type (CRC-stop mode only)
  uint32_t n_cb, llr_stride, out_stride;
  uint32_t crc_len_bits;    // the reference's p_decParams->E handed to check_crc (nrLDPC_decoder.c:858)
  uint8_t numMaxIter, outMode, use_crc, quirks;
  uint8_t latency;          // batch API: 1 = a cluster per code block when the launch fits (nrb200_ldpc_batch_desc_t::latency_mode)
  // low-latency mode (ll_ctrl != nullptr): block b works on staging row ll_rows[b] of llr / out / ll_ctrl instead of row b, `iters` is unused
  LlCtrl *ll_ctrl;
  uint8_t ll_seq;
  uint16_t ll_rows[kLlMaxBatch];
};

// where block cb's data lives
struct BlockIo {
  const int8_t *llr;
  uint8_t *out;
  int32_t *iters;
  LlCtrl *ctrl;                      // nullptr outside the low-latency mode
  const volatile uint8_t *abort;     // nullptr when the caller gave no abort flag
};
// low-latency mode: CTA 0 of a block stamps its start time
__device__ __forceinline__ void block_begin(const BlockIo &io, int rank);

__device__ __forceinline__ BlockIo block_io(const DecodeArgs &a, int cb)
{
  BlockIo io;
  const size_t row = a.ll_ctrl ? (size_t)a.ll_rows[cb] : (size_t)cb;
  io.llr = a.llr + row * a.llr_stride;
  io.out = a.out + row * a.out_stride;
  io.ctrl = a.ll_ctrl ? a.ll_ctrl + row : nullptr;
  io.iters = a.ll_ctrl ? &io.ctrl->iters : a.iters + cb;
  io.abort = a.ll_ctrl ? &io.ctrl->abort : (a.abort_flags ? a.abort_flags + cb : nullptr);
  return io;
}

__device__ __forceinline__ void block_begin(const BlockIo &io, int rank)
{
  if (io.ctrl && rank == 0 && threadIdx.x == 0) { uint64_t t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); io.ctrl->t_begin = t; }
}

// End of a block: hand the result over.  Low-latency mode: every thread has fenced its output stores to host memory (system scope); the
// barrier orders them before thread 0's iteration count and completion byte, which the host thread is spinning on.
__device__ __forceinline__ void block_finish(const BlockIo &io, const DecodeArgs &a, int numIter, int rank)
{
  if (io.ctrl) {
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
      if (rank == 0) {
        io.ctrl->iters = numIter;
        uint64_t t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        io.ctrl->t_end = t;
        __threadfence_system();
      }
      *reinterpret_cast<volatile uint8_t *>(&io.ctrl->done[rank]) = a.ll_seq;
    }
  } else if (threadIdx.x == 0 && rank == 0) {
    *io.iters = numIter;
  }
}

// Hard-decision output (reference nrLDPC_bnProc.h:1321-1380). hd holds one 0/1 byte per LLR position in shared memory.
// BIT mode packs MSB first; BITINT8 writes one bit per byte; LLRINT8 is what the reference actually produces for that
// mode: hard bits again, because llr2bit runs in place over the LLR output (nrLDPC_decoder.c:866-877).
__device__ __forceinline__ void write_output(const DecodeArgs &a, uint8_t *o, const uint8_t *hd, int numLLR)
{
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
  if (a.outMode == 0) {
    for (int i = threadIdx.x; i < n; i += blockDim.x)
      if (hd[i]) rem ^= __ldg(a.crc_tab + (n - 1 - i));
  } else {
    // BITINT8 / LLRINT8: the reference hands check_crc the one-bit-per-byte array itself (nrLDPC_decoder.c:852-858), so message bit 8j+7 is
    // hard bit j and every other message bit is 0
    for (int j = threadIdx.x; 8 * j + 7 < n; j += blockDim.x)
      if (hd[j]) rem ^= __ldg(a.crc_tab + (n - 8 - 8 * j));
  }
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
