// NR LDPC encoder kernels.  Systematic QC encoding through H's dual-diagonal core: the result is bit-identical to the
// reference's generator-matrix XOR networks (nrLDPC_encoder/ldpc_encoder_optim8segmulti.c:46-212,
// ldpc_encode_parity_check.c:90-220) -- any correct systematic encoder of the same code is.  Output layout is the ABI's:
// one bit per byte, K-2Z systematic bits followed by all parity bits (66Z for BG1, 50Z for BG2).
#include "nrb200_ctx.h"
#include "ldpc_common.cuh"
#include <cstdlib>

namespace nrb200 {

// one CTA per code block; x = the whole codeword as 0/1 bytes in shared memory
__global__ void __launch_bounds__(384, 2)
ldpc_encode_kernel(const EncGraphDev *__restrict__ gdev, int K, uint32_t n_cb, const uint8_t *__restrict__ in, uint32_t in_stride,
                   uint8_t *__restrict__ out, uint32_t out_stride, volatile uint8_t *done)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  EncGraphDev &g = *reinterpret_cast<EncGraphDev *>(smem_raw);
  for (int i = threadIdx.x; i < (int)(sizeof(EncGraphDev) / 4); i += blockDim.x)
    reinterpret_cast<int *>(smem_raw)[i] = reinterpret_cast<const int *>(gdev)[i];
  __syncthreads();
  const int Z = g.Z, nsys = g.nsys, ncols = g.ncols, nrows = g.nrows;
  uint8_t *x = smem_raw + ((sizeof(EncGraphDev) + 15) & ~15);
  uint8_t *lam = x + (((size_t)ncols * Z + 15) & ~15);     // 4 x Z systematic partial sums of the core rows

  for (uint32_t cb = blockIdx.x; cb < n_cb; cb += gridDim.x) {
    const uint8_t *src = in + (size_t)cb * in_stride;
    for (int i = threadIdx.x; i < ncols * Z; i += blockDim.x)
      x[i] = i < K ? (uint8_t)((src[i >> 3] >> (7 - (i & 7))) & 1) : (uint8_t)0;
    __syncthreads();
    // lambda_r (r < 4): systematic columns only
    for (int i = threadIdx.x; i < 4 * Z; i += blockDim.x) {
      const int r = i / Z, t = i - r * Z;
      unsigned acc = 0;
      for (int e = g.row_start[r]; e < g.row_start[r + 1]; e++) {
        const int c = g.edge_col[e];
        if (c >= nsys) continue;
        int v = t + g.edge_shift[e]; if (v >= Z) v -= Z;
        acc ^= x[c * Z + v];
      }
      lam[i] = (uint8_t)acc;
    }
    __syncthreads();
    // p0: sum of the four core rows = x^sigma * p0
    for (int t = threadIdx.x; t < Z; t += blockDim.x) {
      int v = t + g.sigma; if (v >= Z) v -= Z;
      x[nsys * Z + v] = lam[t] ^ lam[Z + t] ^ lam[2 * Z + t] ^ lam[3 * Z + t];
    }
    __syncthreads();
    // the other three core parity columns, in dependency order
    for (int n = 0; n < 3; n++) {
      const int r = g.core_row[n], pc = g.core_col[n], psh = g.core_shift[n];
      for (int t = threadIdx.x; t < Z; t += blockDim.x) {
        unsigned acc = lam[r * Z + t];
        for (int e = g.row_start[r]; e < g.row_start[r + 1]; e++) {
          const int c = g.edge_col[e];
          if (c < nsys || c == pc) continue;
          // columns still unknown at this step are all-zero in x, so including them is harmless
          int v = t + g.edge_shift[e]; if (v >= Z) v -= Z;
          acc ^= x[c * Z + v];
        }
        int v = t + psh; if (v >= Z) v -= Z;
        x[pc * Z + v] = (uint8_t)acc;
      }
      __syncthreads();
    }
    // extension rows: the degree-1 diagonal column (shift 0) closes each row
    for (int i = threadIdx.x; i < (nrows - 4) * Z; i += blockDim.x) {
      const int r = 4 + i / Z, t = i % Z;
      unsigned acc = 0;
      for (int e = g.row_start[r]; e < g.row_start[r + 1]; e++) {
        const int c = g.edge_col[e];
        if (c == nsys + r) continue;
        int v = t + g.edge_shift[e]; if (v >= Z) v -= Z;
        acc ^= x[c * Z + v];
      }
      x[(nsys + r) * Z + t] = (uint8_t)acc;
    }
    __syncthreads();
    uint8_t *dst = out + (size_t)cb * out_stride;
    const int nout = (ncols - 2) * Z;
    for (int i = threadIdx.x; i < nout; i += blockDim.x) dst[i] = x[2 * Z + i];
    if (done) __threadfence_system();
    __syncthreads();
    if (done && threadIdx.x == 0) done[cb] = 1;
  }
}

// ---- bit-packed variant for lifting sizes that are a multiple of 32 (the hot Z = 384 and every Z = 32 k): a column of the code word is W = Z / 32 words
//      (bit b of word w = lift 32 w + b), a circular shift is one funnel shift of two neighbouring words, and a row of H is an XOR of a few words.  The whole
//      code word is 68 W words (3.2 KB at Z = 384), so a block costs ~4 k word operations instead of ~120 k byte operations; one CTA of 128 threads per block.
//      Same schedule as above: lambda of the four core rows -> p0 -> the other three core columns -> extension rows -> one-bit-per-byte store.
__device__ __forceinline__ uint32_t rotw(const uint32_t *col, int W, int w, int s)
{
  const int q = s >> 5, r = s & 31;
  int a = w + q; if (a >= W) a -= W;
  int b = a + 1; if (b >= W) b -= W;
  return __funnelshift_r(col[a], col[b], r);       // bits (32 w + s) mod Z ... of the column
}

__global__ void __launch_bounds__(128)
ldpc_encode_packed_kernel(const EncGraphDev *__restrict__ gdev, int K, uint32_t n_cb, const uint8_t *__restrict__ in, uint32_t in_stride,
                          uint8_t *__restrict__ out, uint32_t out_stride, volatile uint8_t *done)
{
  __shared__ EncGraphDev g;
  __shared__ __align__(16) uint32_t x[68 * 12];
  __shared__ uint32_t lam[4 * 12];
  for (int i = threadIdx.x; i < (int)(sizeof(EncGraphDev) / 4); i += blockDim.x)
    reinterpret_cast<int *>(&g)[i] = reinterpret_cast<const int *>(gdev)[i];
  __syncthreads();
  const int Z = g.Z, W = Z >> 5, nsys = g.nsys, ncols = g.ncols, nrows = g.nrows;
  for (uint32_t cb = blockIdx.x; cb < n_cb; cb += gridDim.x) {
    const uint8_t *src = in + (size_t)cb * in_stride;
    for (int i = threadIdx.x; i < ncols * W; i += blockDim.x) {
      uint32_t v = 0;
      const int bit0 = i << 5;
      if (bit0 < K) {                                                        // MSB-first source bits (ldpc_encoder_optim8segmulti.c:132-149)
        const uint32_t be = ((uint32_t)src[(bit0 >> 3)] << 24) | ((uint32_t)src[(bit0 >> 3) + 1] << 16) | ((uint32_t)src[(bit0 >> 3) + 2] << 8) | src[(bit0 >> 3) + 3];
        v = __brev(be);
        if (bit0 + 32 > K) v &= (1u << (K - bit0)) - 1u;
      }
      x[i] = v;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 4 * W; i += blockDim.x) {
      const int r = i / W, w = i - r * W;
      uint32_t acc = 0;
      for (int e = g.row_start[r]; e < g.row_start[r + 1]; e++) {
        const int c = g.edge_col[e];
        if (c < nsys) acc ^= rotw(x + c * W, W, w, g.edge_shift[e]);
      }
      lam[i] = acc;
    }
    __syncthreads();
    if ((int)threadIdx.x < W) {                                             // x^sigma p0 = sum of the four core rows
      const int w = threadIdx.x;
      uint32_t *y = x + (nsys + 4) * W;                                      // scratch: the first extension column is still unused here
      y[w] = lam[w] ^ lam[W + w] ^ lam[2 * W + w] ^ lam[3 * W + w];
    }
    __syncthreads();
    if ((int)threadIdx.x < W) {
      int s = Z - g.sigma; if (s >= Z) s -= Z;
      x[nsys * W + threadIdx.x] = rotw(x + (nsys + 4) * W, W, threadIdx.x, s);
    }
    __syncthreads();
    for (int n = 0; n < 3; n++) {
      const int r = g.core_row[n], pc = g.core_col[n];
      uint32_t *y = x + (nsys + 4) * W;
      if ((int)threadIdx.x < W) {
        const int w = threadIdx.x;
        uint32_t acc = lam[r * W + w];
        for (int e = g.row_start[r]; e < g.row_start[r + 1]; e++) {
          const int c = g.edge_col[e];
          if (c < nsys || c == pc) continue;                                 // columns still unknown at this step are all-zero in x
          acc ^= rotw(x + c * W, W, w, g.edge_shift[e]);
        }
        y[w] = acc;
      }
      __syncthreads();
      if ((int)threadIdx.x < W) {
        int s = Z - g.core_shift[n]; if (s >= Z) s -= Z;
        x[pc * W + threadIdx.x] = rotw(y, W, threadIdx.x, s);
      }
      __syncthreads();
    }
    if ((int)threadIdx.x < W) x[(nsys + 4) * W + threadIdx.x] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < (nrows - 4) * W; i += blockDim.x) {
      const int r = 4 + i / W, w = i % W;
      uint32_t acc = 0;
      for (int e = g.row_start[r]; e < g.row_start[r + 1]; e++) {
        const int c = g.edge_col[e];
        if (c < nsys + 4) acc ^= rotw(x + c * W, W, w, g.edge_shift[e]);     // systematic + core columns; the diagonal column closes the row
      }
      x[(nsys + r) * W + w] = acc;
    }
    __syncthreads();
    uint8_t *dst = out + (size_t)cb * out_stride;
    const int n16 = (ncols - 2) * Z / 16;                                    // 16 code bits -> one 16-byte store, consecutive lanes contiguous
    const bool al16 = ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
    for (int i = threadIdx.x; i < n16; i += blockDim.x) {
      const uint32_t h = (x[2 * W + (i >> 1)] >> ((i & 1) * 16)) & 0xFFFFu;
      uint4 v;
      v.x = ((h & 0xFu) * 0x00204081u) & 0x01010101u;
      v.y = (((h >> 4) & 0xFu) * 0x00204081u) & 0x01010101u;
      v.z = (((h >> 8) & 0xFu) * 0x00204081u) & 0x01010101u;
      v.w = ((h >> 12) * 0x00204081u) & 0x01010101u;
      if (al16) *reinterpret_cast<uint4 *>(dst + 16 * (size_t)i) = v;
      else {
        uint32_t a[4] = {v.x, v.y, v.z, v.w};
        for (int k = 0; k < 16; k++) dst[16 * (size_t)i + k] = (uint8_t)(a[k >> 2] >> (8 * (k & 3)));
      }
    }
    if (done) __threadfence_system();        // low-latency mode: out is mapped host memory, done[cb] is what the calling thread spins on
    __syncthreads();
    if (done && threadIdx.x == 0) done[cb] = 1;
  }
}

// done != nullptr (low-latency mode): in / out / done are mapped host memory, done[cb] := 1 once block cb's output is stored
int launch_encode(const EncGraphDev *d_g, const EncGraphDev &h_g, int K, uint32_t n_cb, const uint8_t *d_in, uint32_t in_stride,
                  uint8_t *d_out, uint32_t out_stride, cudaStream_t stream, uint8_t *done)
{
  if (n_cb == 0) return 0;
  static const bool force_bytes = getenv("NRB200_ENCODE_BYTES") != nullptr;
  if (h_g.Z % 32 == 0 && K % 32 == 0 && !force_bytes) {
    ldpc_encode_packed_kernel<<<n_cb, 128, 0, stream>>>(d_g, K, n_cb, d_in, in_stride, d_out, out_stride, done);
    ctx().launches++;
    NRB200_CUDA_OK(cudaGetLastError(), "packed encode launch");
    return 0;
  }
  auto a16 = [](size_t v) { return (v + 15) & ~(size_t)15; };
  const size_t smem = a16(sizeof(EncGraphDev)) + a16((size_t)h_g.ncols * h_g.Z) + a16((size_t)4 * h_g.Z);
  static std::atomic<size_t> configured[kMaxDevices];
  if (smem > configured[ctx().dev].load()) {
    NRB200_CUDA_OK(cudaFuncSetAttribute(ldpc_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "enc smem attr");
    configured[ctx().dev].store(smem);
  }
  int threads = ((h_g.Z + 31) / 32) * 32;
  if (threads < 64) threads = 64;
  if (threads > 384) threads = 384;
  ldpc_encode_kernel<<<n_cb, threads, smem, stream>>>(d_g, K, n_cb, d_in, in_stride, d_out, out_stride, done);
  ctx().launches++;
  NRB200_CUDA_OK(cudaGetLastError(), "encode launch");
  return 0;
}

// ---- CRC: out = data(x) * x^deg mod g, left-aligned like crc_byte.c:148-312.  One CTA per bit string;
//      remainder = XOR over set bits i of tab[bitlen-1-i+deg].
__global__ void crc_kernel(const uint32_t *__restrict__ tab, int deg, uint32_t n_blk, const uint8_t *__restrict__ in, uint32_t stride,
                           uint32_t bitlen, uint32_t *__restrict__ out)
{
  __shared__ unsigned s_acc;
  for (uint32_t b = blockIdx.x; b < n_blk; b += gridDim.x) {
    if (threadIdx.x == 0) s_acc = 0;
    __syncthreads();
    const uint8_t *src = in + (size_t)b * stride;
    unsigned rem = 0;
    const uint32_t nbytes = (bitlen + 7) / 8;
    for (uint32_t j = threadIdx.x; j < nbytes; j += blockDim.x) {
      unsigned byte = src[j];
      while (byte) {
        const int k = 31 - __clz(byte);          // bit k of the byte = stream bit 8j + (7-k)
        byte &= ~(1u << k);
        const uint32_t i = 8 * j + (7 - k);
        if (i < bitlen) rem ^= __ldg(tab + (bitlen - 1 - i + deg));
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rem ^= __shfl_xor_sync(0xffffffffu, rem, o);
    if ((threadIdx.x & 31) == 0 && rem) atomicXor(&s_acc, rem);
    __syncthreads();
    if (threadIdx.x == 0) out[b] = s_acc;
    __syncthreads();
  }
}

// Long messages: CTA (k, b) takes chunk k of block b -- the bits whose distance from the end lies in [k L, (k+1) L) -- reduces it with the
// short table, then moves the deg-bit remainder to its place with shift[k] and XORs it into out[b] (zeroed by the launcher).
__global__ void crc_long_kernel(const uint32_t *__restrict__ tab, const uint32_t *__restrict__ shift, int deg, const uint8_t *__restrict__ in,
                                uint32_t stride, uint32_t bitlen, uint32_t *__restrict__ out)
{
  __shared__ unsigned s_acc;
  const uint32_t k = blockIdx.x, b = blockIdx.y;
  if (threadIdx.x == 0) s_acc = 0;
  __syncthreads();
  const uint8_t *src = in + (size_t)b * stride;
  // message bit i (0 = first) has distance e = bitlen - 1 - i from the end; this chunk: e in [kL, (k+1)L)
  const uint32_t e_lo = k * kCrcChunk, e_hi = min(bitlen, (k + 1) * (uint32_t)kCrcChunk);      // [e_lo, e_hi)
  const uint32_t i_lo = bitlen - e_hi, i_hi = bitlen - e_lo;                                   // bits [i_lo, i_hi)
  unsigned rem = 0;
  for (uint32_t j = (i_lo >> 3) + threadIdx.x; j <= ((i_hi - 1) >> 3); j += blockDim.x) {
    unsigned byte = src[j];
    while (byte) {
      const int kk = 31 - __clz(byte);
      byte &= ~(1u << kk);
      const uint32_t i = 8 * j + (7 - kk);
      if (i >= i_lo && i < i_hi) rem ^= __ldg(tab + (bitlen - 1 - i - e_lo + deg));
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) rem ^= __shfl_xor_sync(0xffffffffu, rem, o);
  if ((threadIdx.x & 31) == 0 && rem) atomicXor(&s_acc, rem);
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned r = s_acc, folded = 0;
    for (int bb = 0; bb < deg; bb++) if ((r >> (32 - deg + bb)) & 1u) folded ^= __ldg(shift + k * 32 + bb);
    if (folded) atomicXor(out + b, folded);
  }
}

int launch_crc(int poly_id, uint32_t n_blk, const uint8_t *d_in, uint32_t stride, uint32_t bitlen, uint32_t *d_out, cudaStream_t stream)
{
  if (poly_id < 0 || poly_id > 7) return -4;
  const int deg = poly_id <= 2 ? 24 : poly_id == 3 ? 16 : poly_id == 4 ? 12 : poly_id == 5 ? 11 : poly_id == 6 ? 8 : 6;
  if (n_blk == 0) return 0;
  if (bitlen + deg > (uint32_t)kCrcTableLen) {
    const uint32_t nchunks = (bitlen + kCrcChunk - 1) / kCrcChunk;
    if (nchunks > (uint32_t)kCrcMaxChunks || n_blk > 65535) return -4;
    NRB200_CUDA_OK(cudaMemsetAsync(d_out, 0, (size_t)n_blk * 4, stream), "crc memset");
    crc_long_kernel<<<dim3(nchunks, n_blk), 256, 0, stream>>>(ctx().crc_tab[poly_id], ctx().crc_shift[poly_id], deg, d_in, stride, bitlen, d_out);
    ctx().launches++;
    NRB200_CUDA_OK(cudaGetLastError(), "crc launch");
    return 0;
  }
  crc_kernel<<<n_blk, 256, 0, stream>>>(ctx().crc_tab[poly_id], deg, n_blk, d_in, stride, bitlen, d_out);
  ctx().launches++;
  NRB200_CUDA_OK(cudaGetLastError(), "crc launch");
  return 0;
}

// ---- transport block CRC attachment + nr_segmentation on the device (nr_dlsch_coding.c:300-336, nr_segmentation.c:143-175): CTA r copies its (Kprime - L) / 8
// payload bytes (the last segment ends with the TB CRC, read from d_tbcrc), computes CRC24B over them while copying (C > 1), appends it and zeroes the filler
// bytes.  The TB CRC itself is one launch_crc before this kernel (stream ordered).
__global__ void __launch_bounds__(256) segment_kernel(const uint32_t *__restrict__ tab24b, const uint32_t *__restrict__ d_tbcrc, const uint8_t *__restrict__ payload,
                                                      uint32_t a_bytes, uint32_t crc_bytes, uint32_t nbytes, int add_crc, uint32_t kp_bytes, uint32_t k_bytes,
                                                      uint8_t *__restrict__ segs, uint32_t stride)
{
  __shared__ unsigned s_acc;
  const uint32_t r = blockIdx.x, bitlen = nbytes * 8;
  if (threadIdx.x == 0) s_acc = 0;
  __syncthreads();
  uint8_t *dst = segs + (size_t)r * stride;
  const uint32_t tbcrc = *d_tbcrc;                       // left aligned: byte q of the attached CRC = bits 31 - 8q .. 24 - 8q
  unsigned rem = 0;
  for (uint32_t j = threadIdx.x; j < nbytes; j += blockDim.x) {
    const uint32_t q = r * nbytes + j;
    unsigned byte = q < a_bytes ? payload[q] : (q - a_bytes < crc_bytes ? (tbcrc >> (24 - 8 * (q - a_bytes))) & 255u : 0u);
    dst[j] = (uint8_t)byte;
    if (add_crc)
      while (byte) {
        const int k = 31 - __clz(byte);
        byte &= ~(1u << k);
        rem ^= __ldg(tab24b + (bitlen - 1 - (8 * j + (7 - k)) + 24));
      }
  }
  for (uint32_t j = kp_bytes + threadIdx.x; j < k_bytes; j += blockDim.x) dst[j] = 0;      // filler bytes
  if (!add_crc) return;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) rem ^= __shfl_xor_sync(0xffffffffu, rem, o);
  if ((threadIdx.x & 31) == 0 && rem) atomicXor(&s_acc, rem);
  __syncthreads();
  if (threadIdx.x < 3) dst[nbytes + threadIdx.x] = (uint8_t)(s_acc >> (24 - 8 * threadIdx.x));
}

// nr_segmentation's scalar part (nr_segmentation.c:32-141).  out: C, K, Z, F, Kprime, L.  Returns Kb or -1.
int tb_segment_parms(int BG, uint32_t A, uint32_t out[6])
{
  const uint32_t B = A + (A > 3824 ? 24 : 16), Kcb = BG == 1 ? 8448 : 3840;
  uint32_t L = 0, Cn = 1, Bp = B;
  if (B > Kcb) { L = 24; Cn = B / (Kcb - L); if ((Kcb - L) * Cn < B) Cn++; Bp = B + Cn * L; }
  const uint32_t Kp = Bp / Cn;
  const uint32_t Kb = BG == 1 ? 22 : (B > 640 ? 10 : B > 560 ? 9 : B > 192 ? 8 : 6);
  const uint32_t Zmin = Kp / Kb + ((Kp % Kb) ? 1 : 0);
  uint32_t Z;
  if (Zmin <= 2) Z = 2;
  else if (Zmin <= 16) Z = Zmin;
  else {
    uint32_t step = Zmin <= 32 ? 2 : Zmin <= 64 ? 4 : Zmin <= 128 ? 8 : Zmin <= 256 ? 16 : Zmin <= 384 ? 32 : 0;
    if (!step) return -1;
    Z = (Zmin / step) * step;
    if (Z < Zmin) Z += step;
  }
  const uint32_t K = Z * (BG == 1 ? 22 : 10);
  out[0] = Cn; out[1] = K; out[2] = Z; out[3] = K - Kp; out[4] = Kp; out[5] = L;
  return (int)Kb;
}

int launch_tb_segment(int BG, uint32_t A, const uint8_t *d_payload, uint8_t *d_segs, uint32_t seg_stride, uint32_t *d_scratch, cudaStream_t stream)
{
  uint32_t q[6];
  if ((BG != 1 && BG != 2) || A == 0 || (A & 7) || tb_segment_parms(BG, A, q) < 0) return -4;
  const uint32_t Cn = q[0], K = q[1], Kp = q[4], L = q[5];
  if (((Kp - L) & 7) || seg_stride < K / 8 || K + 24 > (uint32_t)kCrcTableLen) return -4;      // the reference's byte copy assumes whole bytes per segment too
  const int tb_poly = A > 3824 ? 0 : 3;
  int rc = launch_crc(tb_poly, 1, d_payload, (A + 7) / 8, A, d_scratch, stream);
  if (rc) return rc;
  segment_kernel<<<Cn, 256, 0, stream>>>(ctx().crc_tab[1], d_scratch, d_payload, A / 8, tb_poly == 0 ? 3 : 2, (Kp - L) / 8, Cn > 1, Kp / 8, K / 8, d_segs, seg_stride);
  ctx().launches++;
  NRB200_CUDA_OK(cudaGetLastError(), "segment launch");
  return 0;
}

int quirks_from_env()
{
  static int q = -1;
  if (q < 0) {
    const char *s = getenv("NRB200_EMULATE_AVX2_BG2R15_DEFECT"), *t = getenv("NRB200_CLUSTER_TIMERS");
    q = ((s && *s == '1') ? 1 : 0) | ((t && *t == '1') ? 2 : 0);   // bit 1: the cluster decoder records clock64() phase marks (tools/cluster_phases.py)
  }
  return q;
}

}  // namespace nrb200
