// Low-latency variant of the packed flooding decoder: ONE code block is decoded by a thread-block CLUSTER of C CTAs (2, 4 or 8 SMs)
// instead of one CTA, for the calls where a handful of code blocks arrive at a time and the time of one block -- not the rate of a
// thousand -- is what the caller sees: the per-segment LDPCdecoder calls of OAI's tpool workers (nr_ulsch_decoding.c:435-468), ldpctest's
// serial loop (ldpctest.c:329-340), a single slot's 28-52 code blocks.  Same arithmetic, same schedule, same results as
// ldpc_decode_packed_kernel (bit exact with nrLDPC_decoder.c:206-881); what changes is where the state lives:
//
//   every CTA   holds a full replica of A (a-posteriori LLRs), L (channel LLRs) and the P rows of the check rows it owns
//   R           (cn->bn messages) of (row, 32-word chunk) lives ONLY in the CTA whose warp owns that work item
//   CN phase    all local: reads the A replica, updates the owned R words in place
//   BN phase    pulls the R words it needs from their owners through distributed shared memory (ld.shared::cluster; consecutive lanes read
//               consecutive words, so the second word of the funnel shift comes from the neighbouring lane by shuffle and only lane 31 loads
//               it), forms A and stores it into every CTA's replica (st.shared::cluster)
//   barriers    two cluster barriers per iteration (R complete -> BN may pull; A complete -> CN may read); the parity-check / abort verdicts
//               travel as one byte per CTA written into every CTA's flag word before the first of them
//
// DSMEM traffic per iteration and cluster: ~105 KB of R pulled + 20 KB x C of A broadcast; at ~20 B/clk per SM that is hidden behind the
// ~4 k cycles of check-node work per CTA at C = 8.
#pragma once
#include "ldpc_cluster.h"
#include "ldpc_decoder_packed.cuh"

namespace nrb200 {

__device__ __forceinline__ uint32_t cl_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cl_map(uint32_t saddr, uint32_t rank)
{
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ uint32_t cl_ld(uint32_t caddr)
{
  uint32_t v;
  asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(caddr) : "memory");
  return v;
}
__device__ __forceinline__ void cl_st(uint32_t caddr, uint32_t v) { asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(caddr), "r"(v) : "memory"); }
__device__ __forceinline__ void cl_st8(uint32_t caddr, uint32_t v) { asm volatile("st.shared::cluster.u8 [%0], %1;" ::"r"(caddr), "r"(v) : "memory"); }
__device__ __forceinline__ void cl_sync()
{
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// A' of column c, word kb/4, from the channel word and the column's R' words pulled from their owner CTAs; stored into all C replicas
template <int ZWC>
__device__ __forceinline__ void bn_col_cluster(const PackedGraph &G, const ClusterSched &S, char *__restrict__ smb, uint32_t sbase, int c, uint32_t kb,
                                               uint32_t lane, int C)
{
  const uint32_t ZB = geo_zb<ZWC>(G);
  const uint32_t lw = lds(smb, G.off_L + c * geo_rsb<ZWC>(G) + kb);
  uint32_t s0 = __dp4a(lw, 0x00000001u, 0u), s1 = __dp4a(lw, 0x00000100u, 0u), s2 = __dp4a(lw, 0x00010000u, 0u), s3 = __dp4a(lw, 0x01000000u, 0u);
  const int i1 = G.col_start[c + 1];
  for (int i = G.col_start[c]; i < i1; i += 4) {
    uint32_t w0[4], x1[4], fa[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {                                      // four edges' pulls in flight before the first is consumed
      const int e = min(i + j, i1 - 1);
      const uint2 d = *reinterpret_cast<const uint2 *>(S.bn_desc[e]);
      const uint32_t lim = d.y & 0xFFFFFu;
      uint32_t wb = kb - (lim >> 8);                                   // byte offset of the word inside its R row: 4 * ((k - qq) mod Zw)
      uint32_t ad = kb + d.x;
      if (((kb << 8) | 0xFFu) < lim) { ad += ZB; wb += ZB; }
      w0[j] = cl_ld(cl_map(sbase + ad, (d.y >> (20u + 3u * (wb >> 7))) & 7u));
      x1[j] = 0u;
      if (lane == 31u) x1[j] = cl_ld(cl_map(sbase + ad + 4u, (d.y >> (20u + 3u * ((wb + 4u) >> 7))) & 7u));
      fa[j] = d.y;
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
      uint32_t w1 = __shfl_down_sync(0xffffffffu, w0[j], 1);           // lane l + 1 holds the word after lane l's (rows are circular)
      if (lane == 31u) w1 = x1[j];
      uint32_t rw = __funnelshift_r(w0[j], w1, fa[j]);
      if (i + j >= i1) rw = 0u;                                        // padding of the last group: byte value 0 adds nothing to the sums
      s0 = __dp4a(rw, 0x00000001u, s0);
      s1 = __dp4a(rw, 0x00000100u, s1);
      s2 = __dp4a(rw, 0x00010000u, s2);
      s3 = __dp4a(rw, 0x01000000u, s3);
    }
  }
  const uint32_t nb = G.col_negbias[c];
  const uint32_t lo = __vmins2(__viaddmax_s16x2(prmt(s0, s1, 0x5410u), nb, 0u), 0x00ff00ffu);
  const uint32_t hi = __vmins2(__viaddmax_s16x2(prmt(s2, s3, 0x5410u), nb, 0u), 0x00ff00ffu);
  const uint32_t a = prmt(lo, hi, 0x6420u);
  const uint32_t ao = sbase + G.off_A + G.col_arow[c] * 2 * ZB + kb;
  for (int r = 0; r < C; r++) {
    const uint32_t ra = cl_map(ao, (uint32_t)r);
    cl_st(ra, a);
    cl_st(ra + ZB, a);
  }
}

// grid = n_cb * C CTAs in clusters of (C, 1, 1): cluster q decodes code block q.  blockDim = 32 * S.T.
template <int ZWC>
__global__ void __launch_bounds__(32 * kClMaxWarps, 1)
ldpc_decode_cluster_kernel(const PackedGraph *__restrict__ gdev, const ClusterSched *__restrict__ sdev, DecodeArgs a)
{
  extern __shared__ __align__(16) uint32_t sm[];
  __shared__ PackedGraph G;
  __shared__ ClusterSched S;
  __shared__ __align__(8) uint8_t s_flags[2][8];
  __shared__ int s_flag;
  char *smb = reinterpret_cast<char *>(sm);
  for (int i = threadIdx.x; i < (int)(sizeof(PackedGraph) / 4); i += blockDim.x)
    reinterpret_cast<int *>(&G)[i] = reinterpret_cast<const int *>(gdev)[i];
  for (int i = threadIdx.x; i < (int)(sizeof(ClusterSched) / 4); i += blockDim.x)
    reinterpret_cast<int *>(&S)[i] = reinterpret_cast<const int *>(sdev)[i];
  if (threadIdx.x < 16) reinterpret_cast<uint8_t *>(s_flags)[threadIdx.x] = 0;
  __syncthreads();
  const int C = S.C;
  const uint32_t rank = cl_rank();
  const int cb = (int)blockIdx.x / C;
  const BlockIo io = block_io(a, cb);
  const int Zw = geo_zw<ZWC>(G);
  const uint32_t ZB = geo_zb<ZWC>(G), RSB = geo_rsb<ZWC>(G);
  const uint32_t lane = threadIdx.x & 31u, kb0 = 4u * lane;
  const int list = (int)rank * S.T + (int)(threadIdx.x >> 5);
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smb);

  // ---- R := 0 for the owned (row, chunk) items (nobody else ever touches those words)
  for (int i = S.cn_start[list]; i < S.cn_start[list + 1]; i++) {
    const int it = S.cn_items[i];
    const PackedRow row = G.rows[it & 0xFF];
    const uint32_t kb = kb0 + 128u * (uint32_t)(it >> 8), rb = row.rbase + kb;
    const int D = (int)((row.e0_deg >> 12) & 0xFFu);
    for (int j = 0; j < D; j++) {
      sts(smb, rb + j * RSB, kH);
      if (kb == 0u) sts(smb, rb + j * RSB + ZB, kH);
    }
  }
  // every CTA of the cluster is resident and past its own initialisation before anybody stores into it
  cl_sync();
  // ---- channel LLRs: each CTA fetches 1/C of them (global, or mapped host memory in the low-latency mode: the block crosses PCIe once)
  //      and stores them, offset binary, into the L rows of all C CTAs
  {
    const int W = G.ncols * Zw;
    const int per = (((W + C - 1) / C) + 31) & ~31;
    const int w0 = (int)rank * per, w1 = min(W, w0 + per);
    const int8_t *gl = io.llr;
    const bool al4 = ((reinterpret_cast<uintptr_t>(gl) & 3) == 0);
    for (int i = w0 + (int)threadIdx.x; i < w1; i += blockDim.x) {
      uint32_t w;
      if (al4) w = __ldg(reinterpret_cast<const uint32_t *>(gl) + i);
      else {
        const uint8_t *b = reinterpret_cast<const uint8_t *>(gl) + 4 * i;
        w = b[0] | (b[1] << 8) | (b[2] << 16) | ((uint32_t)b[3] << 24);
      }
      w ^= kH;
      const int c = i / Zw, k = i - c * Zw;
      const uint32_t lo = sbase + G.off_L + c * RSB + 4 * k;
      for (int r = 0; r < C; r++) {
        const uint32_t ra = cl_map(lo, (uint32_t)r);
        cl_st(ra, w);
        if (k == 0) cl_st(ra + ZB, w);
      }
    }
  }
  cl_sync();
  // ---- local: A := L (stored twice), P rows as in the single-CTA kernel
  for (int i = threadIdx.x; i < G.ncols * Zw; i += blockDim.x) {
    const int c = i / Zw, k = i - c * Zw, ar = G.col_arow[c];
    if (ar < 0) continue;
    const uint32_t w = lds(smb, G.off_L + c * RSB + 4 * k);
    const uint32_t ao = G.off_A + ar * 2 * ZB + 4 * k;
    sts(smb, ao, w);
    sts(smb, ao + ZB, w);
  }
  for (int i = S.cn_start[list]; i < S.cn_start[list + 1]; i++) {
    const int it = S.cn_items[i], r = it & 0xFF;
    const PackedRow row = G.rows[r];
    if (row.lrow == 0xFFFFFFFFu) continue;
    const uint32_t kb = kb0 + 128u * (uint32_t)(it >> 8);
    uint32_t w0 = kb + 4u * (uint32_t)G.row_p_q[r];
    if (w0 >= ZB) w0 -= ZB;
    const uint32_t lp = __funnelshift_r(lds(smb, row.lrow + w0), lds(smb, row.lrow + w0 + 4), (uint32_t)G.row_p_rho[r]);
    const uint32_t dd = vabsdiffu4(lp, kH);
    const uint32_t pa = (row.prow_pcw & 0xFFFFFFu) + kb;
    sts(smb, pa, 0u);
    sts(smb, pa + ZB, lop3<kLutOrAnd>(dd, msb_mask(dd), kL7) | (~lp & kH));
    sts(smb, pa + 2 * ZB, lp);
  }
  __syncthreads();

  const int maxIter = a.numMaxIter;
  int numIter = 0, par = 0;
  bool done = false, wrote = false;
  while (!done) {
    // ---- CN phase of iteration numIter + 1 (yields the syndrome of iteration numIter); rank 0 samples the caller's abort flag meanwhile
    uint32_t bad = 0, ab = 0;
    if (io.abort && rank == 0 && threadIdx.x == 0) ab = *io.abort;
    for (int i = S.cn_start[list]; i < S.cn_start[list + 1]; i++) {
      const int it = S.cn_items[i];
      const uint32_t kb = kb0 + 128u * (uint32_t)(it >> 8);
      cn_dispatch<ZWC, false>(G, smb, it & 0xFF, kb, kb == 0u, numIter == 0, bad);
    }
    const int bad_cta = __syncthreads_or(bad != 0);
    if (threadIdx.x == 0) {
      const uint32_t v = (bad_cta ? 1u : 0u) | (ab ? 2u : 0u);
      const uint32_t fa = (uint32_t)__cvta_generic_to_shared(&s_flags[par][rank]);
      for (int r = 0; r < C; r++) cl_st8(cl_map(fa, (uint32_t)r), v);
    }
    cl_sync();                                                        // every R word of this iteration is in place, every verdict delivered
    const uint2 f = *reinterpret_cast<const uint2 *>(s_flags[par]);
    par ^= 1;
    const uint32_t any = f.x | f.y;
    if (numIter >= 2 && !a.use_crc && (any & 0x01010101u) == 0u) break;   // iteration numIter passed its parity check (nrLDPC_decoder.c:552)
    numIter++;
    if (numIter >= 2 && (any & 0x02020202u)) { numIter = maxIter + 2; break; }   // check_abort at the top of the iteration (:557-560)
    // ---- BN phase
    for (int i = S.bn_start[list]; i < S.bn_start[list + 1]; i++) {
      const int it = S.bn_items[i];
      bn_col_cluster<ZWC>(G, S, smb, sbase, it & 0xFF, kb0 + 128u * (uint32_t)(it >> 8), lane, C);
    }
    cl_sync();                                                        // every replica of A is complete; all pulls of R are done
    if (numIter == 1) {
      if (!(1 <= maxIter)) done = true;
    } else if (a.use_crc) {
      if (numIter > 2) {                                              // :850-862; every CTA evaluates the same CRC on its own replica
        wrote = true;                                                 // the reference rewrites p_out here every time: only the last state is observable
        if (packed_crc_check(G, smb, a, &s_flag)) break;
      }
      if (!(numIter <= maxIter)) done = true;
    } else if (!(numIter <= maxIter)) done = true;
  }
  if (!a.use_crc || wrote) packed_write_output(G, smb, a, io.out, (int)rank, C);
  block_finish(io, a, numIter, (int)rank);
}

}  // namespace nrb200
