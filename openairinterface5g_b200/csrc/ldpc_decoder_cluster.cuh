// Low-latency variant of the packed flooding decoder: ONE code block is decoded by a thread-block CLUSTER of C CTAs (2, 4 or 8 SMs)
// instead of one CTA, for the calls where a handful of code blocks arrive at a time and the time of one block -- not the rate of a
// thousand -- is what the caller sees: the per-segment LDPCdecoder calls of OAI's tpool workers (nr_ulsch_decoding.c:435-468), ldpctest's
// serial loop (ldpctest.c:329-340), a single slot's 28-52 code blocks.  Same arithmetic, same schedule, same results as
// ldpc_decode_packed_kernel (bit exact with nrLDPC_decoder.c:206-881); what changes is where the state lives:
//
//   every CTA   holds a full replica of A (a-posteriori LLRs) and L (channel LLRs); the shared-memory image has the single-CTA layout
//   bit columns are owned whole by one CTA (ClusterSched::col_rank); check-row work items (row, 32-word chunk) by any warp of any CTA
//   CN phase    reads the local A replica and the row's own R words (local), writes each new cn->bn message locally (it is next
//               iteration's R_old) AND into the R row copy of the CTA that owns the edge's bit column (st.shared::cluster, fire and forget)
//   BN phase    all local reads -- the column owner holds every message of its columns -- then the new A word is stored into all C replicas
//   barriers    two cluster barriers per iteration (messages delivered -> BN may read; A complete -> CN may read); the parity-check / abort
//               verdicts travel as one byte per CTA written into every CTA's flag word before the first of them
//
// DSMEM traffic per iteration and cluster: ~105 KB of messages pushed + 20 KB x C of A broadcast.
#pragma once
#include "ldpc_cluster.h"
#include "ldpc_decoder_packed.cuh"

namespace nrb200 {

// Phase marks (DecodeArgs::quirks bit 1, NRB200_CLUSTER_TIMERS=1): thread 0 of every CTA of code block 0 records clock64() at the phase
// boundaries; tools/cluster_phases.py reads them back through nrb200_debug_cluster_marks().
__device__ long long g_cl_marks[kClMaxCtas * 64];
#define NRB200_CL_MARK() do { if (timers && threadIdx.x == 0 && cb == 0) { if (mark_n < 64) g_cl_marks[rank * 64 + mark_n] = clock64(); mark_n++; } } while (0)

// One 16-word piece of a HEAVY check row (D = 2 * DH or 2 * DH - 1 stored edges) on a whole warp: lanes 0-15 take edges 0 .. DH-1, lanes 16-31
// edges DH .. D-1 of the same 16 words; the halves swap their partial two-minimum state, sign product and syndrome by shuffle and merge them
// (twomin_merge), then each half writes back its own edges.  Same results as cn_row<D>, half the dependent chain.
template <int ZWC, int DH>
__device__ __forceinline__ void cn_row_split(const PackedGraph &G, char *__restrict__ smb, const PackedRow &row, uint32_t kb, uint32_t half, bool halo,
                                             bool first_iter, uint32_t &bad, uint32_t sbase)
{
  const uint32_t one = G.one, mone = 0u - one;
  const uint32_t ZB = geo_zb<ZWC>(G), RSB = geo_rsb<ZWC>(G);
  const uint32_t D = (row.e0_deg >> 12) & 0xFFu;
  const bool dummy = half != 0u && (D & 1u);                             // the last slot of the upper half does not exist when D is odd
  const uint32_t e0 = (row.e0_deg & 0xFFFu) + half * DH;
  const uint32_t rb = row.rbase + half * DH * RSB + kb;
  uint32_t q[DH];
  uint32_t sgn = 0u, synd = (half == 0u && (D & 1u)) ? kH : 0u;
  TwoMin tm = twomin_init();
#pragma unroll
  for (int j = 0; j < DH; j++) {
    const bool skip = (j == DH - 1) && dummy;
    const uint32_t e = skip ? e0 + j - 1 : e0 + j;                        // keep the loads in bounds
    const uint2 d = *reinterpret_cast<const uint2 *>(G.cn_desc[e]);
    const uint32_t aa = add_fma(kb, d.x, one);
    uint32_t aw = __funnelshift_r(lds(smb, aa), lds(smb, aa + 4), d.y);
    const uint32_t ro = lds(smb, skip ? rb + (j - 1) * RSB : rb + j * RSB);
    uint32_t mag;
    cn_input(aw, ro, mone, mag, q[j]);
    if (skip) { mag = kL7; q[j] = 0u; aw = 0u; }
    synd ^= aw;
    sgn ^= q[j];
    twomin(mag, tm, one, mone);
  }
  {
    TwoMin o;
    o.n1 = __shfl_xor_sync(0xffffffffu, tm.n1, 16);
    o.n2p = __shfl_xor_sync(0xffffffffu, tm.n2p, 16);
    sgn ^= __shfl_xor_sync(0xffffffffu, sgn, 16);
    synd ^= __shfl_xor_sync(0xffffffffu, synd, 16);
    twomin_merge(tm, o, one, mone);
  }
  cn_row_neighbour<ZWC, false>(G, smb, row, kb, first_iter, 0u, tm, sgn, synd, bad);   // both halves: same values to the same words
  const uint32_t p1 = twomin_min1(tm, mone) | kH, p2 = twomin_min2(tm, mone) | kH;
#pragma unroll
  for (int j = 0; j < DH; j++) {
    if ((j == DH - 1) && dummy) continue;
    const uint32_t rn = make_r(q[j], tm.n1, p1, p2, sgn, one, mone);
    sts(smb, rb + j * RSB, rn);
    const uint32_t owner = prmt(G.cn_desc[e0 + j][1], 0u, 0x4441u);
    const uint32_t ra = cl_map(sbase + rb + j * RSB, owner), mb = cl_mbar_at(0, owner);
    cl_push(ra, rn, mb);
    cl_push_if(halo, ra + ZB, rn, mb);
  }
}

// One 16-word piece of a HEAVY bit column on a whole warp: each half sums half of the column's messages, the partial sums meet by shuffle,
// and each half stores the new word into half of the C replicas.
template <int ZWC>
__device__ __forceinline__ void bn_col_split(const PackedGraph &G, char *__restrict__ smb, int c, uint32_t kb, uint32_t half, int C, uint32_t sbase)
{
  const uint32_t ZB = geo_zb<ZWC>(G);
  const int i0 = G.col_start[c], i1 = G.col_start[c + 1], mid = i0 + ((i1 - i0 + 1) >> 1);
  const uint32_t lw = half ? 0u : lds(smb, G.off_L + c * geo_rsb<ZWC>(G) + kb);
  uint32_t s0 = __dp4a(lw, 0x00000001u, 0u), s1 = __dp4a(lw, 0x00000100u, 0u), s2 = __dp4a(lw, 0x00010000u, 0u), s3 = __dp4a(lw, 0x01000000u, 0u);
  int i = half ? mid : i0;
  const int ie = half ? i1 : mid;
  for (; i + 2 <= ie; i += 2) { bn_edge<ZWC>(G, smb, i, kb, s0, s1, s2, s3); bn_edge<ZWC>(G, smb, i + 1, kb, s0, s1, s2, s3); }
  if (i < ie) bn_edge<ZWC>(G, smb, i, kb, s0, s1, s2, s3);
  s0 += __shfl_xor_sync(0xffffffffu, s0, 16);
  s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
  s2 += __shfl_xor_sync(0xffffffffu, s2, 16);
  s3 += __shfl_xor_sync(0xffffffffu, s3, 16);
  const uint32_t nb = G.col_negbias[c];
  const uint32_t lo = __vmins2(__viaddmax_s16x2(prmt(s0, s1, 0x5410u), nb, 0u), 0x00ff00ffu);
  const uint32_t hi = __vmins2(__viaddmax_s16x2(prmt(s2, s3, 0x5410u), nb, 0u), 0x00ff00ffu);
  const uint32_t a = prmt(lo, hi, 0x6420u);
  const uint32_t ao = sbase + G.off_A + G.col_arow[c] * 2 * ZB + kb;
  for (int r = (int)half; r < C; r += 2) {
    const uint32_t ra = cl_map(ao, (uint32_t)r), mb = cl_mbar_at(1, (uint32_t)r);
    cl_push(ra, a, mb);
    cl_push(ra + ZB, a, mb);
  }
}

// grid = n_cb * C CTAs in clusters of (C, 1, 1): cluster q decodes code block q.  blockDim = 32 * S.T.
// MAXT = launch bound: 512 threads (the 4- and 8-CTA clusters run 12-16 warps per CTA) leave the compiler 128 registers per thread, 768 leave 80.
template <int ZWC, int MAXT>
__global__ void __launch_bounds__(MAXT, 1)
ldpc_decode_cluster_kernel(const PackedGraph *__restrict__ gdev, const ClusterSched *__restrict__ sdev, DecodeArgs a)
{
  extern __shared__ __align__(16) uint32_t sm[];
  __shared__ __align__(16) PackedGraph G;
  __shared__ __align__(16) ClusterSched S;
  __shared__ __align__(16) uint32_t s_flags[2][8];                       // per iteration parity: one verdict word per CTA of the cluster
  __shared__ int s_flag;
  char *smb = reinterpret_cast<char *>(sm);
  static_assert(sizeof(PackedGraph) % 16 == 0 && sizeof(ClusterSched) % 16 == 0, "tables are copied with 16-byte accesses");
  {
    // both tables with 16-byte loads, all of a thread's loads in flight before its first store
    constexpr int nG = (int)(sizeof(PackedGraph) / 16), nS = (int)(sizeof(ClusterSched) / 16);
    const uint4 *g4 = reinterpret_cast<const uint4 *>(gdev), *s4 = reinterpret_cast<const uint4 *>(sdev);
    uint4 v[3];
    const int t = (int)threadIdx.x, n = (int)blockDim.x;
#pragma unroll
    for (int k = 0; k < 3; k++) { const int i = t + k * n; if (i < nG + nS) v[k] = i < nG ? __ldg(g4 + i) : __ldg(s4 + (i - nG)); }
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const int i = t + k * n;
      if (i < nG) reinterpret_cast<uint4 *>(&G)[i] = v[k];
      else if (i < nG + nS) reinterpret_cast<uint4 *>(&S)[i - nG] = v[k];
    }
    for (int i = t + 3 * n; i < nG + nS; i += n) {
      if (i < nG) reinterpret_cast<uint4 *>(&G)[i] = __ldg(g4 + i);
      else reinterpret_cast<uint4 *>(&S)[i - nG] = __ldg(s4 + (i - nG));
    }
  }
  if (threadIdx.x < 16) reinterpret_cast<uint32_t *>(s_flags)[threadIdx.x] = 0;
  if (NRB200_CLUSTER_ASYNC && threadIdx.x == 0) cl_mbar_init();           // before the first cluster barrier: nobody stores into this CTA earlier
  __syncthreads();
  const int C = S.C;
  const uint32_t rank = cl_rank();
  const int cb = (int)blockIdx.x / C;
  const BlockIo io = block_io(a, cb);
  block_begin(io, (int)rank);
  const int Zw = geo_zw<ZWC>(G);
  const uint32_t ZB = geo_zb<ZWC>(G), RSB = geo_rsb<ZWC>(G);
  const uint32_t lane = threadIdx.x & 31u, kb0 = 4u * lane;
  const int list = (int)rank * S.T + (int)(threadIdx.x >> 5);
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smb);
  const bool timers = (a.quirks & 2) != 0;
  int mark_n = 0;
  NRB200_CL_MARK();

  // ---- channel LLRs: each CTA fetches 1/C of them (global, or mapped host memory in the low-latency mode: the block crosses PCIe once).
  //      The loads are issued now and consumed after the first cluster barrier, so their latency hides behind the initialisation.
  const int W = G.ncols * Zw;
  const int per = (((W + C - 1) / C) + 31) & ~31;
  const int w0 = (int)rank * per, w1 = min(W, w0 + per);
  const int8_t *gl = io.llr;
  const bool al4 = ((reinterpret_cast<uintptr_t>(gl) & 3) == 0);
  auto fetch = [&](int i) -> uint32_t {
    if (al4) return __ldg(reinterpret_cast<const uint32_t *>(gl) + i);
    const uint8_t *b = reinterpret_cast<const uint8_t *>(gl) + 4 * i;
    return b[0] | (b[1] << 8) | (b[2] << 16) | ((uint32_t)b[3] << 24);
  };
  constexpr int kPre = 4;
  uint32_t pre[kPre];
#pragma unroll
  for (int k = 0; k < kPre; k++) { const int i = w0 + (int)threadIdx.x + k * (int)blockDim.x; pre[k] = i < w1 ? fetch(i) : 0u; }
  // ---- where each stored edge's messages go: the owner of its bit column (read by cn_row<.., CL = true> from bits[10:8] of cn_desc .y)
  for (int m = threadIdx.x; m < G.nreal; m += blockDim.x) G.cn_desc[m][1] |= (uint32_t)S.edge_rank[m] << 8;
  // ---- R := 0 everywhere (the rows of owned columns receive the producers' messages, the owned row chunks are this CTA's own R_old)
  {
    const uint4 z = make_uint4(kH, kH, kH, kH);
    uint4 *r4 = reinterpret_cast<uint4 *>(smb + G.off_R);
    const int n4 = (G.nreal * (int)RSB) >> 4;
    for (int i = threadIdx.x; i < n4; i += blockDim.x) r4[i] = z;
  }
  // every CTA of the cluster is resident and past its own initialisation before anybody stores into it
  cl_sync();
  NRB200_CL_MARK();
  // ---- the fetched words, offset binary, into the L rows of all C CTAs
  {
    auto spread = [&](int i, uint32_t w) {
      w ^= kH;
      const int c = i / Zw, k = i - c * Zw;
      const uint32_t lo = sbase + G.off_L + c * RSB + 4 * k;
      for (int r = 0; r < C; r++) {
        const uint32_t ra = cl_map(lo, (uint32_t)r);
        cl_st(ra, w);
        if (k == 0) cl_st(ra + ZB, w);
      }
    };
#pragma unroll
    for (int k = 0; k < kPre; k++) { const int i = w0 + (int)threadIdx.x + k * (int)blockDim.x; if (i < w1) spread(i, pre[k]); }
    for (int i = w0 + (int)threadIdx.x + kPre * (int)blockDim.x; i < w1; i += blockDim.x) spread(i, fetch(i));
  }
  cl_sync();
  NRB200_CL_MARK();
  // ---- local: A := L (stored twice, 16 bytes at a time), P rows of the owned row chunks as in the single-CTA kernel
  {
    const int q4 = Zw >> 2;
    for (int i = threadIdx.x; i < G.ncols * q4; i += blockDim.x) {
      const int c = i / q4, k = i - c * q4, ar = G.col_arow[c];
      if (ar < 0) continue;
      const uint4 w = *reinterpret_cast<const uint4 *>(smb + G.off_L + c * RSB + 16 * k);
      char *ao = smb + G.off_A + ar * 2 * ZB + 16 * k;
      *reinterpret_cast<uint4 *>(ao) = w;
      *reinterpret_cast<uint4 *>(ao + ZB) = w;
    }
  }
  for (int i = S.cn_start[list]; i < S.cn_start[list + 1]; i++) {
    const int it = S.cn_items[i], r = it & 0xFF;
    const PackedRow row = G.rows[r];
    if (row.lrow == 0xFFFFFFFFu) continue;
    const uint32_t kb = (it & kClSplitItem) ? 4u * (lane & 15u) + 64u * (uint32_t)((it >> 8) & 7) : kb0 + 128u * (uint32_t)(it >> 8);
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
  NRB200_CL_MARK();

  const int maxIter = a.numMaxIter;
  int numIter = 0, par = 0;
  bool done = false, wrote = false;
  uint32_t phase = 0;                                                    // parity of the two mbarriers' current phase (both complete once per iteration)
  while (!done) {
    // ---- CN phase of iteration numIter + 1 (yields the syndrome of iteration numIter); rank 0 samples the caller's abort flag meanwhile
    uint32_t bad = 0, ab = 0;
    if (io.abort && rank == 0 && threadIdx.x == 0) ab = *io.abort;
    if (NRB200_CLUSTER_ASYNC && threadIdx.x == 0) cl_mbar_expect(0, S.cn_tx[rank]);   // what this phase delivers to this CTA: messages of its columns + C verdicts
    for (int i = S.cn_start[list]; i < S.cn_start[list + 1]; i++) {
      const int it = S.cn_items[i];
      if (it & kClSplitItem) {
        const uint32_t kb = 4u * (lane & 15u) + 64u * (uint32_t)((it >> 8) & 7);
        cn_row_split<ZWC, (kClSplitRowDeg + 1) / 2>(G, smb, G.rows[it & 0xFF], kb, lane >> 4, kb == 0u, numIter == 0, bad, sbase);
      } else {
        const uint32_t kb = kb0 + 128u * (uint32_t)(it >> 8);
        cn_dispatch<ZWC, false, true>(G, smb, it & 0xFF, kb, kb == 0u, numIter == 0, bad, sbase);
      }
    }
    NRB200_CL_MARK();
    const int bad_cta = __syncthreads_or(bad != 0);
    if (threadIdx.x == 0) {
      // the verdict goes out after the CTA-wide barrier above: its arrival tells the receiver that EVERY warp of this CTA is done with the phase
      const uint32_t v = (bad_cta ? 1u : 0u) | (ab ? 2u : 0u);
      const uint32_t fa = (uint32_t)__cvta_generic_to_shared(&s_flags[par][rank]);
      for (int r = 0; r < C; r++) cl_push(cl_map(fa, (uint32_t)r), v, cl_mbar_at(0, (uint32_t)r));
    }
    if (NRB200_CLUSTER_ASYNC) { if (!cl_mbar_wait(0, phase)) __trap(); }
    else cl_sync();                                                   // every message of this iteration is delivered, every verdict too
    NRB200_CL_MARK();
    const uint4 f0 = *reinterpret_cast<const uint4 *>(&s_flags[par][0]), f1 = *reinterpret_cast<const uint4 *>(&s_flags[par][4]);
    par ^= 1;
    const uint32_t any = f0.x | f0.y | f0.z | f0.w | f1.x | f1.y | f1.z | f1.w;
    if (numIter >= 2 && !a.use_crc && (any & 1u) == 0u) break;          // iteration numIter passed its parity check (nrLDPC_decoder.c:552)
    numIter++;
    if (numIter >= 2 && (any & 2u)) { numIter = maxIter + 2; break; }    // check_abort at the top of the iteration (:557-560)
    if (NRB200_CLUSTER_ASYNC && threadIdx.x == 0) cl_mbar_expect(1, S.bn_tx);
    // ---- BN phase: the columns this CTA owns, all reads local, the new word into every replica of A
    for (int i = S.bn_start[list]; i < S.bn_start[list + 1]; i++) {
      const int it = S.bn_items[i];
      if (it & kClSplitItem) bn_col_split<ZWC>(G, smb, it & 0xFF, 4u * (lane & 15u) + 64u * (uint32_t)((it >> 8) & 7), lane >> 4, C, sbase);
      else bn_col<ZWC>(G, smb, it & 0xFF, kb0 + 128u * (uint32_t)(it >> 8), C, sbase);
    }
    NRB200_CL_MARK();
    if (NRB200_CLUSTER_ASYNC) { if (!cl_mbar_wait(1, phase)) __trap(); phase ^= 1u; }
    else cl_sync();                                                   // every replica of A is complete
    NRB200_CL_MARK();
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
  NRB200_CL_MARK();
}

}  // namespace nrb200
