// Packed flooding int8 min-sum decoder for lifting sizes that are a multiple of 4 (every NR Z >= 32 and the hot
// Z = 384 case): one CTA per code block, four lifts per 32-bit register (byte SIMD-in-word), all state in shared memory.
//
// Schedule (bit exact with the reference's two-phase flooding decoder, nrLDPC_decoder.c:206-881):
//   state    R[m][t]   = cn->bn message of edge slot m at check lift t            (cnProcBufRes)
//            A[c][v]   = a-posteriori LLR of bit (c, v), degree >= 2 columns       (llrRes)
//            L[c][v]   = channel LLR                                               (llrProcBuf)
//   CN phase thread (row r, word k): for every edge  Q = subs_epi8(A[c][t+s], R_old)  -- what bnProc/bn2cnProcBuf
//            produced at the end of the previous iteration (nrLDPC_bnProc.h:325), formed on the fly from the rotated A
//            word, so no bn->cn buffer and no circular copies exist -- then exclude-self min / sign product
//            (nrLDPC_cnProc.h:388-877) written back in place.  The sign bytes of the same A words give the previous
//            iteration's syndrome (nrLDPC_cnProc.h:887-1960) for free: sign(adds_epi8(Q, R)) == sign(A)  (DESIGN.md).
//   BN phase thread (column c, word k): A = sat8(L + sum_e R[m_e][v - s_e])       (nrLDPC_bnProc.h:40-263)
// The quantisation points are the reference's: int16 sum -> sat8 -> subs_epi8 -> |.| clipped to 127.
//
// Number format: every byte in shared memory is OFFSET BINARY (value + 128).  Then
//   |A - R|          is one VABSDIFF4.U8 (the only byte-SIMD ALU op sm_100 has in hardware),
//   min(|Q|,127)     == min(|A - R|, 127): saturating the difference first (subs_epi8) never changes the clipped magnitude,
//   sign(Q)          == (A' < R') unsigned, 4 logic ops,
//   sum of messages  is a plain 32-bit add of zero-extended byte pairs (no lane can overflow), bias removed once per column.
#pragma once
#include "ldpc_common.cuh"
#include "ldpc_packed_graph.h"
#include "ldpc_packed_simd.cuh"   // cn_input / twomin / make_r and the LOP3 / PRMT forms (host-checked: tests/host/packed_simd_check.cc)

namespace nrb200 {

__device__ __forceinline__ uint32_t lds(const char *smb, uint32_t off) { return *reinterpret_cast<const uint32_t *>(smb + off); }
__device__ __forceinline__ uint2 lds2(const char *smb, uint32_t off) { return *reinterpret_cast<const uint2 *>(smb + off); }
__device__ __forceinline__ void sts(char *smb, uint32_t off, uint32_t v) { *reinterpret_cast<uint32_t *>(smb + off) = v; }

// ---- distributed shared memory (thread-block clusters): the cluster decoder (ldpc_decoder_cluster.cuh) pushes every new cn->bn message into
//      the CTA that owns the edge's bit column
__device__ __forceinline__ uint32_t cl_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cl_map(uint32_t saddr, uint32_t rank)
{
  uint32_t r;
  asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));      // a pure function of its inputs: the compiler may share it
  return r;
}
__device__ __forceinline__ void cl_st(uint32_t caddr, uint32_t v) { asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(caddr), "r"(v) : "memory"); }
__device__ __forceinline__ void cl_st8(uint32_t caddr, uint32_t v) { asm volatile("st.shared::cluster.u8 [%0], %1;" ::"r"(caddr), "r"(v) : "memory"); }
// Message delivery without fences (compile with -DNRB200_CLUSTER_ASYNC=1): every word that crosses the cluster becomes an ASYNCHRONOUS store that completes
// bytes on an mbarrier of the destination CTA (st.async ... mbarrier::complete_tx::bytes, SASS STAS).  Each CTA knows how many bytes a phase delivers to it
// (ClusterSched::cn_tx / bn_tx), arms its barrier with that count and waits for the phase to complete; no barrier.cluster (MEMBAR.ALL.GPU + CCTL.IVALL each)
// is needed inside the iteration.  g_cl_mbar[0]: the cn->bn messages and the CTAs' verdict words of the check-node phase; [1]: the A words of the bit-node
// phase.  Bit exact (tests/test_gpu_ldpc_lowlat.py passes with it) and MEASURED NO FASTER than the two cluster barriers: 39.3 us per block on 8 SMs either way
// (A/B in one run, r02) -- STAS cannot be predicated, so the halo word costs a divergent branch per edge, and what the barriers cost without the phase
// timers is less than the timers suggest.  The plain-store path stays the default; this one is kept because it removes every fence from the loop, which is
// what a larger cluster or a slower fabric would need.
#ifndef NRB200_CLUSTER_ASYNC
#define NRB200_CLUSTER_ASYNC 0
#endif
__shared__ __align__(8) unsigned long long g_cl_mbar[2];
__device__ __forceinline__ uint32_t cl_mbar_at(int which, uint32_t rank) { return cl_map((uint32_t)__cvta_generic_to_shared(&g_cl_mbar[which]), rank); }
// store v at the cluster address caddr of CTA `rank`; which = the phase's barrier
__device__ __forceinline__ void cl_push(uint32_t caddr, uint32_t v, uint32_t mbar_caddr)
{
#if NRB200_CLUSTER_ASYNC
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(caddr), "r"(v), "r"(mbar_caddr) : "memory");
#else
  (void)mbar_caddr;
  cl_st(caddr, v);
#endif
}
// the same store under a predicate (the halo word: lane 0 of a row's first chunk only) -- written as predicated PTX so that it stays one predicated STAS
// instead of a divergent branch per edge
__device__ __forceinline__ void cl_push_if(bool on, uint32_t caddr, uint32_t v, uint32_t mbar_caddr)
{
#if NRB200_CLUSTER_ASYNC
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\t@p st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];\n\t}"
               ::"r"(caddr), "r"(v), "r"(mbar_caddr), "r"((uint32_t)on) : "memory");
#else
  (void)mbar_caddr;
  if (on) cl_st(caddr, v);
#endif
}
__device__ __forceinline__ void cl_mbar_init()
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&g_cl_mbar[0])) : "memory");
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&g_cl_mbar[1])) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void cl_mbar_expect(int which, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&g_cl_mbar[which])), "r"(bytes) : "memory");
}
// returns false when the phase has not completed after ~2^22 polls (a byte count that does not add up): the caller traps instead of hanging the GPU
__device__ __forceinline__ bool cl_mbar_wait(int which, uint32_t parity)
{
  const uint32_t mb = (uint32_t)__cvta_generic_to_shared(&g_cl_mbar[which]);
  for (int spin = 0; spin < (1 << 22); spin++) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(mb), "r"(parity) : "memory");
    if (ok) return true;
  }
  return false;
}
__device__ __forceinline__ void cl_sync()
{
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// Row geometry: ZWC = Z / 4 fixed at compile time (the hot Z = 384 instantiation: every R / L / P row offset becomes an immediate
// of the LDS / STS instead of a multiply-add per edge), ZWC = 0 reads it from the graph tables (every other lifting size).
template <int ZWC> __device__ __forceinline__ int geo_zw(const PackedGraph &G) { return ZWC ? ZWC : G.Zw; }
template <int ZWC> __device__ __forceinline__ uint32_t geo_zb(const PackedGraph &G) { return ZWC ? 4u * ZWC : (uint32_t)G.ZB; }
template <int ZWC> __device__ __forceinline__ uint32_t geo_rsb(const PackedGraph &G) { return ZWC ? 4u * (ZWC + 4) : (uint32_t)G.RSB; }

// the row's degree-1 neighbour (extension parity column, no stored message: its Q is the channel LLR forever) and the row's parity check
template <int ZWC, bool QUIRK>
__device__ __forceinline__ void cn_row_neighbour(const PackedGraph &G, char *__restrict__ smb, const PackedRow &row, uint32_t kb, bool first_iter,
                                                 uint32_t quirk_zero, TwoMin &tm, uint32_t &sgn, uint32_t &synd, uint32_t &bad)
{
  const uint32_t one = G.one, mone = 0u - one;
  const uint32_t ZB = geo_zb<ZWC>(G);
  if (row.lrow != 0xFFFFFFFFu) {
    const uint32_t pa = (row.prow_pcw & 0xFFFFFFu) + kb;
    const uint32_t qsm = lds(smb, pa + ZB);                              // precomputed sign-magnitude of the channel LLR
    synd ^= lds(smb, pa);                                                  // sign(llr + R_p) of the previous iteration
    sgn ^= qsm;
    twomin(qsm & kL7, tm, one, mone);
    uint32_t rp = make_r(qsm, tm.n1, twomin_min1(tm, mone) | kH, twomin_min2(tm, mone) | kH, sgn, one, mone);
    if (QUIRK) rp = (rp & ~quirk_zero) | (kH & quirk_zero);
    // adds_epi8(llr, R_p) < 0  <=>  L' + R' < 256  <=>  no carry out of the byte
    const uint32_t lp = lds(smb, pa + 2 * ZB);                           // the neighbour's channel LLR + 128, already rotated
    const uint32_t x = (lp & kL7) + (rp & kL7);
    sts(smb, pa, lop3<kLutMajNot>(lp, rp, x) & kH);
  }
  if (!first_iter && (kb >> 2) < (row.prow_pcw >> 24)) bad |= synd & kH;
}

// CL (cluster decoder): sbase = shared-window address of smb; bits[10:8] of an edge's cn_desc .y hold the cluster rank of the CTA that owns the
// edge's bit column, and every new message is also stored into that CTA's copy of the R row (same offset), where its bit-node phase reads it.
template <int ZWC, int D, bool QUIRK, bool CL = false>
__device__ __forceinline__ void cn_row(const PackedGraph &G, char *__restrict__ smb, const PackedRow &row, uint32_t kb, bool halo,
                                       bool first_iter, uint32_t quirk_zero, uint32_t &bad, uint32_t sbase = 0u)
{
  const uint32_t one = G.one, mone = 0u - one;
  const uint32_t ZB = geo_zb<ZWC>(G), RSB = geo_rsb<ZWC>(G);
  const uint32_t e0 = row.e0_deg & 0xFFFu;
  const uint32_t rb = row.rbase + kb;
  uint32_t q[D];
  uint32_t sgn = 0u, synd = (D & 1) ? kH : 0u;                           // sign(A) < 0 <=> bit 7 of A' clear
  TwoMin tm = twomin_init();
  uint32_t aprev = 0u, qprev = 0u;
#pragma unroll
  for (int j = 0; j < D; j++) {
    const uint2 d = *reinterpret_cast<const uint2 *>(G.cn_desc[e0 + j]);
    const uint32_t aa = add_fma(kb, d.x, one);
    const uint32_t aw = __funnelshift_r(lds(smb, aa), lds(smb, aa + 4), d.y);   // A' at lifts t+s .. t+s+3
    const uint32_t ro = lds(smb, rb + j * RSB);
    uint32_t mag;
    cn_input(aw, ro, mone, mag, q[j]);
    if (j & 1) {                                                                 // XORs of two edges at a time: one LOP3 each
      synd = lop3<kLutXor3>(synd, aprev, aw);
      sgn = lop3<kLutXor3>(sgn, qprev, q[j]);
    }
    aprev = aw;
    qprev = q[j];
    twomin(mag, tm, one, mone);
  }
  if (D & 1) { synd ^= aprev; sgn ^= qprev; }
  cn_row_neighbour<ZWC, QUIRK>(G, smb, row, kb, first_iter, quirk_zero, tm, sgn, synd, bad);
  const uint32_t p1 = twomin_min1(tm, mone) | kH, p2 = twomin_min2(tm, mone) | kH;
#pragma unroll
  for (int j = 0; j < D; j++) {
    uint32_t rn = make_r(q[j], tm.n1, p1, p2, sgn, one, mone);
    if (QUIRK) rn = (rn & ~quirk_zero) | (kH & quirk_zero);
    sts(smb, rb + j * RSB, rn);
    if (CL) {
      const uint32_t owner = prmt(G.cn_desc[e0 + j][1], 0u, 0x4441u);
      const uint32_t ra = cl_map(sbase + rb + j * RSB, owner), mb = cl_mbar_at(0, owner);
      cl_push(ra, rn, mb);
      cl_push_if(halo, ra + ZB, rn, mb);
    } else if (halo) sts(smb, rb + j * RSB + ZB, rn);
  }
}

// The same row as a LOOP over its edges: the fallback for row degrees without an unrolled instantiation, and an experiment.  The decoder's
// instruction stream is straight-line code, and the micro-benchmark (tools/ubench/alu_ceiling.cu) sustains 0.76 warp instructions per cycle and
// scheduler on this opcode blend while the code stays below ~32 KB but 0.66 at the ~50 KB the unrolled rows of one iteration add up to -- the rate
// the kernel runs at.  A loop cannot keep the inputs Q in registers between the two passes, so pass 1 parks the sign-magnitude word in the edge's
// own R slot (R_old is dead once read; only this thread touches this word during the CN phase) and pass 2 reads it back.  Measured: the ~25 % more
// instructions cost more than the smaller footprint returns (all rows looped 0.79 ms, degrees 7-19 looped 0.72 ms, all unrolled 0.70 ms per 1024
// blocks), so every NR row degree stays unrolled.
template <int ZWC, bool QUIRK, bool CL = false>
__device__ __forceinline__ void cn_row_loop(const PackedGraph &G, char *__restrict__ smb, const PackedRow &row, uint32_t kb, bool halo,
                                         bool first_iter, uint32_t quirk_zero, uint32_t &bad, uint32_t sbase = 0u)
{
  const uint32_t one = G.one, mone = 0u - one;
  const uint32_t ZB = geo_zb<ZWC>(G), RSB = geo_rsb<ZWC>(G);
  const int D = (int)((row.e0_deg >> 12) & 0xFFu);
  const uint32_t e0 = row.e0_deg & 0xFFFu;
  const uint32_t rb = row.rbase + kb;
  uint32_t sgn = 0u, synd = (D & 1) ? kH : 0u;
  TwoMin tm = twomin_init();
  {
    const uint32_t *dp = G.cn_desc[e0];
    uint32_t ra = rb;
#pragma unroll 1
    for (int j = 0; j < D; j++, dp += 2, ra += RSB) {
      const uint2 d = *reinterpret_cast<const uint2 *>(dp);
      const uint32_t aa = add_fma(kb, d.x, one);
      const uint32_t aw = __funnelshift_r(lds(smb, aa), lds(smb, aa + 4), d.y);
      const uint32_t ro = lds(smb, ra);
      uint32_t mag, qsm;
      cn_input(aw, ro, mone, mag, qsm);
      sts(smb, ra, qsm);
      synd ^= aw;
      sgn ^= qsm;
      twomin(mag, tm, one, mone);
    }
  }
  cn_row_neighbour<ZWC, QUIRK>(G, smb, row, kb, first_iter, quirk_zero, tm, sgn, synd, bad);
  const uint32_t p1 = twomin_min1(tm, mone) | kH, p2 = twomin_min2(tm, mone) | kH;
  uint32_t ra = rb;
#pragma unroll 1
  for (int j = 0; j < D; j++, ra += RSB) {
    uint32_t rn = make_r(lds(smb, ra), tm.n1, p1, p2, sgn, one, mone);
    if (QUIRK) rn = (rn & ~quirk_zero) | (kH & quirk_zero);
    sts(smb, ra, rn);
    if (CL) {
      const uint32_t owner = prmt(G.cn_desc[e0 + j][1], 0u, 0x4441u);
      const uint32_t rr = cl_map(sbase + ra, owner), mb = cl_mbar_at(0, owner);
      cl_push(rr, rn, mb);
      cl_push_if(halo, rr + ZB, rn, mb);
    } else if (halo) sts(smb, ra + ZB, rn);
  }
}

template <int ZWC, bool QUIRK, bool CL = false>
__device__ __forceinline__ void cn_dispatch(const PackedGraph &G, char *smb, int r, uint32_t kb, bool halo, bool first_iter, uint32_t &bad, uint32_t sbase = 0u)
{
  const PackedRow row = G.rows[r];
  uint32_t qz = 0u;
  if (QUIRK && (row.e0_deg >> 20)) {
    // reference AVX2 generator defect (BG2 R15): odd 32-byte vectors of the degree-3 group are never written
    const int gi = (int)(row.e0_deg >> 20) - 1;
#pragma unroll
    for (int b = 0; b < 4; b++) if (((gi * G.Z + (int)kb + b) >> 5) & 1) qz |= 0xFFu << (8 * b);
  }
  // row degrees in NRB200_UNROLL_MASK run unrolled, any other through the loop (see cn_row_loop)
#ifndef NRB200_UNROLL_MASK
#define NRB200_UNROLL_MASK 0x807FCu   /* every row degree the NR base graphs have: measured best (profiles/variants_r01n.txt) */
#endif
  const uint32_t D = (row.e0_deg >> 12) & 0xFFu;
#define NRB200_CN_CASE(d) case d: if ((NRB200_UNROLL_MASK >> d) & 1u) { cn_row<ZWC, d, QUIRK, CL>(G, smb, row, kb, halo, first_iter, qz, bad, sbase); return; } break;
  switch (D) {
    NRB200_CN_CASE(2) NRB200_CN_CASE(3) NRB200_CN_CASE(4) NRB200_CN_CASE(5) NRB200_CN_CASE(6) NRB200_CN_CASE(7) NRB200_CN_CASE(8) NRB200_CN_CASE(9)
    NRB200_CN_CASE(10) NRB200_CN_CASE(19)
    default: break;
  }
#undef NRB200_CN_CASE
  cn_row_loop<ZWC, QUIRK, CL>(G, smb, row, kb, halo, first_iter, qz, bad, sbase);
}

// one bit-node edge: fetch the rotated R' word and add its four bytes to the four running sums (IDP.4A, FMA pipe)
template <int ZWC>
__device__ __forceinline__ void bn_edge(const PackedGraph &G, const char *__restrict__ smb, int i, uint32_t kb, uint32_t &s0, uint32_t &s1,
                                        uint32_t &s2, uint32_t &s3)
{
  const uint2 d = *reinterpret_cast<const uint2 *>(G.bn_desc[i]);
  uint32_t a = kb + d.x;
  if (((kb << 8) | 0xFFu) < d.y) a += geo_zb<ZWC>(G);                        // circular wrap of v - s
  const uint32_t rw = __funnelshift_r(lds(smb, a), lds(smb, a + 4), d.y);
  s0 = __dp4a(rw, 0x00000001u, s0);
  s1 = __dp4a(rw, 0x00000100u, s1);
  s2 = __dp4a(rw, 0x00010000u, s2);
  s3 = __dp4a(rw, 0x01000000u, s3);
}

// A' = clamp(L' + sum R' - 128*deg, 0, 255) for column c, word k   (packs_epi16 of the int16 sum, in offset binary)
// CL > 0 (cluster decoder): the word goes into the A replica of every one of the CL CTAs
template <int ZWC>
__device__ __forceinline__ void bn_col(const PackedGraph &G, char *__restrict__ smb, int c, uint32_t kb, int CL = 0, uint32_t sbase = 0u)
{
  const uint32_t ZB = geo_zb<ZWC>(G);
  const uint32_t lw = lds(smb, G.off_L + c * geo_rsb<ZWC>(G) + kb);
  uint32_t s0 = __dp4a(lw, 0x00000001u, 0u), s1 = __dp4a(lw, 0x00000100u, 0u), s2 = __dp4a(lw, 0x00010000u, 0u), s3 = __dp4a(lw, 0x01000000u, 0u);
  int i = G.col_start[c];
  const int i1 = G.col_start[c + 1];
  for (; i + 2 <= i1; i += 2) { bn_edge<ZWC>(G, smb, i, kb, s0, s1, s2, s3); bn_edge<ZWC>(G, smb, i + 1, kb, s0, s1, s2, s3); }
  if (i < i1) bn_edge<ZWC>(G, smb, i, kb, s0, s1, s2, s3);
  const uint32_t nb = G.col_negbias[c];
  const uint32_t lo = __vmins2(__viaddmax_s16x2(prmt(s0, s1, 0x5410u), nb, 0u), 0x00ff00ffu);
  const uint32_t hi = __vmins2(__viaddmax_s16x2(prmt(s2, s3, 0x5410u), nb, 0u), 0x00ff00ffu);
  const uint32_t a = prmt(lo, hi, 0x6420u);
  const uint32_t ao = G.off_A + G.col_arow[c] * 2 * ZB + kb;
  if (CL > 0) {
    for (int r = 0; r < CL; r++) {
      const uint32_t ra = cl_map(sbase + ao, (uint32_t)r), mb = cl_mbar_at(1, (uint32_t)r);
      cl_push(ra, a, mb);
      cl_push(ra + ZB, a, mb);
    }
  } else {
    sts(smb, ao, a);
    sts(smb, ao + ZB, a);
  }
}

// hard decision of codeword position i (0/1); degree-1 columns read as 0 like the reference's untouched llrRes
__device__ __forceinline__ unsigned hd_bit(const PackedGraph &G, const char *smb, int i)
{
  const int c = i / G.Z, v = i - c * G.Z;
  const int ar = G.col_arow[c];
  if (ar < 0) return 0u;
  return (reinterpret_cast<const uint8_t *>(smb + G.off_A + ar * 2 * G.ZB)[v] >> 7) ^ 1u;
}

// part/nparts: the share of the output this CTA stores (cluster kernel: every CTA of the cluster holds the whole a-posteriori state)
__device__ __forceinline__ void packed_write_output(const PackedGraph &G, const char *smb, const DecodeArgs &a, uint8_t *o, int part = 0, int nparts = 1)
{
  const int numLLR = G.ncols * G.Z;
  const int tid = (int)threadIdx.x + part * (int)blockDim.x, nthr = nparts * (int)blockDim.x;
  if (a.outMode == 0) {
    const int nbytes = (numLLR + 7) >> 3;
    if ((G.Z & 7) == 0) {
      for (int j = tid; j < nbytes; j += nthr) {
        const int i = j * 8, c = i / G.Z, v = i - c * G.Z, ar = G.col_arow[c];
        unsigned b = 0;
        if (ar >= 0) {
          const uint32_t wo = G.off_A + ar * 2 * G.ZB + v;
          const uint32_t t0 = (~lds(smb, wo) & kH) >> 7, t1 = (~lds(smb, wo + 4) & kH) >> 7;
          b = (((t0 * 0x08040201u) >> 24) & 0xFu) << 4 | (((t1 * 0x08040201u) >> 24) & 0xFu);   // lift v first = MSB
        }
        o[j] = (uint8_t)b;
      }
    } else {
      for (int j = tid; j < nbytes; j += nthr) {
        unsigned b = 0;
        for (int kk = 0; kk < 8; kk++) { const int i = j * 8 + kk; if (i < numLLR) b |= hd_bit(G, smb, i) << (7 - kk); }
        o[j] = (uint8_t)b;
      }
    }
  } else {
    for (int i = tid; i < numLLR; i += nthr) o[i] = (uint8_t)hd_bit(G, smb, i);
  }
}

__device__ __forceinline__ int packed_crc_check(const PackedGraph &G, const char *smb, const DecodeArgs &a, int *scratch)
{
  const int n = (int)a.crc_len_bits;
  unsigned rem = 0;
  if (a.outMode == 0) {
    for (int i = threadIdx.x; i < n; i += blockDim.x)
      if (hd_bit(G, smb, i)) rem ^= __ldg(a.crc_tab + (n - 1 - i));
  } else {   // the reference checks the one-bit-per-byte array as it stands (nrLDPC_decoder.c:852-858): message bit 8j+7 = hard bit j
    for (int j = threadIdx.x; 8 * j + 7 < n; j += blockDim.x)
      if (hd_bit(G, smb, j)) rem ^= __ldg(a.crc_tab + (n - 8 - 8 * j));
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

// MAXT = launch bound: 768 threads leave 80 registers per thread, 864 leave 72, 960 leave 64 (no spills in any of them); more bins
// of Zw threads = more warps per scheduler to hide the shared-memory latency of the dependent descriptor -> message loads.
// -DNRB200_DECODER_TMA=1: stage the channel LLRs of the batch path with TMA bulk copies (below).  Bit exact and MEASURED SLOWER than eight register loads in flight
// per thread (0.696 against 0.668 ms per 1024 blocks, r02): the LLRs are not used as they arrive -- offset binary, the halo word and the doubled A rows are made
// on the way into shared memory, which a bulk copy cannot do, so it costs a second pass over the 26 KB.  Off by default.
#ifndef NRB200_DECODER_TMA
#define NRB200_DECODER_TMA 0
#endif
// PLAIN = the batch entry points with no per-block control row and no abort flags (what the bench's 1024-block launches are): the low-latency bookkeeping
// and the per-iteration abort poll are compiled out.
template <int ZWC, int MAXT, bool PLAIN = false>
__global__ void __launch_bounds__(MAXT, 1)
ldpc_decode_packed_kernel(const PackedGraph *__restrict__ gdev, DecodeArgs a)
{
  extern __shared__ __align__(16) uint32_t sm[];
  __shared__ PackedGraph G;
  __shared__ int s_flag, s_abort;
  __shared__ __align__(8) unsigned long long s_mbar;
  char *smb = reinterpret_cast<char *>(sm);
  for (int i = threadIdx.x; i < (int)(sizeof(PackedGraph) / 4); i += blockDim.x)
    reinterpret_cast<int *>(&G)[i] = reinterpret_cast<const int *>(gdev)[i];
  if (PLAIN && threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&s_mbar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  uint32_t tma_phase = 0;
  __syncthreads();
  const int Zw = geo_zw<ZWC>(G);
  // work lists (ldpc_packed_graph.h): one per bin of Zw threads (item = a whole row / column, this thread's word fixed), or one per warp
  // (item = 32 words of a row / column: item >> 8 selects which)
  const bool witems = G.warp_items != 0;
  const int bin = witems ? (int)(threadIdx.x >> 5) : (int)threadIdx.x / Zw;
  const uint32_t kb0 = witems ? 4u * (threadIdx.x & 31u) : 4u * (uint32_t)((int)threadIdx.x - bin * Zw);
  const bool worker = bin < G.nbins;
  const bool quirks = (a.quirks & 1) != 0;

  for (int cb = blockIdx.x; cb < (int)a.n_cb; cb += gridDim.x) {
    // ---- load channel LLRs (global int8, coalesced 32-bit) as offset binary into L rows with halo; A := L; R := 0; P := 0
    BlockIo io;
    if (PLAIN) { io.llr = a.llr + (size_t)cb * a.llr_stride; io.out = a.out + (size_t)cb * a.out_stride; io.iters = a.iters + cb; io.ctrl = nullptr; io.abort = nullptr; }
    else { io = block_io(a, cb); block_begin(io, 0); }
    const int8_t *gl = io.llr;
    const bool al4 = ((reinterpret_cast<uintptr_t>(gl) & 3) == 0);
    auto fetch = [&](int i) -> uint32_t {
      if (al4) return __ldg(reinterpret_cast<const uint32_t *>(gl) + i);
      const uint8_t *b = reinterpret_cast<const uint8_t *>(gl) + 4 * i;
      return b[0] | (b[1] << 8) | (b[2] << 16) | ((uint32_t)b[3] << 24);
    };
    auto place = [&](int i, uint32_t w) {
      const int c = i / Zw, k = i - c * Zw;
      w ^= kH;
      const uint32_t lo = G.off_L + c * G.RSB + 4 * k;
      sts(smb, lo, w);
      if (k == 0) sts(smb, lo + G.ZB, w);
      const int ar = G.col_arow[c];
      if (ar >= 0) { const uint32_t ao = G.off_A + ar * 2 * G.ZB + 4 * k; sts(smb, ao, w); sts(smb, ao + G.ZB, w); }
    };
    // Z = 384 batch path: the block's int8 LLRs are STAGED BY TMA -- one 1-D bulk copy per bit column (384 B) straight into the column's L row, completing on an
    // mbarrier -- while the threads clear the message rows; afterwards one pass over shared memory turns them into offset binary and fills the halo and the A rows.
    const bool tma = PLAIN && ZWC == 96 && NRB200_DECODER_TMA && al4 && ((reinterpret_cast<uintptr_t>(gl) | (uintptr_t)G.off_L | (uintptr_t)G.RSB) & 15) == 0;
    if (tma) {
      if (threadIdx.x == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // the previous block's reads of the L rows are ordered before the bulk writes
        const uint32_t mb = (uint32_t)__cvta_generic_to_shared(&s_mbar);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"((uint32_t)(G.ncols * 4 * Zw)) : "memory");
        const uint32_t l0 = (uint32_t)__cvta_generic_to_shared(smb) + (uint32_t)G.off_L;
        for (int c = 0; c < G.ncols; c++)
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(l0 + (uint32_t)(c * G.RSB)),
                       "l"(gl + (size_t)c * 4 * Zw), "r"((uint32_t)(4 * Zw)), "r"(mb) : "memory");
      }
    } else {
      // eight loads in flight per thread before the first store: the block's 26 KB arrive in two round trips to HBM instead of nine
      const int W = G.ncols * Zw, n = (int)blockDim.x;
      int i = threadIdx.x;
      for (; i + 7 * n < W; i += 8 * n) {
        uint32_t w[8];
#pragma unroll
        for (int k = 0; k < 8; k++) w[k] = fetch(i + k * n);
#pragma unroll
        for (int k = 0; k < 8; k++) place(i + k * n, w[k]);
      }
      for (; i + 3 * n < W; i += 4 * n) {
        const uint32_t w0 = fetch(i), w1 = fetch(i + n), w2 = fetch(i + 2 * n), w3 = fetch(i + 3 * n);
        place(i, w0); place(i + n, w1); place(i + 2 * n, w2); place(i + 3 * n, w3);
      }
      for (; i < W; i += n) place(i, fetch(i));
    }
    {
      // R := 0 (offset binary), 16 bytes per store where the rows allow it (every lifting size that is a multiple of 16)
      if (((G.off_R | (G.nreal * G.RSB)) & 15) == 0) {
        const uint4 z = make_uint4(kH, kH, kH, kH);
        uint4 *r4 = reinterpret_cast<uint4 *>(smb + G.off_R);
        const int n4 = (G.nreal * G.RSB) >> 4;
        for (int i = threadIdx.x; i < n4; i += blockDim.x) r4[i] = z;
      } else {
        for (int i = threadIdx.x; i < G.nreal * (G.RSB >> 2); i += blockDim.x) sts(smb, G.off_R + 4 * i, kH);
      }
    }
    if (tma) {
      const uint32_t mb = (uint32_t)__cvta_generic_to_shared(&s_mbar);
      uint32_t ok = 0;
      while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(mb), "r"(tma_phase) : "memory");
      tma_phase ^= 1u;
      for (int i = threadIdx.x; i < G.ncols * Zw; i += blockDim.x) {
        const int c = i / Zw, k = i - c * Zw;
        place(i, lds(smb, G.off_L + c * G.RSB + 4 * k));
      }
    }
    __syncthreads();
    // P rows: word k = sign(llr_p + R_p) flags (start clear), word Zw+k = sign-magnitude of the degree-1 neighbour's channel LLR,
    // word 2Zw+k = that LLR + 128 itself (rotated)
    for (int i = threadIdx.x; i < G.nrows * Zw; i += blockDim.x) {
      const int r = i / Zw, k = i - r * Zw;
      const PackedRow row = G.rows[r];
      if (row.lrow == 0xFFFFFFFFu) continue;
      uint32_t w0 = 4u * (uint32_t)(k + G.row_p_q[r]);
      if (w0 >= (uint32_t)G.ZB) w0 -= G.ZB;
      const uint32_t lp = __funnelshift_r(lds(smb, row.lrow + w0), lds(smb, row.lrow + w0 + 4), (uint32_t)G.row_p_rho[r]);
      const uint32_t dd = vabsdiffu4(lp, kH);
      const uint32_t pa = (row.prow_pcw & 0xFFFFFFu) + 4u * (uint32_t)k;
      sts(smb, pa, 0u);
      sts(smb, pa + G.ZB, lop3<kLutOrAnd>(dd, msb_mask(dd), kL7) | (~lp & kH));
      sts(smb, pa + 2 * G.ZB, lp);
    }
    __syncthreads();

    const int maxIter = a.numMaxIter;
    // Reference control flow (nrLDPC_decoder.c:541-863), restated for the fused schedule: after BN phase n the state
    // equals the reference's after iteration n.  The parity check of iteration n is a by-product of CN phase n+1.
    int numIter = 0;
    bool done = false;
    while (!done) {
      // CN phase of iteration numIter+1 (also yields the syndrome of iteration numIter when numIter >= 2)
      uint32_t bad = 0;
      // check_abort(ab) is polled at the top of every iteration from the second on (nrLDPC_decoder.c:557-560): thread 0 samples the flag while
      // the check-node phase runs (the load's latency hides behind it), everybody reads it after the barrier
      if (!PLAIN && io.abort && threadIdx.x == 0) s_abort = *io.abort;
      if (worker) {
        for (int i = G.cn_bin_start[bin]; i < G.cn_bin_start[bin + 1]; i++) {
          const int it = G.cn_bin_rows[i];
          const uint32_t kb = kb0 + 128u * (uint32_t)(it >> 8);
          // the defect-emulation variant (BG2 R15 only) lives in the generic instantiation alone: launch_decode() routes such calls there,
          // and the Z = 384 kernels carry half the code
          if (ZWC == 0 && quirks) cn_dispatch<ZWC, ZWC == 0>(G, smb, it & 0xFF, kb, kb == 0u, numIter == 0, bad);
          else cn_dispatch<ZWC, false>(G, smb, it & 0xFF, kb, kb == 0u, numIter == 0, bad);
        }
      }
      const int pcRes = __syncthreads_or(bad != 0);   // also the CN->BN barrier
      if (numIter >= 2 && !a.use_crc && pcRes == 0) break;          // iteration numIter passed its parity check (:552)
      numIter++;
      if (!PLAIN && numIter >= 2 && io.abort && s_abort) { numIter = maxIter + 2; break; }   // the state (A) is still that of the previous iteration
      // BN phase
      if (worker)
        for (int i = G.bn_bin_start[bin]; i < G.bn_bin_start[bin + 1]; i++) {
          const int it = G.bn_bin_cols[i];
          const uint32_t kb = kb0 + 128u * (uint32_t)(it >> 8);
          bn_col<ZWC>(G, smb, it & 0xFF, kb);
        }
      __syncthreads();
      // loop control, mirroring `while (numIter <= numMaxIter && pcRes != 0)` evaluated before each further iteration
      if (numIter == 1) {
        if (!(1 <= maxIter)) done = true;                            // the while condition fails straight away
      } else {
        if (a.use_crc) {
          if (numIter > 2) {                                         // :850-862
            packed_write_output(G, smb, a, io.out);
            if (packed_crc_check(G, smb, a, &s_flag)) break;
          }
          if (!(numIter <= maxIter)) done = true;
        } else {
          if (!(numIter <= maxIter)) done = true;                    // last allowed iteration: its parity check no longer matters
        }
      }
    }
    if (!a.use_crc) packed_write_output(G, smb, a, io.out);           // :865-877
    if (PLAIN) { if (threadIdx.x == 0) *io.iters = numIter; }
    else block_finish(io, a, numIter, 0);
    __syncthreads();
  }
}

}  // namespace nrb200
