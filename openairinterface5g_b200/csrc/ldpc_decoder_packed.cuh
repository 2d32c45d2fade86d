// Packed flooding int8 min-sum decoder for lifting sizes that are a multiple of 4 (every NR Z >= 32 and the hot
// Z = 384 case): one CTA per code block, four lifts per 32-bit register (byte SIMD-in-word), all state in shared memory.
//
// Schedule (bit exact with the reference's two-phase flooding decoder, nrLDPC_decoder.c:206-881):
//   state    Rn[m][t]  = cn->bn message of edge slot m at check lift t           (cnProcBufRes)
//            A[c][v]   = a-posteriori LLR of bit (c, v), degree >= 2 columns      (llrRes)
//            L[c][v]   = channel LLR                                              (llrProcBuf)
//   CN phase thread (row r, word k): for every edge  Q = subs_epi8(A[c][t+s], R_old)  -- what bnProc/bn2cnProcBuf
//            produced at the end of the previous iteration (nrLDPC_bnProc.h:325), formed on the fly from the shifted A
//            word, so no bn->cn buffer and no circular copies exist -- then exclude-self min / sign product
//            (nrLDPC_cnProc.h:388-877) written back in place.  The sign bytes of the same A words give the previous
//            iteration's syndrome (nrLDPC_cnProc.h:887-1960) for free: sign(adds_epi8(Q, R)) == sign(A)  (DESIGN.md).
//   BN phase thread (column c, word k): A = sat8(L + sum_e R[m_e][v - s_e])      (nrLDPC_bnProc.h:40-263)
// The quantisation points are the reference's: int16 sum -> sat8 -> subs_epi8 -> |.| clipped to 127.  Clipping Q to
// [-127,127] instead of [-128,127] is exact because the check node only ever uses min(|Q|,127) and sign(Q).
//
// Rows carry one halo word (word Zw repeats word 0) so a circularly shifted 4-lift group is always two consecutive
// words + one funnel shift; row stride is Zw+4 words to keep every row 16-byte aligned for the bulk (TMA) load of L.
#pragma once
#include "ldpc_common.cuh"
#include "ldpc_packed_graph.h"

namespace nrb200 {

// ---------------------------------------------------------------------------------------------- byte SIMD helpers
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel)
{
  uint32_t r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
  return r;
}
// 0xFF in every byte whose bit 7 is set
__device__ __forceinline__ uint32_t msb_mask(uint32_t x) { return prmt(x, 0u, 0xba98u); }
__device__ __forceinline__ uint32_t sel4(uint32_t m, uint32_t a, uint32_t b) { return (a & m) | (b & ~m); }   // one LOP3

// per byte, operands in [0,127]: 0xFF where a >= b
__device__ __forceinline__ uint32_t ge7(uint32_t a, uint32_t b) { return msb_mask((a | 0x80808080u) - b); }

// two's complement bytes of +-mag (mag in [0,127]) where neg has bit 7 of a byte set for "negative"; -0 -> 0
__device__ __forceinline__ uint32_t apply_sign7(uint32_t mag, uint32_t neg)
{
  const uint32_t n = msb_mask(neg);
  const uint32_t c = (mag ^ (n & 0x7f7f7f7fu)) + (n & 0x01010101u);   // (127-mag)+1 = 128-mag on negative bytes
  return c ^ (n & 0x80808080u);
}

__device__ __forceinline__ void unpack_s16x2(uint32_t w, uint32_t &lo, uint32_t &hi)
{
  lo = prmt(w, 0u, 0x9180u);   // {sext(b1), sext(b0)}
  hi = prmt(w, 0u, 0xb3a2u);   // {sext(b3), sext(b2)}
}

template <int D>
__device__ __forceinline__ void cn_row(const PackedGraph &G, uint32_t *__restrict__ sm, int r, int k, bool first_iter, uint32_t quirk_zero,
                                       uint32_t &bad)
{
  const int e0 = G.row_start[r];
  uint32_t q[D];
  uint32_t min1 = 0x7f7f7f7fu, min2 = 0x7f7f7f7fu, sgn = 0u, synd = 0u;
  uint32_t *Rrow = sm + G.off_R + e0 * G.RS + k;
#pragma unroll
  for (int j = 0; j < D; j++) {
    const int m = e0 + j;
    int w0 = k + G.cn_q[m];
    if (w0 >= G.Zw) w0 -= G.Zw;
    const uint32_t *ap = sm + G.cn_abase[m] + w0;
    const uint32_t aw = __funnelshift_r(ap[0], ap[1], G.cn_rho[m]);       // A at lifts t+s .. t+s+3
    const uint32_t rold = Rrow[j * G.RS];
    synd ^= aw;
    const uint32_t qq = __vsubss4(aw, rold);                               // subs_epi8(llrRes, cn->bn)  (bnProc)
    q[j] = qq;
    const uint32_t mag = __vabsss4(qq);                                    // min(|Q|,127)
    sgn ^= qq;
    const uint32_t m1 = ge7(mag, min1);                                    // mag >= min1
    const uint32_t t = sel4(m1, mag, min1);                                // max(mag, min1)
    min1 = sel4(m1, min1, mag);
    min2 = sel4(ge7(t, min2), min2, t);
  }
  uint32_t qp = 0u;
  const int pc = G.row_p_col[r];
  if (pc >= 0) {                                                           // degree-1 neighbour: Q is the channel LLR forever
    int w0 = k + G.row_p_q[r];
    if (w0 >= G.Zw) w0 -= G.Zw;
    const uint32_t *lp = sm + G.off_L + pc * G.RS + w0;
    qp = __funnelshift_r(lp[0], lp[1], G.row_p_rho[r]);
    uint32_t *P = sm + G.off_P + G.row_p_idx[r] * G.Zw + k;
    synd ^= *P;                                                            // sign(llr + R_p) of the previous iteration
    const uint32_t mag = __vabsss4(qp);
    sgn ^= qp;
    const uint32_t m1 = ge7(mag, min1);
    const uint32_t t = sel4(m1, mag, min1);
    min1 = sel4(m1, min1, mag);
    min2 = sel4(ge7(t, min2), min2, t);
    // R_p of this iteration -> sign of adds_epi8(llr, R_p) for the next syndrome
    const uint32_t isMin = ~msb_mask(((mag ^ min1) & 0x7f7f7f7fu) + 0x7f7f7f7fu);   // 0xFF where mag == min1
    const uint32_t rp = apply_sign7(sel4(isMin, min2, min1), sgn ^ qp) & ~quirk_zero;
    *P = __vaddss4(qp, rp) & 0x80808080u;
  }
  if (!first_iter && k < G.row_pc_words[r]) bad |= synd & 0x80808080u;
#pragma unroll
  for (int j = 0; j < D; j++) {
    const uint32_t mag = __vabsss4(q[j]);
    const uint32_t isMin = ~msb_mask(((mag ^ min1) & 0x7f7f7f7fu) + 0x7f7f7f7fu);
    const uint32_t rn = apply_sign7(sel4(isMin, min2, min1), sgn ^ q[j]) & ~quirk_zero;
    Rrow[j * G.RS] = rn;
    if (k == 0) Rrow[j * G.RS + G.Zw] = rn;                                // halo
  }
}

__device__ __forceinline__ void cn_dispatch(const PackedGraph &G, uint32_t *sm, int r, int k, bool first_iter, uint32_t qz, uint32_t &bad)
{
  switch (G.row_start[r + 1] - G.row_start[r]) {
    case 2: cn_row<2>(G, sm, r, k, first_iter, qz, bad); break;
    case 3: cn_row<3>(G, sm, r, k, first_iter, qz, bad); break;
    case 4: cn_row<4>(G, sm, r, k, first_iter, qz, bad); break;
    case 5: cn_row<5>(G, sm, r, k, first_iter, qz, bad); break;
    case 6: cn_row<6>(G, sm, r, k, first_iter, qz, bad); break;
    case 7: cn_row<7>(G, sm, r, k, first_iter, qz, bad); break;
    case 8: cn_row<8>(G, sm, r, k, first_iter, qz, bad); break;
    case 9: cn_row<9>(G, sm, r, k, first_iter, qz, bad); break;
    case 10: cn_row<10>(G, sm, r, k, first_iter, qz, bad); break;
    case 19: cn_row<19>(G, sm, r, k, first_iter, qz, bad); break;
    default: break;   // build_packed_graph() refuses graphs with other row degrees
  }
}

// A = sat8(L + sum R) for column c, word k
__device__ __forceinline__ void bn_col(const PackedGraph &G, uint32_t *__restrict__ sm, int c, int k)
{
  const uint32_t lw = sm[G.off_L + c * G.RS + k];
  uint32_t lo, hi;
  unpack_s16x2(lw, lo, hi);
  for (int i = G.col_start[c]; i < G.col_start[c + 1]; i++) {
    int w0 = k - G.bn_qq[i];
    if (w0 < 0) w0 += G.Zw;
    const uint32_t *rp = sm + G.bn_rbase[i] + w0;
    const uint32_t rw = __funnelshift_r(rp[0], rp[1], G.bn_sh[i]);
    uint32_t rl, rh;
    unpack_s16x2(rw, rl, rh);
    lo = __vadd2(lo, rl);
    hi = __vadd2(hi, rh);
  }
  // packs_epi16: saturate each 16-bit lane to int8
  lo = __vmaxs2(__vmins2(lo, 0x007f007fu), 0xff80ff80u);
  hi = __vmaxs2(__vmins2(hi, 0x007f007fu), 0xff80ff80u);
  const uint32_t a = prmt(lo, hi, 0x6420u);
  uint32_t *ap = sm + G.off_A + G.col_arow[c] * G.RS;
  ap[k] = a;
  if (k == 0) ap[G.Zw] = a;
}

// hard decision of codeword position i (0/1); degree-1 columns read as 0 like the reference's untouched llrRes
__device__ __forceinline__ unsigned hd_bit(const PackedGraph &G, const uint32_t *sm, int i)
{
  const int c = i / G.Z, v = i - c * G.Z;
  const int ar = G.col_arow[c];
  if (ar < 0) return 0u;
  const uint8_t *row = reinterpret_cast<const uint8_t *>(sm + G.off_A + ar * G.RS);
  return row[v] >> 7;
}

__device__ __forceinline__ void packed_write_output(const PackedGraph &G, const uint32_t *sm, const DecodeArgs &a, int cb)
{
  uint8_t *o = a.out + (size_t)cb * a.out_stride;
  const int numLLR = G.ncols * G.Z;
  if (a.outMode == 0) {
    const int nbytes = (numLLR + 7) >> 3;
    if ((G.Z & 7) == 0) {
      for (int j = threadIdx.x; j < nbytes; j += blockDim.x) {
        const int i = j * 8, c = i / G.Z, v = i - c * G.Z, ar = G.col_arow[c];
        unsigned b = 0;
        if (ar >= 0) {
          const uint32_t *w = sm + G.off_A + ar * G.RS + (v >> 2);
          const uint32_t t0 = (w[0] & 0x80808080u) >> 7, t1 = (w[1] & 0x80808080u) >> 7;
          b = (((t0 * 0x08040201u) >> 24) & 0xFu) << 4 | (((t1 * 0x08040201u) >> 24) & 0xFu);   // lift v first = MSB
        }
        o[j] = (uint8_t)b;
      }
    } else {
      for (int j = threadIdx.x; j < nbytes; j += blockDim.x) {
        unsigned b = 0;
        for (int kk = 0; kk < 8; kk++) { const int i = j * 8 + kk; if (i < numLLR) b |= hd_bit(G, sm, i) << (7 - kk); }
        o[j] = (uint8_t)b;
      }
    }
  } else {
    for (int i = threadIdx.x; i < numLLR; i += blockDim.x) o[i] = (uint8_t)hd_bit(G, sm, i);
  }
}

__device__ __forceinline__ int packed_crc_check(const PackedGraph &G, const uint32_t *sm, const DecodeArgs &a, int *scratch)
{
  const int n = (int)a.crc_len_bits;
  unsigned rem = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x)
    if (hd_bit(G, sm, i)) rem ^= __ldg(a.crc_tab + (n - 1 - i));
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

__global__ void __launch_bounds__(kPackedMaxThreads, 1)
ldpc_decode_packed_kernel(const PackedGraph *__restrict__ gdev, DecodeArgs a)
{
  extern __shared__ __align__(16) uint32_t sm[];
  __shared__ PackedGraph G;
  __shared__ int s_flag;
  for (int i = threadIdx.x; i < (int)(sizeof(PackedGraph) / 4); i += blockDim.x)
    reinterpret_cast<int *>(&G)[i] = reinterpret_cast<const int *>(gdev)[i];
  __syncthreads();
  const int Zw = G.Zw, RS = G.RS;
  const int bin = threadIdx.x / Zw, kw = threadIdx.x - bin * Zw;
  const bool worker = bin < G.nbins;

  for (int cb = blockIdx.x; cb < (int)a.n_cb; cb += gridDim.x) {
    // ---- load channel LLRs (global int8, coalesced 32-bit) into L rows with halo; A := L for degree>=2 columns; R := 0
    const int8_t *gl = a.llr + (size_t)cb * a.llr_stride;
    const bool al4 = ((reinterpret_cast<uintptr_t>(gl) & 3) == 0);
    for (int i = threadIdx.x; i < G.ncols * Zw; i += blockDim.x) {
      const int c = i / Zw, k = i - c * Zw;
      uint32_t w;
      if (al4) w = __ldg(reinterpret_cast<const uint32_t *>(gl) + i);
      else {
        const uint8_t *b = reinterpret_cast<const uint8_t *>(gl) + 4 * i;
        w = b[0] | (b[1] << 8) | (b[2] << 16) | ((uint32_t)b[3] << 24);
      }
      sm[G.off_L + c * RS + k] = w;
      if (k == 0) sm[G.off_L + c * RS + Zw] = w;
      const int ar = G.col_arow[c];
      if (ar >= 0) { sm[G.off_A + ar * RS + k] = w; if (k == 0) sm[G.off_A + ar * RS + Zw] = w; }
    }
    for (int i = threadIdx.x; i < G.nreal * RS; i += blockDim.x) sm[G.off_R + i] = 0u;
    for (int i = threadIdx.x; i < G.nrowP * Zw; i += blockDim.x) sm[G.off_P + i] = 0u;
    __syncthreads();

    const int maxIter = a.numMaxIter;
    const bool abort_in = a.abort_flags != nullptr && a.abort_flags[cb] != 0;
    // Reference control flow (nrLDPC_decoder.c:541-863), restated for the fused schedule: after BN phase n the state
    // equals the reference's after iteration n.  The parity check of iteration n is a by-product of CN phase n+1.
    int numIter = 0;
    bool done = false;
    while (!done) {
      // CN phase of iteration numIter+1 (also yields the syndrome of iteration numIter when numIter >= 2)
      uint32_t bad = 0;
      if (worker) for (int i = G.cn_bin_start[bin]; i < G.cn_bin_start[bin + 1]; i++) {
        const int r = G.cn_bin_rows[i], k = kw;
        uint32_t qz = 0u;
        if ((a.quirks & 1) && G.row_deg3_idx[r] >= 0) {
          // reference AVX2 generator defect (BG2 R15): odd 32-byte vectors of the degree-3 group are never written
#pragma unroll
          for (int b = 0; b < 4; b++) if (((G.row_deg3_idx[r] * G.Z + 4 * k + b) >> 5) & 1) qz |= 0xFFu << (8 * b);
        }
        cn_dispatch(G, sm, r, k, numIter == 0, qz, bad);
      }
      const int pcRes = __syncthreads_or(bad != 0);   // also the CN->BN barrier
      if (numIter >= 2 && !a.use_crc && pcRes == 0) break;          // iteration numIter passed its parity check (:552)
      numIter++;
      // BN phase
      if (worker) for (int i = G.bn_bin_start[bin]; i < G.bn_bin_start[bin + 1]; i++) bn_col(G, sm, G.bn_bin_cols[i], kw);
      __syncthreads();
      // loop control, mirroring `while (numIter <= numMaxIter && pcRes != 0)` evaluated before each further iteration
      if (numIter == 1) {
        if (!(1 <= maxIter)) done = true;                            // the while condition fails straight away
        else if (abort_in) { numIter = maxIter + 2; done = true; }   // :557-560, checked when entering iteration 2
      } else {
        if (a.use_crc) {
          if (numIter > 2) {                                         // :850-862
            packed_write_output(G, sm, a, cb);
            if (packed_crc_check(G, sm, a, &s_flag)) break;
          }
          if (!(numIter <= maxIter)) done = true;
        } else {
          if (!(numIter <= maxIter)) done = true;                    // last allowed iteration: its parity check no longer matters
        }
      }
    }
    if (!a.use_crc) packed_write_output(G, sm, a, cb);               // :865-877
    if (threadIdx.x == 0) a.iters[cb] = numIter;
    __syncthreads();
  }
}

}  // namespace nrb200
