// Generic (any NR lifting size) flooding int8 min-sum decoder: one CTA per code block, one message byte per
// thread-operation.  This is the correctness anchor for every (BG, Z, R) and the only path for the 15 lifting sizes
// that are not a multiple of 4 (all <= 30); the hot sizes go through ldpc_decoder_packed.cuh.
//
// Arithmetic restated from the reference (all saturating int8, deterministic => bit exact):
//   cnProc    nrLDPC_cnProc.h:388-877     R_e = prod_{k!=e} sgn(Q_k) * min(min_{k!=e}|Q_k|, 127)   (|-128| = 128)
//   bnProcPc  nrLDPC_bnProc.h:40-263      A_c = sat8(llr_c + sum_e R_e)  (int16 accumulation)
//   bnProc    nrLDPC_bnProc.h:271-1313    Q_e = subs_epi8(A_c, R_e); degree-1 bit nodes keep Q = llr
//   cnProcPc  nrLDPC_cnProc.h:887-1960    syndrome over sign(adds_epi8(Q_e, R_e)) == sign(A_c) ("syndrome identity")
//   control   nrLDPC_decoder.c:206-881
#pragma once
#include "ldpc_common.cuh"

namespace nrb200 {

__device__ __forceinline__ int sat8i(int v) { return max(-128, min(127, v)); }

// dynamic smem layout: GraphDev | llr[ncols*Z] | msg[nreal*Z] | hd[ncols*Z] (0/1) | hdp[nrows*Z] (0/1)
__global__ void __launch_bounds__(512, 1)
ldpc_decode_generic_kernel(const GraphDev *__restrict__ gdev, DecodeArgs a)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  GraphDev &g = *reinterpret_cast<GraphDev *>(smem_raw);
  for (int i = threadIdx.x; i < (int)(sizeof(GraphDev) / 4); i += blockDim.x)
    reinterpret_cast<int *>(smem_raw)[i] = reinterpret_cast<const int *>(gdev)[i];
  __syncthreads();
  const int Z = g.Z, ncols = g.ncols, nrows = g.nrows, nreal = g.nreal;
  const int numLLR = ncols * Z;
  int8_t *llr = reinterpret_cast<int8_t *>(smem_raw + ((sizeof(GraphDev) + 15) & ~15));
  int8_t *msg = llr + ((numLLR + 15) & ~15);
  uint8_t *hd = reinterpret_cast<uint8_t *>(msg + ((nreal * Z + 15) & ~15));
  uint8_t *hdp = hd + ((numLLR + 15) & ~15);
  __shared__ int s_flag;

  for (int cb = blockIdx.x; cb < (int)a.n_cb; cb += gridDim.x) {
    const BlockIo io = block_io(a, cb);
    block_begin(io, 0);
    const int8_t *gl = io.llr;
    for (int i = threadIdx.x; i < numLLR; i += blockDim.x) { llr[i] = gl[i]; hd[i] = 0; }
    __syncthreads();
    // llr2CnProcBuf (nrLDPC_mPass.h:128-169)
    for (int i = threadIdx.x; i < nreal * Z; i += blockDim.x) {
      const int m = i / Z, t = i - m * Z;
      int v = t + g.edge_shift[m]; if (v >= Z) v -= Z;
      msg[i] = llr[g.edge_col[m] * Z + v];
    }
    __syncthreads();

    const int maxIter = a.numMaxIter;
    int numIter = 0, pc = 1;
    bool wrote_out = false;
    for (;;) {
      if (numIter >= 1 && !(numIter <= maxIter && pc != 0)) break;      // nrLDPC_decoder.c:552
      numIter++;
      if (numIter > 1 && io.abort) {                                    // check_abort(ab), polled at the top of every iteration (:557-560)
        if (threadIdx.x == 0) s_flag = *io.abort;
        __syncthreads();
        const int ab = s_flag;
        __syncthreads();
        if (ab) { numIter = maxIter + 2; break; }
      }

      // ---- check nodes
      for (int i = threadIdx.x; i < nrows * Z; i += blockDim.x) {
        const int r = i / Z, t = i - r * Z;
        const int e0 = g.row_start[r], e1 = g.row_start[r + 1];
        int min1 = 255, min2 = 255, idx = -1, sgn = 0, zero = 0;
        for (int m = e0; m < e1; m++) {
          const int q = msg[m * Z + t];
          const int mag = abs(q);
          sgn ^= (q < 0); zero += (q == 0);
          if (mag < min1) { min2 = min1; min1 = mag; idx = m; } else if (mag < min2) min2 = mag;
        }
        const int pc_ = g.row_p_col[r];
        int qp = 0;
        if (pc_ >= 0) {
          int v = t + g.row_p_shift[r]; if (v >= Z) v -= Z;
          qp = llr[pc_ * Z + v];
          const int mag = abs(qp);
          sgn ^= (qp < 0); zero += (qp == 0);
          if (mag < min1) { min2 = min1; min1 = mag; idx = -2; } else if (mag < min2) min2 = mag;
        }
        const bool quirk = a.quirks & 1 ? (g.row_deg3_idx[r] >= 0 && (((g.row_deg3_idx[r] * Z + t) >> 5) & 1)) : false;
        for (int m = e0; m < e1; m++) {
          const int q = msg[m * Z + t];
          int mag = min(m == idx ? min2 : min1, 127);
          int s = sgn ^ (q < 0);
          // sign_epi8 chain: any *other* zero input forces 0; then min is 0 anyway unless that zero is q itself
          int r_ = (zero - (q == 0)) > 0 ? 0 : (s ? -mag : mag);
          msg[m * Z + t] = (int8_t)(quirk ? 0 : r_);
        }
        if (pc_ >= 0) {
          int mag = min(idx == -2 ? min2 : min1, 127);
          int s = sgn ^ (qp < 0);
          int rp = (zero - (qp == 0)) > 0 ? 0 : (s ? -mag : mag);
          if (quirk) rp = 0;
          hdp[i] = (uint8_t)(sat8i(qp + rp) < 0);   // sign(adds_epi8(cnProcBuf, cnProcBufRes)) of the degree-1 edge
        }
      }
      __syncthreads();
      // ---- bit nodes (degree >= 2)
      for (int i = threadIdx.x; i < numLLR; i += blockDim.x) {
        const int c = i / Z, v = i - c * Z;
        const int k0 = g.col_start[c], k1 = g.col_start[c + 1];
        if (k1 == k0) continue;
        int acc = llr[i];
        for (int k = k0; k < k1; k++) {
          const int m = g.col_edges[k];
          int t = v - g.edge_shift[m]; if (t < 0) t += Z;
          acc += msg[m * Z + t];
        }
        const int A = sat8i(acc);
        hd[i] = (uint8_t)(A < 0);
        for (int k = k0; k < k1; k++) {
          const int m = g.col_edges[k];
          int t = v - g.edge_shift[m]; if (t < 0) t += Z;
          msg[m * Z + t] = (int8_t)sat8i(A - msg[m * Z + t]);
        }
      }
      __syncthreads();
      if (numIter == 1) continue;

      if (!a.use_crc) {
        int bad = 0;
        for (int i = threadIdx.x; i < nrows * Z; i += blockDim.x) {
          const int r = i / Z, t = i - r * Z;
          if (t >= g.row_pc_from[r]) continue;   // lifts the reference's cnProcPc never tests (nrb200_graph.cc)
          int par = g.row_p_col[r] >= 0 ? hdp[i] : 0;
          for (int m = g.row_start[r]; m < g.row_start[r + 1]; m++) {
            int v = t + g.edge_shift[m]; if (v >= Z) v -= Z;
            par ^= hd[g.edge_col[m] * Z + v];
          }
          bad |= par;
        }
        pc = __syncthreads_or(bad);
      } else if (numIter > 2) {                                          // :850-862
        write_output(a, io.out, hd, numLLR);
        wrote_out = true;
        const int ok = crc_check_block(a, hd, &s_flag);
        if (ok) break;
      }
    }
    if (!a.use_crc) write_output(a, io.out, hd, numLLR);                 // :865-877
    (void)wrote_out;
    block_finish(io, a, numIter, 0);
    __syncthreads();
  }
}

}  // namespace nrb200
