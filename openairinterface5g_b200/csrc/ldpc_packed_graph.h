// Host/device POD tables of the packed decoder (built on the host by ldpc_packed_graph.cc).
#pragma once
#include <cstdint>
#include "nrb200_graph.h"

namespace nrb200 {

constexpr int kPackedMaxThreads = 768;
constexpr int kMaxBins = 24;

// Host-built tables for the packed kernel (copied to shared memory at kernel start).
struct PackedGraph {
  int32_t Z, Zw, RS;                 // lifts, words per row, row stride in words (Zw + 4)
  int32_t ncols, nrows, nreal, ncolA, nrowP;
  int32_t off_R, off_A, off_L, off_P, total_words;   // region offsets (words) inside the dynamic shared buffer
  int32_t nbins, nthreads;           // thread t works for bin t / Zw on word t % Zw; bins own whole rows / columns
  int16_t cn_bin_start[kMaxBins + 1]; // LPT-balanced row lists per bin (heavy rows first)
  int16_t cn_bin_rows[kMaxRows];
  int16_t bn_bin_start[kMaxBins + 1]; // LPT-balanced column lists per bin
  int16_t bn_bin_cols[kMaxCols];
  int16_t row_start[kMaxRows + 1];   // slot range per row
  int16_t row_p_col[kMaxRows];       // degree-1 column of the row or -1
  int16_t row_p_q[kMaxRows], row_p_rho[kMaxRows];
  int16_t row_p_idx[kMaxRows];       // row index inside the P (degree-1 sign) region
  int16_t row_deg3_idx[kMaxRows];
  int16_t row_pc_words[kMaxRows];    // words of the row that take part in the parity check (reference cnProcPc coverage)
  int16_t col_start[kMaxCols + 1];
  int16_t col_arow[kMaxCols];        // row of column c inside the A region, -1 for degree-1 columns
  // per slot, CN side: where the edge's bit node lives in A and how far it is rotated
  int32_t cn_abase[kMaxEdges];       // word offset of A row of the edge's column
  int16_t cn_q[kMaxEdges], cn_rho[kMaxEdges];
  // per column-edge entry, BN side
  int32_t bn_rbase[kMaxEdges];       // word offset of R row of the slot
  int16_t bn_qq[kMaxEdges], bn_sh[kMaxEdges];   // word back-shift and funnel byte shift*8
};

bool build_packed_graph(const GraphDev &g, PackedGraph *p, int max_threads = kPackedMaxThreads);

}  // namespace nrb200
