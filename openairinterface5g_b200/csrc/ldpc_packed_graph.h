// Host/device POD tables of the packed decoder (built on the host by ldpc_packed_graph.cc).
#pragma once
#include <cstdint>
#include "nrb200_graph.h"

namespace nrb200 {

constexpr int kPackedMaxThreads = 768;       // launch bound of the generic instantiation
constexpr int kPackedMaxThreadsZ384 = 960;   // Z = 384 instantiations exist for 768 / 864 / 960 threads (8 / 9 / 10 bins of 96)
constexpr int kMaxBins = 32;                 // work lists: bins of Zw threads, or single warps (see PackedGraph::warp_items)

// Per check row, one 16-byte record (read with a single LDS.128).
struct PackedRow {
  uint32_t e0_deg;     // bits[11:0] first slot, bits[19:12] number of stored edges D, bits[27:20] degree-3 group index + 1
  uint32_t rbase;      // byte offset of R row of the first slot
  uint32_t lrow;       // byte offset of the L row of the degree-1 neighbour column, 0xFFFFFFFF when the row has none
  uint32_t prow_pcw;   // bits[23:0] byte offset of the row's P (degree-1 sign) words, bits[31:24] words taking part in the parity check
};

// Shared-memory image (all offsets in bytes from the start of the dynamic shared buffer):
//   A  ncolA rows of 2*Zw words : a-posteriori LLRs + 128 (offset binary), stored twice so a rotation never wraps
//   R  nreal rows of Zw+4 words : cn->bn messages + 128, word Zw repeats word 0 (halo)
//   L  ncols rows of Zw+4 words : channel LLRs + 128, with halo
//   P  nrowP rows of 3*Zw words : word k: 0x80 per byte where sat8(llr_p + R_p) < 0; word Zw+k: sign-magnitude of the neighbour's
//                                 channel LLR; word 2Zw+k: that LLR + 128 (rotated)
struct PackedGraph {
  int32_t Z, Zw, ZB, RSB;            // lifts, words per row, 4*Zw, R/L row stride in bytes
  int32_t ncols, nrows, nreal, ncolA, nrowP;
  int32_t off_A, off_R, off_L, off_P, total_bytes;
  int32_t nbins, nthreads;           // warp_items == 0: thread t works for bin t / Zw on word t % Zw; bins own whole rows / columns
  int32_t warp_items;                // 1 (Zw a multiple of 32): a work item is 32 words of a row / column, every WARP owns a list of items --
                                     //   138 + 78 items on 24 warps for BG1 Z = 384 balance to ~1 % where 46 rows on 8 bins leave ~9 % at the barriers
  // list l = items cn_bin_start[l] .. cn_bin_start[l + 1]; item = row (column) | chunk << 8, chunk = which 32 words of the row
  int16_t cn_bin_start[kMaxBins + 1];
  int16_t cn_bin_rows[3 * kMaxRows];
  int16_t bn_bin_start[kMaxBins + 1];
  int16_t bn_bin_cols[3 * kMaxCols];
  int16_t row_p_q[kMaxRows], row_p_rho[kMaxRows];   // rotation of the degree-1 neighbour (0 for every NR base graph)
  int16_t col_start[kMaxCols + 1];
  int16_t col_arow[kMaxCols];        // row of column c inside the A region, -1 for degree-1 columns
  uint32_t col_negbias[kMaxCols];    // two 16-bit lanes of -(128 * degree): removes the offset-binary bias of the summed messages
  alignas(16) PackedRow rows[kMaxRows];
  // per slot, CN side: .x = byte offset of A row + 4*(shift / 4), .y = 8*(shift % 4) (funnel amount)
  alignas(8) uint32_t cn_desc[kMaxEdges][2];
  // per column-edge entry, BN side: .x = byte offset of R row - 4*qq, .y = (4*qq << 8) | 8*((4 - shift%4) & 3)
  // (.y is compared against (kb << 8) | 0xFF for the circular wrap and used as-is as the funnel amount: only bits[4:0] count).
  // (A 16-byte record with the wrap as one unsigned min(x, x + ZB) -- VIADDMNMX -- and the adds on the FMA pipe measured 3 % slower.)
  alignas(8) uint32_t bn_desc[kMaxEdges][2];
  uint32_t one;                      // == 1, read at run time so that selected adds are emitted as IMAD (FMA pipe), see DESIGN.md
};

bool build_packed_graph(const GraphDev &g, PackedGraph *p, int max_threads = kPackedMaxThreads);

}  // namespace nrb200
