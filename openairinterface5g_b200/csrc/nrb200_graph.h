// Lifted-graph descriptors for the NR LDPC kernels (host side builder + the POD the kernels read).
//
// One descriptor per (BG, Z, decoder-rate selector R).  R keeps the first ncols(R) - nsys base-graph rows, exactly the
// LUT sets the reference selects in nrLDPC_init (reference openair1/PHY/CODING/nrLDPC_decoder/nrLDPC_init.h:58-...,
// nrLDPCdecoder_defs.h:53-84).  Unlike the reference there are no per-degree buffers and no circular memcpy tables:
// an edge is (row, column, shift) and every message lives at [edge slot][check-node lift].
#pragma once
#include <cstdint>
#include <vector>

namespace nrb200 {

constexpr int kMaxRows = 46;
constexpr int kMaxCols = 68;
constexpr int kMaxEdges = 316;
constexpr int kMaxRowDeg = 19;  // real (stored) edges per row
constexpr int kMaxColDeg = 30;

// Device-visible graph (POD, copied verbatim to global memory, staged to shared memory by the kernels).
// "real" edges connect bit nodes of degree >= 2 and own a message slot; an edge to a degree-1 bit node (the
// extension parity columns) has no slot: its bn->cn message is the channel LLR forever (reference
// nrLDPC_mPass.h:350-351 skips them) and only the sign of llr + cn->bn is needed for the syndrome.
struct GraphDev {
  int32_t BG, Z, R, ils;
  int32_t ncols, nrows, nsys;
  int32_t nreal;               // stored edges
  int32_t quirk_deg3_rows;     // bit r set => row r is in the BG2 degree-3 group (AVX2-defect emulation only)
  int16_t row_start[kMaxRows + 1];   // real-edge slot range of row r: [row_start[r], row_start[r+1])
  int16_t row_p_col[kMaxRows];       // degree-1 column attached to row r, or -1
  int16_t row_p_shift[kMaxRows];     // its shift (mod Z)
  int16_t row_deg3_idx[kMaxRows];    // index of the row inside the degree-3 group (quirk emulation), else -1
  int16_t row_pc_from[kMaxRows];     // first lift of row r the reference's parity check does NOT test (Z = all tested), see nrb200_graph.cc
  int16_t edge_col[kMaxEdges];       // per slot
  int16_t edge_shift[kMaxEdges];     // per slot, already reduced mod Z
  int16_t col_start[kMaxCols + 1];   // range into col_edges of column c
  int16_t col_edges[kMaxEdges];      // slots of column c, ascending row
  int16_t col_deg[kMaxCols];         // total degree incl. degree-1 edges
};

// Encoder view: the full base graph split in core / extension parts.
struct EncGraphDev {
  int32_t BG, Z, ils, ncols, nrows, nsys;
  int32_t sigma;                     // sum of the 4 core rows leaves x^sigma * p0
  int16_t row_start[kMaxRows + 1];   // all edges, row-major
  int16_t edge_col[kMaxEdges];
  int16_t edge_shift[kMaxEdges];
  // core solve order for parity columns nsys+1..nsys+3: row to use, the column it yields and that column's shift in the row
  int16_t core_row[3], core_col[3], core_shift[3];
};

int ils_of_z(int Z);                       // -1 when Z is not an NR lifting size
int ncols_for_rate(int BG, int R);         // -1 when (BG, R) is not a decoder LUT
bool build_graph(int BG, int Z, int R, GraphDev *g);      // false on invalid parameters
bool build_enc_graph(int BG, int Z, EncGraphDev *g);

}  // namespace nrb200
