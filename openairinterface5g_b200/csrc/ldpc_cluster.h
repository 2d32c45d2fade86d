// Work partition of ONE code block over a thread-block cluster (ldpc_decoder_cluster.cuh): which warp of which CTA owns which
// 32-word piece of which check row / bit column, and -- for the bit-node phase -- which CTA holds the cn->bn messages it has to pull.
#pragma once
#include <cstdint>
#include "ldpc_packed_graph.h"

namespace nrb200 {

constexpr int kClMaxCtas = 8;        // portable cluster size limit
constexpr int kClMaxWarps = 24;      // warps per CTA (launch bound 768 threads)
constexpr int kClMaxLists = kClMaxCtas * 16;   // C * T <= 128 work lists

struct ClusterSched {
  int32_t C, T, nthreads, chunks;    // CTAs per code block, warps per CTA, 32 * T, Zw / 32
  // list l = rank * T + warp; items cn_start[l] .. cn_start[l + 1]; item = row (column) | chunk << 8 as in PackedGraph
  int16_t cn_start[kClMaxLists + 1];
  int16_t cn_items[3 * kMaxRows];
  int16_t bn_start[kClMaxLists + 1];
  int16_t bn_items[3 * kMaxCols];
  // per column-edge entry (index as PackedGraph::bn_desc): .x = byte offset of the R row - 4*qq (same as PackedGraph), .y = bits[4:0] funnel
  // amount, bits[16:8] 4*qq, bits[31:20] four 3-bit cluster ranks: the CTA that owns words 32*k .. 32*k+31 of this edge's R row for k = 0, 1, 2
  // and, in field `chunks`, the owner of word 0 again (the halo word Zw)
  alignas(8) uint32_t bn_desc[kMaxEdges][2];
};

// C CTAs of T warps each; false when the configuration cannot be split this way (Zw not a multiple of 32, too many lists).
bool build_cluster_sched(const GraphDev &g, const PackedGraph &p, int C, int T, ClusterSched *s);

}  // namespace nrb200
