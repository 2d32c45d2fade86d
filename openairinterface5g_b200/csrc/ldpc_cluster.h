// Work partition of ONE code block over a thread-block cluster (ldpc_decoder_cluster.cuh): which warp of which CTA owns which
// 32-word piece of which check row / bit column, and which CTA owns each bit column (the check-node phase pushes messages there).
#pragma once
#include <cstdint>
#include "ldpc_packed_graph.h"

namespace nrb200 {

constexpr int kClMaxCtas = 8;        // portable cluster size limit
constexpr int kClMaxWarps = 24;      // warps per CTA (launch bound 768 threads)
constexpr int kClMaxLists = kClMaxCtas * 16;   // C * T <= 128 work lists
constexpr int kClSplitItem = 0x1000;          // item flag: 16-word piece shared by the two halves of a warp
constexpr int kClSplitRowDeg = 19, kClSplitColDeg = 16;

struct alignas(16) ClusterSched {
  int32_t C, T, nthreads, chunks;    // CTAs per code block, warps per CTA, 32 * T, Zw / 32
  // list l = rank * T + warp; items cn_start[l] .. cn_start[l + 1]; item = row (column) | chunk << 8 as in PackedGraph, or -- the heavy rows
  // (19 stored edges) and columns (>= kClSplitColDeg edges) -- row (column) | piece << 8 | kClSplitItem: 16 words, the two halves of the warp
  // take one half of the edges each and combine their partial results by shuffle, which halves the longest dependent chain of the phase
  int16_t cn_start[kClMaxLists + 1];
  int16_t cn_items[6 * kMaxRows];
  int16_t bn_start[kClMaxLists + 1];
  int16_t bn_items[6 * kMaxCols];
  // bit columns are owned whole (all chunks) by one CTA: col_rank[c]; edge_rank[m] = col_rank[column of stored edge m] is where the check-node
  // phase pushes the edge's new cn->bn message (patched into bits[10:8] of the kernel's copy of PackedGraph::cn_desc[m][1])
  uint8_t col_rank[kMaxCols];
  uint8_t edge_rank[kMaxEdges];
  uint8_t pad[16 - (kMaxCols + kMaxEdges) % 16];
  // bytes each phase delivers into CTA r through asynchronous stores (what its mbarrier is armed with): check-node phase = (ZB + 4) per stored edge whose
  // column r owns (Zw message words + the halo word) + one 4-byte verdict word from every CTA; bit-node phase = 2 ZB per broadcast column (the doubled A row)
  uint32_t cn_tx[kClMaxCtas];
  uint32_t bn_tx, pad2[3];
};

// C CTAs of T warps each; false when the configuration cannot be split this way (Zw not a multiple of 32, too many lists).
bool build_cluster_sched(const GraphDev &g, const PackedGraph &p, int C, int T, ClusterSched *s);

}  // namespace nrb200
