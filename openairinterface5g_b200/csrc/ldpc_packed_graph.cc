// Host-side construction of the packed decoder's tables (ldpc_decoder_packed.cuh: PackedGraph).
#include <algorithm>
#include <cstring>
#include <vector>
#include "ldpc_packed_graph.h"

namespace nrb200 {

// Longest-processing-time assignment of weighted items to nbins bins; emits per-bin lists (heaviest first).
static void lpt(const std::vector<std::pair<int, int>> &items /* (weight, id) */, int nbins, int16_t *bin_start, int16_t *out)
{
  std::vector<std::pair<int, int>> s(items);
  std::stable_sort(s.begin(), s.end(), [](const std::pair<int, int> &a, const std::pair<int, int> &b) { return a.first > b.first; });
  std::vector<std::vector<int>> bins(nbins);
  std::vector<int> load(nbins, 0);
  for (auto &it : s) {
    int b = 0;
    for (int i = 1; i < nbins; i++) if (load[i] < load[b]) b = i;
    bins[b].push_back(it.second);
    load[b] += it.first;
  }
  int n = 0;
  for (int b = 0; b < nbins; b++) {
    bin_start[b] = (int16_t)n;
    for (int id : bins[b]) out[n++] = (int16_t)id;
  }
  bin_start[nbins] = (int16_t)n;
}

bool build_packed_graph(const GraphDev &g, PackedGraph *p, int max_threads)
{
  if (g.Z % 4) return false;
  std::memset(p, 0, sizeof(*p));
  p->Z = g.Z; p->Zw = g.Z / 4; p->RS = p->Zw + 4;
  p->ncols = g.ncols; p->nrows = g.nrows; p->nreal = g.nreal;
  // A rows for degree>=2 columns
  int na = 0;
  for (int c = 0; c < g.ncols; c++) p->col_arow[c] = (int16_t)(g.col_deg[c] >= 2 ? na++ : -1);
  p->ncolA = na;
  int np = 0;
  for (int r = 0; r < g.nrows; r++) {
    p->row_start[r] = g.row_start[r];
    p->row_p_col[r] = g.row_p_col[r];
    p->row_deg3_idx[r] = g.row_deg3_idx[r];
    p->row_pc_words[r] = (int16_t)(g.row_pc_from[r] / 4);
    if (g.row_pc_from[r] % 4) return false;
    p->row_p_idx[r] = -1;
    if (g.row_p_col[r] >= 0) {
      p->row_p_idx[r] = (int16_t)np++;
      p->row_p_q[r] = (int16_t)(g.row_p_shift[r] / 4);
      p->row_p_rho[r] = (int16_t)(8 * (g.row_p_shift[r] % 4));
    }
    const int d = g.row_start[r + 1] - g.row_start[r];
    if (!((d >= 2 && d <= 10) || d == 19)) return false;   // cn_dispatch() instantiations
  }
  p->row_start[g.nrows] = g.row_start[g.nrows];
  p->nrowP = np;
  p->off_R = 0;
  p->off_A = p->off_R + g.nreal * p->RS;
  p->off_L = p->off_A + na * p->RS;
  p->off_P = p->off_L + g.ncols * p->RS;
  p->total_words = p->off_P + np * p->Zw;
  for (int m = 0; m < g.nreal; m++) {
    const int c = g.edge_col[m], s = g.edge_shift[m];
    p->cn_abase[m] = p->off_A + p->col_arow[c] * p->RS;
    p->cn_q[m] = (int16_t)(s / 4);
    p->cn_rho[m] = (int16_t)(8 * (s % 4));
  }
  for (int c = 0; c <= g.ncols; c++) p->col_start[c] = g.col_start[c];
  for (int i = 0; i < g.nreal; i++) {
    const int m = g.col_edges[i], s = g.edge_shift[m], q = s / 4, rho = s % 4;
    p->bn_rbase[i] = p->off_R + m * p->RS;
    p->bn_qq[i] = (int16_t)(q + (rho ? 1 : 0));
    p->bn_sh[i] = (int16_t)(8 * ((4 - rho) & 3));
  }
  // thread geometry: bins of Zw threads
  int nbins = max_threads / p->Zw;
  nbins = std::max(1, std::min(nbins, std::min(kMaxBins, g.nrows)));
  p->nbins = nbins;
  p->nthreads = std::min(max_threads, std::max(32, ((nbins * p->Zw + 31) / 32) * 32));
  std::vector<std::pair<int, int>> rows, cols;
  for (int r = 0; r < g.nrows; r++) rows.push_back({g.row_start[r + 1] - g.row_start[r] + (g.row_p_col[r] >= 0 ? 1 : 0) + 1, r});
  for (int c = 0; c < g.ncols; c++) if (g.col_deg[c] >= 2) cols.push_back({g.col_deg[c] + 1, c});
  lpt(rows, nbins, p->cn_bin_start, p->cn_bin_rows);
  lpt(cols, nbins, p->bn_bin_start, p->bn_bin_cols);
  return true;
}

}  // namespace nrb200
