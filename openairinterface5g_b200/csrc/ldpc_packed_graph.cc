// Host-side construction of the packed decoder's tables (ldpc_packed_graph.h).
#include <algorithm>
#include <cstring>
#include <vector>
#include "ldpc_packed_graph.h"
#include "ldpc_cluster.h"

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
  // LPT leaves the heaviest list up to ~10 % above the mean here (identical chunks of a row land in lockstep); refine by moving an item
  // off the heaviest list, or swapping it against a lighter item of another list, while that lowers the pair's maximum.
  std::vector<int> w(1 << 16, 0);
  for (auto &it : s) w[(size_t)it.second & 0xFFFF] = it.first;
  for (int round = 0; round < 4000; round++) {
    int a = 0;
    for (int i = 1; i < nbins; i++) if (load[i] > load[a]) a = i;
    bool improved = false;
    for (int b = 0; b < nbins && !improved; b++) {
      if (b == a) continue;
      const int pair_max = load[a];
      for (size_t i = 0; i < bins[a].size() && !improved; i++) {
        const int wi = w[(size_t)bins[a][i] & 0xFFFF];
        if (std::max(load[a] - wi, load[b] + wi) < pair_max) {              // move
          load[a] -= wi; load[b] += wi;
          bins[b].push_back(bins[a][i]); bins[a].erase(bins[a].begin() + (long)i);
          improved = true;
          break;
        }
        for (size_t j = 0; j < bins[b].size(); j++) {                       // swap
          const int wj = w[(size_t)bins[b][j] & 0xFFFF];
          if (wj < wi && std::max(load[a] - wi + wj, load[b] + wi - wj) < pair_max) {
            load[a] += wj - wi; load[b] += wi - wj;
            std::swap(bins[a][i], bins[b][j]);
            improved = true;
            break;
          }
          for (size_t k = j + 1; k < bins[b].size(); k++) {                 // one item against two lighter ones
            const int wjk = wj + w[(size_t)bins[b][k] & 0xFFFF];
            if (wjk < wi && std::max(load[a] - wi + wjk, load[b] + wi - wjk) < pair_max) {
              load[a] += wjk - wi; load[b] += wi - wjk;
              const int x = bins[a][i], y = bins[b][j], z = bins[b][k];
              bins[a][i] = y; bins[a].push_back(z);
              bins[b].erase(bins[b].begin() + (long)k); bins[b][j] = x;
              improved = true;
              break;
            }
          }
          if (improved) break;
        }
      }
    }
    if (!improved) break;
  }
  for (auto &bl : bins) std::stable_sort(bl.begin(), bl.end(), [&](int x, int y) { return w[(size_t)x & 0xFFFF] > w[(size_t)y & 0xFFFF]; });
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
  p->Z = g.Z; p->Zw = g.Z / 4; p->ZB = 4 * p->Zw; p->RSB = 4 * (p->Zw + 4);
  p->ncols = g.ncols; p->nrows = g.nrows; p->nreal = g.nreal;
  int na = 0;
  for (int c = 0; c < g.ncols; c++) {
    p->col_arow[c] = (int16_t)(g.col_deg[c] >= 2 ? na++ : -1);
    const uint32_t nb = (uint32_t)(-(128 * g.col_deg[c])) & 0xFFFFu;
    p->col_negbias[c] = nb | (nb << 16);
  }
  p->ncolA = na;
  int np = 0;
  for (int r = 0; r < g.nrows; r++) if (g.row_p_col[r] >= 0) np++;
  p->nrowP = np;
  p->off_A = 0;
  p->off_R = p->off_A + na * 2 * p->ZB;
  p->off_L = p->off_R + g.nreal * p->RSB;
  p->off_P = p->off_L + g.ncols * p->RSB;
  p->total_bytes = p->off_P + np * 3 * p->ZB;
  p->one = 1u;
  np = 0;
  for (int r = 0; r < g.nrows; r++) {
    const int d = g.row_start[r + 1] - g.row_start[r];
    if (!((d >= 2 && d <= 10) || d == 19)) return false;   // cn_dispatch() instantiations
    if (g.row_pc_from[r] % 4) return false;
    PackedRow &pr = p->rows[r];
    pr.e0_deg = (uint32_t)g.row_start[r] | ((uint32_t)d << 12) | ((uint32_t)(g.row_deg3_idx[r] + 1) << 20);
    pr.rbase = (uint32_t)(p->off_R + g.row_start[r] * p->RSB);
    pr.lrow = 0xFFFFFFFFu;
    pr.prow_pcw = (uint32_t)(g.row_pc_from[r] / 4) << 24;
    if (g.row_p_col[r] >= 0) {
      pr.lrow = (uint32_t)(p->off_L + g.row_p_col[r] * p->RSB);
      pr.prow_pcw |= (uint32_t)(p->off_P + np * 3 * p->ZB);
      p->row_p_q[r] = (int16_t)(g.row_p_shift[r] / 4);
      p->row_p_rho[r] = (int16_t)(8 * (g.row_p_shift[r] % 4));
      np++;
    }
  }
  for (int m = 0; m < g.nreal; m++) {
    const int c = g.edge_col[m], s = g.edge_shift[m];
    const uint32_t aoff = (uint32_t)(p->off_A + p->col_arow[c] * 2 * p->ZB + 4 * (s / 4));
    p->cn_desc[m][0] = aoff;
    p->cn_desc[m][1] = (uint32_t)(8 * (s % 4));
  }
  for (int c = 0; c <= g.ncols; c++) p->col_start[c] = g.col_start[c];
  for (int i = 0; i < g.nreal; i++) {
    const int m = g.col_edges[i], s = g.edge_shift[m], q = s / 4, rho = s % 4;
    const int qq4 = 4 * (q + (rho ? 1 : 0));
    const uint32_t base_minus = (uint32_t)(p->off_R + m * p->RSB - qq4);   // off_R > 4*(Zw+1) always (A region precedes)
    p->bn_desc[i][0] = base_minus;
    p->bn_desc[i][1] = ((uint32_t)qq4 << 8) | (uint32_t)(8 * ((4 - rho) & 3));
  }
  // thread geometry.  Costs are warp instructions per item measured on the Z = 384 kernel (profiles/r01n_*): a row costs 18 (dispatch) +
  // 32 per stored edge + 51 with a degree-1 neighbour (11 without); a column 65 + 13.5 per edge.
  auto row_cost = [&](int r) { return 2 * (18 + 32 * (g.row_start[r + 1] - g.row_start[r]) + (g.row_p_col[r] >= 0 ? 51 : 11)); };
  auto col_cost = [&](int c) { return 130 + 27 * g.col_deg[c]; };
  std::vector<std::pair<int, int>> rows, cols;
  if (p->Zw % 32 == 0 && max_threads >= 32) {
    // work item = 32 words of one row / column, one list per warp
    const int chunks = p->Zw / 32;
    int nwarps = std::min(max_threads / 32, kMaxBins);
    nwarps = std::max(1, std::min(nwarps, g.nrows * chunks));
    p->warp_items = 1;
    p->nbins = nwarps;
    p->nthreads = 32 * nwarps;
    for (int r = 0; r < g.nrows; r++) for (int k = 0; k < chunks; k++) rows.push_back({row_cost(r), r | (k << 8)});
    for (int c = 0; c < g.ncols; c++) if (g.col_deg[c] >= 2) for (int k = 0; k < chunks; k++) cols.push_back({col_cost(c), c | (k << 8)});
  } else {
    // bins of Zw threads owning whole rows / columns
    int nbins = max_threads / p->Zw;
    nbins = std::max(1, std::min(nbins, std::min(kMaxBins, g.nrows)));
    p->nbins = nbins;
    p->nthreads = std::min(max_threads, std::max(32, ((nbins * p->Zw + 31) / 32) * 32));
    for (int r = 0; r < g.nrows; r++) rows.push_back({row_cost(r), r});
    for (int c = 0; c < g.ncols; c++) if (g.col_deg[c] >= 2) cols.push_back({col_cost(c), c});
  }
  lpt(rows, p->nbins, p->cn_bin_start, p->cn_bin_rows);
  lpt(cols, p->nbins, p->bn_bin_start, p->bn_bin_cols);
  return true;
}

bool build_cluster_sched(const GraphDev &g, const PackedGraph &p, int C, int T, ClusterSched *s)
{
  if (p.Zw % 32 || C < 2 || C > kClMaxCtas || T < 1 || T > kClMaxWarps || C * T > kClMaxLists) return false;
  std::memset(s, 0, sizeof(*s));
  const int chunks = p.Zw / 32;
  if (chunks > 3) return false;
  s->C = C; s->T = T; s->nthreads = 32 * T; s->chunks = chunks;
  // cost model of the single-CTA lists; a row additionally pays the push of every message (4 per edge), a column the broadcast of its word
  auto row_cost = [&](int r) { return 2 * (18 + 36 * (g.row_start[r + 1] - g.row_start[r]) + (g.row_p_col[r] >= 0 ? 51 : 11)); };
  auto col_cost = [&](int c) { return 130 + 27 * g.col_deg[c] + 6 * C; };
  // check rows: any warp of any CTA (their messages travel to the column owners anyway)
  const int pieces = p.Zw / 16;
  std::vector<std::pair<int, int>> rows;
  for (int r = 0; r < g.nrows; r++) {
    if (g.row_start[r + 1] - g.row_start[r] == kClSplitRowDeg)
      for (int k = 0; k < pieces; k++) rows.push_back({row_cost(r) / 2 + 80, r | (k << 8) | kClSplitItem});
    else
      for (int k = 0; k < chunks; k++) rows.push_back({row_cost(r), r | (k << 8)});
  }
  lpt(rows, C * T, s->cn_start, s->cn_items);
  // bit columns: whole columns to CTAs (longest first onto the lightest CTA), then each CTA's (column, chunk) items onto its T warps
  std::vector<std::pair<int, int>> cols;
  for (int c = 0; c < g.ncols; c++) if (g.col_deg[c] >= 2) cols.push_back({col_cost(c), c});
  std::stable_sort(cols.begin(), cols.end(), [](const std::pair<int, int> &a, const std::pair<int, int> &b) { return a.first > b.first; });
  std::vector<int> load(C, 0);
  std::vector<std::vector<std::pair<int, int>>> mine(C);
  for (auto &cc : cols) {
    int b = 0;
    for (int i = 1; i < C; i++) if (load[i] < load[b]) b = i;
    load[b] += cc.first;
    s->col_rank[cc.second] = (uint8_t)b;
    if (g.col_start[cc.second + 1] - g.col_start[cc.second] >= kClSplitColDeg)
      for (int k = 0; k < pieces; k++) mine[b].push_back({cc.first / 2 + 40, cc.second | (k << 8) | kClSplitItem});
    else
      for (int k = 0; k < chunks; k++) mine[b].push_back({cc.first, cc.second | (k << 8)});
  }
  int n = 0;
  for (int r = 0; r < C; r++) {
    int16_t st[kClMaxWarps + 1], items[6 * kMaxCols];
    lpt(mine[r], T, st, items);
    for (int w = 0; w < T; w++) {
      s->bn_start[r * T + w] = (int16_t)n;
      for (int i = st[w]; i < st[w + 1]; i++) s->bn_items[n++] = items[i];
    }
  }
  s->bn_start[C * T] = (int16_t)n;
  for (int m = 0; m < g.nreal; m++) s->edge_rank[m] = s->col_rank[g.edge_col[m]];
  const uint32_t ZB = 4u * (uint32_t)p.Zw;
  for (int r = 0; r < C; r++) s->cn_tx[r] = 4u * (uint32_t)C;
  for (int m = 0; m < g.nreal; m++) s->cn_tx[s->edge_rank[m]] += ZB + 4u;
  s->bn_tx = 2u * ZB * (uint32_t)cols.size();
  return true;
}

}  // namespace nrb200
