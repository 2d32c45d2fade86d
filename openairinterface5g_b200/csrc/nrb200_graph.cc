// Host-side construction of the lifted-graph descriptors (see nrb200_graph.h).
#include "nrb200_graph.h"
#include "nr_bg_tables.h"
#include <cstring>

namespace nrb200 {

int ils_of_z(int Z)
{
  // Z = a * 2^j, a in {2,3,5,7,9,11,13,15} (TS 38.212 Table 5.3.2-1)
  if (Z < 2 || Z > 384) return -1;
  int a = Z, j = 0;
  while ((a & 1) == 0) { a >>= 1; j++; }
  static const int odd[8] = {1, 3, 5, 7, 9, 11, 13, 15};
  static const int jmax[8] = {8, 7, 6, 5, 5, 5, 4, 4};
  for (int i = 0; i < 8; i++)
    if (a == odd[i]) return (j <= jmax[i] && (i != 0 || j >= 1)) ? i : -1;
  return -1;
}

int ncols_for_rate(int BG, int R)
{
  if (BG == 1) return R == 13 ? 68 : R == 23 ? 35 : R == 89 ? 27 : -1;
  if (BG == 2) return R == 15 ? 52 : R == 13 ? 32 : R == 23 ? 17 : -1;
  return -1;
}

namespace {
struct Bg { int nrows, ncols, nedges; const uint8_t *row, *col; const uint16_t *shift; };
Bg bg_of(int BG, int ils)
{
  if (BG == 1) return {NRB200_BG1_NROWS, NRB200_BG1_NCOLS, NRB200_BG1_NEDGES, NRB200_BG1_ROW, NRB200_BG1_COL, NRB200_BG1_SHIFT[ils]};
  return {NRB200_BG2_NROWS, NRB200_BG2_NCOLS, NRB200_BG2_NEDGES, NRB200_BG2_ROW, NRB200_BG2_COL, NRB200_BG2_SHIFT[ils]};
}
}  // namespace

bool build_graph(int BG, int Z, int R, GraphDev *g)
{
  const int ils = ils_of_z(Z);
  const int ncols = ncols_for_rate(BG, R);
  if (ils < 0 || ncols < 0) return false;
  const Bg b = bg_of(BG, ils);
  std::memset(g, 0, sizeof(*g));
  g->BG = BG; g->Z = Z; g->R = R; g->ils = ils;
  g->ncols = ncols; g->nsys = BG == 1 ? 22 : 10; g->nrows = ncols - g->nsys;
  int ne = 0;
  while (ne < b.nedges && b.row[ne] < g->nrows) ne++;
  int deg[kMaxCols] = {0};
  for (int e = 0; e < ne; e++) deg[b.col[e]]++;
  for (int c = 0; c < ncols; c++) g->col_deg[c] = (int16_t)deg[c];
  // slots in row-major order, degree-1 edges excluded
  int slot_of[kMaxEdges];
  int m = 0, deg3 = 0;
  for (int r = 0; r < g->nrows; r++) { g->row_p_col[r] = -1; g->row_deg3_idx[r] = -1; }
  int e = 0;
  for (int r = 0; r < g->nrows; r++) {
    g->row_start[r] = (int16_t)m;
    int rowdeg = 0;
    for (; e < ne && b.row[e] == r; e++, rowdeg++) {
      const int c = b.col[e], s = b.shift[e] % Z;
      if (deg[c] >= 2) { slot_of[e] = m; g->edge_col[m] = (int16_t)c; g->edge_shift[m] = (int16_t)s; m++; }
      else {
        if (g->row_p_col[r] >= 0) return false;  // at most one degree-1 neighbour per check row in NR graphs
        slot_of[e] = -1; g->row_p_col[r] = (int16_t)c; g->row_p_shift[r] = (int16_t)s;
      }
    }
    if (rowdeg == 3) { g->row_deg3_idx[r] = (int16_t)deg3++; }
  }
  g->row_start[g->nrows] = (int16_t)m;
  g->nreal = m;
  // Parity-check coverage of the reference (nrLDPC_cnProc.h:887-1960): each check-node degree group is scanned as
  // ceil(n_g*Z/32) 32-byte vectors, the last of which is only tested `if (Mrem)` (:964-965) -- when n_g*Z is a multiple
  // of 32 (always for Z = 384) the final 32 check nodes of the group never take part in the early-stop decision.
  // Rows sit in ascending order inside a group (lut_startAddrCnGroups layout).  Replicated for bit-exact iteration counts.
  {
    int rowdeg[kMaxRows] = {0};
    for (int ee = 0; ee < ne; ee++) rowdeg[b.row[ee]]++;
    for (int r = 0; r < g->nrows; r++) g->row_pc_from[r] = (int16_t)Z;
    for (int d = 1; d <= kMaxRowDeg; d++) {
      int members[kMaxRows], n = 0;
      for (int r = 0; r < g->nrows; r++) if (rowdeg[r] == d) members[n++] = r;
      if (n == 0 || (n * Z) % 32 != 0) continue;
      for (int idx = n * Z - 32; idx < n * Z; idx++) {
        const int r = members[idx / Z], t = idx % Z;
        if (t < g->row_pc_from[r]) g->row_pc_from[r] = (int16_t)t;
      }
    }
  }
  // column lists
  int k = 0;
  for (int c = 0; c < ncols; c++) {
    g->col_start[c] = (int16_t)k;
    if (deg[c] < 2) continue;
    for (int ee = 0; ee < ne; ee++) if (b.col[ee] == c) g->col_edges[k++] = (int16_t)slot_of[ee];
  }
  g->col_start[ncols] = (int16_t)k;
  return k == m;
}

bool build_enc_graph(int BG, int Z, EncGraphDev *g)
{
  const int ils = ils_of_z(Z);
  if (ils < 0 || (BG != 1 && BG != 2)) return false;
  const Bg b = bg_of(BG, ils);
  std::memset(g, 0, sizeof(*g));
  g->BG = BG; g->Z = Z; g->ils = ils; g->ncols = b.ncols; g->nrows = b.nrows; g->nsys = BG == 1 ? 22 : 10;
  int e = 0;
  for (int r = 0; r < b.nrows; r++) {
    g->row_start[r] = (int16_t)e;
    for (; e < b.nedges && b.row[e] == r; e++) { g->edge_col[e] = b.col[e]; g->edge_shift[e] = (int16_t)(b.shift[e] % Z); }
  }
  g->row_start[b.nrows] = (int16_t)e;
  // sigma: the shift occurring an odd number of times in column nsys over the 4 core rows
  int cnt[384] = {0};
  for (int i = 0; i < g->row_start[4]; i++) if (g->edge_col[i] == g->nsys) cnt[g->edge_shift[i]] ^= 1;
  g->sigma = -1;
  for (int s = 0; s < Z; s++) if (cnt[s]) { if (g->sigma >= 0) return false; g->sigma = s; }
  if (g->sigma < 0) return false;
  // order in which the remaining core parity columns become solvable
  bool known[kMaxCols] = {false};
  for (int c = 0; c <= g->nsys; c++) known[c] = true;
  int n = 0;
  for (int pass = 0; pass < 3 && n < 3; pass++)
    for (int r = 0; r < 4 && n < 3; r++) {
      int unk = -1, nunk = 0, sh = 0;
      for (int i = g->row_start[r]; i < g->row_start[r + 1]; i++)
        if (!known[g->edge_col[i]]) { unk = g->edge_col[i]; sh = g->edge_shift[i]; nunk++; }
      if (nunk != 1) continue;
      g->core_row[n] = (int16_t)r; g->core_col[n] = (int16_t)unk; g->core_shift[n] = (int16_t)sh; n++;
      known[unk] = true;
    }
  return n == 3;
}

}  // namespace nrb200
