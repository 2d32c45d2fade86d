/*
 * TEST INFRASTRUCTURE ONLY -- see nrb200_oracle.h.  Scalar C restatement of the reference's arithmetic,
 * indexed by base-graph edges (row, column, shift) instead of the reference's degree-grouped SIMD buffers.
 * Parity status: PINNED against oracle/_ref (the compiled reference) and tests/golden/.
 */
#include "nrb200_oracle.h"
#include "../openairinterface5g_b200/csrc/nr_bg_tables.h"
#include <stdlib.h>
#include <string.h>

/* ---- lifting sizes: Z = a * 2^j, a in {2,3,5,7,9,11,13,15} (TS 38.212 Table 5.3.2-1; ldpctest.c:238-240) ---- */
int orc_ils_of_z(int Z)
{
  if (Z < 2 || Z > 384) return -1;
  int a = Z, j = 0;
  while ((a & 1) == 0) { a >>= 1; j++; }
  static const int odd[8] = {1, 3, 5, 7, 9, 11, 13, 15};
  for (int i = 0; i < 8; i++) {
    if (a != odd[i]) continue;
    if (i == 0) { if (j < 1 || j > 8) return -1; return 0; }      /* 2..256 */
    static const int jmax[8] = {0, 7, 6, 5, 5, 5, 4, 4};           /* 3*2^7=384 5*2^6=320 7*2^5=224 9*2^5=288 11*2^5=352 13*2^4=208 15*2^4=240 */
    if (j > jmax[i]) return -1;
    return i;
  }
  return -1;
}

/* nrLDPCdecoder_defs.h:53-57,80-84 */
int orc_ncols_for_rate(int BG, int R)
{
  if (BG == 1) return R == 13 ? 68 : R == 23 ? 35 : R == 89 ? 27 : -1;
  if (BG == 2) return R == 15 ? 52 : R == 13 ? 32 : R == 23 ? 17 : -1;
  return -1;
}

/* ---- CRC (crc_byte.c:46-58 polynomials, :148-312 table drivers) ---- */
static const uint32_t orc_poly[8] = {0x864cfb00u, 0x80006300u, 0xb2b11700u, 0x10210000u, 0x80F00000u, 0xc4200000u, 0x9B000000u, 0x84000000u};

uint32_t orc_crc(int poly_id, const uint8_t *data, uint32_t bitlen)
{
  /* crcbit() (crc_byte.c:64-98): shift register, MSB first, zero initial state, result left-aligned */
  const uint32_t poly = orc_poly[poly_id];
  uint32_t crc = 0;
  for (uint32_t i = 0; i < bitlen; i++) {
    uint32_t bit = (data[i >> 3] >> (7 - (i & 7))) & 1u;
    uint32_t top = (crc >> 31) ^ bit;
    crc <<= 1;
    if (top) crc ^= poly;
  }
  return crc;
}

int orc_check_crc(const uint8_t *d, uint32_t n, int crc_type)
{
  /* crc_byte.c:314-379 */
  int crc_len = crc_type <= 1 ? 3 : crc_type == 2 ? 2 : 1;
  uint32_t oldcrc = 0, crc;
  for (int i = 0; i < crc_len; i++) oldcrc |= (uint32_t)d[(n >> 3) - crc_len + i] << ((crc_len - 1 - i) << 3);
  switch (crc_type) {
    case 0: crc = orc_crc(0, d, n - 24) >> 8; break;
    case 1: crc = orc_crc(1, d, n - 24) >> 8; break;
    case 2: crc = orc_crc(3, d, n - 16) >> 16; break;
    default: crc = orc_crc(6, d, n - 8) >> 24; break;
  }
  return crc == oldcrc;
}

/* ---- graph access ---- */
typedef struct { int nrows, ncols, nedges; const uint8_t *row, *col; const uint16_t *shift; } orc_bg_t;
static orc_bg_t orc_bg(int BG, int ils)
{
  orc_bg_t g;
  if (BG == 1) { g.nrows = NRB200_BG1_NROWS; g.ncols = NRB200_BG1_NCOLS; g.nedges = NRB200_BG1_NEDGES; g.row = NRB200_BG1_ROW; g.col = NRB200_BG1_COL; g.shift = NRB200_BG1_SHIFT[ils]; }
  else         { g.nrows = NRB200_BG2_NROWS; g.ncols = NRB200_BG2_NCOLS; g.nedges = NRB200_BG2_NEDGES; g.row = NRB200_BG2_ROW; g.col = NRB200_BG2_COL; g.shift = NRB200_BG2_SHIFT[ils]; }
  return g;
}

/* Reference-defect emulation switches (default 0 = the arithmetic every other rate and the AVX512 build implement).
 * bit 0: the AVX2 code generator unrolls the BG2 degree-3 check-node group with `i+=2`
 *        (nrLDPC_tools/generator_cnProc/cnProc_gen_BG2_avx2.c -> cnProc/nrLDPC_cnProc_BG2_R15_AVX2.h:10,23,36), so in the
 *        AVX2 build of BG2 R=15 every odd 32-byte vector of that group's cn->bn messages stays 0.  */
static int orc_quirks = 0;
void orc_set_quirks(int q) { orc_quirks = q; }

static inline int sat8(int v) { return v > 127 ? 127 : v < -128 ? -128 : v; }

/* ---- decoder (nrLDPC_decoder.c:206-881) ---- */
int orc_ldpc_decode(int BG, int Z, int R, int numMaxIter, int outMode, const int8_t *llr, int8_t *out, int use_crc,
                    uint32_t crc_len_bits, int crc_type, int abort_in)
{
  const int ils = orc_ils_of_z(Z);
  const int ncols = orc_ncols_for_rate(BG, R);
  if (ils < 0 || ncols < 0) return -1;
  const orc_bg_t g = orc_bg(BG, ils);
  const int nsys = BG == 1 ? 22 : 10;
  const int nrows = ncols - nsys;          /* rate LUT R keeps the first nrows base-graph rows (nrLDPCdecoder_defs.h:53-84) */
  int ne = 0;
  while (ne < g.nedges && g.row[ne] < nrows) ne++;
  const int numLLR = ncols * Z;
  int coldeg[68] = {0};
  for (int e = 0; e < ne; e++) coldeg[g.col[e]]++;

  int8_t *Q = malloc((size_t)ne * Z);      /* bn->cn, indexed [edge][check lift]  (cnProcBuf)    */
  int8_t *Rm = malloc((size_t)ne * Z);     /* cn->bn, same indexing                (cnProcBufRes) */
  int8_t *llrRes = calloc((size_t)numLLR, 1);
  uint8_t *bits = calloc((size_t)numLLR / 8 + 8, 1);

  /* llr2CnProcBuf (nrLDPC_mPass.h:128-169): every edge starts with the channel LLR of its bit node */
  for (int e = 0; e < ne; e++) {
    const int c = g.col[e], s = g.shift[e] % Z;
    for (int t = 0; t < Z; t++) Q[(size_t)e * Z + t] = llr[c * Z + (t + s) % Z];
  }

  /* rows grouped by degree, ascending row order inside a group (the reference's lut_startAddrCnGroups layout) */
  int pc_from[46], rowdeg[46] = {0};
  for (int e = 0; e < ne; e++) rowdeg[g.row[e]]++;
  for (int r = 0; r < nrows; r++) pc_from[r] = Z;
  for (int d = 1; d <= 19; d++) {
    int members[46], n = 0;
    for (int r = 0; r < nrows; r++) if (rowdeg[r] == d) members[n++] = r;
    if (n == 0 || (n * Z) % 32 != 0) continue;
    for (int idx = n * Z - 32; idx < n * Z; idx++) { const int r = members[idx / Z], t = idx % Z; if (t < pc_from[r]) pc_from[r] = t; }
  }

  int numIter = 0, pcRes = 1, crc_ok_break = 0;
  for (;;) {
    if (numIter >= 1) {                   /* while ((numIter <= numMaxIter) && (pcRes != 0)) (nrLDPC_decoder.c:552) */
      if (!(numIter <= numMaxIter && pcRes != 0)) break;
    }
    numIter++;
    if (numIter > 1 && abort_in) { numIter = numMaxIter + 2; break; }   /* :557-560 */

    /* cnProc (nrLDPC_cnProc.h:388-877): exclude-self sign product x min magnitude, clipped to 127;
       abs_epi8(-128) stays 0x80 = 128 unsigned, sign_epi8(x, 0) = 0 */
    int e0 = 0, deg3_idx = -1;
    while (e0 < ne) {
      int e1 = e0;
      while (e1 < ne && g.row[e1] == g.row[e0]) e1++;
      if (e1 - e0 == 3) deg3_idx++;      /* position of this check row inside the degree-3 group (rows in ascending order) */
      for (int t = 0; t < Z; t++) {
        for (int j = e0; j < e1; j++) {
          int mn = 255, sg = 1;
          for (int k = e0; k < e1; k++) {
            if (k == j) continue;
            int v = Q[(size_t)k * Z + t];
            int a = v < 0 ? -v : v;      /* 128 for -128 */
            if (a < mn) mn = a;
            sg = v < 0 ? -sg : v == 0 ? 0 : sg;
          }
          if (mn > 127) mn = 127;
          Rm[(size_t)j * Z + t] = (int8_t)(sg < 0 ? -mn : sg == 0 ? 0 : mn);
          if ((orc_quirks & 1) && BG == 2 && R == 15 && e1 - e0 == 3 && (((deg3_idx * Z + t) >> 5) & 1)) Rm[(size_t)j * Z + t] = 0;
        }
      }
      e0 = e1;
    }

    /* bnProcPc (nrLDPC_bnProc.h:40-263): llrRes = sat8(sum16(cn->bn) + llr); the generated (UNROLL_BN_PROC_PC) code
       skips degree-1 bit nodes, whose llrRes therefore stays 0 (nrLDPC_decoder.c:222-227 zero-initialises it).
       bnProc (:271-1313): bn->cn = subs_epi8(llrRes, cn->bn); degree-1 edges keep the channel LLR
       (bn2cnProcBuf skips them, nrLDPC_mPass.h:350-351). */
    for (int c = 0; c < ncols; c++) {
      if (coldeg[c] < 2) continue;
      for (int v = 0; v < Z; v++) {
        int acc = llr[c * Z + v];
        for (int e = 0; e < ne; e++)
          if (g.col[e] == c) acc += Rm[(size_t)e * Z + ((v - g.shift[e] % Z + Z) % Z)];
        llrRes[c * Z + v] = (int8_t)sat8(acc);
      }
    }
    for (int e = 0; e < ne; e++) {
      const int c = g.col[e], s = g.shift[e] % Z;
      if (coldeg[c] < 2) continue;
      for (int t = 0; t < Z; t++)
        Q[(size_t)e * Z + t] = (int8_t)sat8(llrRes[c * Z + (t + s) % Z] - Rm[(size_t)e * Z + t]);
    }

    if (numIter == 1) continue;          /* no parity check after the first iteration (:541-547) */

    if (!use_crc) {
      /* cnProcPc (nrLDPC_cnProc.h:887-1960): XOR over a check's edges of sign(adds_epi8(cnProcBuf, cnProcBufRes)).
         The reference walks each check-node degree group as ceil(n_g*Z/32) 32-byte vectors, checks the first M32-1 of them and
         ORs in the last one only `if (Mrem)` (:964-965, same in every group) -- so when n_g*Z is a multiple of 32 (always for
         Z = 384) the final 32 check nodes of the group are never tested.  pc_from[r] = first unchecked lift of row r. */
      pcRes = 0;
      int a0 = 0;
      while (a0 < ne && !pcRes) {
        int a1 = a0;
        while (a1 < ne && g.row[a1] == g.row[a0]) a1++;
        for (int t = 0; t < pc_from[g.row[a0]] && !pcRes; t++) {
          int par = 0;
          for (int k = a0; k < a1; k++) par ^= (sat8(Q[(size_t)k * Z + t] + Rm[(size_t)k * Z + t]) < 0);
          pcRes |= par;
        }
        a0 = a1;
      }
    } else if (numIter > 2) {            /* :850-862 */
      for (int i = 0; i < numLLR; i++) if (llrRes[i] < 0) bits[i >> 3] |= (uint8_t)(0x80 >> (i & 7)); else bits[i >> 3] &= (uint8_t)~(0x80 >> (i & 7));
      if (outMode == 0) memcpy(out, bits, (size_t)numLLR / 8);
      else for (int i = 0; i < numLLR; i++) out[i] = llrRes[i] < 0;   /* LLRINT8: see note at the final hard decision */
      if (orc_check_crc(outMode == 0 ? bits : (const uint8_t *)out, crc_len_bits, crc_type)) { crc_ok_break = 1; break; }
    }
  }
  (void)crc_ok_break;
  if (!use_crc) {                        /* :865-877 */
    memset(bits, 0, (size_t)numLLR / 8 + 8);
    for (int i = 0; i < numLLR; i++) if (llrRes[i] < 0) bits[i >> 3] |= (uint8_t)(0x80 >> (i & 7));
    if (outMode == 0) memcpy(out, bits, ((size_t)numLLR + 7) / 8);
    else for (int i = 0; i < numLLR; i++) out[i] = llrRes[i] < 0;
    /* outMode LLRINT8: the reference aliases p_llrOut = p_out and then still runs nrLDPC_llr2bit(p_out, p_llrOut)
       (nrLDPC_decoder.c:866-877), so the "LLR" output is overwritten in place by 0/1 hard bits == BITINT8. */
  }
  free(Q); free(Rm); free(llrRes); free(bits);
  return numIter;
}

/* ---- encoder: systematic QC encoding through H's dual-diagonal core (result identical to the reference's
 *      generator-matrix XOR networks, ldpc_encode_parity_check.c:90-220 / ldpc_generate_coefficient.c:363-428) ---- */
int orc_ldpc_encode(int BG, int Z, int K, const uint8_t *in, uint8_t *out)
{
  const int ils = orc_ils_of_z(Z);
  if (ils < 0) return -1;
  const orc_bg_t g = orc_bg(BG, ils);
  const int nsys = BG == 1 ? 22 : 10;
  if (K != nsys * Z) return -1;
  uint8_t *x = calloc((size_t)g.ncols * Z, 1);
  for (int i = 0; i < K; i++) x[i] = (in[i >> 3] >> (7 - (i & 7))) & 1;
  uint8_t *lam = calloc((size_t)4 * Z, 1);
  /* lambda_i = sum over systematic columns of row i, i = 0..3 */
  for (int e = 0; e < g.nedges; e++) {
    const int r = g.row[e], c = g.col[e], s = g.shift[e] % Z;
    if (r >= 4 || c >= nsys) continue;
    for (int t = 0; t < Z; t++) lam[r * Z + t] ^= x[c * Z + (t + s) % Z];
  }
  /* first core parity column p0 (= column nsys): summing the 4 core rows cancels the other core parity columns and
     leaves x^sigma * p0, sigma = the shift that occurs an odd number of times in column nsys */
  int cnt[384] = {0}, sigma = -1;
  for (int e = 0; e < g.nedges; e++) if (g.row[e] < 4 && g.col[e] == nsys) cnt[g.shift[e] % Z] ^= 1;
  for (int s = 0; s < Z; s++) if (cnt[s]) { if (sigma >= 0) { free(x); free(lam); return -2; } sigma = s; }
  for (int t = 0; t < Z; t++) x[nsys * Z + (t + sigma) % Z] = lam[t] ^ lam[Z + t] ^ lam[2 * Z + t] ^ lam[3 * Z + t];
  /* remaining 3 core parity columns: repeatedly take a core row with exactly one unknown parity column */
  int known[68] = {0};
  for (int c = 0; c <= nsys; c++) known[c] = 1;
  for (int pass = 0; pass < 3; pass++) {
    for (int r = 0; r < 4; r++) {
      int unk = -1, nunk = 0, sunk = 0;
      for (int e = 0; e < g.nedges; e++) if (g.row[e] == r && !known[g.col[e]]) { unk = g.col[e]; sunk = g.shift[e] % Z; nunk++; }
      if (nunk != 1) continue;
      for (int t = 0; t < Z; t++) {
        int acc = 0;
        for (int e = 0; e < g.nedges; e++) if (g.row[e] == r && known[g.col[e]]) acc ^= x[g.col[e] * Z + (t + g.shift[e] % Z) % Z];
        x[unk * Z + (t + sunk) % Z] = (uint8_t)acc;
      }
      known[unk] = 1;
    }
  }
  /* extension rows: degree-1 parity column nsys+r with shift 0 */
  for (int r = 4; r < g.nrows; r++) {
    for (int t = 0; t < Z; t++) {
      int acc = 0;
      for (int e = 0; e < g.nedges; e++) if (g.row[e] == r && g.col[e] != nsys + r) acc ^= x[g.col[e] * Z + (t + g.shift[e] % Z) % Z];
      x[(nsys + r) * Z + t] = (uint8_t)acc;
    }
  }
  /* output: K-2Z systematic then all parity (ldpc_encoder_optim8segmulti.c:175-207) */
  memcpy(out, x + 2 * Z, (size_t)(g.ncols - 2) * Z);
  free(x); free(lam);
  return 0;
}

/* ---- segmentation (nr_segmentation.c:32-180) ---- */
int orc_segmentation(const uint8_t *in, uint8_t **seg_out, unsigned B, unsigned *C, unsigned *K, unsigned *Zout, unsigned *F, int BG)
{
  unsigned L, Bprime, Z, Kcb = BG == 1 ? 8448 : 3840, Kb, Kprime;
  if (B <= Kcb) { L = 0; *C = 1; Bprime = B; }
  else { L = 24; *C = B / (Kcb - L); if ((Kcb - L) * (*C) < B) (*C)++; Bprime = B + (*C) * L; }
  Kprime = Bprime / (*C);
  if (BG == 1) Kb = 22; else Kb = B > 640 ? 10 : B > 560 ? 9 : B > 192 ? 8 : 6;
  Z = (Kprime % Kb) ? Kprime / Kb + 1 : Kprime / Kb;
  unsigned step;
  if (Z <= 2) { *K = 2; step = 0; }
  else if (Z <= 16) { *K = Z; step = 0; }
  else if (Z <= 32) step = 2; else if (Z <= 64) step = 4; else if (Z <= 128) step = 8; else if (Z <= 256) step = 16; else if (Z <= 384) step = 32; else return -1;
  if (step) { *K = (Z / step) * step; if (*K < Z) *K += step; }
  *Zout = *K;
  *K = *K * (BG == 1 ? 22 : 10);
  *F = *K - Kprime;
  if (in && seg_out) {
    unsigned s = 0;
    for (unsigned r = 0; r < *C; r++) {
      unsigned k = 0;
      while (k < ((Kprime - L) >> 3)) seg_out[r][k++] = in[s++];
      if (*C > 1) {
        uint32_t crc = orc_crc(1, seg_out[r], Kprime - L) >> 8;
        seg_out[r][(Kprime - L) >> 3] = (uint8_t)(crc >> 16);
        seg_out[r][1 + ((Kprime - L) >> 3)] = (uint8_t)(crc >> 8);
        seg_out[r][2 + ((Kprime - L) >> 3)] = (uint8_t)crc;
      }
      if (*F > 0) for (k = Kprime >> 3; k < (*K) >> 3; k++) seg_out[r][k] = 0;
    }
  }
  return (int)Kb;
}

/* ---- rate matching (nr_rate_matching.c) ---- */
static const uint8_t orc_k0[2][4] = {{0, 17, 33, 56}, {0, 13, 25, 43}};   /* :34 */

int orc_rate_matching_tx(uint32_t Tbslbrm, int BG, int Z, const uint8_t *w, uint8_t *e, int C, uint32_t F, uint32_t Foffset, int rv, uint32_t E)
{
  if (C == 0) return -1;
  uint32_t N = (BG == 1 ? 66 : 50) * (uint32_t)Z, Ncb = N, k = 0;
  if (Tbslbrm) { uint32_t Nref = 3 * Tbslbrm / (2 * C); Ncb = N < Nref ? N : Nref; }
  uint32_t ind = (orc_k0[BG - 1][rv] * Ncb / N) * Z;
  if (Foffset > E || Foffset > Ncb) return -1;
  if (ind >= Foffset && ind < F + Foffset) ind = F + Foffset;
  if (ind < Foffset) {
    memcpy(e, w + ind, Foffset - ind);
    if (E + F <= Ncb - ind) { memcpy(e + Foffset - ind, w + Foffset + F, E - Foffset + ind); k = E; }
    else { memcpy(e + Foffset - ind, w + Foffset + F, Ncb - Foffset - F); k = Ncb - F - ind; }
  } else {
    if (E <= Ncb - ind) { memcpy(e, w + ind, E); k = E; }
    else { memcpy(e, w + ind, Ncb - ind); k = Ncb - ind; }
  }
  while (k < E)
    for (ind = 0; ind < Ncb && k < E; ind++) if (w[ind] != 2 /* NR_NULL */) e[k++] = w[ind];
  return 0;
}

int orc_rate_matching_rx(uint32_t Tbslbrm, int BG, int Z, int16_t *w, const int16_t *soft, int C, int rv, int clear, uint32_t E, uint32_t F, uint32_t Foffset)
{
  if (C == 0) return -1;
  uint32_t N = (BG == 1 ? 66 : 50) * (uint32_t)Z, Ncb = N, k = 0;
  if (Tbslbrm) { uint32_t Nref = 3 * Tbslbrm / (2 * C); Ncb = N < Nref ? N : Nref; }
  uint32_t ind = (orc_k0[BG - 1][rv] * Ncb / N) * Z;
  if (Foffset > E || Foffset > Ncb) return -1;
  if (clear == 1) memset(w, 0, Ncb * sizeof(int16_t));
  if (ind < Foffset) for (; ind < Foffset && k < E; ind++) w[ind] = (int16_t)(w[ind] + soft[k++]);
  if (ind >= Foffset && ind < Foffset + F) ind = Foffset + F;
  for (; ind < Ncb && k < E; ind++) w[ind] = (int16_t)(w[ind] + soft[k++]);
  while (k < E) {
    for (ind = 0; ind < Foffset && k < E; ind++) w[ind] = (int16_t)(w[ind] + soft[k++]);
    for (ind = Foffset + F; ind < Ncb && k < E; ind++) w[ind] = (int16_t)(w[ind] + soft[k++]);
  }
  return 0;
}

void orc_interleave(uint32_t E, int Qm, const uint8_t *e, uint8_t *f)
{
  const uint32_t EQm = E / Qm;
  memset(f, 0, E);
  for (uint32_t j = 0; j < EQm; j++) for (int i = 0; i < Qm; i++) f[j * Qm + i] = e[i * EQm + j];
}

void orc_deinterleave(uint32_t E, int Qm, int16_t *e, const int16_t *f)
{
  const uint32_t EQm = E / Qm;
  for (uint32_t j = 0; j < EQm; j++) for (int i = 0; i < Qm; i++) e[i * EQm + j] = f[j * Qm + i];
}

int orc_get_R_ldpc_decoder(int rv, int E, int BG, int Z, int *llrLen, int round)
{
  int Ncb = (BG == 1 ? 66 : 50) * Z;
  int infoBits = orc_k0[BG - 1][rv] * Z + E;
  if (round == 0) *llrLen = infoBits;
  if (infoBits > Ncb) infoBits = Ncb;
  if (infoBits > *llrLen) *llrLen = infoBits;
  int sysBits = (BG == 1 ? 22 : 10) * Z;
  float decoderR = (float)sysBits / (infoBits + 2 * Z);
  if (BG == 2) return decoderR < 0.3333 ? 15 : decoderR < 0.6667 ? 13 : 23;
  return decoderR < 0.6667 ? 13 : decoderR < 0.8889 ? 23 : 89;
}

/* ---- PUSCH single-layer max-log LLRs (openair1/PHY/NR_TRANSPORT/nr_ulsch_llr_computation.c:45-312) ----
 * rxF, mag*: nb_re complex int16 {re, im}; out: nb_re * Qm int16.  abs_epi16(-32768) stays -32768; subs_epi16 saturates. */
static inline int orc_abs16(int v) { return v == -32768 ? -32768 : (v < 0 ? -v : v); }
static inline int orc_subs16(int a, int b) { int d = a - b; return d > 32767 ? 32767 : d < -32768 ? -32768 : d; }
void orc_ulsch_llr(int Qm, const int16_t *rxF, const int16_t *maga, const int16_t *magb, const int16_t *magc, int16_t *out, uint32_t nb_re)
{
  for (uint32_t i = 0; i < nb_re; i++) {
    const int yr = rxF[2 * i], yi = rxF[2 * i + 1];
    int16_t *o = out + (size_t)i * Qm;
    if (Qm == 2) { o[0] = (int16_t)(yr >> 3); o[1] = (int16_t)(yi >> 3); continue; }   /* :45-58 */
    const int ar = orc_subs16(maga[2 * i], orc_abs16(yr)), ai = orc_subs16(maga[2 * i + 1], orc_abs16(yi));
    o[0] = (int16_t)yr; o[1] = (int16_t)yi; o[2] = (int16_t)ar; o[3] = (int16_t)ai;
    if (Qm == 4) continue;                                                                /* :64-118 */
    const int br = orc_subs16(magb[2 * i], orc_abs16(ar)), bi = orc_subs16(magb[2 * i + 1], orc_abs16(ai));
    o[4] = (int16_t)br; o[5] = (int16_t)bi;
    if (Qm == 6) continue;                                                                /* :124-196 */
    o[6] = (int16_t)orc_subs16(magc[2 * i], orc_abs16(br)); o[7] = (int16_t)orc_subs16(magc[2 * i + 1], orc_abs16(bi));   /* :198-282 */
  }
}

/* ---- Gold sequence, (un)scrambling, QAM mapper ----
 * lte_gold_generic (openair1/PHY/LTE_TRANSPORT/transport_proto.h:633-680): c(n) = x1(n+1600) ^ x2(n+1600), produced 32 bits per call
 * (bit i of the word = c(32w + i)); nr_codeword_scrambling / _unscrambling (NR_TRANSPORT/nr_scrambling.c:30-96);
 * nr_modulation (MODULATION/nr_modulation.c:115-244) with the tables of NR_REFSIG/nr_gen_mod_table.c:33-98. */
static void orc_gold_step(uint32_t *x1, uint32_t *x2)
{
  *x1 = (*x1 >> 1) ^ (*x1 >> 4);
  *x1 = *x1 ^ (*x1 << 31) ^ (*x1 << 28);
  *x2 = (*x2 >> 1) ^ (*x2 >> 2) ^ (*x2 >> 3) ^ (*x2 >> 4);
  *x2 = *x2 ^ (*x2 << 31) ^ (*x2 << 30) ^ (*x2 << 29) ^ (*x2 << 28);
}
void orc_gold_words(uint32_t c_init, uint32_t n_words, uint32_t *out)
{
  uint32_t x1 = 1u + (1u << 31), x2 = c_init;
  x2 = x2 ^ ((x2 ^ (x2 >> 1) ^ (x2 >> 2) ^ (x2 >> 3)) << 31);
  for (int n = 1; n < 50; n++) orc_gold_step(&x1, &x2);
  for (uint32_t w = 0; w < n_words; w++) { orc_gold_step(&x1, &x2); out[w] = x1 ^ x2; }
}
/* in: one bit per byte (the rate matcher's output); out: ceil(size/32) words, bit i of word w = in[32w+i] ^ c(32w+i) */
void orc_scramble(const uint8_t *in, uint32_t size, uint32_t q, uint32_t Nid, uint32_t n_RNTI, uint32_t *out)
{
  const uint32_t nw = (size + 31) >> 5;
  orc_gold_words((n_RNTI << 15) + (q << 14) + Nid, nw, out);
  for (uint32_t w = 0; w < nw; w++) {
    uint32_t v = 0;
    for (int i = 0; i < 32; i++) if (32 * w + i < size) v |= (uint32_t)(in[32 * w + i] & 1) << i;   /* movemask(slli_epi16(c, 7)) */
    out[w] ^= v;
  }
}
/* llr[i] *= (1 - 2 c(i)) with mullo_epi16 (so -32768 stays -32768) */
void orc_unscramble_llr(int16_t *llr, uint32_t size, uint32_t q, uint32_t Nid, uint32_t n_RNTI)
{
  const uint32_t nw = (size + 31) >> 5;
  uint32_t *c = malloc(4 * (size_t)nw + 4);
  orc_gold_words((n_RNTI << 15) + (q << 14) + Nid, nw, c);
  for (uint32_t i = 0; i < size; i++) if ((c[i >> 5] >> (i & 31)) & 1) llr[i] = (int16_t)(uint16_t)(0u - (uint16_t)llr[i]);
  free(c);
}
/* bits: packed LSB-first (what orc_scramble produces); out: length/Qm symbols {re, im} */
void orc_modulate(const uint8_t *bits, uint32_t length, int Qm, int16_t *out)
{
  const float val = 32768.0f, s2 = 0.70711f, s10 = 0.31623f, s42 = 0.15430f, s170 = 0.076696f;
  for (uint32_t i = 0; i < length / Qm; i++) {
    int b[8];
    for (int j = 0; j < Qm; j++) { const uint32_t n = i * Qm + j; b[j] = 1 - 2 * ((bits[n >> 3] >> (n & 7)) & 1); }
    short lr, li;
    float sc;
    if (Qm == 2) { lr = (short)b[0]; li = (short)b[1]; sc = s2; }
    else if (Qm == 4) { lr = (short)(b[0] * (2 - b[2])); li = (short)(b[1] * (2 - b[3])); sc = s10; }
    else if (Qm == 6) { lr = (short)(b[0] * (4 - b[2] * (2 - b[4]))); li = (short)(b[1] * (4 - b[3] * (2 - b[5]))); sc = s42; }
    else { lr = (short)(b[0] * (8 - b[2] * (4 - b[4] * (2 - b[6])))); li = (short)(b[1] * (8 - b[3] * (4 - b[5] * (2 - b[7])))); sc = s170; }
    out[2 * i] = (short)(lr * val * sc * s2);        /* float32, evaluated left to right like nr_gen_mod_table.c */
    out[2 * i + 1] = (short)(li * val * sc * s2);
  }
}
