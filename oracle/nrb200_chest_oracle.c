/* TEST INFRASTRUCTURE ONLY (see nrb200_oracle.h).  CPU restatement of the PUSCH channel estimator of the reference for DMRS
 * configuration types 1 and 2, with frequency-domain interpolation (chest_freq == 0) and with one average per PRB (chest_freq == 1),
 * transform precoding disabled:
 *   nr_pusch_channel_estimation    openair1/PHY/NR_ESTIMATION/nr_ul_channel_estimation.c:67-487
 *   nr_gold_pusch / nr_pusch_dmrs_rx  NR_REFSIG/nr_gold.c:99-116, NR_REFSIG/nr_dmrs_rx.c:44-116
 *   nr_est_delay, get_delay_idx, init_delay_table   common/utils/nr/nr_common.c:906-990
 *   c16multaddVectRealComplex + filt16_ul_*          PHY/TOOLS/tools_defs.h:266-297, NR_UE_ESTIMATION/filt16a_32.h:242-249
 * Pinned against the compiled reference (oracle/_ref/libref_chest.so) by tests/test_oracle_vs_reference.py. */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "nrb200_oracle.h"

static inline int16_t sat16(int32_t v) { return v > 32767 ? 32767 : v < -32768 ? -32768 : (int16_t)v; }
static inline int16_t wrap16(int32_t v) { return (int16_t)(uint16_t)(uint32_t)v; }
static inline int16_t mulhrs16(int a, int b) { return wrap16(((a * b) + 0x4000) >> 15); }

/* the pointer shift of some branches can address a few samples past the symbol: the next symbol's first sub-carriers, as in the reference; past the
 * slot's last symbol (the reference reads the next slot of its ring there) the restatement and the product read 0 */
#define RX_AT(rx, idx, c) (((idx) >= N && p->symbol >= 13) ? 0 : (rx)[2 * (idx) + (c)])

static const int16_t F_P0[16] = {4096, 4096, 4096, 4096, 4096, 4096, 4096, 4096, 0, 0, 0, 0, 0, 0, 0, 0};
static const int16_t F_P1P2[16] = {4096, 4096, 4096, 4096, 2048, 2048, 2048, 2048, 2048, 2048, 2048, 2048, 0, 0, 0, 0};
static const int16_t F_MID[16] = {2048, 2048, 2048, 2048, 2048, 2048, 2048, 2048, 2048, 2048, 2048, 2048, 2048, 2048, 2048, 2048};
static const int16_t F_LAST[16] = {4096, 4096, 4096, 4096, 8192, 8192, 8192, 8192, 0, 0, 0, 0, 0, 0, 0, 0};

/* conjugated DMRS of one symbol: 6 * rb_size c16 (type 1) or 4 * rb_size (type 2) */
/* Transform precoding: the estimator correlates with the conjugate of the low-PAPR type-1 sequence of (u, v), whatever the port, starting at element 0
 * whatever rb_start (nr_ul_channel_estimation.c:122-133 -> nr_pusch_lowpaprtype1_dmrs_rx(p = 1000, re_offset = 0), nr_dmrs_rx.c:258-300). */
static const int16_t *g_lowpapr_seq;
void orc_chest_set_lowpapr(const int16_t *seq) { g_lowpapr_seq = seq; }
/* base sequences of TS 38.211 5.2.2 as ul_ref_seq_nr.c:55-196 computes them (double precision, floor): M_ZC = 30 (5.2.2.2, closed form) and M_ZC >= 36
 * (5.2.2.1, Zadoff-Chu of the largest prime below M_ZC, cyclically extended).  The shorter ones (6, 12, 18, 24) are table look-ups of the specification
 * (phi tables) and are not restated: returns -1. */
int orc_lowpapr_seq(int u, int v, int M_ZC, int scaling, int16_t *out)
{
  if (M_ZC == 30) {
    for (int n = 0; n < M_ZC; n++) {
      const double x = -(M_PI * (u + 1) * (n + 1) * (n + 2)) / (double)31;
      out[2 * n] = (int16_t)floor(scaling * cos(x)); out[2 * n + 1] = (int16_t)floor(scaling * sin(x));
    }
    return 0;
  }
  if (M_ZC < 36) return -1;
  int N_ZC = M_ZC - 1;
  for (;; N_ZC--) { int pr = 1; for (int d = 2; d * d <= N_ZC; d++) if (N_ZC % d == 0) { pr = 0; break; } if (pr) break; }
  const double q_overbar = N_ZC * (u + 1) / (double)31;
  unsigned q = (((int)floor(2 * q_overbar)) & 1) == 0 ? (unsigned)((int)floor(q_overbar + .5) - v) : (unsigned)((int)floor(q_overbar + .5) + v);
  for (unsigned n = 0; n < (unsigned)M_ZC; n++) {
    const unsigned m = n % (unsigned)N_ZC;
    const double x = (double)q * m * (m + 1) / N_ZC;
    out[2 * n] = (int16_t)floor(scaling * cos(M_PI * x));
    out[2 * n + 1] = (int16_t)-(int16_t)floor(scaling * sin(M_PI * x));
  }
  return 0;
}

void orc_pusch_dmrs_pilots(const orc_chest_t *p, int16_t *pil)
{
  if (g_lowpapr_seq) {
    for (int k = 0; k < 6 * p->rb_size; k++) { pil[2 * k] = g_lowpapr_seq[2 * k]; pil[2 * k + 1] = (int16_t)-g_lowpapr_seq[2 * k + 1]; }
    return;
  }
  static const int wf1[8][2] = {{1, 1}, {1, -1}, {1, 1}, {1, -1}, {1, 1}, {1, -1}, {1, 1}, {1, -1}};
  const uint32_t nid = (uint32_t)p->dmrs_scrambling_id;
  const uint64_t t = ((1ULL << 17) * (uint64_t)(14 * p->slot + p->symbol + 1) * ((nid << 1) + 1) + ((nid << 1) + (uint32_t)p->scid));
  const uint32_t x2 = (uint32_t)(t % (1ULL << 31));
  const int dmrs_offset = ((p->bwp_start + p->rb_start) * 12) / (p->dmrs_type ? 3 : 2), n = (p->dmrs_type ? 4 : 6) * p->rb_size;   /* wf2 == wf1 for ports 0..7 */
  const uint32_t nw = (uint32_t)((2 * (dmrs_offset + n) + 31) / 32 + 1);
  uint32_t *g = malloc(4 * (size_t)nw);
  orc_gold_words(x2, nw, g);
  for (int i = dmrs_offset; i < dmrs_offset + n; i++) {
    const int w = wf1[p->port][i & 1];                          /* wt1[p][l' = 0] = 1 */
    const int b0 = (g[(2 * i) >> 5] >> ((2 * i) & 31)) & 1, b1 = (g[(2 * i + 1) >> 5] >> ((2 * i + 1) & 31)) & 1;
    const int idx = (b0 << 1) ^ b1;
    /* nr_rx_mod_table QPSK entries = conj of the transmitted symbol: idx 0 (+,-) 1 (+,+) 2 (-,-) 3 (-,+), negated when w == -1 */
    static const int sr[4] = {1, 1, -1, -1}, si[4] = {-1, 1, -1, 1};
    pil[2 * (i - dmrs_offset)] = (int16_t)(w * sr[idx] * 23170);
    pil[2 * (i - dmrs_offset) + 1] = (int16_t)(w * si[idx] * 23170);
  }
  free(g);
}

static void multadd16(const int16_t *filt, int16_t ar, int16_t ai, int16_t *y)
{
  for (int t = 0; t < 16; t++) {
    const int16_t mr = mulhrs16(ar, filt[t]), mi = mulhrs16(ai, filt[t]);
    y[2 * t] = sat16((int32_t)y[2 * t] + sat16(2 * (int32_t)mr));
    y[2 * t + 1] = sat16((int32_t)y[2 * t + 1] + sat16(2 * (int32_t)mi));
  }
}

/* rxdataF [nb_rx][14 N] c16; ul_ch_est [nb_rx][14 N] c16 (symbol p->symbol is rewritten, N entries + up to 8 beyond the allocation stay 0);
 * out: max_ch, nvar, est_delay, delay_max_pos, delay_max_val */
static int chest_type2_freq(const orc_chest_t *p, const int16_t *rxdataF, int16_t *ul_ch_est, int32_t *out);
static int chest_prb_average(const orc_chest_t *p, const int16_t *rxdataF, int16_t *ul_ch_est, int32_t *out);

int orc_pusch_channel_estimation(const orc_chest_t *p, const int16_t *rxdataF, int16_t *ul_ch_est, int32_t *out)
{
  if (p->chest_freq) return chest_prb_average(p, rxdataF, ul_ch_est, out);
  if (p->dmrs_type) return chest_type2_freq(p, rxdataF, ul_ch_est, out);
  static const int delta1[8] = {0, 0, 1, 1, 0, 0, 1, 1};
  const int N = p->fft_size, nb = p->rb_size, np = 6 * nb, delta = delta1[p->port];
  const int k0 = ((p->rb_start + p->bwp_start) * 12 + p->first_carrier_offset) % N;
  int16_t *pil = malloc(4 * (size_t)np), *ls = malloc(4 * (size_t)N), *tim = malloc(4 * (size_t)N), *acc = malloc(4 * (size_t)(N + 16));
  orc_pusch_dmrs_pilots(p, pil);
  int max_ch = 0, max_pos = 0, max_val = 0;
  uint64_t noise = 0;
  int nest = 0;
  for (int a = 0; a < p->nb_rx; a++) {
    const int16_t *rx = rxdataF + 2 * ((size_t)a * 14 + p->symbol) * N;
    int16_t *ul = ul_ch_est + 2 * ((size_t)a * 14 + p->symbol) * N;
    memset(ls, 0, 4 * (size_t)N);
    memset(acc, 0, 4 * (size_t)(N + 16));
    for (int n = 0; n < 3 * nb; n++) {                                       /* LS estimate: average of two pilots, held over 4 REs */
      int32_t cr = 0, ci = 0;
      for (int kl = 0; kl < 2; kl++) {
        const int re = (k0 + (n << 2) + (kl << 1) + delta) % N;
        const int32_t pr = pil[2 * (2 * n + kl)], pi = pil[2 * (2 * n + kl) + 1], yr = rx[2 * re], yi = rx[2 * re + 1];
        cr += (pr * yr - pi * yi) >> 16;
        ci += (pr * yi + pi * yr) >> 16;
      }
      const int acr = cr < 0 ? -cr : cr, aci = ci < 0 ? -ci : ci;
      if (acr > max_ch) max_ch = acr;
      if (aci > max_ch) max_ch = aci;
      for (int k = 4 * n; k < 4 * n + 4; k++) { ls[2 * k] = (int16_t)cr; ls[2 * k + 1] = (int16_t)ci; }
    }
    orc_dft(N, 1, ls, tim, 1);                                                /* nr_est_delay: peak of the impulse response */
    for (int i = 0; i < N; i++) {
      const int temp = (int)(((uint32_t)((int32_t)tim[2 * i] * tim[2 * i] + (int32_t)tim[2 * i + 1] * tim[2 * i + 1])) >> 1);
      if (temp > max_val) { max_pos = i; max_val = temp; }
    }
    if (max_pos > N / 2) max_pos -= N;
    const int est_delay = max_pos;
    int d_idx = 20 + est_delay; d_idx = d_idx < 0 ? 0 : d_idx > 40 ? 40 : d_idx;
    int i_idx = 20 - est_delay; i_idx = i_idx < 0 ? 0 : i_idx > 40 ? 40 : i_idx;
    const int dly = d_idx - 20, idly = i_idx - 20;
    int base = 0;
    for (int pc = 0; pc < np; pc++) {                                         /* delay compensation + 16-tap interpolation, overlap-added */
      const int k = pc << 1;
      const double ang = 2.0 * M_PI * k * dly / N;
      const int16_t tr = (int16_t)round(256 * cos(ang)), ti = (int16_t)round(256 * sin(ang));
      const int32_t lr = ls[2 * k], li = ls[2 * k + 1];
      const int16_t cr = (int16_t)((lr * tr - li * ti) >> 8), ci = (int16_t)((lr * ti + li * tr) >> 8);
      if (pc == 0) multadd16(F_P0, cr, ci, acc + 2 * base);
      else if (pc == 1 || pc == 2) multadd16(F_P1P2, cr, ci, acc + 2 * base);
      else if (pc == np - 1) multadd16(F_LAST, cr, ci, acc + 2 * base);
      else { multadd16(F_MID, cr, ci, acc + 2 * base); if (pc % 2 == 0) base += 4; }
    }
    for (int k = 0; k < 12 * nb; k++) {                                       /* revert the delay, accumulate the noise estimate */
      const double ang = 2.0 * M_PI * k * idly / N;
      const int16_t tr = (int16_t)round(256 * cos(ang)), ti = (int16_t)round(256 * sin(ang));
      const int32_t ar = acc[2 * k], ai = acc[2 * k + 1];
      const int16_t cr = (int16_t)((ar * tr - ai * ti) >> 8), ci = (int16_t)((ar * ti + ai * tr) >> 8);
      acc[2 * k] = cr; acc[2 * k + 1] = ci;
      const int16_t dr = (int16_t)(ls[2 * k] - cr), di = (int16_t)(ls[2 * k + 1] - ci);
      noise += (uint32_t)((int32_t)dr * dr + (int32_t)di * di);
    }
    nest += 12 * nb;
    memset(ul, 0, 4 * (size_t)N);
    memcpy(ul, acc, 4 * (size_t)(12 * nb + 8 <= N ? 12 * nb + 8 : N));
  }
  out[0] = max_ch; out[1] = nest > 0 ? (int32_t)(uint32_t)(noise / (uint64_t)nest) : 0; out[2] = max_pos; out[3] = max_pos; out[4] = max_val;
  free(pil); free(ls); free(tim); free(acc);
  return 0;
}

/* ---- DMRS type 2, frequency-domain "interpolation" (nr_ul_channel_estimation.c:258-283): one least-squares value per CDM pair
 * (two adjacent pilots, >> 15 each, averaged), held over the pair's 6 sub-carriers, delay estimated from the IDFT peak and compensated with
 * ul_delay_table[n % 6] -- only the first six entries of the table are ever used.  As the reference does:
 *  - nushift = (p >> 1) & 1 moves the symbol POINTER (ports 2, 3 read one sub-carrier up, not two; the index wraps before the shift is added);
 *  - ul_ls_est is cleared once per call, and the first four REs of every group are written with a saturating ADD
 *    (multadd_real_four_symbols_vector_complex_scalar with filt8_rep4 = ch & ~3), so antenna a holds the running sum over antennas 0..a
 *    there, and the plain value in the last two REs;
 *  - the noise estimate compares the first pilot's product with the pair's average. */
static int chest_type2_freq(const orc_chest_t *p, const int16_t *rxdataF, int16_t *ul_ch_est, int32_t *out)
{
  const int N = p->fft_size, nb = p->rb_size, nushift = (p->port >> 1) & 1;
  const int k0 = ((p->rb_start + p->bwp_start) * 12 + p->first_carrier_offset) % N;
  int16_t *pil = malloc(4 * (size_t)(4 * nb)), *ls = calloc((size_t)N, 4), *tim = malloc(4 * (size_t)N);
  orc_pusch_dmrs_pilots(p, pil);
  int max_ch = 0, max_pos = 0, max_val = 0, nest = 0;
  uint64_t noise = 0;
  for (int a = 0; a < p->nb_rx; a++) {
    const int16_t *rx = rxdataF + 2 * ((size_t)a * 14 + p->symbol) * N;
    int16_t *ul = ul_ch_est + 2 * ((size_t)a * 14 + p->symbol) * N;
    for (int n = 0, m = 0; n < 12 * nb; n += 6, m += 2) {
      int16_t c[2][2];
      for (int kl = 0; kl < 2; kl++) {
        const int re = (k0 + n + kl) % N;
        const int32_t pr = pil[2 * (m + kl)], pi = pil[2 * (m + kl) + 1], yr = RX_AT(rx, re + nushift, 0), yi = RX_AT(rx, re + nushift, 1);
        c[kl][0] = (int16_t)((pr * yr - pi * yi) >> 15);
        c[kl][1] = (int16_t)((pr * yi + pi * yr) >> 15);
      }
      const int16_t cr = (int16_t)((c[0][0] + c[1][0]) >> 1), ci = (int16_t)((c[0][1] + c[1][1]) >> 1);
      const int acr = cr < 0 ? -cr : cr, aci = ci < 0 ? -ci : ci;
      if (acr > max_ch) max_ch = acr;
      if (aci > max_ch) max_ch = aci;
      const int16_t fr = wrap16((cr >> 2) << 2), fi = wrap16((ci >> 2) << 2);   /* mulhi_s1_int16(ch, 16384) */
      for (int k = n; k < n + 4; k++) { ls[2 * k] = sat16((int32_t)ls[2 * k] + fr); ls[2 * k + 1] = sat16((int32_t)ls[2 * k + 1] + fi); }
      for (int k = n + 4; k < n + 6; k++) { ls[2 * k] = cr; ls[2 * k + 1] = ci; }
      const int16_t dr = (int16_t)(c[0][0] - cr), di = (int16_t)(c[0][1] - ci);
      noise += (uint32_t)((int32_t)dr * dr + (int32_t)di * di);
      nest++;
    }
    orc_dft(N, 1, ls, tim, 1);
    for (int i = 0; i < N; i++) {
      const int temp = (int)(((uint32_t)((int32_t)tim[2 * i] * tim[2 * i] + (int32_t)tim[2 * i + 1] * tim[2 * i + 1])) >> 1);
      if (temp > max_val) { max_pos = i; max_val = temp; }
    }
    if (max_pos > N / 2) max_pos -= N;
    int i_idx = 20 - max_pos; i_idx = i_idx < 0 ? 0 : i_idx > 40 ? 40 : i_idx;   /* get_delay_idx(-est_delay, MAX_DELAY_COMP) */
    memset(ul, 0, 4 * (size_t)N);
    for (int n = 0; n < 12 * nb; n++) {
      const double ang = 2.0 * M_PI * (n % 6) * (i_idx - 20) / N;
      const int16_t tr = (int16_t)round(256 * cos(ang)), ti = (int16_t)round(256 * sin(ang));
      const int32_t lr = ls[2 * n], li = ls[2 * n + 1];
      ul[2 * n] = (int16_t)((lr * tr - li * ti) >> 8); ul[2 * n + 1] = (int16_t)((lr * ti + li * tr) >> 8);
    }
  }
  out[0] = max_ch; out[1] = nest > 0 ? (int32_t)(uint32_t)(noise / (uint64_t)nest) : 0; out[2] = max_pos; out[3] = max_pos; out[4] = max_val;
  free(pil); free(ls); free(tim);
  return 0;
}

/* ---- chest_freq == 1: one value per PRB, no delay estimation, no noise estimate (nr_ul_channel_estimation.c:285-343 type 1, :343-460
 * type 2).  NO_INTERP is defined to 1 at the top of that file, so each PRB's 12 sub-carriers simply take the PRB's average.
 * Type 1: average of the 6 pilot products (>> 15 each, 32-bit sum, C division by 6).  Type 2: (sum of 4 products) / 4, where, as the
 * reference does, the first PRB uses its third pilot twice (the pointer is not advanced), so PRB j >= 1 correlates with pilots
 * 4 j - 1 ... 4 j + 2; every access but the very first omits `soffset`, i.e. reads the slot-ring position 0 -- the restatement (and the
 * product) therefore requires slot % 4 == 0 for type 2.  max_ch only sees the PRBs between the first and the last.  rb_size >= 2 (with one
 * PRB the reference runs the "last PRB" code on a second, non-existent PRB). */
static int chest_prb_average(const orc_chest_t *p, const int16_t *rxdataF, int16_t *ul_ch_est, int32_t *out)
{
  const int N = p->fft_size, nb = p->rb_size, nushift = (p->port >> 1) & 1, t2 = p->dmrs_type != 0;
  const int k0 = ((p->rb_start + p->bwp_start) * 12 + p->first_carrier_offset) % N;
  if (nb < 2 || (t2 && (p->slot & 3))) return -1;
  int16_t *pil = malloc(4 * (size_t)(6 * nb));
  orc_pusch_dmrs_pilots(p, pil);
  int max_ch = 0;
  for (int a = 0; a < p->nb_rx; a++) {
    const int16_t *rx = rxdataF + 2 * ((size_t)a * 14 + p->symbol) * N;
    int16_t *ul = ul_ch_est + 2 * ((size_t)a * 14 + p->symbol) * N;
    memset(ul, 0, 4 * (size_t)N);
    for (int j = 0; j < nb; j++) {
      int32_t sr = 0, si = 0;
      const int cnt = t2 ? 4 : 6;
      for (int i = 0; i < cnt; i++) {
        int pi_idx, re;
        if (!t2) { pi_idx = 6 * j + i; re = (k0 + 12 * j + 2 * i) % N; }
        else {
          static const int off[4] = {0, 1, 6, 7};
          pi_idx = j == 0 ? (i < 3 ? i : 2) : 4 * j - 1 + i;
          re = (k0 + 12 * j + off[i]) % N;
        }
        const int32_t pr = pil[2 * pi_idx], pim = pil[2 * pi_idx + 1], yr = RX_AT(rx, re + nushift, 0), yi = RX_AT(rx, re + nushift, 1);
        sr += (pr * yr - pim * yi) >> 15;
        si += (pr * yi + pim * yr) >> 15;
      }
      const int16_t cr = (int16_t)(sr / cnt), ci = (int16_t)(si / cnt);
      if (j > 0 && j < nb - 1) {
        const int acr = cr < 0 ? -cr : cr, aci = ci < 0 ? -ci : ci;
        if (acr > max_ch) max_ch = acr;
        if (aci > max_ch) max_ch = aci;
      }
      for (int k = 12 * j; k < 12 * j + 12; k++) { ul[2 * k] = cr; ul[2 * k + 1] = ci; }
    }
  }
  out[0] = max_ch; out[1] = 0; out[2] = 0; out[3] = 0; out[4] = 0;
  free(pil);
  return 0;
}

/* ---- UE side: nr_pdsch_channel_estimation with NFAPI_NR_DMRS_TYPE1_linear_interp (NR_UE_ESTIMATION/nr_dl_channel_estimation.c:1305-1385,
 * 1614-1735), DMRS type 1, chest_freq == 0.  Same pilots (nr_pdsch_dmrs_rx / nr_gold_pdsch use the PUSCH formulas), same delay estimation,
 * filters and delay reversal as the gNB estimator; the least-squares step differs: 16-bit accumulation with >> 15 per product and a final >> 1,
 * the port's comb offset is added to the symbol pointer (not wrapped with the sub-carrier index), and there is no max_ch / noise output.
 * p->slot, symbol, port, scid, dmrs_scrambling_id as for the gNB; rb_start + bwp_start = rb_offset of the PDSCH.  dl_ch_est [nb_rx][14 N]. */
/* UE, chest_freq == 1 (NFAPI_NR_DMRS_TYPE1_average_prb / TYPE2_average_prb, nr_dl_channel_estimation.c:1378-1612; NO_INTERP is 1 there too): every PRB's 12
 * sub-carriers take the average of the PRB's pilot products.  Type 1 is the gNB's arithmetic.  Type 2 walks the sub-carriers its own way: the first PRB reads four
 * CONSECUTIVE sub-carriers k0 .. k0+3, every later PRB four sub-carriers 5 apart continuing from there (k0 + 4 + 20 (j - 1) + 5 i) -- reproduced as written.
 * The pointer shift is (p >> 1) & 1 for type 1 and delta2[p] (0, 2, 4) for type 2. */
static int ue_prb_average(const orc_chest_t *p, const int16_t *rxdataF, int16_t *dl_ch_est)
{
  static const int delta2[6] = {0, 0, 2, 2, 4, 4};
  const int N = p->fft_size, nb = p->rb_size, t2 = p->dmrs_type != 0, nushift = t2 ? delta2[p->port] : (p->port >> 1) & 1, cnt = t2 ? 4 : 6;
  const int k0 = ((p->rb_start + p->bwp_start) * 12 + p->first_carrier_offset) % N;
  if (nb < 2) return -1;
  int16_t *pil = malloc(4 * (size_t)(6 * nb));
  orc_pusch_dmrs_pilots(p, pil);
  for (int a = 0; a < p->nb_rx; a++) {
    const int16_t *rx = rxdataF + 2 * ((size_t)a * 14 + p->symbol) * N;
    int16_t *dl = dl_ch_est + 2 * ((size_t)a * 14 + p->symbol) * N;
    memset(dl, 0, 4 * (size_t)N);
    for (int j = 0; j < nb; j++) {
      int32_t sr = 0, si = 0;
      for (int i = 0; i < cnt; i++) {
        const int re = !t2 ? (k0 + 12 * j + 2 * i) % N : j == 0 ? (k0 + i) % N : (k0 + 4 + 20 * (j - 1) + 5 * i) % N;
        const int32_t pr = pil[2 * (cnt * j + i)], pim = pil[2 * (cnt * j + i) + 1], yr = RX_AT(rx, re + nushift, 0), yi = RX_AT(rx, re + nushift, 1);
        sr += (pr * yr - pim * yi) >> 15;
        si += (pr * yi + pim * yr) >> 15;
      }
      const int16_t cr = (int16_t)(sr / cnt), ci = (int16_t)(si / cnt);
      for (int k = 12 * j; k < 12 * j + 12; k++) { dl[2 * k] = cr; dl[2 * k + 1] = ci; }
    }
  }
  free(pil);
  return 0;
}

int orc_pdsch_channel_estimation(const orc_chest_t *p, const int16_t *rxdataF, int16_t *dl_ch_est)
{
  static const int delta2[6] = {0, 0, 2, 2, 4, 4};
  if (p->chest_freq) return ue_prb_average(p, rxdataF, dl_ch_est);
  /* DMRS type 2 (NFAPI_NR_DMRS_TYPE2_linear_interp, :1463-1528): one least-squares value per CDM pair held over the pair's 6 sub-carriers, then the TYPE 1
   * machinery on top of it -- 6 "pilots" per PRB, three of them sharing a value (k = (pilot / 3) * 6), the same four 16-tap filters, delay compensation and reversal */
  const int t2 = p->dmrs_type != 0;
  const int N = p->fft_size, nb = p->rb_size, np = 6 * nb, nushift = t2 ? delta2[p->port] : (p->port >> 1) & 1;
  const int k0 = ((p->rb_start + p->bwp_start) * 12 + p->first_carrier_offset) % N;
  int16_t *pil = malloc(4 * (size_t)np), *ls = malloc(4 * (size_t)N), *tim = malloc(4 * (size_t)N), *acc = malloc(4 * (size_t)(N + 16));
  orc_pusch_dmrs_pilots(p, pil);
  int max_pos = 0, max_val = 0;
  for (int a = 0; a < p->nb_rx; a++) {
    const int16_t *rx = rxdataF + 2 * ((size_t)a * 14 + p->symbol) * N;
    int16_t *dl = dl_ch_est + 2 * ((size_t)a * 14 + p->symbol) * N;
    memset(ls, 0, 4 * (size_t)N);
    memset(acc, 0, 4 * (size_t)(N + 16));
    int re = k0;
    if (!t2) {
      for (int pc = 0; pc < np; pc += 2) {
        const int32_t p0r = pil[2 * pc], p0i = pil[2 * pc + 1], p1r = pil[2 * pc + 2], p1i = pil[2 * pc + 3];
        const int32_t y0r = RX_AT(rx, re + nushift, 0), y0i = RX_AT(rx, re + nushift, 1);
        re = (re + 2) % N;
        const int32_t y1r = RX_AT(rx, re + nushift, 0), y1i = RX_AT(rx, re + nushift, 1);
        re = (re + 2) % N;
        int16_t cr = (int16_t)((p0r * y0r - p0i * y0i) >> 15), ci = (int16_t)((p0r * y0i + p0i * y0r) >> 15);          /* c16mulShift */
        cr = (int16_t)(((p1r * y1r - p1i * y1i) >> 15) + cr); ci = (int16_t)(((p1r * y1i + p1i * y1r) >> 15) + ci);     /* c16maddShift */
        cr = (int16_t)(cr >> 1); ci = (int16_t)(ci >> 1);                                                               /* c16Shift */
        for (int k = 2 * pc; k < 2 * pc + 4; k++) { ls[2 * k] = cr; ls[2 * k + 1] = ci; }
      }
    } else {
      for (int pc = 0; pc < 4 * nb; pc += 2) {
        const int32_t p0r = pil[2 * pc], p0i = pil[2 * pc + 1], p1r = pil[2 * pc + 2], p1i = pil[2 * pc + 3];
        const int32_t y0r = RX_AT(rx, re + nushift, 0), y0i = RX_AT(rx, re + nushift, 1);
        re = (re + 1) % N;
        const int32_t y1r = RX_AT(rx, re + nushift, 0), y1i = RX_AT(rx, re + nushift, 1);
        re = (re + 5) % N;
        const int16_t lr = (int16_t)((p0r * y0r - p0i * y0i) >> 15), li = (int16_t)((p0r * y0i + p0i * y0r) >> 15);
        const int16_t rr = (int16_t)((p1r * y1r - p1i * y1i) >> 15), ri = (int16_t)((p1r * y1i + p1i * y1r) >> 15);
        const int16_t cr = (int16_t)((lr + rr) >> 1), ci = (int16_t)((li + ri) >> 1);                                   /* c16addShift */
        for (int k = 3 * pc; k < 3 * pc + 6; k++) { ls[2 * k] = cr; ls[2 * k + 1] = ci; }
      }
    }
    orc_dft(N, 1, ls, tim, 1);
    for (int i = 0; i < N; i++) {
      const int temp = (int)(((uint32_t)((int32_t)tim[2 * i] * tim[2 * i] + (int32_t)tim[2 * i + 1] * tim[2 * i + 1])) >> 1);
      if (temp > max_val) { max_pos = i; max_val = temp; }
    }
    if (max_pos > N / 2) max_pos -= N;
    int d_idx = 20 + max_pos; d_idx = d_idx < 0 ? 0 : d_idx > 40 ? 40 : d_idx;
    int i_idx = 20 - max_pos; i_idx = i_idx < 0 ? 0 : i_idx > 40 ? 40 : i_idx;
    const int dly = d_idx - 20, idly = i_idx - 20;
    int base = 0;
    for (int pc = 0; pc < np; pc++) {
      const int k = t2 ? (pc / 3) * 6 : pc << 1;
      const double ang = 2.0 * M_PI * k * dly / N;
      const int16_t tr = (int16_t)round(256 * cos(ang)), ti = (int16_t)round(256 * sin(ang));
      const int32_t lr = ls[2 * k], li = ls[2 * k + 1];
      const int16_t cr = (int16_t)((lr * tr - li * ti) >> 8), ci = (int16_t)((lr * ti + li * tr) >> 8);
      if (pc == 0) multadd16(F_P0, cr, ci, acc + 2 * base);
      else if (pc == 1 || pc == 2) multadd16(F_P1P2, cr, ci, acc + 2 * base);
      else if (pc == np - 1) multadd16(F_LAST, cr, ci, acc + 2 * base);
      else { multadd16(F_MID, cr, ci, acc + 2 * base); if (pc % 2 == 0) base += 4; }
    }
    for (int k = 0; k < 12 * nb; k++) {
      const double ang = 2.0 * M_PI * k * idly / N;
      const int16_t tr = (int16_t)round(256 * cos(ang)), ti = (int16_t)round(256 * sin(ang));
      const int32_t ar = acc[2 * k], ai = acc[2 * k + 1];
      acc[2 * k] = (int16_t)((ar * tr - ai * ti) >> 8); acc[2 * k + 1] = (int16_t)((ar * ti + ai * tr) >> 8);
    }
    memset(dl, 0, 4 * (size_t)N);
    memcpy(dl, acc, 4 * (size_t)(12 * nb + 8 <= N ? 12 * nb + 8 : N));
  }
  free(pil); free(ls); free(tim); free(acc);
  return 0;
}


/* ---- nr_chest_time_domain_avg (openair1/PHY/NR_REFSIG/dmrs_nr.c:343-417; gNB: nr_rx_pusch_tp with chest_time == 1, nr_ulsch_demodulation.c:1527-1538;
 * UE: phy_procedures_nr_ue.c:560): the estimates of the slot's DMRS symbols are summed (saturating) into the FIRST DMRS symbol over the first
 * 12 * num_rbs entries of the symbol (from the symbol's start, whatever the allocation's first sub-carrier is) and divided by their number:
 * 2 -> >> 1, 4 -> >> 2 (arithmetic shifts), 3 -> C division (towards zero), 1 -> unchanged.  Only planes 0 .. nb_rx-1 are touched (with several
 * layers: layer 0 only, as in the reference).  est [nb_rx][14][N] c16 in place.  Returns -1 for 0 or more than 4 DMRS symbols (AssertFatal there). */
int orc_chest_time_domain_avg(int N, int nb_rx, int num_symbols, int start_symbol, int dmrs_bitmap, int num_rbs, int16_t *est)
{
  const int total = start_symbol + num_symbols;
  int ndmrs = 0, first = -1;
  for (int s = 0; s < total; s++) ndmrs += (dmrs_bitmap >> s) & 1;
  for (int s = start_symbol; s < total; s++) if ((dmrs_bitmap >> s) & 1) { first = s; break; }
  if (first < 0 || ndmrs < 1 || ndmrs > 4) return -1;
  for (int a = 0; a < nb_rx; a++) {
    int16_t *d = est + 2 * ((size_t)a * 14 + first) * N;
    for (int s = first + 1; s < total; s++) {
      if (!((dmrs_bitmap >> s) & 1)) continue;
      const int16_t *x = est + 2 * ((size_t)a * 14 + s) * N;
      for (int k = 0; k < 24 * num_rbs; k++) d[k] = sat16((int32_t)d[k] + x[k]);
    }
    for (int k = 0; k < 24 * num_rbs; k++) {
      if (ndmrs == 2) d[k] = (int16_t)(d[k] >> 1);
      else if (ndmrs == 4) d[k] = (int16_t)(d[k] >> 2);
      else if (ndmrs == 3) d[k] = (int16_t)(d[k] / 3);
    }
  }
  return 0;
}
