/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the gNB PDSCH transmitter after the encoder: nr_generate_pdsch
 * (openair1/PHY/NR_TRANSPORT/nr_dlsch.c:56-583) from scrambling to txdataF -- nr_pdsch_codeword_scrambling (:46-54), nr_modulation, nr_layer_mapping
 * (MODULATION/nr_modulation.c:246-270), DMRS generation (nr_init_pdsch_dmrs, NR_REFSIG/nr_gold.c:78-96; port tables NR_TRANSPORT/nr_sch_dmrs.c:35-100;
 * allowed_xlsch_re_in_dmrs_symbol NR_REFSIG/dmrs_nr.c:37-62), resource mapping (:236-478) and identity precoding (:490-530).  Pinned bit-exactly against the
 * compiled reference through oracle/ref_harness_pdschtx.c (tests/test_oracle_vs_reference.py).  Only tests/, smoke() and bench.py's cpu_baseline leg may
 * link this.  One code word, 1..4 layers, PT-RS, one wideband precoding matrix; DMRS ports 0..3 (type 1) / 0..5 (type 2) whose CDM group is below numDmrsCdmGrpsNoData.
 * Reference behaviour restated literally: in symbols without DMRS the data are scaled with mulhrs in groups of four REs per contiguous piece of the allocation
 * and the 1..3 REs left over at the end of a piece get ((x * amp) >> 14) + 1, i.e. twice the amplitude (:421-426, :453-459); in DMRS symbols the scaling
 * truncates ((x * amp) >> 15). */
#include <stdlib.h>
#include <string.h>
#include "nrb200_oracle.h"

static inline int16_t wrap16_(int32_t v) { return (int16_t)(uint16_t)(uint32_t)v; }
static const int8_t dmrs1[8][7] = {{0, 0, 0, 1, 1, 1, 1}, {1, 0, 0, 1, -1, 1, 1}, {2, 1, 1, 1, 1, 1, 1}, {3, 1, 1, 1, -1, 1, 1},
                                   {4, 0, 0, 1, 1, 1, -1}, {5, 0, 0, 1, -1, 1, -1}, {6, 1, 1, 1, 1, 1, -1}, {7, 1, 1, 1, -1, 1, -1}};
static const int8_t dmrs2[12][7] = {{0, 0, 0, 1, 1, 1, 1}, {1, 0, 0, 1, -1, 1, 1}, {2, 1, 2, 1, 1, 1, 1}, {3, 1, 2, 1, -1, 1, 1}, {4, 2, 4, 1, 1, 1, 1}, {5, 2, 4, 1, -1, 1, 1},
                                    {6, 0, 0, 1, 1, 1, -1}, {7, 0, 0, 1, -1, 1, -1}, {8, 1, 2, 1, 1, 1, -1}, {9, 1, 2, 1, -1, 1, -1}, {10, 2, 4, 1, 1, 1, -1}, {11, 2, 4, 1, -1, 1, -1}};

static int allowed_re(int k, int start_sc, int N, int cdm, int type)
{
  const int diff = k > start_sc ? k - start_sc : (N - start_sc) + k;
  for (int i = 0; i < cdm; i++) {
    if (type == 0) { if ((diff % 2) == i) return 0; }
    else { const int d = i << 1; if ((diff % 6) == d || (diff % 6) == d + 1) return 0; }
  }
  return 1;
}

int orc_pdsch_tx_slot(const orc_pdsch_tx_t *p, const uint8_t *bits, int16_t *txdataF)
{
  const int N = p->fft_size, nl = p->nrOfLayers, Qm = p->Qm, type = p->dmrs_config_type, cdm = p->num_dmrs_cdm_grps_no_data, amp = (int16_t)p->amp;
  const int nb_re_dmrs = cdm * (type == 0 ? 6 : 4);
  int n_dmrs_sym = 0;
  for (int s = 0; s < 14; s++) n_dmrs_sym += (p->dl_dmrs_symb_pos >> s) & 1;
  /* PT-RS (:98-111, :287-352): on a PT-RS symbol (set_ptrs_symb_idx) every layer carries the QPSK symbols of the first 2 n_ptrs bits of the symbol's DMRS Gold
   * sequence on the PT-RS sub-carriers (is_ptrs_subcarrier relative to start_sc), the data skip them, and the whole symbol takes the per-RE branch whose scaling
   * truncates.  harq->unav_res = PT-RS REs of the slot: the encoder produces that many modulation symbols less per layer. */
  uint32_t ptrs_pos = 0;
  int n_ptrs = 0, n_ptrs_sym = 0, k_rb_ref = 0;
  if (p->ptrs_on) {
    ptrs_pos = orc_ptrs_symbols(p->start_symbol, p->nr_of_symbols, 1 << p->ptrs_L, (uint32_t)p->dl_dmrs_symb_pos);
    n_ptrs = (p->rb_size + p->ptrs_K - 1) / p->ptrs_K;
    for (int s = 0; s < 14; s++) n_ptrs_sym += (ptrs_pos >> s) & 1;
    k_rb_ref = (p->rb_size % p->ptrs_K == 0) ? (p->rnti & 0xFFFF) % p->ptrs_K : (p->rnti & 0xFFFF) % (p->rb_size % p->ptrs_K);
  }
  const int nb_re = ((12 * p->nr_of_symbols - nb_re_dmrs * n_dmrs_sym) * p->rb_size - n_ptrs * n_ptrs_sym) * nl, G = nb_re * Qm;
  const int n_dmrs = (p->bwp_start + p->rb_start + p->rb_size) * nb_re_dmrs;
  uint32_t *scr = calloc((size_t)(G >> 5) + 8, 4);
  int16_t *mod = malloc(4 * (size_t)nb_re + 64);
  orc_scramble(bits, (uint32_t)G, 0, p->data_scrambling_id, p->rnti, scr);
  orc_modulate((const uint8_t *)scr, (uint32_t)G, Qm, mod);
  int start_sc = p->first_carrier_offset + (p->rb_start + p->bwp_start) * 12;
  if (start_sc >= N) start_sc -= N;
  int16_t *mod_dmrs = malloc(4 * (size_t)n_dmrs + 64);
  uint32_t *gold = malloc(4 * ((size_t)(n_dmrs >> 4) + 16));
  for (int layer = 0; layer < nl; layer++) {
    int port = 0;
    if (p->dmrs_ports) { int found = -1; port = -1; for (int i = 0; i < 12; i++) if ((p->dmrs_ports >> i) & 1) { if (++found == layer) { port = i; break; } } if (port < 0) return -1; }
    const int8_t *row = type == 0 ? dmrs1[port] : dmrs2[port];
    const int delta = row[2], Wf[2] = {row[3], row[4]}, Wt[2] = {row[5], row[6]};
    /* not restated: ports whose own CDM group carries data, and the type-2 configuration in which allowed_xlsch_re_in_dmrs_symbol wrongly admits the first
     * sub-carrier (k == start_sc gives diff = N, and N % 6 == 4 passes both group tests): that layer then maps one RE too many and the reference reads one
     * modulation symbol beyond its buffer */
    if (row[1] >= cdm || (type == 1 && delta != 0 && cdm == 2 && N % 6 == 4)) { free(scr); free(mod); free(mod_dmrs); free(gold); return -2; }
    int l_prime = 0, l_overline = 0;
    while (l_overline < 14 && !((p->dl_dmrs_symb_pos >> l_overline) & 1)) l_overline++;
    int m = 0;
    for (int l = p->start_symbol; l < p->start_symbol + p->nr_of_symbols; l++) {
      int16_t *out = txdataF + 2 * ((size_t)layer * 14 + l) * N;
      const int is_dmrs = (p->dl_dmrs_symb_pos >> l) & 1;
      int k = start_sc;
      if (is_dmrs) {
        int dmrs_idx = (p->rb_start + p->bwp_start) * (type == 0 ? 6 : 4), k_prime = 0, n = 0;
        if (l == l_overline + 1) l_prime = 1;
        else if (l > l_overline + 1) { l_overline = l; l_prime = 0; }
        const uint64_t x2 = ((1ULL << 17) * (14 * p->slot + l + 1) * ((p->dl_dmrs_scrambling_id << 1) + 1) + ((p->dl_dmrs_scrambling_id << 1) + p->scid));
        orc_gold_words((uint32_t)(x2 % (1ULL << 31)), (uint32_t)((2 * n_dmrs + 31) >> 5), gold);
        orc_modulate((const uint8_t *)gold, (uint32_t)(2 * n_dmrs), 2, mod_dmrs);
        for (int i = 0; i < p->rb_size * 12; i++) {
          const int fidx = type ? 6 * n + k_prime + delta : (n << 2) + (k_prime << 1) + delta;
          if (k == (start_sc + fidx) % N) {
            const int w = Wt[l_prime] * Wf[k_prime] * amp;
            out[2 * k] = wrap16_((mod_dmrs[2 * dmrs_idx] * w) >> 15); out[2 * k + 1] = wrap16_((mod_dmrs[2 * dmrs_idx + 1] * w) >> 15);
            dmrs_idx++; k_prime++; k_prime &= 1; n += k_prime ? 0 : 1;
          } else if (allowed_re(k, start_sc, N, cdm, type)) {
            const int16_t *x = mod + 2 * ((size_t)m * nl + layer);
            out[2 * k] = wrap16_((x[0] * amp) >> 15); out[2 * k + 1] = wrap16_((x[1] * amp) >> 15);
            m++;
          } else { out[2 * k] = 0; out[2 * k + 1] = 0; }
          if (++k >= N) k -= N;
        }
      } else if ((ptrs_pos >> l) & 1) {
        int16_t mod_ptrs[2 * 140];
        int ptrs_idx = 0;
        const uint64_t x2 = ((1ULL << 17) * (14 * p->slot + l + 1) * ((p->dl_dmrs_scrambling_id << 1) + 1) + ((p->dl_dmrs_scrambling_id << 1) + p->scid));
        orc_gold_words((uint32_t)(x2 % (1ULL << 31)), 10, gold);
        orc_modulate((const uint8_t *)gold, (uint32_t)(2 * n_ptrs), 2, mod_ptrs);
        for (int i = 0; i < p->rb_size * 12; i++) {
          if ((i - p->ptrs_re_offset - k_rb_ref * 12) % (p->ptrs_K * 12) == 0) {
            out[2 * k] = wrap16_((mod_ptrs[2 * ptrs_idx] * amp) >> 15); out[2 * k + 1] = wrap16_((mod_ptrs[2 * ptrs_idx + 1] * amp) >> 15);
            ptrs_idx++;
          } else {
            const int16_t *x = mod + 2 * ((size_t)m * nl + layer);
            out[2 * k] = wrap16_((x[0] * amp) >> 15); out[2 * k + 1] = wrap16_((x[1] * amp) >> 15);
            m++;
          }
          if (++k >= N) k -= N;
        }
      } else {
        int upper = p->rb_size * 12, rem = 0;
        if (start_sc + upper > N) { rem = upper + start_sc - N; upper = N - start_sc; }
        for (int piece = 0; piece < 2; piece++) {
          const int len = piece ? rem : upper, base = piece ? 0 : start_sc;
          for (int i = 0; i < len; i++) {
            const int16_t *x = mod + 2 * ((size_t)(m + i) * nl + layer);
            for (int c = 0; c < 2; c++)
              out[2 * (base + i) + c] = i < ((len >> 2) << 2) ? wrap16_((x[c] * amp + 0x4000) >> 15) : wrap16_(((x[c] * amp) >> 14) + 1);
          }
          m += len;
        }
      }
    }
  }
  if (p->pm_idx > 0) {
    /* non-identity precoding (:536-590, nr_layer_precoder_simd / nr_layer_precoder_cm, MODULATION/nr_modulation.c:702-815): with one PRG the RBs are taken two at a
     * time (the last one alone when rb_size is odd); a group whose last sub-carrier stays below the end of the symbol (subCarrier + re_cnt < N) goes through
     * the SIMD routine -- per-layer products truncated to 16 bits, SATURATING accumulation over the layers -- any other group through the scalar one, whose
     * c16maddShift accumulation WRAPS.  (The SIMD routine conjugates the weight in 16 bits; a weight with Im = -32768 is not restated.) */
    int16_t *lay = malloc(4 * (size_t)nl * 14 * N);
    memcpy(lay, txdataF, 4 * (size_t)nl * 14 * N);
    for (int ant = 0; ant < p->nb_tx; ant++)
      for (int l = p->start_symbol; l < p->start_symbol + p->nr_of_symbols; l++)
        for (int i = 0; i < p->rb_size * 12; i++) {
          const int g = i / 24, cnt = (p->rb_size * 12 - 24 * g) >= 24 ? 24 : 12, sc_g = (start_sc + 24 * g) % N, wraps = sc_g + cnt >= N;
          const int k = (start_sc + i) % N;
          int32_t yr = 0, yi = 0;
          for (int al = 0; al < nl; al++) {
            const int32_t xr = lay[2 * (((size_t)al * 14 + l) * N + k)], xi = lay[2 * (((size_t)al * 14 + l) * N + k) + 1];
            const int32_t wr = p->pm_weights[al][ant][0], wi = p->pm_weights[al][ant][1];
            const int16_t pr = wrap16_((xr * wr - xi * wi) >> 15), pi = wrap16_((xr * wi + xi * wr) >> 15);
            if (wraps) { yr = wrap16_(yr + pr); yi = wrap16_(yi + pi); }
            else { yr = yr + pr > 32767 ? 32767 : yr + pr < -32768 ? -32768 : yr + pr; yi = yi + pi > 32767 ? 32767 : yi + pi < -32768 ? -32768 : yi + pi; }
          }
          txdataF[2 * (((size_t)ant * 14 + l) * N + k)] = (int16_t)yr; txdataF[2 * (((size_t)ant * 14 + l) * N + k) + 1] = (int16_t)yi;
        }
    free(lay); free(scr); free(mod); free(mod_dmrs); free(gold);
    return G;
  }
  /* antennas beyond the layers: zero over the allocation (identity precoding, :497-523) */
  for (int ant = nl; ant < p->nb_tx; ant++)
    for (int l = p->start_symbol; l < p->start_symbol + p->nr_of_symbols; l++)
      for (int i = 0; i < p->rb_size * 12; i++) { const int k = (start_sc + i) % N; txdataF[2 * (((size_t)ant * 14 + l) * N + k)] = 0; txdataF[2 * (((size_t)ant * 14 + l) * N + k) + 1] = 0; }
  free(scr); free(mod); free(mod_dmrs); free(gold);
  return G;
}
