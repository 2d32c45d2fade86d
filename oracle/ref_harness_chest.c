/* TEST INFRASTRUCTURE ONLY -- never linked into the product.
 * Flat C entry point around the UNMODIFIED reference PUSCH channel estimator (nr_pusch_channel_estimation,
 * openair1/PHY/NR_ESTIMATION/nr_ul_channel_estimation.c:67-495, with nr_dmrs_rx.c, nr_gold.c, common/utils/nr/nr_common.c compiled from
 * /root/reference by build_ref.sh).  The harness allocates the parts of PHY_VARS_gNB the function touches and owns the dft/idft
 * function-pointer globals, bound to the compiled reference libref_dfts.so. */
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "PHY/defs_gNB.h"
#include "PHY/NR_ESTIMATION/nr_ul_estimation.h"

void init_delay_table(uint16_t ofdm_symbol_size, int max_delay_comp, int max_ofdm_symbol_size, c16_t delay_table[][max_ofdm_symbol_size]);

dftfunc_t dft;
idftfunc_t idft;

int refh_chest_init(const char *dfts_so)
{
  void *h = dlopen(dfts_so, RTLD_NOW | RTLD_LOCAL);
  if (!h) { fprintf(stderr, "refh_chest_init: %s\n", dlerror()); return -1; }
  int (*autoinit)(void) = (int (*)(void))dlsym(h, "dfts_autoinit");
  dft = (dftfunc_t)dlsym(h, "dft");
  idft = (idftfunc_t)dlsym(h, "idft");
  if (!autoinit || !dft || !idft) return -2;
  autoinit();
  return 0;
}

/* transform precoding (DFT-s-OFDM): the estimator then correlates with the low-PAPR type-1 sequence of group u, base sequence v
 * (nr_ul_channel_estimation.c:122-133) instead of the Gold-sequence DMRS.  The sequences come from the reference's own generator
 * (ul_ref_seq_nr.c, compiled in). */
#include "PHY/NR_REFSIG/ul_ref_seq_nr.h"
static int g_tp_on, g_tp_u, g_tp_v;
void refh_chest_set_transform_precoding(int on, int u, int v)
{
  generate_lowpapr_typ1_refsig_sequences(SHRT_MAX);
  g_tp_on = on; g_tp_u = u; g_tp_v = v;
}
/* the 2 * n_re int16 of gNB_dmrs_lowpaprtype1_sequence[u][v][index(n_re)]; returns the index or -1 (n_re not of the form 6 * 2^a 3^b 5^c) */
int refh_lowpapr_seq(int u, int v, int n_re, int16_t *out)
{
  generate_lowpapr_typ1_refsig_sequences(SHRT_MAX);
  const int idx = get_index_for_dmrs_lowpapr_seq((int16_t)n_re);
  if (idx < 0 || !gNB_dmrs_lowpaprtype1_sequence[u][v][idx]) return -1;
  memcpy(out, gNB_dmrs_lowpaprtype1_sequence[u][v][idx], 4 * (size_t)n_re);
  return idx;
}

enum { C_N, C_NB_RX, C_N_RB_UL, C_SLOT, C_SYMBOL, C_PORT, C_RB_START, C_BWP_START, C_RB_SIZE, C_FCO, C_SCID, C_DMRS_ID, C_DMRS_TYPE, C_CHEST_FREQ, C_COUNT };

/* rxdataF: [nb_rx][14*N] c16 (the slot).  ul_ch_est out: [nb_rx][14*N] c16 (only symbol `symbol` is written).
 * out[0] = max_ch, out[1] = nvar, out[2] = est_delay, out[3] = delay_max_pos, out[4] = delay_max_val.
 * pilots_out (optional): 6 * rb_size c16, the conjugated DMRS the estimator used (regenerated with nr_pusch_dmrs_rx). */
int refh_pusch_chest(const int32_t *p, const int16_t *rxdataF, int16_t *ul_ch_est, int32_t *out, int16_t *pilots_out)
{
  const int N = p[C_N], nrx = p[C_NB_RX], Ns = p[C_SLOT];
  PHY_VARS_gNB *gNB = calloc(1, sizeof(*gNB));
  NR_DL_FRAME_PARMS *fp = &gNB->frame_parms;
  fp->ofdm_symbol_size = N; fp->symbols_per_slot = 14; fp->nb_antennas_rx = nrx; fp->N_RB_UL = p[C_N_RB_UL]; fp->slots_per_frame = 20;
  fp->Ncp = NORMAL; fp->first_carrier_offset = p[C_FCO];
  init_delay_table(N, MAX_DELAY_COMP, NR_MAX_OFDM_SYMBOL_SIZE, fp->delay_table);
  gNB->chest_freq = p[C_CHEST_FREQ];
  gNB->pusch_vars = calloc(1, sizeof(NR_gNB_PUSCH));
  gNB->ulsch = calloc(1, sizeof(NR_gNB_ULSCH_t));
  NR_gNB_PUSCH *pv = &gNB->pusch_vars[0];
  const int nports = p[C_PORT] + 1;
  pv->ul_ch_estimates = calloc(nports * nrx, sizeof(int32_t *));
  pv->ul_ch_estimates_time = calloc(nrx, sizeof(int32_t *));
  gNB->common_vars.rxdataF = calloc(nrx, sizeof(c16_t *));
  const int soffset = (Ns & 3) * 14 * N;
  for (int i = 0; i < nports * nrx; i++) { posix_memalign((void **)&pv->ul_ch_estimates[i], 32, 4 * (size_t)(14 * N + 64)); memset(pv->ul_ch_estimates[i], 0, 4 * (size_t)(14 * N + 64)); }
  for (int a = 0; a < nrx; a++) {
    posix_memalign((void **)&pv->ul_ch_estimates_time[a], 32, 4 * (size_t)N);
    memset(pv->ul_ch_estimates_time[a], 0, 4 * (size_t)N);
    posix_memalign((void **)&gNB->common_vars.rxdataF[a], 32, 4 * (size_t)(4 * 14 * N));
    memset(gNB->common_vars.rxdataF[a], 0, 4 * (size_t)(4 * 14 * N));
    memcpy(&gNB->common_vars.rxdataF[a][soffset], rxdataF + 2 * (size_t)a * 14 * N, 4 * (size_t)14 * N);
  }
  /* the DMRS Gold sequences: [scid][slot][symbol][word] as init_nr_transport / nr_init.c allocate them */
  const int words = ((fp->N_RB_UL * 12) >> 5) + 1;
  gNB->nr_gold_pusch_dmrs = calloc(2, sizeof(uint32_t ***));
  for (int s = 0; s < 2; s++) {
    gNB->nr_gold_pusch_dmrs[s] = calloc(fp->slots_per_frame, sizeof(uint32_t **));
    for (int ns = 0; ns < fp->slots_per_frame; ns++) {
      gNB->nr_gold_pusch_dmrs[s][ns] = calloc(14, sizeof(uint32_t *));
      for (int l = 0; l < 14; l++) gNB->nr_gold_pusch_dmrs[s][ns][l] = calloc(words + 2, 4);
    }
  }
  gNB->pusch_gold_init[0] = gNB->pusch_gold_init[1] = -1;
  nfapi_nr_pusch_pdu_t pdu;
  memset(&pdu, 0, sizeof(pdu));
  pdu.rb_size = p[C_RB_SIZE]; pdu.rb_start = p[C_RB_START]; pdu.bwp_start = p[C_BWP_START];
  pdu.scid = p[C_SCID]; pdu.ul_dmrs_scrambling_id = p[C_DMRS_ID]; pdu.dmrs_config_type = p[C_DMRS_TYPE];
  pdu.transform_precoding = g_tp_on ? transformPrecoder_enabled : transformPrecoder_disabled;
  pdu.dfts_ofdm.low_papr_group_number = (uint8_t)g_tp_u; pdu.dfts_ofdm.low_papr_sequence_number = (uint8_t)g_tp_v;
  int max_ch = 0;
  uint32_t nvar = 0;
  const unsigned short k0 = ((p[C_RB_START] + p[C_BWP_START]) * 12 + p[C_FCO]) % N;
  nr_pusch_channel_estimation(gNB, (unsigned char)Ns, (unsigned short)p[C_PORT], (unsigned char)p[C_SYMBOL], 0, k0, &pdu, &max_ch, &nvar);
  for (int a = 0; a < nrx; a++) memcpy(ul_ch_est + 2 * (size_t)a * 14 * N, pv->ul_ch_estimates[p[C_PORT] * nrx + a], 4 * (size_t)14 * N);
  out[0] = max_ch; out[1] = (int32_t)nvar; out[2] = gNB->ulsch[0].delay.est_delay; out[3] = gNB->ulsch[0].delay.delay_max_pos; out[4] = gNB->ulsch[0].delay.delay_max_val;
  if (pilots_out)
    nr_pusch_dmrs_rx(gNB, Ns, gNB->nr_gold_pusch_dmrs[pdu.scid][Ns][p[C_SYMBOL]], (int32_t *)pilots_out, 1000 + p[C_PORT], 0, pdu.rb_size,
                     (pdu.bwp_start + pdu.rb_start) * 12, pdu.dmrs_config_type);
  for (int s = 0; s < 2; s++) { for (int ns = 0; ns < fp->slots_per_frame; ns++) { for (int l = 0; l < 14; l++) free(gNB->nr_gold_pusch_dmrs[s][ns][l]); free(gNB->nr_gold_pusch_dmrs[s][ns]); } free(gNB->nr_gold_pusch_dmrs[s]); }
  free(gNB->nr_gold_pusch_dmrs);
  for (int i = 0; i < nports * nrx; i++) free(pv->ul_ch_estimates[i]);
  for (int a = 0; a < nrx; a++) { free(pv->ul_ch_estimates_time[a]); free(gNB->common_vars.rxdataF[a]); }
  free(pv->ul_ch_estimates); free(pv->ul_ch_estimates_time); free(gNB->common_vars.rxdataF); free(gNB->pusch_vars); free(gNB->ulsch); free(gNB);
  return 0;
}
