/* TEST INFRASTRUCTURE ONLY -- never linked into the product.
 * Flat C entry points around the UNMODIFIED reference PUSCH inner receiver.  The per-symbol functions of
 * openair1/PHY/NR_TRANSPORT/nr_ulsch_demodulation.c are `static`, so this harness textually includes that file from /root/reference at
 * build time (build_ref.sh; nothing is copied into the repo) and calls inner_rx / nr_ulsch_extract_rbs / nr_ulsch_scale_channel /
 * nr_ulsch_channel_level with the few fields of PHY_VARS_gNB, NR_gNB_PUSCH, nfapi_nr_pusch_pdu_t and NR_DL_FRAME_PARMS they read. */
#include "PHY/NR_TRANSPORT/nr_ulsch_demodulation.c"

/* transform precoding: inner_rx then runs nr_freq_equalization (Qm > 2) and nr_idft on the compensated symbol (:1326-1336); dft / idft are bound to
 * the compiled reference libref_dfts.so */
#include <dlfcn.h>
void nr_init_fde(void);
static int g_tp_on;
int refh_pusch_set_transform_precoding(int on, const char *dfts_so)
{
  static int bound;
  if (on && !bound) {
    void *h = dlopen(dfts_so, RTLD_NOW | RTLD_LOCAL);
    if (!h) { fprintf(stderr, "refh_pusch_set_transform_precoding: %s\n", dlerror()); return -1; }
    int (*autoinit)(void) = (int (*)(void))dlsym(h, "dfts_autoinit");
    dft = (dftfunc_t)dlsym(h, "dft");
    idft = (idftfunc_t)dlsym(h, "idft");
    if (!autoinit || !dft || !idft) return -2;
    autoinit();
    nr_init_fde();
    bound = 1;
  }
  g_tp_on = on;
  return 0;
}

enum { P_N, P_NB_RX, P_NB_LAYER, P_RB_START, P_BWP_START, P_RB_SIZE, P_FCO, P_QM, P_SYMBOL, P_DMRS_SYMBOL, P_DMRS_POS, P_CDM_NO_DATA, P_DMRS_TYPE,
       P_SHIFT, P_NVAR, P_VALID_RE, P_COUNT };

static void fill(const int32_t *p, NR_DL_FRAME_PARMS *fp, nfapi_nr_pusch_pdu_t *pdu)
{
  memset(fp, 0, sizeof(*fp));
  memset(pdu, 0, sizeof(*pdu));
  fp->ofdm_symbol_size = p[P_N];
  fp->first_carrier_offset = p[P_FCO];
  fp->nb_antennas_rx = p[P_NB_RX];
  fp->symbols_per_slot = 14;
  fp->N_RB_UL = 275;                                   /* only read by an assertion of nr_freq_equalization */
  pdu->rb_start = p[P_RB_START]; pdu->bwp_start = p[P_BWP_START]; pdu->rb_size = p[P_RB_SIZE];
  pdu->qam_mod_order = p[P_QM]; pdu->nrOfLayers = p[P_NB_LAYER];
  pdu->ul_dmrs_symb_pos = p[P_DMRS_POS]; pdu->dmrs_config_type = p[P_DMRS_TYPE]; pdu->num_dmrs_cdm_grps_no_data = p[P_CDM_NO_DATA];
  pdu->transform_precoding = g_tp_on ? transformPrecoder_enabled : transformPrecoder_disabled;
}

/* One OFDM symbol through inner_rx.  rxdataF: [nb_rx][14*N] c16.  ch_est: [nb_layer*nb_rx][14*N] c16 (ul_ch_estimates layout).
 * llr: [nb_layer][valid_re*Qm] int16 out.  comp: [nb_layer][buffer_length] c16 out (rxdataF_comp of the first rx index of each layer). */
int refh_pusch_inner_rx(const int32_t *p, const int16_t *rxdataF, const int16_t *ch_est, int16_t *llr, int16_t *comp)
{
  static PHY_VARS_gNB *gNB;
  if (!gNB) gNB = calloc(1, sizeof(*gNB));
  NR_DL_FRAME_PARMS fp;
  nfapi_nr_pusch_pdu_t pdu;
  fill(p, &fp, &pdu);
  const int N = p[P_N], nrx = p[P_NB_RX], nl = p[P_NB_LAYER], symbol = p[P_SYMBOL];
  const int buffer_length = (p[P_RB_SIZE] * 12 + 15) & ~15;
  NR_gNB_PUSCH pv;
  memset(&pv, 0, sizeof(pv));
  int32_t *est[8], *cmp[8];
  int16_t valid[14] = {0};
  c16_t *rxF[8];
  int16_t *llrp[4];
  for (int i = 0; i < nl * nrx; i++) {
    est[i] = (int32_t *)ch_est + (size_t)i * 14 * N;
    posix_memalign((void **)&cmp[i], 32, sizeof(int32_t) * 14 * buffer_length);
    memset(cmp[i], 0, sizeof(int32_t) * 14 * buffer_length);
  }
  for (int a = 0; a < nrx; a++) rxF[a] = (c16_t *)rxdataF + (size_t)a * 14 * N;
  const size_t llr_n = (size_t)p[P_VALID_RE] * p[P_QM];
  for (int l = 0; l < nl; l++) { posix_memalign((void **)&llrp[l], 64, 2 * llr_n + 1024); memset(llrp[l], 0, 2 * llr_n + 1024); }   /* the LLR kernels use aligned vector stores */
  valid[symbol] = (int16_t)p[P_VALID_RE];
  pv.ul_ch_estimates = est; pv.rxdataF_comp = cmp; pv.ul_valid_re_per_slot = valid;
  pv.llr_layers = llrp;                                                   /* nr_ulsch_shift_llr (QPSK ML path) works on these; llr_offset stays 0 */
  pv.dmrs_symbol = (uint8_t)p[P_DMRS_SYMBOL]; pv.log2_maxh = (int16_t)p[P_SHIFT];
  inner_rx(gNB, 0, 0, &fp, &pv, &pdu, rxF, NULL, llrp, 0, p[P_VALID_RE], symbol, p[P_SHIFT], (uint32_t)p[P_NVAR]);
  for (int l = 0; l < nl; l++) memcpy(comp + (size_t)l * 2 * buffer_length, &cmp[l * nrx][symbol * buffer_length], 4 * (size_t)buffer_length);
  for (int l = 0; l < nl; l++) { memcpy(llr + l * llr_n, llrp[l], 2 * llr_n); free(llrp[l]); }
  for (int i = 0; i < nl * nrx; i++) free(cmp[i]);
  return buffer_length;
}

/* log2_maxh as nr_rx_pusch_tp derives it (nr_ulsch_demodulation.c:1595-1640): extract the measurement symbol, scale, level, max, log2/2.
 * max_ch = the channel estimator's max_ch output (only used for 2 layers). */
int refh_pusch_log2_maxh(const int32_t *p, int max_ch, const int16_t *rxdataF, const int16_t *ch_est, int32_t *avg_out)
{
  NR_DL_FRAME_PARMS fp;
  nfapi_nr_pusch_pdu_t pdu;
  fill(p, &fp, &pdu);
  const int N = p[P_N], nrx = p[P_NB_RX], nl = p[P_NB_LAYER], meas_symbol = p[P_SYMBOL];
  int nb_re_pusch = get_nb_re_pusch(&fp, &pdu, meas_symbol);
  nb_re_pusch = (nb_re_pusch + 15) & ~15;
  int32_t *ext[8], *rxext[8];
  for (int i = 0; i < nl * nrx; i++) { posix_memalign((void **)&ext[i], 32, 4 * 14 * (size_t)nb_re_pusch); memset(ext[i], 0, 4 * 14 * (size_t)nb_re_pusch); }
  for (int a = 0; a < nrx; a++) { posix_memalign((void **)&rxext[a], 32, 4 * 14 * (size_t)nb_re_pusch); memset(rxext[a], 0, 4 * 14 * (size_t)nb_re_pusch); }
  for (int aarx = 0; aarx < nrx; aarx++)
    for (int aatx = 0; aatx < nl; aatx++)
      nr_ulsch_extract_rbs((c16_t *)rxdataF + (size_t)aarx * 14 * N, (c16_t *)ch_est + (size_t)(aatx * nrx + aarx) * 14 * N,
                           (c16_t *)&rxext[aarx][meas_symbol * nb_re_pusch], (c16_t *)&ext[aatx * nrx + aarx][meas_symbol * nb_re_pusch],
                           meas_symbol * N, p[P_DMRS_SYMBOL] * N, aarx, (p[P_DMRS_POS] >> meas_symbol) & 1, &pdu, &fp);
  int avg[8] = {0}, avgs = 0;
  const uint8_t shift_ch_ext = nl > 1 ? log2_approx(max_ch >> 11) : 0;
  nr_ulsch_scale_channel(ext, &fp, meas_symbol, (p[P_DMRS_POS] >> meas_symbol) & 1, nb_re_pusch, nl, p[P_RB_SIZE], shift_ch_ext);
  nr_ulsch_channel_level(ext, &fp, avg, meas_symbol, nb_re_pusch, nl);
  for (int i = 0; i < nl * nrx; i++) { avgs = cmax(avgs, avg[i]); if (avg_out) avg_out[i] = avg[i]; }
  for (int i = 0; i < nl * nrx; i++) free(ext[i]);
  for (int a = 0; a < nrx; a++) free(rxext[a]);
  return log2_approx(avgs) >> 1;
}
