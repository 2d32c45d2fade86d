/* TEST INFRASTRUCTURE ONLY -- never linked into the product.
 * The reference-side CALLER of nr_rx_pusch_tp (openair1/PHY/NR_TRANSPORT/nr_ulsch_demodulation.c:1447): allocates the parts of PHY_VARS_gNB the function and its
 * callers touch -- frame parameters, the 4-slot rxdataF ring, pusch_vars (estimates, LLR buffer, per-symbol bookkeeping), the ULSCH's PDU -- the way
 * init_nr_transport / phy_init_nr_gNB do, and calls the function by name like phy_procedures_gNB_uespec_RX does (SCHED_NR/phy_procedures_nr_gNB.c).  Linked against
 * integration/oai_shim_rx_pusch.c + oai_shim_pusch_chest.c (integration/build_shims.sh -> oracle/_ref/libshimtest_rxpusch.so) the call lands in libldpc_b200.so. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "PHY/defs_gNB.h"
#include "PHY/NR_TRANSPORT/nr_transport_proto.h"

/* the function-pointer globals of PHY/TOOLS/tools_defs.h: nr_common.c (nr_est_delay) references them; with the interposed estimator nobody calls them here */
dftfunc_t dft;
idftfunc_t idft;

void init_delay_table(uint16_t ofdm_symbol_size, int max_delay_comp, int max_ofdm_symbol_size, c16_t delay_table[][max_ofdm_symbol_size]);

/* DFT-s-OFDM: the PDU's transform-precoding fields; the low-PAPR sequence table is built like nr_init.c:249 does */
#include "PHY/NR_REFSIG/ul_ref_seq_nr.h"
static int g_tp_on, g_tp_u, g_tp_v;
void refh_rxpusch_set_transform_precoding(int on, int u, int v)
{
  generate_lowpapr_typ1_refsig_sequences(SHRT_MAX);
  g_tp_on = on; g_tp_u = u; g_tp_v = v;
}

enum { X_N, X_NB_RX, X_N_RB_UL, X_SLOT, X_RB_START, X_BWP_START, X_RB_SIZE, X_FCO, X_QM, X_START_SYMBOL, X_NR_SYMBOLS, X_DMRS_POS, X_DMRS_TYPE, X_CDM, X_NL, X_DMRS_PORTS,
       X_SCID, X_DMRS_ID, X_RNTI, X_DATA_ID, X_CHEST_FREQ, X_CHEST_TIME, X_COUNT };

/* rxdataF: [nb_rx][14 N] c16 (the slot).  Outputs: llr (G int16), est ([nl * nb_rx][14 N] c16), info[0] = log2_maxh, [1] = dmrs_symbol, [2..15] = ul_valid_re_per_slot,
 * [16..29] = llr_offset, [30..37] = ulsch_power, [38] = unav_res.  Returns nr_rx_pusch_tp's return value. */
int refh_rx_pusch(const int32_t *p, const int16_t *rxdataF, int G, int16_t *llr_out, int16_t *est_out, int32_t *info)
{
  const int N = p[X_N], nrx = p[X_NB_RX], nl = p[X_NL], slot = p[X_SLOT];
  PHY_VARS_gNB *gNB = calloc(1, sizeof(*gNB));
  NR_DL_FRAME_PARMS *fp = &gNB->frame_parms;
  fp->ofdm_symbol_size = N; fp->symbols_per_slot = 14; fp->nb_antennas_rx = nrx; fp->N_RB_UL = p[X_N_RB_UL]; fp->slots_per_frame = 20;
  fp->Ncp = NORMAL; fp->first_carrier_offset = p[X_FCO];
  init_delay_table(N, MAX_DELAY_COMP, NR_MAX_OFDM_SYMBOL_SIZE, fp->delay_table);
  gNB->chest_freq = p[X_CHEST_FREQ]; gNB->chest_time = p[X_CHEST_TIME]; gNB->num_pusch_symbols_per_thread = 1;
  gNB->pusch_vars = calloc(1, sizeof(NR_gNB_PUSCH));
  gNB->ulsch = calloc(1, sizeof(NR_gNB_ULSCH_t));
  gNB->ulsch[0].harq_process = calloc(1, sizeof(NR_UL_gNB_HARQ_t));
  nfapi_nr_pusch_pdu_t *u = &gNB->ulsch[0].harq_process->ulsch_pdu;
  u->rb_start = p[X_RB_START]; u->bwp_start = p[X_BWP_START]; u->rb_size = p[X_RB_SIZE]; u->qam_mod_order = p[X_QM];
  u->start_symbol_index = p[X_START_SYMBOL]; u->nr_of_symbols = p[X_NR_SYMBOLS]; u->ul_dmrs_symb_pos = p[X_DMRS_POS];
  u->dmrs_config_type = p[X_DMRS_TYPE]; u->num_dmrs_cdm_grps_no_data = p[X_CDM]; u->nrOfLayers = nl; u->dmrs_ports = p[X_DMRS_PORTS];
  u->scid = p[X_SCID]; u->ul_dmrs_scrambling_id = p[X_DMRS_ID]; u->rnti = p[X_RNTI]; u->data_scrambling_id = p[X_DATA_ID];
  u->transform_precoding = g_tp_on ? transformPrecoder_enabled : transformPrecoder_disabled; u->pdu_bit_map = 0;
  u->dfts_ofdm.low_papr_group_number = (uint8_t)g_tp_u; u->dfts_ofdm.low_papr_sequence_number = (uint8_t)g_tp_v;
  NR_gNB_PUSCH *pv = &gNB->pusch_vars[0];
  pv->ul_ch_estimates = calloc(nl * nrx, sizeof(int32_t *));
  pv->ul_ch_estimates_time = calloc(nrx, sizeof(int32_t *));
  gNB->common_vars.rxdataF = calloc(nrx, sizeof(c16_t *));
  const int soffset = (slot & 3) * 14 * N;
  for (int i = 0; i < nl * nrx; i++) { posix_memalign((void **)&pv->ul_ch_estimates[i], 32, 4 * (size_t)(14 * N + 64)); memset(pv->ul_ch_estimates[i], 0, 4 * (size_t)(14 * N + 64)); }
  for (int a = 0; a < nrx; a++) {
    posix_memalign((void **)&pv->ul_ch_estimates_time[a], 32, 4 * (size_t)N);
    memset(pv->ul_ch_estimates_time[a], 0, 4 * (size_t)N);
    posix_memalign((void **)&gNB->common_vars.rxdataF[a], 32, 4 * (size_t)(4 * 14 * N));
    memset(gNB->common_vars.rxdataF[a], 0, 4 * (size_t)(4 * 14 * N));
    memcpy(&gNB->common_vars.rxdataF[a][soffset], rxdataF + 2 * (size_t)a * 14 * N, 4 * (size_t)14 * N);
  }
  posix_memalign((void **)&pv->llr, 64, 2 * (size_t)G + 4096);
  memset(pv->llr, 0, 2 * (size_t)G + 4096);
  pv->ul_valid_re_per_slot = calloc(14, sizeof(int16_t));
  const int rc = nr_rx_pusch_tp(gNB, 0, 0, (uint8_t)slot, 0);
  memcpy(llr_out, pv->llr, 2 * (size_t)G);
  for (int i = 0; i < nl * nrx; i++) memcpy(est_out + 2 * (size_t)i * 14 * N, pv->ul_ch_estimates[i], 4 * (size_t)14 * N);
  info[0] = pv->log2_maxh; info[1] = pv->dmrs_symbol;
  for (int s = 0; s < 14; s++) { info[2 + s] = pv->ul_valid_re_per_slot[s]; info[16 + s] = pv->llr_offset[s]; }
  for (int a = 0; a < 8; a++) info[30 + a] = pv->ulsch_power[a];
  info[38] = (int32_t)gNB->ulsch[0].unav_res;
  return rc;
}
