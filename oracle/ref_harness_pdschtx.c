/* TEST INFRASTRUCTURE ONLY -- never linked into the product.
 * Flat C entry point around the UNMODIFIED reference PDSCH transmitter nr_generate_pdsch (openair1/PHY/NR_TRANSPORT/nr_dlsch.c:56-583): scrambling,
 * modulation, layer mapping, DMRS generation, resource mapping and (identity) precoding into txdataF run from the reference's own control flow.  The one
 * callee replaced is nr_dlsch_encoding (nr_dlsch_coding.c, pinned separately through libref_coding.so / libref_ldpc_enc.so): the stub below hands the
 * caller's already rate-matched, interleaved bits (one per byte) to nr_generate_pdsch.  The harness fills the fields of PHY_VARS_gNB / NR_gNB_DLSCH_t /
 * nfapi_nr_dl_tti_pdsch_pdu_rel15_t the function reads. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "PHY/defs_gNB.h"
#include "PHY/NR_TRANSPORT/nr_dlsch.h"
#include "PHY/NR_REFSIG/nr_refsig.h"
#include "PHY/NR_REFSIG/nr_mod_table.h"
#include <time.h>
/* wall time of the last call into the reference function(s), excluding the harness's own allocation and copying (cpu_baseline of the DL slot chain) */
static double g_last_s;
static inline double now_s(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }


static const uint8_t *g_bits;
static uint32_t g_nbits;
int nr_dlsch_encoding(PHY_VARS_gNB *gNB, int frame, uint8_t slot, NR_DL_gNB_HARQ_t *harq, NR_DL_FRAME_PARMS *frame_parms, unsigned char *output,
                      time_stats_t *tinput, time_stats_t *tprep, time_stats_t *tparity, time_stats_t *toutput, time_stats_t *dlsch_rate_matching_stats,
                      time_stats_t *dlsch_interleaving_stats, time_stats_t *dlsch_segmentation_stats)
{
  (void)gNB; (void)frame; (void)slot; (void)harq; (void)frame_parms; (void)tinput; (void)tprep; (void)tparity; (void)toutput;
  (void)dlsch_rate_matching_stats; (void)dlsch_interleaving_stats; (void)dlsch_segmentation_stats;
  memcpy(output, g_bits, g_nbits);
  return 0;
}

double refh_pdschtx_last_seconds(void) { return g_last_s; }

/* wideband precoding for the next refh_pdsch_tx_slot calls: pm_idx = 0 restores the identity; weights [layer][antenna]{re, im} (nfapi_nr_pm_pdu_t.weights).
 * One PRG spanning the allocation: precodingAndBeamforming.prgs_list has a single entry in the reference's nFAPI structures (nfapi_nr_interface_scf.h:704). */
static int g_pm_idx;
static int16_t g_pm_w[4][4][2];
void refh_pdschtx_set_precoding(int pm_idx, const int16_t *weights) { g_pm_idx = pm_idx; if (weights) memcpy(g_pm_w, weights, sizeof(g_pm_w)); }

/* PT-RS for the next refh_pdsch_tx_slot calls (pduBitmap bit 0; PTRSTimeDensity = log2 of L, PTRSFreqDensity = K, PTRSReOffset); on = 0 switches it off */
static int g_ptrs[4];
void refh_pdschtx_set_ptrs(int on, int L, int K, int re_offset) { g_ptrs[0] = on; g_ptrs[1] = L; g_ptrs[2] = K; g_ptrs[3] = re_offset; }

enum { X_N, X_N_RB_DL, X_NB_TX, X_SLOT, X_RB_START, X_BWP_START, X_RB_SIZE, X_FCO, X_QM, X_NL, X_START_SYMBOL, X_NR_SYMBOLS, X_DMRS_POS, X_DMRS_TYPE,
       X_CDM_GROUPS, X_DMRS_PORTS, X_SCID, X_DMRS_ID, X_DATA_ID, X_RNTI, X_AMP, X_COUNT };

/* bits: G bytes (0/1), G = nb_re * Qm as nr_generate_pdsch computes it; txdataF out: [nb_tx][14 N] c16 of the slot (zero where nothing is mapped) */
int refh_pdsch_tx_slot(const int32_t *p, const uint8_t *bits, uint32_t nbits, int16_t *txdataF_out)
{
  const int N = p[X_N], ntx = p[X_NB_TX], slot = p[X_SLOT];
  static int tables_done;
  if (!tables_done) { nr_generate_modulation_table(); tables_done = 1; }      /* what init_nr_transport / the softmodem start-up does once */
  PHY_VARS_gNB *gNB = calloc(1, sizeof(*gNB));
  NR_DL_FRAME_PARMS *fp = &gNB->frame_parms;
  fp->ofdm_symbol_size = N; fp->symbols_per_slot = 14; fp->slots_per_frame = 20; fp->nb_antennas_tx = ntx; fp->N_RB_DL = p[X_N_RB_DL];
  fp->first_carrier_offset = p[X_FCO]; fp->samples_per_slot_wCP = 14 * N; fp->Ncp = NORMAL;
  gNB->TX_AMP = (int16_t)p[X_AMP];
  gNB->pdsch_gold_init[0] = gNB->pdsch_gold_init[1] = -1;
  const int words = ((fp->N_RB_DL * 12) >> 5) + 1;
  gNB->nr_gold_pdsch_dmrs = calloc(fp->slots_per_frame, sizeof(uint32_t ***));
  for (int s = 0; s < fp->slots_per_frame; s++) {
    gNB->nr_gold_pdsch_dmrs[s] = calloc(14, sizeof(uint32_t **));
    for (int l = 0; l < 14; l++) {
      gNB->nr_gold_pdsch_dmrs[s][l] = calloc(2, sizeof(uint32_t *));
      for (int q = 0; q < 2; q++) gNB->nr_gold_pdsch_dmrs[s][l][q] = calloc(words + 2, 4);
    }
  }
  gNB->common_vars.txdataF = calloc(ntx, sizeof(c16_t *));
  for (int a = 0; a < ntx; a++) gNB->common_vars.txdataF[a] = calloc((size_t)20 * 14 * N, sizeof(c16_t));
  gNB->common_vars.beam_id = calloc(1, sizeof(uint8_t *));
  gNB->common_vars.beam_id[0] = calloc((size_t)20 * 14, 1);
  NR_gNB_DLSCH_t *dl = calloc(1, sizeof(*dl));
  NR_gNB_DLSCH_t *dlv[1] = {dl};
  NR_DL_gNB_HARQ_t *harq = &dl->harq_process;
  static uint8_t pdu_dummy[16];
  harq->pdu = pdu_dummy;
  harq->f = calloc(nbits + 64, 1);
  nfapi_nr_dl_tti_pdsch_pdu_rel15_t *rel15 = &harq->pdsch_pdu.pdsch_pdu_rel15;
  rel15->BWPStart = p[X_BWP_START]; rel15->BWPSize = p[X_N_RB_DL]; rel15->rbStart = p[X_RB_START]; rel15->rbSize = p[X_RB_SIZE];
  rel15->StartSymbolIndex = p[X_START_SYMBOL]; rel15->NrOfSymbols = p[X_NR_SYMBOLS]; rel15->dlDmrsSymbPos = p[X_DMRS_POS]; rel15->dmrsConfigType = p[X_DMRS_TYPE];
  rel15->numDmrsCdmGrpsNoData = p[X_CDM_GROUPS]; rel15->dmrsPorts = p[X_DMRS_PORTS]; rel15->SCID = p[X_SCID]; rel15->dlDmrsScramblingId = p[X_DMRS_ID];
  rel15->dataScramblingId = p[X_DATA_ID]; rel15->rnti = p[X_RNTI]; rel15->nrOfLayers = p[X_NL]; rel15->NrOfCodewords = 1; rel15->qamModOrder[0] = p[X_QM];
  rel15->pduBitmap = 0; rel15->precodingAndBeamforming.prg_size = 0;
  if (g_ptrs[0]) { rel15->pduBitmap = 1; rel15->PTRSTimeDensity = g_ptrs[1]; rel15->PTRSFreqDensity = g_ptrs[2]; rel15->PTRSReOffset = g_ptrs[3]; }
  nfapi_nr_pm_pdu_t *pm_pdus = NULL;
  if (g_pm_idx > 0) {
    rel15->precodingAndBeamforming.num_prgs = 1; rel15->precodingAndBeamforming.prg_size = rel15->rbSize; rel15->precodingAndBeamforming.prgs_list[0].pm_idx = g_pm_idx;
    pm_pdus = calloc(g_pm_idx, sizeof(*pm_pdus));
    gNB->gNB_config.pmi_list.num_pm_idx = g_pm_idx; gNB->gNB_config.pmi_list.pmi_pdu = pm_pdus;
    nfapi_nr_pm_pdu_t *e = &pm_pdus[g_pm_idx - 1];
    e->pm_idx = g_pm_idx; e->numLayers = rel15->nrOfLayers; e->num_ant_ports = ntx;
    for (int l = 0; l < 4; l++) for (int a = 0; a < 4; a++) { e->weights[l][a].precoder_weight_Re = g_pm_w[l][a][0]; e->weights[l][a].precoder_weight_Im = g_pm_w[l][a][1]; }
  }
  processingData_L1tx_t *msgTx = calloc(1, sizeof(*msgTx));
  msgTx->gNB = gNB; msgTx->dlsch = dlv; msgTx->num_pdsch_slot = 1; msgTx->slot = slot;
  g_bits = bits; g_nbits = nbits;
  const double t0 = now_s();
  nr_generate_pdsch(msgTx, 0, slot);
  g_last_s = now_s() - t0;
  for (int a = 0; a < ntx; a++) memcpy(txdataF_out + 2 * (size_t)a * 14 * N, gNB->common_vars.txdataF[a] + (size_t)slot * 14 * N, sizeof(c16_t) * 14 * N);
  for (int s = 0; s < fp->slots_per_frame; s++) { for (int l = 0; l < 14; l++) { for (int q = 0; q < 2; q++) free(gNB->nr_gold_pdsch_dmrs[s][l][q]); free(gNB->nr_gold_pdsch_dmrs[s][l]); } free(gNB->nr_gold_pdsch_dmrs[s]); }
  free(gNB->nr_gold_pdsch_dmrs);
  for (int a = 0; a < ntx; a++) free(gNB->common_vars.txdataF[a]);
  free(gNB->common_vars.txdataF); free(gNB->common_vars.beam_id[0]); free(gNB->common_vars.beam_id); free(harq->f); free(dl); free(msgTx); free(gNB); free(pm_pdus);
  return 0;
}
