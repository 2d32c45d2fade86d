/* TEST INFRASTRUCTURE ONLY -- never linked into the product.
 * Flat C entry point around the UNMODIFIED reference PDSCH receiver of the UE: nr_rx_pdsch (openair1/PHY/NR_UE_TRANSPORT/nr_dlsch_demodulation.c:241-684)
 * is called symbol by symbol exactly like nr_ue_pdsch_procedures does (SCHED_NR_UE/phy_procedures_nr_ue.c), so extraction, scaling, level, compensation,
 * MRC and -- at the last symbol -- the LLRs of the whole slot come from the reference's own control flow.  The harness fills the fields of
 * PHY_VARS_NR_UE / NR_UE_DLSCH_t / NR_DL_UE_HARQ_t that function reads. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "PHY/defs_nr_UE.h"
#include "PHY/NR_UE_TRANSPORT/nr_transport_proto_ue.h"
#include <time.h>
/* wall time of the last call into the reference function(s), excluding the harness's own allocation and copying (cpu_baseline of the DL slot chain) */
static double g_last_s;
static inline double now_s(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }


double refh_pdsch_last_seconds(void) { return g_last_s; }

#ifdef REFH_PTRS
/* PT-RS (libref_pdsch_ptrs.so only: the same harness linked with the real nr_pdsch_ptrs_processing of NR_UE_ESTIMATION/nr_dl_channel_estimation.c:1765-1907
 * and NR_REFSIG/ptrs_nr.c).  q: enabled, PTRSTimeDensity (log2 of L), PTRSFreqDensity (K), PTRSReOffset, rnti, nr_slot_rx, nscid, scramblingID_dlsch, N_RB_DL */
enum { Q_ON, Q_L, Q_K, Q_REOFF, Q_RNTI, Q_SLOT, Q_NSCID, Q_NID, Q_NRB, Q_COUNT };
static int32_t g_ptrs[Q_COUNT];
static int16_t g_phase[14 * 2];
static int32_t g_ptrs_re[14];
void refh_pdsch_set_ptrs(const int32_t *q) { if (q) memcpy(g_ptrs, q, sizeof(g_ptrs)); else memset(g_ptrs, 0, sizeof(g_ptrs)); }
/* what the last slot left in ptrs_phase_per_slot[0] / ptrs_re_per_slot[0] */
void refh_pdsch_get_ptrs(int16_t *phase28, int32_t *re14) { memcpy(phase28, g_phase, sizeof(g_phase)); memcpy(re14, g_ptrs_re, sizeof(g_ptrs_re)); }
void nr_gold_pdsch(PHY_VARS_NR_UE *ue, int nscid, uint32_t nid);
#endif

enum { D_N, D_NB_RX, D_RB_START, D_BWP_START, D_RB_SIZE, D_FCO, D_QM, D_START_SYMBOL, D_NR_SYMBOLS, D_DMRS_POS, D_DMRS_TYPE, D_CDM_GROUPS, D_G, D_NL, D_COUNT };

/* rxdataF: [nb_rx][14 N] c16; dl_ch_est: [nl * nb_rx][14 N] c16 (plane layer * nb_rx + rx); llr out: G int16 (layer de-mapped).  Returns log2_maxh; valid_re_out (14) and comp_out (nb_rb*12*14 c16 of rx 0) optional. */
int refh_pdsch_rx_slot(const int32_t *p, const int16_t *rxdataF, const int16_t *dl_ch_est, int16_t *llr_out, int32_t *valid_re_out, int16_t *comp_out)
{
  const int N = p[D_N], nrx = p[D_NB_RX], nb_rb = p[D_RB_SIZE], nl = p[D_NL] > 1 ? p[D_NL] : 1;
  PHY_VARS_NR_UE *ue = calloc(1, sizeof(*ue));
  NR_DL_FRAME_PARMS *fp = &ue->frame_parms;
  fp->ofdm_symbol_size = N; fp->symbols_per_slot = 14; fp->nb_antennas_rx = nrx; fp->N_RB_DL = 273; fp->first_carrier_offset = p[D_FCO];
  fp->samples_per_slot_wCP = 14 * N; fp->Ncp = NORMAL;
  ue->chest_time = 0;
  NR_UE_DLSCH_t dlsch[2];
  memset(dlsch, 0, sizeof(dlsch));
  dlsch[0].Nl = nl; dlsch[0].active = true; dlsch[0].rnti_type = 0;
  fapi_nr_dl_config_dlsch_pdu_rel15_t *c = &dlsch[0].dlsch_config;
  c->BWPStart = p[D_BWP_START]; c->start_rb = p[D_RB_START]; c->number_rbs = nb_rb; c->start_symbol = p[D_START_SYMBOL]; c->number_symbols = p[D_NR_SYMBOLS];
  c->dlDmrsSymbPos = p[D_DMRS_POS]; c->dmrsConfigType = p[D_DMRS_TYPE]; c->n_dmrs_cdm_groups = p[D_CDM_GROUPS]; c->qamModOrder = p[D_QM]; c->pduBitmap = 0;
  const int harq_pid = 0;
  ue->dl_harq_processes[0][harq_pid].status = ACTIVE; ue->dl_harq_processes[0][harq_pid].codeword = 0; ue->dl_harq_processes[0][harq_pid].G = p[D_G];
  ue->dl_harq_processes[1][harq_pid].status = SCH_IDLE;
  UE_nr_rxtx_proc_t proc;
  memset(&proc, 0, sizeof(proc));
#ifdef REFH_PTRS
  if (g_ptrs[Q_ON]) {
    c->pduBitmap = 1; dlsch[0].rnti_type = TYPE_C_RNTI_; dlsch[0].rnti = (uint16_t)g_ptrs[Q_RNTI];
    c->PTRSTimeDensity = g_ptrs[Q_L]; c->PTRSFreqDensity = g_ptrs[Q_K]; c->PTRSReOffset = g_ptrs[Q_REOFF]; c->nscid = g_ptrs[Q_NSCID];
    proc.nr_slot_rx = g_ptrs[Q_SLOT]; proc.gNB_id = 0;
    ue->scramblingID_dlsch[g_ptrs[Q_NSCID]] = (uint16_t)g_ptrs[Q_NID];
    fp->N_RB_DL = g_ptrs[Q_NRB]; fp->slots_per_frame = 20;
    const int words = ((fp->N_RB_DL * 12) >> 5) + 1;
    ue->nr_gold_pdsch[0] = calloc(fp->slots_per_frame, sizeof(uint32_t ***));
    for (int ns = 0; ns < fp->slots_per_frame; ns++) {
      ue->nr_gold_pdsch[0][ns] = calloc(14, sizeof(uint32_t **));
      for (int l = 0; l < 14; l++) {
        ue->nr_gold_pdsch[0][ns][l] = calloc(2, sizeof(uint32_t *));
        for (int s = 0; s < 2; s++) ue->nr_gold_pdsch[0][ns][l][s] = calloc(words + 2, 4);
      }
    }
    nr_gold_pdsch(ue, g_ptrs[Q_NSCID], (uint32_t)g_ptrs[Q_NID]);        /* the reference's own generator (NR_REFSIG/nr_gold_ue.c:75-93) */
  }
#endif
  const int est_size = 14 * N, rx_size_symbol = (nb_rb * 12 + 15) & ~15;
  int32_t (*est)[est_size] = calloc((size_t)nl * nrx, sizeof(int32_t) * est_size);
  c16_t (*rx)[est_size] = calloc(nrx, sizeof(c16_t) * est_size);
  memcpy(est, dl_ch_est, (size_t)nl * nrx * est_size * 4);
  memcpy(rx, rxdataF, (size_t)nrx * est_size * 4);
  int32_t (*comp)[nrx][rx_size_symbol * 14];
  posix_memalign((void **)&comp, 32, sizeof(int32_t) * nl * nrx * rx_size_symbol * 14);
  memset(comp, 0, sizeof(int32_t) * nl * nrx * rx_size_symbol * 14);
  int16_t *llr[2];
  posix_memalign((void **)&llr[0], 64, 2 * (size_t)p[D_G] + 4096); memset(llr[0], 0, 2 * (size_t)p[D_G] + 4096);
  llr[1] = NULL;
  uint32_t dl_valid_re_buf[16] = {0}, *dl_valid_re = dl_valid_re_buf + 1;   /* the reference writes dl_valid_re[symbol - 1] */
  uint32_t llr_offset_buf[16] = {0}, *llr_offset = llr_offset_buf + 1;
  int32_t log2_maxh = 0;
  c16_t ptrs_phase[nrx][14];
  int32_t ptrs_re[nrx][14];
  memset(ptrs_phase, 0, sizeof(ptrs_phase)); memset(ptrs_re, 0, sizeof(ptrs_re));
  /* first_symbol_flag as nr_ue_pdsch_procedures derives it (phy_procedures_nr_ue.c:568-590) */
  int first_symbol_with_data = p[D_START_SYMBOL];
  const int dmrs_data_re = p[D_DMRS_TYPE] == 0 ? 12 - 6 * p[D_CDM_GROUPS] : 12 - 4 * p[D_CDM_GROUPS];
  while (dmrs_data_re == 0 && (p[D_DMRS_POS] & (1 << first_symbol_with_data))) first_symbol_with_data++;
  const double t0 = now_s();
  for (int m = p[D_START_SYMBOL]; m < p[D_START_SYMBOL] + p[D_NR_SYMBOLS]; m++) {
    const int first_symbol_flag = m == first_symbol_with_data;
    if (nr_rx_pdsch(ue, &proc, dlsch, (unsigned char)m, (unsigned char)first_symbol_flag, harq_pid, est_size, est, llr, dl_valid_re, rx, llr_offset, &log2_maxh,
                    rx_size_symbol, nrx, comp, ptrs_phase, ptrs_re) < 0) { fprintf(stderr, "nr_rx_pdsch failed at symbol %d\n", m); break; }
  }
  g_last_s = now_s() - t0;
  memcpy(llr_out, llr[0], 2 * (size_t)p[D_G]);
  if (valid_re_out) for (int m = 0; m < 14; m++) valid_re_out[m] = (int32_t)dl_valid_re_buf[m];     /* index m holds symbol m (stored at [symbol - 1] + 1) */
  if (comp_out) memcpy(comp_out, comp[0][0], 4 * (size_t)rx_size_symbol * 14);
#ifdef REFH_PTRS
  memcpy(g_phase, ptrs_phase[0], sizeof(g_phase));
  memcpy(g_ptrs_re, ptrs_re[0], sizeof(g_ptrs_re));
  if (g_ptrs[Q_ON]) {
    for (int ns = 0; ns < fp->slots_per_frame; ns++) { for (int l = 0; l < 14; l++) { for (int s = 0; s < 2; s++) free(ue->nr_gold_pdsch[0][ns][l][s]); free(ue->nr_gold_pdsch[0][ns][l]); } free(ue->nr_gold_pdsch[0][ns]); }
    free(ue->nr_gold_pdsch[0]);
  }
#endif
  free(est); free(rx); free(comp); free(llr[0]); free(ue);
  return log2_maxh;
}


/* nr_chest_time_domain_avg (openair1/PHY/NR_REFSIG/dmrs_nr.c:343-417), the real function: est is [nb_rx][14][N] c16, rewritten in place. */
void nr_chest_time_domain_avg(NR_DL_FRAME_PARMS *frame_parms, int32_t **ch_estimates, uint8_t num_symbols, uint8_t start_symbol, uint16_t dmrs_bitmap, uint16_t num_rbs);
int refh_chest_time_avg(int N, int nb_rx, int num_symbols, int start_symbol, int dmrs_bitmap, int num_rbs, int16_t *est)
{
  NR_DL_FRAME_PARMS *fp = calloc(1, sizeof(*fp));
  fp->ofdm_symbol_size = N; fp->nb_antennas_rx = nb_rx;
  int32_t **planes = calloc(nb_rx, sizeof(int32_t *));
  for (int a = 0; a < nb_rx; a++) { posix_memalign((void **)&planes[a], 32, 4 * (size_t)14 * N); memcpy(planes[a], est + 2 * (size_t)a * 14 * N, 4 * (size_t)14 * N); }
  nr_chest_time_domain_avg(fp, planes, (uint8_t)num_symbols, (uint8_t)start_symbol, (uint16_t)dmrs_bitmap, (uint16_t)num_rbs);
  for (int a = 0; a < nb_rx; a++) { memcpy(est + 2 * (size_t)a * 14 * N, planes[a], 4 * (size_t)14 * N); free(planes[a]); }
  free(planes); free(fp);
  return 0;
}
