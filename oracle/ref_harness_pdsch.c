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
