/* TEST INFRASTRUCTURE ONLY -- never linked into the product.
 * Flat C entry point around the UNMODIFIED reference PDSCH channel estimator of the UE (nr_pdsch_channel_estimation,
 * openair1/PHY/NR_UE_ESTIMATION/nr_dl_channel_estimation.c:1614-1735, with NFAPI_NR_DMRS_TYPE1_linear_interp :1305-1385, nr_dmrs_rx.c,
 * nr_gold_ue.c, dmrs_nr.c, common/utils/nr/nr_common.c compiled from /root/reference by build_ref.sh).  The harness allocates the parts of
 * PHY_VARS_NR_UE the function touches and binds the dft/idft function-pointer globals to the compiled reference libref_dfts.so. */
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "PHY/defs_nr_UE.h"
#include "PHY/NR_UE_ESTIMATION/nr_estimation.h"
#include <time.h>
/* wall time of the last call into the reference function(s), excluding the harness's own allocation and copying (cpu_baseline of the DL slot chain) */
static double g_last_s;
static inline double now_s(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }


void init_delay_table(uint16_t ofdm_symbol_size, int max_delay_comp, int max_ofdm_symbol_size, c16_t delay_table[][max_ofdm_symbol_size]);

dftfunc_t dft;
idftfunc_t idft;

int refh_uechest_init(const char *dfts_so)
{
  void *h = dlopen(dfts_so, RTLD_NOW | RTLD_LOCAL);
  if (!h) { fprintf(stderr, "refh_uechest_init: %s\n", dlerror()); return -1; }
  int (*autoinit)(void) = (int (*)(void))dlsym(h, "dfts_autoinit");
  dft = (dftfunc_t)dlsym(h, "dft");
  idft = (idftfunc_t)dlsym(h, "idft");
  if (!autoinit || !dft || !idft) return -2;
  autoinit();
  return 0;
}

double refh_uechest_last_seconds(void) { return g_last_s; }

enum { U_N, U_NB_RX, U_N_RB_DL, U_SLOT, U_SYMBOL, U_PORT, U_RB_START, U_BWP_START, U_RB_SIZE, U_FCO, U_SCID, U_DMRS_ID, U_DMRS_TYPE, U_CHEST_FREQ, U_COUNT };

/* rxdataF: [nb_rx][14*N] c16; dl_ch_est out: [nb_rx][14*N] c16 (symbol `symbol` of port `port` is written) */
int refh_pdsch_chest(const int32_t *p, const int16_t *rxdataF, int16_t *dl_ch_est)
{
  const int N = p[U_N], nrx = p[U_NB_RX], Ns = p[U_SLOT], port = p[U_PORT];
  PHY_VARS_NR_UE *ue = calloc(1, sizeof(*ue));
  NR_DL_FRAME_PARMS *fp = &ue->frame_parms;
  fp->ofdm_symbol_size = N; fp->symbols_per_slot = 14; fp->nb_antennas_rx = nrx; fp->N_RB_DL = p[U_N_RB_DL]; fp->slots_per_frame = 20;
  fp->Ncp = NORMAL; fp->first_carrier_offset = p[U_FCO];
  init_delay_table(N, MAX_DELAY_COMP, NR_MAX_OFDM_SYMBOL_SIZE, fp->delay_table);
  ue->chest_freq = p[U_CHEST_FREQ];
  ue->scramblingID_dlsch[0] = ue->scramblingID_dlsch[1] = (uint16_t)(p[U_DMRS_ID] ^ 1);   /* forces nr_gold_pdsch to run */
  const int words = ((fp->N_RB_DL * 12) >> 5) + 1;
  ue->nr_gold_pdsch[0] = calloc(fp->slots_per_frame, sizeof(uint32_t ***));
  for (int ns = 0; ns < fp->slots_per_frame; ns++) {
    ue->nr_gold_pdsch[0][ns] = calloc(14, sizeof(uint32_t **));
    for (int l = 0; l < 14; l++) {
      ue->nr_gold_pdsch[0][ns][l] = calloc(2, sizeof(uint32_t *));
      for (int s = 0; s < 2; s++) ue->nr_gold_pdsch[0][ns][l][s] = calloc(words + 2, 4);
    }
  }
  UE_nr_rxtx_proc_t proc;
  memset(&proc, 0, sizeof(proc));
  proc.gNB_id = 0; proc.nr_slot_rx = Ns;
  const int est_size = 14 * N;
  int32_t (*est)[est_size] = calloc((size_t)(port + 1) * nrx, sizeof(int32_t) * est_size);
  c16_t (*rx)[est_size] = calloc(nrx, sizeof(c16_t) * est_size);
  memcpy(rx, rxdataF, (size_t)nrx * est_size * 4);
  /* arguments as nr_ue_pdsch_procedures passes them (phy_procedures_nr_ue.c): BWPStart, rb_offset, bwp_start_subcarrier */
  const unsigned short k0 = ((p[U_RB_START] + p[U_BWP_START]) * 12 + p[U_FCO]) % N;
  const double t0 = now_s();
  nr_pdsch_channel_estimation(ue, &proc, (unsigned short)port, (unsigned char)p[U_SYMBOL], (unsigned char)p[U_SCID], (unsigned short)p[U_DMRS_ID],
                              (unsigned short)p[U_BWP_START], (uint8_t)p[U_DMRS_TYPE], (uint16_t)(p[U_RB_START] + p[U_BWP_START]), k0, (unsigned short)p[U_RB_SIZE], est_size, est,
                              est_size, rx);
  g_last_s = now_s() - t0;
  for (int a = 0; a < nrx; a++) memcpy(dl_ch_est + 2 * (size_t)a * est_size, est[port * nrx + a], 4 * (size_t)est_size);
  for (int ns = 0; ns < fp->slots_per_frame; ns++) { for (int l = 0; l < 14; l++) { for (int s = 0; s < 2; s++) free(ue->nr_gold_pdsch[0][ns][l][s]); free(ue->nr_gold_pdsch[0][ns][l]); } free(ue->nr_gold_pdsch[0][ns]); }
  free(ue->nr_gold_pdsch[0]); free(est); free(rx); free(ue);
  return 0;
}

/* UE OFDM front end: the real nr_slot_fep (MODULATION/slot_fep_nr.c:37-113) for the 14 symbols of slot Ns of a synchronised UE: dft + apply_nr_rotation_RX
 * with the DL rotation table.  rxdata: [nb_rx][samples] c16 starting at sample 0 of the frame; rot_dl: 224 pairs; timeshift: N pairs; rxdataF out [nb_rx][14 N]. */
static uint32_t u_samples_per_slot(int slot, const NR_DL_FRAME_PARMS *fp)
{
  if (fp->numerology_index == 0) return fp->samples_per_subframe;
  return (slot % (fp->slots_per_subframe / 2)) ? fp->samples_per_slotN0 : fp->samples_per_slot0;
}
static uint32_t u_slot_timestamp(int slot, const NR_DL_FRAME_PARMS *fp, uint8_t ahead)
{
  uint32_t s = 0;
  for (int i = ahead ? slot : 0; i < (ahead ? slot + ahead : slot); i++) s += u_samples_per_slot(i, fp);
  return s;
}
int refh_ue_slot_fep(int N, int mu, int nb_rb, int nrx, int Ns, int divisor, const int16_t *rot_dl, const int16_t *timeshift, const int16_t *rxdata, uint32_t n_samples,
                     int16_t *rxdataF)
{
  PHY_VARS_NR_UE *ue = calloc(1, sizeof(*ue));
  NR_DL_FRAME_PARMS *fp = &ue->frame_parms;
  fp->ofdm_symbol_size = N; fp->numerology_index = mu; fp->slots_per_subframe = 1 << mu; fp->slots_per_frame = 10 << mu; fp->symbols_per_slot = 14;
  fp->N_RB_DL = fp->N_RB_UL = nb_rb; fp->first_carrier_offset = N - nb_rb * 6; fp->nb_antennas_rx = nrx;
  fp->nb_prefix_samples = N / 128 * 9; fp->nb_prefix_samples0 = N / 128 * (9 + (1 << mu));
  fp->samples_per_slotN0 = (fp->nb_prefix_samples + N) * 14; fp->samples_per_slot0 = fp->nb_prefix_samples0 + 13 * fp->nb_prefix_samples + 14 * N;
  fp->samples_per_subframe = (fp->nb_prefix_samples0 + N) * 2 + (fp->nb_prefix_samples + N) * (14 * fp->slots_per_subframe - 2);
  fp->samples_per_frame = 10 * fp->samples_per_subframe; fp->samples_per_slot_wCP = 14 * N;
  fp->get_samples_per_slot = u_samples_per_slot; fp->get_samples_slot_timestamp = u_slot_timestamp; fp->ofdm_offset_divisor = divisor;
  memcpy(fp->symbol_rotation[0], rot_dl, 224 * 4);
  memcpy(fp->timeshift_symbol_rotation, timeshift, (size_t)N * 4);
  ue->is_synchronized = 1;
  ue->common_vars.rxdata = calloc(nrx, sizeof(c16_t *));
  for (int a = 0; a < nrx; a++) { posix_memalign((void **)&ue->common_vars.rxdata[a], 32, 4 * (size_t)n_samples + 64); memcpy(ue->common_vars.rxdata[a], rxdata + 2 * (size_t)a * n_samples, 4 * (size_t)n_samples); }
  UE_nr_rxtx_proc_t proc;
  memset(&proc, 0, sizeof(proc));
  proc.nr_slot_rx = Ns;
  c16_t (*rxF)[14 * N];
  posix_memalign((void **)&rxF, 32, sizeof(c16_t) * (size_t)nrx * 14 * N);
  memset(rxF, 0, sizeof(c16_t) * (size_t)nrx * 14 * N);
  const double t0 = now_s();
  for (int l = 0; l < 14; l++) nr_slot_fep(ue, &proc, (unsigned char)l, rxF);
  g_last_s = now_s() - t0;
  memcpy(rxdataF, rxF, sizeof(c16_t) * (size_t)nrx * 14 * N);
  for (int a = 0; a < nrx; a++) free(ue->common_vars.rxdata[a]);
  free(ue->common_vars.rxdata); free(rxF); free(ue);
  return 0;
}
