/* Link-time interposer for the OAI UE: nr_pdsch_channel_estimation on the GPU with host C unchanged (the UE-side twin of
 * oai_shim_pusch_chest.c; see there for how it is used).  Same prototype as openair1/PHY/NR_UE_ESTIMATION/nr_dl_channel_estimation.c:1614-1628, compiled
 * against OAI's headers; called once per DMRS symbol and port from nr_ue_pdsch_procedures (SCHED_NR_UE/phy_procedures_nr_ue.c:527-543).  It reads what the
 * reference function reads (frame_parms, the slot's rxdataF planes, ue->chest_freq) and rewrites the DMRS symbol of dl_ch_estimates[p * nb_rx + aarx]. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "PHY/defs_nr_UE.h"
#include "PHY/NR_UE_ESTIMATION/nr_estimation.h"
#include "nfapi_nr_interface.h"
#define NRB200_NO_OAI_LOADER_PROTOTYPES
#include "nrb200_ldpc.h"

int nr_pdsch_channel_estimation(PHY_VARS_NR_UE *ue, const UE_nr_rxtx_proc_t *proc, unsigned short p, unsigned char symbol, unsigned char nscid,
                                unsigned short scrambling_id, unsigned short BWPStart, uint8_t config_type, uint16_t rb_offset, unsigned short bwp_start_subcarrier,
                                unsigned short nb_rb_pdsch, uint32_t pdsch_est_size, int32_t dl_ch_estimates[][pdsch_est_size], int rxdataFsize,
                                c16_t rxdataF[][rxdataFsize])
{
  const NR_DL_FRAME_PARMS *fp = &ue->frame_parms;
  const int N = fp->ofdm_symbol_size, nrx = fp->nb_antennas_rx;
  (void)BWPStart;
  nrb200_pusch_chest_t d;
  memset(&d, 0, sizeof(d));
  d.fft_size = N; d.nb_rx = nrx; d.slot = proc->nr_slot_rx; d.symbol = symbol; d.port = p;
  /* the DMRS sequence is indexed from common resource block rb_offset; the first sub-carrier is given directly */
  d.rb_start = rb_offset; d.bwp_start = 0; d.rb_size = nb_rb_pdsch;
  d.first_carrier_offset = (uint32_t)(((int)bwp_start_subcarrier - 12 * (int)rb_offset) % N + N) % N;
  d.scid = nscid; d.ul_dmrs_scrambling_id = scrambling_id;
  d.n_ports = 1; d.pdsch_ue = 1; d.dmrs_config_type = config_type == NFAPI_NR_DMRS_TYPE1 ? 0 : 1; d.chest_freq = ue->chest_freq;
  const size_t plane = (size_t)14 * N;
  int16_t *rx = malloc(4 * plane * nrx), *est = malloc(4 * plane * nrx);
  if (!rx || !est) abort();
  for (int a = 0; a < nrx; a++) {
    memcpy(rx + 2 * plane * a, rxdataF[a], 4 * plane);
    memcpy(est + 2 * plane * a, dl_ch_estimates[p * nrx + a], 4 * plane);
  }
  int32_t state[5];
  const int rc = nrb200_pusch_chest_host(&d, rx, est, state);
  if (rc != 0) { fprintf(stderr, "nrb200 shim: nrb200_pusch_chest_host (pdsch_ue) failed (rc = %d)\n", rc); abort(); }
  for (int a = 0; a < nrx; a++)
    memcpy(&dl_ch_estimates[p * nrx + a][symbol * N], est + 2 * (plane * a + (size_t)symbol * N), 4 * (size_t)N);
  free(rx); free(est);
  return 0;
}
