/* Link-time interposer for OpenAirInterface: nr_pusch_channel_estimation on the GPU with host C unchanged.
 *
 * OAI has no plug-in boundary for channel estimation: the function is called directly from nr_rx_pusch_tp
 * (openair1/PHY/NR_TRANSPORT/nr_ulsch_demodulation.c:1473-1492).  This file DEFINES the same symbol with the same prototype
 * (openair1/PHY/NR_ESTIMATION/nr_ul_estimation.h, nr_ul_channel_estimation.c:67-75) and forwards to libldpc_b200.so, so that
 * linking it ahead of libPHY_NR with -Wl,--allow-multiple-definition (first definition wins) or with -Wl,--wrap routes every call to the B200 library.
 * (OAI links the PHY statically into the softmodem, and a definition inside the executable wins over LD_PRELOAD: preloading only works for a build that keeps
 * the PHY in a shared library.)  It is compiled against the reference's own headers (integration/build_shims.sh), reads exactly the
 * fields the reference function reads (frame_parms, common_vars.rxdataF, pusch_vars[ul_id].ul_ch_estimates, the PDU) and writes exactly what it
 * writes (the DMRS symbol of ul_ch_estimates for every rx antenna, *max_ch, *nvar, gNB->ulsch[ul_id].delay).  Transform precoding is served with OAI's
 * own low-PAPR sequence table.  A configuration the library refuses aborts loudly like any other AssertFatal in this code base: there is no CPU fallback.
 * Test: tests/test_gpu_interpose.py drives it through the reference-side caller harness and compares with the pinned oracle. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "PHY/defs_gNB.h"
#include "PHY/NR_ESTIMATION/nr_ul_estimation.h"
#include "PHY/NR_REFSIG/ul_ref_seq_nr.h"
#define NRB200_NO_OAI_LOADER_PROTOTYPES
#include "nrb200_ldpc.h"

int nr_pusch_channel_estimation(PHY_VARS_gNB *gNB, unsigned char Ns, unsigned short p, unsigned char symbol, int ul_id, unsigned short bwp_start_subcarrier,
                                nfapi_nr_pusch_pdu_t *pusch_pdu, int *max_ch, uint32_t *nvar)
{
  const NR_DL_FRAME_PARMS *fp = &gNB->frame_parms;
  const int N = fp->ofdm_symbol_size, nrx = fp->nb_antennas_rx;
  const int soffset = (Ns & 3) * fp->symbols_per_slot * N;
  (void)bwp_start_subcarrier;   /* = ((rb_start + bwp_start) * 12 + first_carrier_offset) % N, recomputed by the library from the PDU */
  nrb200_pusch_chest_t d;
  memset(&d, 0, sizeof(d));
  d.fft_size = N; d.nb_rx = nrx; d.slot = Ns; d.symbol = symbol; d.port = p;
  d.rb_start = pusch_pdu->rb_start; d.bwp_start = pusch_pdu->bwp_start; d.rb_size = pusch_pdu->rb_size; d.first_carrier_offset = fp->first_carrier_offset;
  d.scid = pusch_pdu->scid; d.ul_dmrs_scrambling_id = pusch_pdu->ul_dmrs_scrambling_id;
  d.n_ports = 1; d.dmrs_config_type = pusch_pdu->dmrs_config_type; d.chest_freq = gNB->chest_freq;
  if (pusch_pdu->transform_precoding != transformPrecoder_disabled) {
    /* DFT-s-OFDM: the low-PAPR type-1 sequence OAI generated at start-up (generate_lowpapr_typ1_refsig_sequences), selected as the reference
     * selects it (nr_ul_channel_estimation.c:125-131) */
    const int16_t index = get_index_for_dmrs_lowpapr_seq(pusch_pdu->rb_size * (NR_NB_SC_PER_RB / 2));
    const int16_t *dmrs_seq = index >= 0 ? gNB_dmrs_lowpaprtype1_sequence[pusch_pdu->dfts_ofdm.low_papr_group_number][pusch_pdu->dfts_ofdm.low_papr_sequence_number][index] : NULL;
    if (!dmrs_seq) { fprintf(stderr, "nrb200 shim: no low-PAPR DMRS sequence for %d PRBs\n", pusch_pdu->rb_size); abort(); }
    d.transform_precoding = 1; d.lowpapr_seq = (uint64_t)(uintptr_t)dmrs_seq;
  }
  /* the library's host entry point takes the slot as [nb_rx][14][N] c16; OAI keeps one buffer per antenna with a 4-slot ring */
  const size_t plane = (size_t)14 * N;
  int16_t *rx = malloc(4 * plane * nrx), *est = malloc(4 * plane * nrx);
  if (!rx || !est) abort();
  NR_gNB_PUSCH *pv = &gNB->pusch_vars[ul_id];
  for (int a = 0; a < nrx; a++) {
    memcpy(rx + 2 * plane * a, &gNB->common_vars.rxdataF[a][soffset], 4 * plane);
    memcpy(est + 2 * plane * a, pv->ul_ch_estimates[p * nrx + a], 4 * plane);
  }
  int32_t state[5];
  const int rc = nrb200_pusch_chest_host(&d, rx, est, state);
  if (rc != 0) { fprintf(stderr, "nrb200 shim: nrb200_pusch_chest_host failed (rc = %d)\n", rc); abort(); }
  for (int a = 0; a < nrx; a++)
    memcpy(&pv->ul_ch_estimates[p * nrx + a][symbol * N], est + 2 * (plane * a + (size_t)symbol * N), 4 * (size_t)N);
  /* *max_ch is a running maximum over the caller's ports and symbols (max(*max_ch, ...) in the reference); *nvar is written only when
   * the estimator counted noise samples (chest_freq == 0) */
  if (state[0] > *max_ch) *max_ch = state[0];
  if (nvar && gNB->chest_freq == 0) *nvar = (uint32_t)state[1];
  delay_t *delay = &gNB->ulsch[ul_id].delay;
  memset(delay, 0, sizeof(*delay));
  delay->est_delay = state[2]; delay->delay_max_pos = state[3]; delay->delay_max_val = state[4];
  free(rx); free(est);
  return 0;
}
