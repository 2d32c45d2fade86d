/* Link-time interposer for the OAI gNB: nr_rx_pusch_tp on the GPU with host C unchanged (see oai_shim_pusch_chest.c for how the
 * interposers are used).  Same prototype as openair1/PHY/NR_TRANSPORT/nr_ulsch_demodulation.c:1447-1451, compiled against OAI's headers.
 *
 * The reference function (:1447-1700) estimates the channel on every DMRS symbol and layer, measures powers, derives log2_maxh from the first symbol that
 * carries data, and pushes one thread-pool job per group of symbols (extraction, compensation / MMSE / ML, LLRs, layer de-mapping, unscrambling).  Here:
 *   - channel estimation: nr_pusch_channel_estimation per DMRS symbol and layer, called BY NAME as the reference does -- with oai_shim_pusch_chest.c linked
 *     that is the GPU estimator, otherwise the reference's; *max_ch / *nvar accumulate the same way
 *   - measurements (nr_gnb_measurements, signal_energy_nodc, n0_subband_power): OAI's own control-plane helpers, called like the reference calls them -- the
 *     caller decides DTX from ulsch_power / ulsch_noise_power (phy_procedures_nr_gNB.c), so they must be there
 *   - everything after that -- level measurement, the inner receiver of all symbols, layer de-mapping, unscrambling -- is ONE call into libldpc_b200.so
 * It writes what the reference writes for its caller: pusch_vars->llr, ->log2_maxh, ->ul_valid_re_per_slot[], ->llr_offset[], ->dmrs_symbol, ->ulsch_power[],
 * ->ulsch_noise_power[], ulsch->unav_res, and (through the estimator) ul_ch_estimates and the delay.  rxdataF_ext / ul_ch_estimates_ext / rxdataF_comp are the
 * reference's scratch and stay untouched.  Not served: PT-RS, transform precoding, more than two layers -- the call aborts loudly, there is no CPU fallback.
 * Test: tests/test_gpu_interpose.py through oracle/ref_harness_rxpusch.c (reference-side caller) against the pinned oracle. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "PHY/defs_gNB.h"
#include "PHY/NR_ESTIMATION/nr_ul_estimation.h"
#include "PHY/NR_TRANSPORT/nr_transport_proto.h"
#include "PHY/NR_REFSIG/dmrs_nr.h"
#include "PHY/NR_TRANSPORT/nr_sch_dmrs.h"
#include "common/utils/nr/nr_common.h"
#define NRB200_NO_OAI_LOADER_PROTOTYPES
#include "nrb200_ldpc.h"

#define INVALID_VALUE 255   /* nr_ulsch_demodulation.c:14 */

int nr_rx_pusch_tp(PHY_VARS_gNB *gNB, uint8_t ulsch_id, uint32_t frame, uint8_t slot, unsigned char harq_pid)
{
  (void)frame; (void)harq_pid;
  NR_DL_FRAME_PARMS *fp = &gNB->frame_parms;
  nfapi_nr_pusch_pdu_t *pdu = &gNB->ulsch[ulsch_id].harq_process->ulsch_pdu;
  NR_gNB_PUSCH *pv = &gNB->pusch_vars[ulsch_id];
  const int N = fp->ofdm_symbol_size, nrx = fp->nb_antennas_rx, nl = pdu->nrOfLayers;
  if ((pdu->pdu_bit_map & PUSCH_PDU_BITMAP_PUSCH_PTRS) || nl < 1 || nl > 2) {
    fprintf(stderr, "nrb200 shim: nr_rx_pusch_tp: PT-RS / %d layers are not served by libldpc_b200\n", nl);
    abort();
  }
  pv->dmrs_symbol = INVALID_VALUE;
  gNB->nbSymb = 0;
  const uint32_t bwp_start_subcarrier = ((pdu->rb_start + pdu->bwp_start) * NR_NB_SC_PER_RB + fp->first_carrier_offset) % N;
  const int first = pdu->start_symbol_index, end = pdu->start_symbol_index + pdu->nr_of_symbols;
  /* ---- channel estimation + the measurements the caller's DTX decision reads (:1470-1522) */
  int max_ch = 0;
  uint32_t nvar = 0;
  for (int symbol = first; symbol < end; symbol++) {
    if (!((pdu->ul_dmrs_symb_pos >> symbol) & 1)) continue;
    if (pv->dmrs_symbol == INVALID_VALUE) pv->dmrs_symbol = symbol;
    for (int l = 0; l < nl; l++) {
      uint32_t nvar_tmp = 0;
      nr_pusch_channel_estimation(gNB, slot, get_dmrs_port(l, pdu->dmrs_ports), symbol, ulsch_id, bwp_start_subcarrier, pdu, &max_ch, &nvar_tmp);
      nvar += nvar_tmp;
    }
    nr_gnb_measurements(gNB, &gNB->ulsch[ulsch_id], pv, symbol, nl);
    allocCast2D(n0_subband_power, unsigned int, gNB->measurements.n0_subband_power, nrx, fp->N_RB_UL, false);
    for (int a = 0; a < nrx; a++) {
      if (symbol == first) { pv->ulsch_power[a] = 0; pv->ulsch_noise_power[a] = 0; }
      for (int l = 0; l < nl; l++) pv->ulsch_power[a] += signal_energy_nodc(&pv->ul_ch_estimates[l * nrx + a][symbol * N], pdu->rb_size * 12);
      for (int rb = 0; rb < pdu->rb_size; rb++) pv->ulsch_noise_power[a] += n0_subband_power[a][pdu->bwp_start + pdu->rb_start + rb] / pdu->rb_size;
    }
  }
  nvar /= (pdu->nr_of_symbols * nl * nrx);
  if (gNB->chest_time == 1) {   /* averaging across the DMRS symbols (:1527-1538) */
    nr_chest_time_domain_avg(fp, pv->ul_ch_estimates, pdu->nr_of_symbols, pdu->start_symbol_index, pdu->ul_dmrs_symb_pos, pdu->rb_size);
    pv->dmrs_symbol = get_next_dmrs_symbol_in_slot(pdu->ul_dmrs_symb_pos, pdu->start_symbol_index, pdu->nr_of_symbols);
  }
  /* ---- per-symbol bookkeeping the decoder's caller reads (:1545-1570, :1651-1660) */
  const int nb_re_dmrs = pdu->dmrs_config_type == pusch_dmrs_type1 ? 6 * pdu->num_dmrs_cdm_grps_no_data : 4 * pdu->num_dmrs_cdm_grps_no_data;
  gNB->ulsch[ulsch_id].unav_res = 0;
  for (int s = first; s < end; s++) {
    const int dm = (pdu->ul_dmrs_symb_pos >> s) & 1;
    pv->ul_valid_re_per_slot[s] = pdu->rb_size * (dm ? 12 - nb_re_dmrs : 12);
    pv->llr_offset[s] = s == first ? 0 : pv->llr_offset[s - 1] + pv->ul_valid_re_per_slot[s - 1] * pdu->qam_mod_order;
  }
  /* ---- level measurement + inner receiver of the whole slot + layer de-mapping + unscrambling: one library call */
  nrb200_pusch_rx_t d;
  memset(&d, 0, sizeof(d));
  d.fft_size = N; d.nb_rx = nrx; d.rb_start = pdu->rb_start; d.bwp_start = pdu->bwp_start; d.rb_size = pdu->rb_size; d.first_carrier_offset = fp->first_carrier_offset;
  d.qam_mod_order = pdu->qam_mod_order; d.start_symbol_index = pdu->start_symbol_index; d.nr_of_symbols = pdu->nr_of_symbols; d.ul_dmrs_symb_pos = pdu->ul_dmrs_symb_pos;
  d.dmrs_config_type = pdu->dmrs_config_type == pusch_dmrs_type1 ? 0 : 1; d.num_dmrs_cdm_grps_no_data = pdu->num_dmrs_cdm_grps_no_data;
  d.log2_maxh = 0xFFFFFFFFu; d.unscramble = 1; d.rnti = pdu->rnti; d.data_scrambling_id = pdu->data_scrambling_id;
  d.nrOfLayers = nl; d.noise_var = nvar; d.max_ch = (uint32_t)max_ch;
  d.transform_precoding = pdu->transform_precoding == transformPrecoder_enabled;   /* DFT-s-OFDM: equalisation + nr_idft inside the library call (inner_rx :1326-1336) */
  const uint32_t G = nrb200_pusch_num_llr(&d);
  const size_t plane = (size_t)14 * N;
  const int soffset = (slot % RU_RX_SLOT_DEPTH) * fp->symbols_per_slot * N;
  int16_t *rx = malloc(4 * plane * nrx), *est = malloc(4 * plane * nrx * nl);
  if (!rx || !est || G == 0) { fprintf(stderr, "nrb200 shim: nr_rx_pusch_tp: configuration not served (G = %u)\n", G); abort(); }
  for (int a = 0; a < nrx; a++) memcpy(rx + 2 * plane * a, &gNB->common_vars.rxdataF[a][soffset], 4 * plane);
  for (int p = 0; p < nl * nrx; p++) memcpy(est + 2 * plane * p, pv->ul_ch_estimates[p], 4 * plane);
  int32_t shift = 0;
  const int rc = nrb200_pusch_inner_rx_host(&d, rx, est, pv->llr, &shift);
  if (rc != 0) { fprintf(stderr, "nrb200 shim: nrb200_pusch_inner_rx_host failed (rc = %d)\n", rc); abort(); }
  pv->log2_maxh = (int16_t)shift;
  free(rx); free(est);
  return 0;
}
