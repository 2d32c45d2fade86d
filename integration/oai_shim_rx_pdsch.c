/* Link-time interposer for the OAI UE: nr_rx_pdsch on the GPU with host C unchanged (see oai_shim_pusch_chest.c for how the interposers
 * are used).  Same prototype as openair1/PHY/NR_UE_TRANSPORT/nr_dlsch_demodulation.c:241-258, compiled against OAI's headers.
 *
 * nr_ue_pdsch_procedures calls the function once per PDSCH symbol (SCHED_NR_UE/phy_procedures_nr_ue.c:568-600).  The reference extracts, scales and
 * compensates each symbol as it is called and computes the LLRs of the WHOLE slot in the call for the last symbol (:576-616), from the buffers the earlier
 * calls left behind.  The B200 library does the slot in one go, so this interposer keeps the per-symbol bookkeeping the caller can see (dl_valid_re,
 * llr_offset) and runs the receiver when the last symbol arrives: by then rxdataF and dl_ch_estimates hold the whole slot.  It writes what the reference
 * writes for its caller: llr[0] (layer de-mapped, not yet unscrambled), dl_valid_re[], llr_offset[], *log2_maxh (here: with the last symbol's call, the
 * reference sets it at the first).  rxdataF_comp and ptrs_phase_per_slot are the reference's internal scratch and stay untouched.
 * PT-RS (pduBitmap bit 0 with a C-RNTI, one layer): dlsch[0].ptrs_symbols, ptrs_re_per_slot[][symbol] and the reduced dl_valid_re are kept per symbol like
 * nr_pdsch_ptrs_processing does (:569-574); estimation, interpolation and rotation happen inside the library's slot receiver.
 * Not served: PT-RS with two layers, two code words -- the call aborts loudly like an AssertFatal, there is no CPU fallback.
 * Test: tests/test_gpu_interpose.py drives it through the reference-side caller harness (oracle/ref_harness_pdsch.c) and compares with the pinned oracle. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "PHY/defs_nr_UE.h"
#include "PHY/NR_UE_TRANSPORT/nr_transport_proto_ue.h"
#include "nfapi_nr_interface.h"
#define NRB200_NO_OAI_LOADER_PROTOTYPES
#include "nrb200_ldpc.h"

int nr_rx_pdsch(PHY_VARS_NR_UE *ue, const UE_nr_rxtx_proc_t *proc, NR_UE_DLSCH_t dlsch[2], unsigned char symbol, unsigned char first_symbol_flag,
                unsigned char harq_pid, uint32_t pdsch_est_size, int32_t dl_ch_estimates[][pdsch_est_size], int16_t *llr[2],
                uint32_t dl_valid_re[NR_SYMBOLS_PER_SLOT], c16_t rxdataF[][ue->frame_parms.samples_per_slot_wCP], uint32_t llr_offset[NR_SYMBOLS_PER_SLOT],
                int32_t *log2_maxh, int rx_size_symbol, int nbRx, int32_t rxdataF_comp[][nbRx][rx_size_symbol * NR_SYMBOLS_PER_SLOT],
                c16_t ptrs_phase_per_slot[][NR_SYMBOLS_PER_SLOT], int32_t ptrs_re_per_slot[][NR_SYMBOLS_PER_SLOT])
{
  (void)first_symbol_flag; (void)rxdataF_comp; (void)ptrs_phase_per_slot;
  const NR_DL_FRAME_PARMS *fp = &ue->frame_parms;
  const fapi_nr_dl_config_dlsch_pdu_rel15_t *c = &dlsch[0].dlsch_config;
  const NR_DL_UE_HARQ_t *h0 = &ue->dl_harq_processes[0][harq_pid];
  if (h0->status != ACTIVE) { fprintf(stderr, "nrb200 shim: nr_rx_pdsch without an active DLSCH\n"); return -1; }
  if (NR_MAX_NB_LAYERS > 4 && ue->dl_harq_processes[1][harq_pid].status == ACTIVE) {
    fprintf(stderr, "nrb200 shim: nr_rx_pdsch with two code words is not served by libldpc_b200\n");
    abort();
  }
  const int N = fp->ofdm_symbol_size, nrx = fp->nb_antennas_rx, nl = dlsch[0].Nl, nb_rb = c->number_rbs, Qm = c->qamModOrder;
  nrb200_pusch_rx_t d;
  memset(&d, 0, sizeof(d));
  d.fft_size = N; d.nb_rx = nrx; d.rb_start = c->start_rb; d.bwp_start = c->BWPStart; d.rb_size = nb_rb; d.first_carrier_offset = fp->first_carrier_offset;
  d.qam_mod_order = Qm; d.start_symbol_index = c->start_symbol; d.nr_of_symbols = c->number_symbols; d.ul_dmrs_symb_pos = c->dlDmrsSymbPos;
  d.dmrs_config_type = c->dmrsConfigType == NFAPI_NR_DMRS_TYPE1 ? 0 : 1; d.num_dmrs_cdm_grps_no_data = c->n_dmrs_cdm_groups;
  d.log2_maxh = 0xFFFFFFFFu;                           /* measured by the library like the reference does at the first symbol */
  d.unscramble = 0; d.nrOfLayers = nl; d.pdsch_ue = 1;
  uint32_t ptrs_mask = 0, ptrs_n = 0;
  if ((c->pduBitmap & 0x1) && dlsch[0].rnti_type == TYPE_C_RNTI_) {
    d.ptrs = 1; d.rnti = dlsch[0].rnti; d.ptrs_time_density = c->PTRSTimeDensity; d.ptrs_freq_density = c->PTRSFreqDensity; d.ptrs_re_offset = c->PTRSReOffset;
    d.ptrs_slot = proc->nr_slot_rx; d.ptrs_nscid = c->nscid; d.ptrs_dmrs_scrambling_id = ue->scramblingID_dlsch[c->nscid];
    if (nrb200_pdsch_ptrs_layout(&d, &ptrs_mask, &ptrs_n) != 0) { fprintf(stderr, "nrb200 shim: nr_rx_pdsch: this PT-RS configuration is not served by libldpc_b200\n"); abort(); }
    dlsch[0].ptrs_symbols = (uint16_t)ptrs_mask;
  }
  /* ---- what the caller sees after every symbol (:393-404, :558): the symbol's number of PDSCH resource elements */
  const int pilots = (c->dlDmrsSymbPos >> symbol) & 1;
  const uint32_t nb_re = pilots ? (c->dmrsConfigType == NFAPI_NR_DMRS_TYPE1 ? nb_rb * (12 - 6 * c->n_dmrs_cdm_groups) : nb_rb * (12 - 4 * c->n_dmrs_cdm_groups))
                                : (uint32_t)nb_rb * 12;
  dl_valid_re[symbol - 1] = nb_re;
  if (d.ptrs) {                                        /* :569-574 */
    const int32_t n = ((ptrs_mask >> symbol) & 1) ? (int32_t)ptrs_n : 0;
    for (int a = 0; a < nrx; a++) ptrs_re_per_slot[a][symbol] = n;
    dlsch[0].ptrs_symbol_index = n ? symbol : 0;
    dl_valid_re[symbol - 1] -= n;
  }
  const int first = c->start_symbol, last = c->start_symbol + c->number_symbols - 1;
  if (symbol != last) return 0;
  /* ---- last symbol: llr_offset as nr_dlsch_llr leaves it (:1932-1936), then the slot's receiver */
  for (int i = first; i <= last; i++) {
    if (i == first && i < 3) llr_offset[i - 1] = 0;
    llr_offset[i] = dl_valid_re[i - 1] * Qm + llr_offset[i - 1];
  }
  const size_t plane = (size_t)14 * N;
  int16_t *rx = malloc(4 * plane * nrx), *est = malloc(4 * plane * nrx * nl);
  const uint32_t G = nrb200_pusch_num_llr(&d);
  int16_t *out = malloc(2 * (size_t)G + 64);
  if (!rx || !est || !out || G == 0) { fprintf(stderr, "nrb200 shim: nr_rx_pdsch: configuration not served (G = %u)\n", G); abort(); }
  for (int a = 0; a < nrx; a++) memcpy(rx + 2 * plane * a, &rxdataF[a][0], 4 * plane);
  for (int p = 0; p < nl * nrx; p++) memcpy(est + 2 * plane * p, &dl_ch_estimates[p][0], 4 * plane);
  int32_t shift = 0;
  const int rc = nrb200_pusch_inner_rx_host(&d, rx, est, out, &shift);
  if (rc != 0) { fprintf(stderr, "nrb200 shim: nrb200_pusch_inner_rx_host failed (rc = %d)\n", rc); abort(); }
  memcpy(llr[0], out, 2 * (size_t)(G < (uint32_t)h0->G ? G : (uint32_t)h0->G));
  *log2_maxh = shift;
  free(rx); free(est); free(out);
  return 0;
}
