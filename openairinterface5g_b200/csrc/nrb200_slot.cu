// Slot-level entry points (include/nrb200_slot.h): the per-stage launches of one PUSCH / PDSCH slot issued back to back on the caller's stream.
// The sequencing is the reference's (nr_ulsch_demodulation.c:1447-1700 -> nr_ulsch_decoding.c:320-470 -> phy_procedures_nr_gNB.c:271-300 on the gNB,
// phy_procedures_nr_ue.c:520-760 on the UE, nr_dlsch.c:56-583 + nr_ru_procedures.c:55-140 for the transmitter); every stage is the entry point its
// stand-alone tests pin against the oracle, so the chain adds no arithmetic of its own.
#include "../../include/nrb200_slot.h"
#include "nrb200_ctx.h"
#include <cstring>

#define NRB200_EXPORT extern "C" __attribute__((visibility("default")))
using namespace nrb200;

// the slot-level OFDM front end is compiled into this library too (dfts_internal.cu), with hidden visibility
extern "C" int32_t nrb200_ofdm_mod_slot_dev(const nrb200_ofdm_slot_t *d, const int16_t *d_txdataF, int16_t *d_txdata, void *stream);
extern "C" int32_t nrb200_ofdm_demod_slot_dev(const nrb200_ofdm_slot_t *d, const int16_t *d_rxdata, const int16_t *d_timeshift, int16_t *d_rxdataF, void *stream);

NRB200_EXPORT int32_t nrb200_sch_slot_rx_dev(const nrb200_sch_rx_slot_t *d, const nrb200_sch_rx_bufs_t *b, void *stream)
{
  if (!d || !b) return -4;
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  // ---- nr_fep_full / nr_slot_fep: FFT windows -> rxdataF (rotation + time shift fused into the transform)
  if ((rc = nrb200_ofdm_demod_slot_dev(&d->ofdm, b->d_rxdata, b->d_timeshift, b->d_rxdataF, stream)) != 0) return rc;
  ctx().launches++;
  // ---- channel estimation on every DMRS port of the PDU
  nrb200_pusch_rx_t rx = d->rx;
  if (!d->use_estimates) {
    if ((rc = nrb200_pusch_chest_dev(&d->chest, b->d_rxdataF, b->d_est, b->d_chest_scratch, b->d_chest_state, stream)) != 0) return rc;
    if (rx.nrOfLayers == 2 && !rx.pdsch_ue) {   // the MMSE receiver takes max_ch / nvar from the estimator's state on the device (:1470-1524)
      rx.d_est_state = (uint64_t)(uintptr_t)b->d_chest_state;
      rx.est_state_ports = 2;
    }
  }
  // ---- level measurement + inner receiver (compensation / MMSE / zero forcing, LLRs, layer de-mapping, unscrambling)
  if ((rc = nrb200_pusch_log2_maxh_dev(&rx, b->d_est, b->d_level, stream)) != 0) return rc;
  if ((rc = nrb200_pusch_inner_rx_dev(&rx, b->d_rxdataF, b->d_est, b->d_level + 8, b->d_llr16, stream)) != 0) return rc;
  // ---- de-interleaving + rate recovery + HARQ combining + decoder-input packing
  if ((rc = nrb200_ldpc_rm_rx_batch_dev(&d->rm, b->d_llr16, b->d_E, b->d_Eoff, b->d_harq, b->harq_stride, b->d_llr8, b->llr8_stride, stream)) != 0) return rc;
  // ---- LDPC decode with the per-segment CRC stop
  nrb200_ldpc_batch_desc_t dd;
  std::memset(&dd, 0, sizeof(dd));
  dd.BG = d->rm.BG; dd.Z = d->rm.Z; dd.R = d->R; dd.numMaxIter = d->numMaxIter; dd.outMode = NRB200_OUTMODE_BIT;
  dd.use_crc = 1; dd.crc_type = (uint8_t)d->seg_crc_type; dd.crc_len_bits = d->crc_len_bits; dd.latency_mode = d->latency_mode;
  dd.n_cb = d->rm.n_seg; dd.llr_stride = b->llr8_stride; dd.out_stride = b->hard_stride;
  if ((rc = nrb200_ldpc_decode_batch_dev(&dd, b->d_llr8, b->d_hard, b->d_iters, stream)) != 0) return rc;
  // ---- nr_postDecode: the segments' payload bytes back to back, then the transport block's own CRC
  if (cudaMemcpy2DAsync(b->d_tb, d->seg_payload_bytes, b->d_hard, b->hard_stride, d->seg_payload_bytes, d->rm.n_seg, cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
    ctx().set_error("slot rx: transport block assembly", cudaGetLastError());
    return -2;
  }
  const uint32_t tb_bits = d->A + d->tb_crc_bits;
  return nrb200_crc_batch_dev(d->tb_crc_bits == 24 ? 0 : 3, 1, b->d_tb, (tb_bits + 7) / 8, tb_bits, b->d_tbcrc, stream);
}

NRB200_EXPORT int32_t nrb200_pdsch_slot_tx_dev(const nrb200_pdsch_tx_slot_t *d, const nrb200_pdsch_tx_bufs_t *b, void *stream)
{
  if (!d || !b) return -4;
  int rc;
  // ---- nr_dlsch_encoding: TB CRC + segmentation + per-segment CRC, LDPC encode, rate matching + interleaving
  if ((rc = nrb200_tb_segment_dev(d->rm.BG, d->A, b->d_payload, b->d_segs, b->seg_stride, b->d_seg_scratch, stream)) != 0) return rc;
  if ((rc = nrb200_ldpc_encode_batch_dev(d->rm.BG, d->rm.Z, (int)d->K, d->rm.n_seg, b->d_segs, b->seg_stride, b->d_cw, b->cw_stride, stream)) != 0) return rc;
  if ((rc = nrb200_ldpc_rm_tx_batch_dev(&d->rm, b->d_cw, b->cw_stride, b->d_E, b->d_Eoff, b->d_f, stream)) != 0) return rc;
  // ---- nr_generate_pdsch after the encoder, in one launch: scrambling, modulation, layer mapping, DMRS, resource mapping, precoding
  if ((rc = nrb200_pdsch_tx_slot_dev(&d->tx, b->d_f, b->d_txdataF, stream)) != 0) return rc;
  // ---- nr_feptx0: rotation + IDFT + cyclic prefix
  if ((rc = nrb200_ofdm_mod_slot_dev(&d->ofdm, b->d_txdataF, b->d_txdata, stream)) != 0) return rc;
  ctx().launches++;
  return 0;
}
