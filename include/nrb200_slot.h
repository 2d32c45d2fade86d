/*
 * nrb200_slot.h -- slot-level entry points of libldpc_b200.so: a whole PUSCH / PDSCH slot per call, device resident, stream ordered.
 *
 * SURVEY.md 8(f) items 2 and 3.  One call replaces what the reference sequences per slot
 *   receive  (gNB PUSCH / UE PDSCH): nr_fep_full | nr_slot_fep (MODULATION/slot_fep_nr.c:37-332) -> nr_pusch_channel_estimation |
 *            nr_pdsch_channel_estimation per DMRS port -> nr_rx_pusch_tp | nr_rx_pdsch (level, compensation / MMSE / zero forcing, LLRs, layer
 *            de-mapping, unscrambling; NR_TRANSPORT/nr_ulsch_demodulation.c:1447-1700, NR_UE_TRANSPORT/nr_dlsch_demodulation.c:241-684) ->
 *            nr_ulsch_decoding | nr_dlsch_decoding (de-interleave, rate recovery, HARQ combine, LDPC decode with the CRC24B stop;
 *            nr_ulsch_decoding.c:320-470) -> nr_postDecode (segments -> transport block, TB CRC; SCHED_NR/phy_procedures_nr_gNB.c:271-300)
 *   transmit (gNB PDSCH): nr_dlsch_encoding (TB CRC, nr_segmentation, LDPC encode, rate matching + interleaving; nr_dlsch_coding.c:280-420) ->
 *            nr_generate_pdsch from scrambling to txdataF (nr_dlsch.c:56-583) -> nr_feptx0 (rotation + IDFT + CP, SCHED_NR/nr_ru_procedures.c:55-140)
 * with every intermediate (rxdataF, estimates, LLRs, soft buffers, code words) staying in device memory.  The descriptors are the per-stage descriptors
 * of nrb200_ldpc.h / nrb200_dfts.h, filled from the same OAI structures (NR_DL_FRAME_PARMS, nfapi_nr_pusch_pdu_t, nfapi_nr_dl_tti_pdsch_pdu_rel15_t) their
 * stand-alone entry points take; the buffers are caller-owned device memory (allocate once per UE / HARQ process like OAI's pusch_vars).
 * All launches go to `stream` in order; nothing synchronises, so a slot can be captured into a CUDA graph.  Returns 0 or the first stage's error.
 */
#ifndef NRB200_SLOT_H
#define NRB200_SLOT_H
#include "nrb200_ldpc.h"
#include "nrb200_dfts.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct nrb200_sch_rx_slot_s {
  nrb200_ofdm_slot_t ofdm;        /* the slot's FFT windows (RX form of the descriptor) */
  nrb200_pusch_chest_t chest;     /* every DMRS port of the PDU in one call (n_ports); pdsch_ue selects the UE's estimator */
  nrb200_pusch_rx_t rx;           /* inner receiver; its log2_maxh is measured on the device, max_ch / nvar are read from the estimator's state.
                                   * DFT-s-OFDM: chest.transform_precoding + chest.lowpapr_seq (device) and rx.transform_precoding + rx.d_tp_scratch */
  nrb200_rm_desc_t rm;            /* rate recovery of the C segments (n_seg = C) */
  uint8_t R, numMaxIter;          /* decoder LUT (nr_get_R_ldpc_decoder) and iteration cap */
  uint8_t use_estimates;          /* 1: skip channel estimation, d_est holds the caller's estimates */
  uint8_t latency_mode;           /* decoder: 1 = a cluster of SMs per code block (one slot alone); 0 = one CTA per block (several slots in flight) */
  uint32_t crc_len_bits;          /* K - F: what check_crc sees per segment (CRC24B stop; a single segment carries the TB's CRC24A / CRC16) */
  uint32_t seg_crc_type;          /* crc_type of the per-segment check (CRC24_B = 1 when C > 1, else the TB's type) */
  uint32_t A;                     /* transport block payload bits */
  uint32_t tb_crc_bits;           /* 24 (A > 3824) or 16 */
  uint32_t seg_payload_bytes;     /* (K' - L) / 8: bytes of each decoded segment that belong to the transport block */
} nrb200_sch_rx_slot_t;

typedef struct nrb200_sch_rx_bufs_s {
  const int16_t *d_rxdata;        /* time-domain samples, [n_ant][t_stride] c16 */
  const int16_t *d_timeshift;     /* fp->timeshift_symbol_rotation, fft_size c16 (NULL when ofdm.rotate == 0) */
  int16_t *d_rxdataF;             /* [n_ant][14 * fft_size] c16 */
  int16_t *d_est;                 /* [n_ports * nb_rx][14 * fft_size] c16 */
  void *d_chest_scratch;          /* nrb200_pusch_chest_scratch_bytes(&chest) */
  int32_t *d_chest_state;         /* 18 int32 per port */
  int32_t *d_level;               /* 9 int32 */
  int16_t *d_llr16;               /* G */
  const uint32_t *d_E, *d_Eoff;   /* per segment: rate-matched length and offset into d_llr16 */
  int16_t *d_harq;                /* [C][harq_stride] int16 soft buffers (kept across HARQ rounds by the caller) */
  int8_t *d_llr8;                 /* [C][llr8_stride] decoder input */
  uint8_t *d_hard;                /* [C][hard_stride] decoder output (BIT mode) */
  int32_t *d_iters;               /* [C] */
  uint8_t *d_tb;                  /* (A + tb_crc_bits) / 8 bytes: payload followed by its CRC */
  uint32_t *d_tbcrc;              /* 1 word: remainder of payload + CRC, 0 when the transport block is intact */
  uint32_t harq_stride, llr8_stride, hard_stride, reserved;
} nrb200_sch_rx_bufs_t;

/* one PUSCH (gNB) or PDSCH (UE) slot: samples in, transport block + per-segment iteration counts + TB CRC verdict out */
int32_t nrb200_sch_slot_rx_dev(const nrb200_sch_rx_slot_t *d, const nrb200_sch_rx_bufs_t *b, void *stream);

typedef struct nrb200_pdsch_tx_slot_s {
  nrb200_pdsch_tx_t tx;           /* scrambling ... resource mapping + precoding */
  nrb200_ofdm_slot_t ofdm;        /* TX form: rotation + IDFT + CP */
  nrb200_rm_desc_t rm;            /* rate matching of the C segments */
  uint32_t A;                     /* transport block payload bits */
  uint32_t K;                     /* segment size incl. filler */
} nrb200_pdsch_tx_slot_t;

typedef struct nrb200_pdsch_tx_bufs_s {
  const uint8_t *d_payload;       /* A / 8 bytes */
  uint8_t *d_segs;                /* [C][seg_stride] */
  uint32_t *d_seg_scratch;        /* 1 word (TB CRC) */
  uint8_t *d_cw;                  /* [C][cw_stride] code words, one bit per byte */
  const uint32_t *d_E, *d_Eoff;
  uint8_t *d_f;                   /* G rate-matched bits */
  int16_t *d_txdataF;             /* [nb_tx][14 * fft_size] c16, in/out (other channels' REs are kept) */
  int16_t *d_txdata;              /* [nb_tx][t_stride] c16 */
  uint32_t seg_stride, cw_stride;
} nrb200_pdsch_tx_bufs_t;

/* one PDSCH slot of the gNB: payload in, time-domain samples out */
int32_t nrb200_pdsch_slot_tx_dev(const nrb200_pdsch_tx_slot_t *d, const nrb200_pdsch_tx_bufs_t *b, void *stream);

/* ---- transport-block level, HOST buffers: what nr_ulsch_decoding does for one PUSCH (nr_ulsch_decoding.c:320-470 + nr_processULSegment :121-230) in ONE call.
 * The reference queues one job per code-block segment on its thread pool -- de-interleaving, rate recovery with HARQ combining into harq_process->d[r],
 * decoder-input packing, one blocking LDPCdecoder call with the CRC stop -- and the caller collects the C results.  Here the C segments are one rate-recovery
 * launch and one decode launch (a cluster of SMs per segment: a TB's 1..56 segments finish in the time of one), then the decoded segments are copied out.
 * The soft buffers are LIBRARY-OWNED DEVICE MEMORY keyed by `harq_key` (the address of the reference's NR_UL_gNB_HARQ_t is a natural key): they stay on the GPU
 * across HARQ rounds, like the accelerator-side HARQ memory of the reference's offload path (nrLDPC_decoder_offload.c:545-546); d_mirror returns a copy for
 * callers (and tests) that want to see them.  integration/oai_shim_ulsch_decoding.c is the interposer that calls this with OAI's structures.
 * The reference's abort flag (a failed segment stops its siblings early) is an optimisation of a block that is lost anyway and is not reproduced: every
 * segment is decoded, iters[r] is its own verdict. */
typedef struct nrb200_ulsch_tb_s {
  nrb200_rm_desc_t rm;            /* BG, Z, Qm, rv, C, Tbslbrm, F, K as nr_ulsch_decoding sets them; n_seg = C; `clear` is ignored (per segment below) */
  uint32_t numMaxIter;            /* ulsch->max_ldpc_iterations */
  uint32_t crc_type, crc_len_bits;/* crcType(C, A), lenWithCrc(C, A): the per-segment check_crc stop (:178-179) */
  uint64_t harq_key;              /* identifies the TB's soft buffers inside the library */
  uint32_t llr_pinned;            /* 1: ulsch_llr lies in page-locked memory (nrb200_host_register / cudaHostAlloc): the GPU reads it in place, no staging copy */
  uint32_t reserved;
} nrb200_ulsch_tb_t;
/* ulsch_llr: the PUSCH's G int16 LLRs (unscrambled, as nr_rx_pusch_tp leaves them); E[r]: nr_get_E per segment; R[r]: nr_get_R_ldpc_decoder per segment;
 * clear[r]: harq_process->d_to_be_cleared[r] (1 = new data).  c[r] receives K / 8 bytes when segment r decoded (iters[r] <= numMaxIter) and is left alone
 * otherwise, as in the reference; iters[r] = the decoder's return value.  d_mirror (optional): C pointers, each receives the segment's soft buffer
 * (66 Z | 50 Z int16) after combining.  Returns 0, or a negative error (-4 invalid arguments, -5 out of memory, -2 CUDA failure). */
int32_t nrb200_ulsch_decode_tb_host(const nrb200_ulsch_tb_t *d, const int16_t *ulsch_llr, const uint32_t *E, const uint8_t *R, const uint8_t *clear,
                                    uint8_t *const *c, int32_t *iters, int16_t *const *d_mirror);
/* page-locks (cudaHostRegister) / releases a caller-owned host buffer that lives as long as the process, e.g. OAI's pusch_vars->llr, so that transfers from it are
 * direct DMA; 0 on success.  The buffer must not be freed while registered. */
int32_t nrb200_host_register(void *p, uint64_t bytes);
int32_t nrb200_host_unregister(void *p);
/* frees the soft buffers of one key (free_gNB_ulsch) -- or of every key when harq_key == 0 */
int32_t nrb200_ulsch_harq_release(uint64_t harq_key);

#ifdef __cplusplus
}
#endif
#endif
