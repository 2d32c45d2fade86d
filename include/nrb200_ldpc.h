/*
 * nrb200_ldpc.h -- C ABI of libldpc_b200.so, the B200-native drop-in for OpenAirInterface's
 * loadable LDPC codec ("libldpc<_version>.so").
 *
 * Part 1 declares EXACTLY the four symbols OAI's loader dlsym()s
 *   (reference openair1/PHY/CODING/nrLDPC_load.c:46-71, prototypes nrLDPC_defs.h:68-87,
 *    nrLDPC_extern.h:27-44) with layout-compatible parameter structs, so the unmodified
 *   ldpctest / nr_dlsim / nr_ulsim / nr-softmodem binaries bind them with
 *   `--loader.ldpc.shlibversion _b200` (or `ldpctest -v _b200`).
 * Part 2 declares the batched extension entry points (BASELINE config 2 needs batch=1024; the
 *   per-code-block blocking ABI cannot express a batch).
 *
 * Plain C: pointers, sizes and PODs only.  No CUDA or torch types appear in any signature;
 * `stream` arguments are an opaque `void *` carrying a cudaStream_t (NULL = legacy default stream).
 */
#ifndef NRB200_LDPC_H
#define NRB200_LDPC_H

#include <stdint.h>
#include <stdbool.h>
#include <pthread.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------
 * Part 1: the OAI loadable-codec ABI
 * ---------------------------------------------------------------------------------------- */

/* reference common/utils/time_meas.h:61-75 (time_stats_t); only .total of the decoder stats and the
 * four encoder timers are ever touched by this library (and only when non-NULL). */
typedef struct nrb200_time_stats {
  long long in;
  long long diff;
  long long p_time;
  double diff_square;
  long long max;
  int trials;
  int meas_flag;
  char *meas_name;
  int meas_index;
  int meas_enabled;
  void *tpoolmsg;
  void *tstatptr;
} nrb200_time_stats_t;

/* reference nrLDPC_decoder/nrLDPC_types.h:115-127 (t_nrLDPC_time_stats) */
typedef struct nrb200_ldpc_time_stats {
  nrb200_time_stats_t llr2llrProcBuf, llr2CnProcBuf, cnProc, cnProcPc, bnProcPc, bnProc, cn2bnProcBuf, bn2cnProcBuf,
      llrRes2llrOut, llr2bit, total;
} nrb200_ldpc_time_stats_t;

/* reference nrLDPC_types.h:75-79 (e_nrLDPC_outMode) */
typedef enum nrb200_ldpc_outmode {
  NRB200_OUTMODE_BIT = 0,     /* numLLR/8 bytes, MSB-first packed hard bits (nrLDPC_bnProc.h:1344-1380) */
  NRB200_OUTMODE_BITINT8 = 1, /* one hard bit per int8 */
  NRB200_OUTMODE_LLRINT8 = 2  /* one a-posteriori LLR per int8 */
} nrb200_ldpc_outmode_t;

/* reference nrLDPC_types.h:84-97 (t_nrLDPC_dec_params); field order and types are ABI. */
typedef struct nrb200_ldpc_dec_params {
  uint8_t BG;         /* base graph 1|2 */
  uint16_t Z;         /* lifting size */
  uint8_t R;          /* decoder rate LUT selector: BG1 {13,23,89}, BG2 {15,13,23} */
  uint16_t F;         /* filler bits (offload convention only) */
  uint8_t Qm;         /* modulation order (offload convention only) */
  uint8_t rv;         /* redundancy version (offload convention only) */
  uint8_t numMaxIter; /* iteration cap */
  int E;              /* CPU convention: payload length in bits handed to check_crc; offload: rate-matched length */
  nrb200_ldpc_outmode_t outMode;
  int crc_type;       /* CRC24_A=0 CRC24_B=1 CRC16=2 CRC8=3 (coding_defs.h:33-36) */
  int (*check_crc)(uint8_t *decoded_bytes, uint32_t n, uint8_t crc_type); /* NULL => parity-check early stop */
  uint8_t setCombIn;  /* offload: combine with the stored HARQ soft buffer */
} nrb200_ldpc_dec_params_t;

/* reference openair1/PHY/defs_common.h:996-1027 (decode_abort_t) */
typedef struct nrb200_decode_abort {
  pthread_mutex_t mutex_failure;
  bool failed;
} nrb200_decode_abort_t;

/* reference nrLDPC_defs.h:40-66 (encoder_implemparams_t); field order and types are ABI. */
typedef struct nrb200_ldpc_enc_params {
  unsigned int n_segments;
  unsigned int macro_num; /* which group of 8 segments this call encodes */
  unsigned char gen_code;
  nrb200_time_stats_t *tinput, *tprep, *tparity, *toutput; /* each may be NULL */
  int Kr;
  uint32_t Kb;
  uint32_t Zc;
  void *harq;
  uint8_t BG;
  unsigned char *output;
  uint32_t K;
  uint32_t F;
  uint8_t Qm;
  uint32_t E;
  unsigned int G;
  uint8_t rv;
} nrb200_ldpc_enc_params_t;

/* A translation unit that also includes OAI's own nrLDPC_extern.h (an interposer compiled against OAI headers, integration/) defines
 * NRB200_NO_OAI_LOADER_PROTOTYPES: the four loader symbols are then declared by OAI, with OAI's layout-identical types. */
#ifndef NRB200_NO_OAI_LOADER_PROTOTYPES
/* replaces LDPCinit (nrLDPC_decoder.c:162): creates the CUDA context, per-thread streams, pinned staging and
 * uploads the lifted-graph tables.  Idempotent (ldpctest calls it per segment, ldpctest.c:326).
 * Returns 0; returns -1 (loader then AssertFatal()s, nrLDPC_load.c:68) when no CUDA device is usable --
 * there is NO CPU fallback. */
int32_t LDPCinit(void);
/* replaces LDPCshutdown (nrLDPC_decoder.c:167) */
int32_t LDPCshutdown(void);
/* replaces LDPCdecoder (nrLDPC_decoder.c:172-195), "CPU-compatible" convention: p_llr holds ncol(R)*Z int8 LLRs
 * (first 2Z punctured = 0, fillers = +127); p_out receives per outMode; returns iterations used,
 * > numMaxIter means failure (and *ab is set).  harq_pid/ulsch_id/C are ignored in this convention.
 * An internal failure (no device, invalid parameters, CUDA error) is reported the same way: numMaxIter + 1 with *ab set, cause in
 * nrb200_last_error() -- never a negative value, OAI's callers test `<= numMaxIter` only.
 * Blocking, re-entrant.  Low-latency path (csrc/nrb200_ll.cu): the LLRs are staged in mapped pinned memory that the kernel reads and
 * answers into directly (no copy engine, no stream synchronisation); callers that arrive while another caller is issuing a launch are
 * combined into the next launch; one code block runs on a thread-block cluster of up to 8 SMs; *ab is re-read while the call waits and
 * polled by the kernel once per iteration (check_abort, nrLDPC_decoder.c:557-560).  NRB200_LL=0 selects the copy-engine path. */
int32_t LDPCdecoder(nrb200_ldpc_dec_params_t *p_decParams, uint8_t harq_pid, uint8_t ulsch_id, uint8_t C, int8_t *p_llr,
                    int8_t *p_out, nrb200_ldpc_time_stats_t *p_profiler, nrb200_decode_abort_t *ab);
/* replaces LDPCencoder (nrLDPC_encoder/ldpc_encoder_optim8segmulti.c:46-212): encodes segments
 * 8*macro_num .. min(8*macro_num+8, n_segments) of input[] (K/8 packed bytes each, MSB first) into
 * output[] as one bit per byte, K-2Z systematic + parity = 66Z (BG1) / 50Z (BG2) bytes.  Returns 0. */
int32_t LDPCencoder(uint8_t **input, uint8_t **output, nrb200_ldpc_enc_params_t *impp);
#endif

/* ------------------------------------------------------------------------------------------
 * Part 2: batched extension (same arithmetic, many code blocks per launch)
 * ---------------------------------------------------------------------------------------- */

/* Shape of one homogeneous batch of code blocks. */
typedef struct nrb200_ldpc_batch_desc {
  uint8_t BG;
  uint16_t Z;
  uint8_t R;          /* decoder LUT selector as above */
  uint8_t numMaxIter;
  uint8_t outMode;    /* nrb200_ldpc_outmode_t */
  uint8_t crc_type;   /* used when use_crc != 0 */
  uint8_t use_crc;    /* 0: parity-check stop (check_crc == NULL); 1: in-kernel CRC stop with reference semantics */
  uint8_t latency_mode; /* 0: throughput -- one CTA per code block, the fewest SM-cycles per block (what a caller with many batches / slots in flight wants);
                         * 1: latency -- when the whole batch fits on the GPU at once, every code block gets a thread-block cluster of 2 / 4 / 8 SMs
                         * (<= 74 / 33 / 15 blocks): a lone slot's 28-52 code blocks finish sooner, at more SM-cycles per block */
  uint32_t crc_len_bits; /* the `E` the reference passes to check_crc (payload incl. CRC, bits) */
  uint32_t n_cb;      /* code blocks in the batch */
  uint32_t llr_stride; /* bytes between consecutive code blocks' LLR arrays (>= ncol(R)*Z) */
  uint32_t out_stride; /* bytes between consecutive outputs (BIT: >= ncol(R)*Z/8, else >= ncol(R)*Z) */
} nrb200_ldpc_batch_desc_t;

/* number of input LLRs per code block for (BG, Z, R): ncol(R)*Z (nrLDPC_init.h:58, nrLDPCdecoder_defs.h:53-84); -1 if invalid */
int32_t nrb200_ldpc_num_llr(int BG, int Z, int R);

/* Device-resident batch decode: d_llr/d_out/d_iters are device pointers; asynchronous on `stream`.
 * d_iters[i] receives what LDPCdecoder would return for block i.  Returns 0, or a negative error. */
int32_t nrb200_ldpc_decode_batch_dev(const nrb200_ldpc_batch_desc_t *desc, const int8_t *d_llr, uint8_t *d_out,
                                     int32_t *d_iters, void *stream);
/* Host-buffer batch decode: H2D (pinned staging if the buffers are not pinned), kernel, D2H; blocking. */
int32_t nrb200_ldpc_decode_batch_host(const nrb200_ldpc_batch_desc_t *desc, const int8_t *llr, uint8_t *out, int32_t *iters);
/* The same call in two halves, so that a caller can keep several batches in flight and the copies of one batch overlap the
 * kernels of another -- the enqueue / dequeue shape of the bbdev calls in the reference's T2 offload
 * (openair1/PHY/CODING/nrLDPC_decoder/nrLDPC_decoder_offload.c:1048-1100: rte_bbdev_enqueue_ldpc_dec_ops / dequeue).
 * submit() stages pageable buffers, enqueues every H2D copy, kernel and D2H copy and returns a ticket; llr / out / iters must stay
 * valid and untouched until wait(ticket), which blocks until that batch is complete and frees the ticket.  out and iters hold the
 * results only after wait() returned 0.  Every ticket must be waited for exactly once. */
int32_t nrb200_ldpc_decode_batch_host_submit(const nrb200_ldpc_batch_desc_t *desc, const int8_t *llr, uint8_t *out, int32_t *iters,
                                             void **ticket);
int32_t nrb200_ldpc_decode_batch_host_wait(void *ticket);

/* Host arithmetic only (no GPU needed): the work schedule of the packed decoder kernel for (BG, Z, R) with at most max_threads threads per CTA.
 * info (8 int32): threads per CTA, number of work lists, 1 if a list belongs to a warp (items = 32 words of a row or column) / 0 if to a bin of
 * Z / 4 threads (items = whole rows or columns), heaviest and mean check-node list, heaviest and mean bit-node list (modelled warp instructions),
 * 1 if every item appears exactly once.  Returns 0, or -4 when the packed kernel does not serve the configuration (Z not a multiple of 4). */
int32_t nrb200_ldpc_packed_schedule_info(int BG, int Z, int R, int max_threads, int32_t *info);

/* Batch encode: in = n_cb x K/8 packed bytes (stride in_stride), out = n_cb x (66Z|50Z) bytes, one bit per byte
 * (stride out_stride).  Same output as LDPCencoder per block. */
int32_t nrb200_ldpc_encode_batch_dev(int BG, int Z, int K, uint32_t n_cb, const uint8_t *d_in, uint32_t in_stride,
                                     uint8_t *d_out, uint32_t out_stride, void *stream);
int32_t nrb200_ldpc_encode_batch_host(int BG, int Z, int K, uint32_t n_cb, const uint8_t *in, uint32_t in_stride, uint8_t *out,
                                      uint32_t out_stride);

/* CRC of n_blk bit strings of bitlen bits each (MSB first, stride bytes apart): out[i] = reference crc24a/crc24b/
 * crc24c/crc16/crc12/crc11/crc8/crc6 value (left-aligned in 32 bits exactly as crc_byte.c:148-312 returns it).
 * poly_id: 0=24A 1=24B 2=24C 3=16 4=12 5=11 6=8 7=6. */
int32_t nrb200_crc_batch_dev(int poly_id, uint32_t n_blk, const uint8_t *d_in, uint32_t stride, uint32_t bitlen, uint32_t *d_out,
                             void *stream);
int32_t nrb200_crc_batch_host(int poly_id, uint32_t n_blk, const uint8_t *in, uint32_t stride, uint32_t bitlen, uint32_t *out);

/* Transport-block CRC attachment + code block segmentation on the device: what nr_dlsch_encoding does before the encoder (nr_dlsch_coding.c:300-336: crc24a
 * for A > 3824, else crc16) followed by nr_segmentation (nr_segmentation.c:32-180: equal payload pieces, CRC24B per segment when C > 1, zero filler bytes).
 * payload: A / 8 bytes (A a multiple of 8); segs: C rows of K / 8 bytes, seg_stride apart -- exactly the encoder's input.  d_scratch: 4 bytes on the device.
 * nrb200_tb_segment_parms is the scalar part (host arithmetic): out = {C, K, Z, F, Kprime, L}, returns Kb or -1. */
int32_t nrb200_tb_segment_parms(int BG, uint32_t A, uint32_t out[6]);
int32_t nrb200_tb_segment_dev(int BG, uint32_t A, const uint8_t *d_payload, uint8_t *d_segs, uint32_t seg_stride, uint32_t *d_scratch, void *stream);
int32_t nrb200_tb_segment_host(int BG, uint32_t A, const uint8_t *payload, uint8_t *segs, uint32_t seg_stride);

/* ------------------------------------------------------------------------------------------
 * Part 3: rate matching / interleaving around the codec (one transport block = n_seg code block segments per call)
 * ---------------------------------------------------------------------------------------- */

/* Parameters shared by all segments of a transport block (reference nr_rate_matching.c:424-603 argument lists). */
typedef struct nrb200_rm_desc {
  uint8_t BG;
  uint16_t Z;
  uint8_t Qm;        /* modulation order 2|4|6|8 */
  uint8_t rv;        /* redundancy version 0..3 */
  uint8_t clear;     /* RX: 1 = new data, zero the first Ncb soft values before accumulating (d_to_be_cleared) */
  uint32_t C;        /* number of segments of the TB (enters Nref = 3*Tbslbrm/(2C)) */
  uint32_t Tbslbrm;  /* 0 = no limited-buffer rate matching */
  uint32_t F;        /* filler bits per segment */
  uint32_t K;        /* segment length incl. fillers (22Z | 10Z); Foffset = K - F - 2Z */
  uint32_t n_seg;    /* segments handled by this call (normally == C) */
} nrb200_rm_desc_t;

/* TX: replaces nr_rate_matching_ldpc + nr_interleaving_ldpc (nr_dlsch_coding.c:204-245).  d = n_seg encoder outputs (one bit per
 * byte, 66Z|50Z each, stride d_stride); E[r] = rate-matched length of segment r; f receives the segments back to back
 * (offset of segment r = sum of E[0..r)), one bit per byte. */
int32_t nrb200_ldpc_rm_tx_batch_dev(const nrb200_rm_desc_t *desc, const uint8_t *d_d, uint32_t d_stride, const uint32_t *d_E,
                                    const uint32_t *d_foff, uint8_t *d_f, void *stream);
int32_t nrb200_ldpc_rm_tx_batch_host(const nrb200_rm_desc_t *desc, const uint8_t *d, uint32_t d_stride, const uint32_t *E, uint8_t *f);
/* RX: replaces nr_deinterleaving_ldpc + nr_rate_matching_ldpc_rx + the decoder-input packing of nr_ulsch_decoding.c:153-210.
 * soft = the segments' E[r] int16 LLRs back to back; harq = n_seg persistent int16 soft buffers (>= 66Z|50Z each, stride
 * harq_stride elements), updated in place; llr receives 68Z|52Z int8 decoder inputs per segment (stride llr_stride bytes). */
int32_t nrb200_ldpc_rm_rx_batch_dev(const nrb200_rm_desc_t *desc, const int16_t *d_soft, const uint32_t *d_E, const uint32_t *d_soff,
                                    int16_t *d_harq, uint32_t harq_stride, int8_t *d_llr, uint32_t llr_stride, void *stream);
int32_t nrb200_ldpc_rm_rx_batch_host(const nrb200_rm_desc_t *desc, const int16_t *soft, const uint32_t *E, int16_t *harq, uint32_t harq_stride,
                                     int8_t *llr, uint32_t llr_stride);

/* ------------------------------------------------------------------------------------------
 * Part 4: demodulation -- single-layer max-log LLRs (no plug-in boundary exists for this in OAI: a maintainer interposes
 * nr_ulsch_compute_llr / nr_dlsch_*_llr)
 * ---------------------------------------------------------------------------------------- */

/* replaces nr_ulsch_compute_llr (nr_ulsch_llr_computation.c:316-373): rxF = nb_re compensated symbols {re, im} int16, mag_a/b/c =
 * the |h|^2-scaled decision thresholds (unused planes may be NULL), llr = nb_re * Qm int16, Qm in {2, 4, 6, 8}.
 * All pointers 16-byte aligned for the _dev variant. */
int32_t nrb200_pusch_llr_dev(int Qm, uint32_t nb_re, const int16_t *d_rxF, const int16_t *d_mag_a, const int16_t *d_mag_b,
                             const int16_t *d_mag_c, int16_t *d_llr, void *stream);
int32_t nrb200_pusch_llr_host(int Qm, uint32_t nb_re, const int16_t *rxF, const int16_t *mag_a, const int16_t *mag_b, const int16_t *mag_c,
                              int16_t *llr);

/* ---- Part 5: scrambling and the QAM mapper ---------------------------------------------------------------------------
 * replaces nr_codeword_scrambling (nr_scrambling.c:30-51): in = `size` bits, one per byte (the rate matcher's f); out = ceil(size/32)
 * words, bit i of word w = in[32w+i] ^ c(32w+i), c = Gold sequence with c_init = (n_RNTI << 15) + (q << 14) + Nid. */
int32_t nrb200_scramble_dev(const uint8_t *d_in, uint32_t size, uint32_t q, uint32_t Nid, uint32_t n_RNTI, uint32_t *d_out, void *stream);
int32_t nrb200_scramble_host(const uint8_t *in, uint32_t size, uint32_t q, uint32_t Nid, uint32_t n_RNTI, uint32_t *out);
/* replaces nr_codeword_unscrambling (nr_scrambling.c:53-80): llr[i] *= (1 - 2 c(i)) in place (mullo_epi16: -32768 stays -32768) */
int32_t nrb200_unscramble_llr_dev(int16_t *d_llr, uint32_t size, uint32_t q, uint32_t Nid, uint32_t n_RNTI, void *stream);
int32_t nrb200_unscramble_llr_host(int16_t *llr, uint32_t size, uint32_t q, uint32_t Nid, uint32_t n_RNTI);
/* replaces nr_modulation (nr_modulation.c:115-244): bits packed LSB-first (what the scrambler writes; the device buffer must be
 * readable one byte past the last bit), length_bits / Qm symbols {re, im} int16 out. */
int32_t nrb200_modulate_dev(const uint32_t *d_bits, uint32_t length_bits, int Qm, int16_t *d_out, void *stream);
int32_t nrb200_modulate_host(const uint32_t *bits, uint32_t length_bits, int Qm, int16_t *out);

/* ---- Part 6: single-layer PUSCH inner receiver ------------------------------------------------------------------------
 * One launch per slot replaces, for nrOfLayers == 1 without PT-RS / transform precoding, the per-symbol chain
 *   nr_ulsch_extract_rbs -> nr_ulsch_channel_compensation (MRC over rx antennas, QAM magnitude thresholds) -> nr_ulsch_compute_llr
 *   -> descrambling                                   (nr_ulsch_demodulation.c inner_rx :1262-1384, nr_pusch_symbol_processing :1386-1436)
 * and nrb200_pusch_log2_maxh_* replaces the measurement that precedes it (nr_ulsch_scale_channel + nr_ulsch_channel_level + the
 * log2_maxh rule, :1595-1647).  Field names follow nfapi_nr_pusch_pdu_t / NR_DL_FRAME_PARMS.  rxdataF: [nb_rx][14][fft_size] c16 (the
 * slot's rxdataF, soffset applied by the caller); ul_ch_estimates: [nb_rx][14][fft_size] c16, the estimates of DMRS symbol s stored at
 * symbol s from index 0 for the first allocated sub-carrier (nr_pusch_channel_estimation's layout).  LLR order and the per-symbol offsets
 * are pusch_vars->llr_offset[] (:1659-1664); the symbol of the estimates is the latest DMRS symbol at or before the data symbol. */
typedef struct nrb200_pusch_rx_s {
  uint32_t fft_size, nb_rx;                 /* ofdm_symbol_size, nb_antennas_rx (1..8) */
  uint32_t rb_start, bwp_start, rb_size, first_carrier_offset;
  uint32_t qam_mod_order;                   /* 2 4 6 8 */
  uint32_t start_symbol_index, nr_of_symbols, ul_dmrs_symb_pos, dmrs_config_type /* 0 = type 1 */, num_dmrs_cdm_grps_no_data;
  uint32_t log2_maxh;                       /* compensation shift (ignored by _dev when d_log2_maxh != NULL) */
  uint32_t rx_stride, ch_stride;            /* _dev: c16 between antennas of rxdataF / ul_ch_estimates */
  uint32_t unscramble, rnti, data_scrambling_id;   /* unscramble != 0: LLRs are multiplied by 1 - 2 c(i), c_init = (rnti << 15) + id */
  uint32_t nrOfLayers;                      /* 0 or 1: one layer.  2: MMSE receiver (nr_ulsch_mmse_2layers), qam_mod_order >= 6, nb_rx 2 or 4;
                                             * ul_ch_estimates then holds [2 * nb_rx] planes, index layer * nb_rx + rx, LLRs are layer de-mapped */
  uint32_t noise_var;                       /* 2 layers: nvar of the channel estimator, added to the diagonal of H^H H */
  uint32_t max_ch;                          /* 2 layers: the estimator's max_ch (scales the level measurement, nr_ulsch_scale_channel) */
  uint32_t pdsch_ue;                        /* 1: the UE's PDSCH receiver instead (nr_rx_pdsch, NR_UE_TRANSPORT/nr_dlsch_demodulation.c:241-684): its own
                                             * extraction patterns, estimate scaling, saturating MRC, thresholds and log2_maxh rule; ul_dmrs_symb_pos = dlDmrsSymbPos,
                                             * num_dmrs_cdm_grps_no_data = n_dmrs_cdm_groups, the estimates' symbol = get_valid_dmrs_idx_for_channel_est; nb_rx <= 4.
                                             * nrOfLayers == 2, 3 or 4 (nb_rx >= 2, any qam_mod_order): per-layer MRC + nr_zero_forcing_rx (:1726-1869, with the
                                             * recursive fixed-point nr_determin / nr_matrix_inverse :1460-1610 for 3 and 4 layers) + layer de-mapping;
                                             * dl_ch_estimates holds [nrOfLayers * nb_rx] planes, index layer * nb_rx + rx */
  uint64_t d_est_state;                     /* _dev, 2 layers, optional: DEVICE address of the channel estimator's state (nrb200_pusch_chest_dev's d_state, 18 int32 per
                                             * port).  When non-zero, max_ch and noise_var are taken from it ON THE DEVICE (max over the ports' max_ch; sum of the ports'
                                             * nvar / (nr_of_symbols * nrOfLayers * nb_rx), nr_ulsch_demodulation.c:1470-1524) and the two fields above are ignored:
                                             * estimator -> level -> receiver then run stream ordered with no host round trip (and can be captured in a CUDA graph). */
  uint32_t est_state_ports;                 /* number of ports in d_est_state (1 or 2) */
  uint32_t transform_precoding;             /* 1: pusch_pdu->transform_precoding == transformPrecoder_enabled (DFT-s-OFDM).  For one layer and qam_mod_order <= 6 the
                                             * receiver then equalises the compensated symbol (nr_freq_equalization, Qm > 2) and takes the 12 * rb_size point transform
                                             * of nr_idft across it before the LLRs (nr_ulsch_demodulation.c:1326-1336, :16-265); two layers and 256QAM are untouched,
                                             * as in the reference.  Needs num_dmrs_cdm_grps_no_data = 2 (no data on DMRS symbols) and 12 * rb_size one of nr_idft's
                                             * sizes other than 768 and 2304 (the reference's own output is not reproducible there): anything else returns -4. */
  uint64_t d_tp_scratch;                    /* _dev, transform precoding: DEVICE scratch of nrb200_pusch_tp_scratch_bytes() bytes (the host entry point has its own) */
  uint32_t ptrs;                            /* 1: PT-RS present (pduBitmap & 1 with a C-RNTI), pdsch_ue = 1 and one layer only: nr_pdsch_ptrs_processing
                                             * (NR_UE_ESTIMATION/nr_dl_channel_estimation.c:1765-1907, called at nr_dlsch_demodulation.c:569-574) -- common phase error per
                                             * PT-RS symbol (nr_ptrs_cpe_estimation, NR_REFSIG/ptrs_nr.c:181-263), PT-RS REs removed from the LLR stream, interpolation over
                                             * the other symbols (nr_ptrs_process_slot :281-337), rotation of every non-DMRS symbol before the LLRs.  `rnti` above is
                                             * dlsch[0].rnti.  Any other combination (gNB side: DESIGN.md defect 19; two layers) returns -4. */
  uint32_t ptrs_time_density;               /* dlsch_config.PTRSTimeDensity: 0 1 2 (L_PTRS = 1 << value) */
  uint32_t ptrs_freq_density;               /* dlsch_config.PTRSFreqDensity: K_PTRS 2 or 4 */
  uint32_t ptrs_re_offset;                  /* dlsch_config.PTRSReOffset, used as k_RE_ref like the reference does (< 12) */
  uint32_t ptrs_slot, ptrs_nscid, ptrs_dmrs_scrambling_id;   /* proc->nr_slot_rx, dlsch_config.nscid, ue->scramblingID_dlsch[nscid]: the Gold sequence of nr_gold_pdsch */
  uint32_t ptrs_reserved;
  uint64_t d_ptrs_state;                    /* _dev: 128 bytes of DEVICE scratch; afterwards words [0..13] = ptrs_phase_per_slot[0] {re, im} packed, [14] = status */
} nrb200_pusch_rx_t;
uint64_t nrb200_pusch_tp_scratch_bytes(const nrb200_pusch_rx_t *d);
/* PT-RS bookkeeping of a descriptor with ptrs = 1: *ptrs_symbols = dlsch->ptrs_symbols as set_ptrs_symb_idx leaves it (NR_REFSIG/ptrs_nr.c:53-86), *ptrs_re_per_symbol =
 * ptrs_re_per_slot[0][l] of every PT-RS symbol (nr_ptrs_cpe_estimation's re_cnt).  0, or -4 for a PT-RS configuration the library does not reproduce. */
int32_t nrb200_pdsch_ptrs_layout(const nrb200_pusch_rx_t *d, uint32_t *ptrs_symbols, uint32_t *ptrs_re_per_symbol);
uint32_t nrb200_pusch_num_llr(const nrb200_pusch_rx_t *d);                     /* int16 LLRs the slot produces (G for one layer), 0 if invalid */
/* d_out: 9 int32 on the device: [0..min(8, nb_rx * layers)) = avg per (layer, antenna), [8] = log2_maxh.  The kernel is stream ordered: pass d_out + 8 as d_log2_maxh. */
int32_t nrb200_pusch_log2_maxh_dev(const nrb200_pusch_rx_t *d, const int16_t *d_ul_ch_estimates, int32_t *d_out, void *stream);
int32_t nrb200_pusch_inner_rx_dev(const nrb200_pusch_rx_t *d, const int16_t *d_rxdataF, const int16_t *d_ul_ch_estimates, const int32_t *d_log2_maxh,
                                  int16_t *d_llr, void *stream);
/* host buffers, contiguous [nb_rx][14][fft_size]; computes log2_maxh itself when d->log2_maxh == 0xFFFFFFFF and returns it in *log2_maxh_out */
int32_t nrb200_pusch_inner_rx_host(const nrb200_pusch_rx_t *d, const int16_t *rxdataF, const int16_t *ul_ch_estimates, int16_t *llr,
                                   int32_t *log2_maxh_out);

/* ---- Part 7: PUSCH channel estimation, DMRS configuration type 1, frequency-domain interpolation -------------------------
 * replaces nr_pusch_channel_estimation (nr_ul_channel_estimation.c:67-243) for one DMRS symbol and one antenna port with
 * transform precoding disabled and chest_freq == 0: DMRS generation, least-squares estimate, delay estimation (IDFT peak, running
 * maximum across the rx antennas), delay compensation, 16-tap interpolation, delay reversal, noise variance.  Configurations outside the
 * ones described here return -4: the library never falls back.  DMRS type 2, chest_freq == 1 and transform precoding (low-PAPR DMRS): see the last fields.
 * rxdataF [nb_rx][14][fft_size] c16; ul_ch_estimates [nb_rx][14][fft_size] c16: symbol `symbol` of every antenna is rewritten.
 * state (5 int32): max_ch, nvar, est_delay, delay_max_pos, delay_max_val -- what the reference returns through *max_ch, *nvar and delay_t. */
typedef struct nrb200_pusch_chest_s {
  uint32_t fft_size, nb_rx;
  uint32_t slot, symbol;                    /* Ns and l of the DMRS symbol */
  uint32_t port;                            /* antenna port p - 1000 (0..3): get_dmrs_port(nl, dmrs_ports) */
  uint32_t rb_start, bwp_start, rb_size, first_carrier_offset;
  uint32_t scid, ul_dmrs_scrambling_id;
  uint32_t rx_stride, ch_stride;            /* _dev: c16 between antennas */
  uint32_t n_ports;                         /* 0 or 1: port `port` only.  2: ports `port` and `port + 1` in the same launches; estimates of port q go to
                                             * plane (q * nb_rx + rx), state of port q to d_state + 18 q (state5 + 5 q for the host entry point) */
  uint32_t pdsch_ue;                        /* 1: the UE's PDSCH estimator instead (nr_pdsch_channel_estimation + NFAPI_NR_DMRS_TYPE1_linear_interp,
                                             * NR_UE_ESTIMATION/nr_dl_channel_estimation.c:1305-1385, 1614-1735): same DMRS, delay handling and filters, its own
                                             * least-squares arithmetic; max_ch and nvar are not produced (0).  rb_start + bwp_start = the PDSCH's rb_offset. */
  uint32_t dmrs_config_type;                /* pusch_pdu->dmrs_config_type: 0 = type 1, 1 = type 2 (nr_ul_channel_estimation.c:258-283) */
  uint32_t chest_freq;                      /* gNB->chest_freq: 0 = frequency-domain interpolation, 1 = one average per PRB (:285-460; NO_INTERP build).
                                             * The three variants (type 2, chest_freq 1 of either type) serve one port per call (n_ports <= 1), for the gNB estimator
                                             * and, with pdsch_ue = 1, for the UE's (NFAPI_NR_DMRS_TYPE2_linear_interp / TYPE1_average_prb / TYPE2_average_prb,
                                             * nr_dl_channel_estimation.c:1378-1612; type 2: ports 0..5).  chest_freq = 1 needs rb_size >= 2 and, for the gNB's type 2,
                                             * slot % 4 == 0: the reference reads slot-ring position 0 there.  Anything else returns -4. */
  uint32_t transform_precoding;             /* 1: pusch_pdu->transform_precoding == transformPrecoder_enabled: the estimator correlates with the conjugate of the
                                             * low-PAPR type-1 sequence below instead of the Gold-sequence DMRS, from element 0 whatever rb_start and with w = +1
                                             * whatever the port (nr_ul_channel_estimation.c:122-133, nr_dmrs_rx.c:258-300).  DMRS type 1, gNB only. */
  uint64_t lowpapr_seq;                     /* transform precoding: address of the 6 * rb_size c16 of gNB_dmrs_lowpaprtype1_sequence[u][v][index] -- DEVICE memory for
                                             * the _dev entry point, host memory for _host.  nrb200_lowpapr_sequence_host() computes it for 6 * rb_size >= 30. */
} nrb200_pusch_chest_t;
/* base sequence r_{u,v} of TS 38.211 5.2.2 as the reference generates it (ul_ref_seq_nr.c:55-196: double precision, floor(scaling * cos / sin)), n_re = 30 or
 * >= 36 elements, {re, im} int16 each.  The lengths 6, 12, 18 and 24 are look-ups in the specification's phi tables and stay with the caller (OAI builds all of
 * them at start-up: generate_lowpapr_typ1_refsig_sequences); returns -4 for them.  Host arithmetic, no GPU needed. */
int32_t nrb200_lowpapr_sequence_host(uint32_t u, uint32_t v, uint32_t n_re, uint32_t scaling, int16_t *seq);
/* the 6 * rb_size conjugated DMRS symbols {re, im} the estimator correlates with (nr_pusch_dmrs_rx output); host arithmetic, no GPU needed */
int32_t nrb200_pusch_dmrs_pilots_host(const nrb200_pusch_chest_t *d, int16_t *pilots);
uint64_t nrb200_pusch_chest_scratch_bytes(const nrb200_pusch_chest_t *d);
/* d_scratch: nrb200_pusch_chest_scratch_bytes() bytes; d_state: 18 int32 per port on the device, [0..5) of each as described above */
int32_t nrb200_pusch_chest_dev(const nrb200_pusch_chest_t *d, const int16_t *d_rxdataF, int16_t *d_ul_ch_estimates, void *d_scratch, int32_t *d_state,
                               void *stream);
int32_t nrb200_pusch_chest_host(const nrb200_pusch_chest_t *d, const int16_t *rxdataF, int16_t *ul_ch_estimates, int32_t *state5);
/* nr_chest_time_domain_avg (openair1/PHY/NR_REFSIG/dmrs_nr.c:343-417; called with gNB->chest_time == 1 from nr_rx_pusch_tp, nr_ulsch_demodulation.c:1527-1538,
 * and by the UE, SCHED_NR_UE/phy_procedures_nr_ue.c:560): the estimates of the allocation's DMRS symbols are averaged into the first DMRS symbol (first
 * 12 * rb_size entries of the symbol; saturating sums; / 2 and / 4 as arithmetic shifts, / 3 towards zero).  est: [nb_rx][14][fft_size] c16, planes ch_stride
 * c16 apart (_dev), rewritten in place; with several layers only layer 0's planes are averaged, as in the reference.  Returns the index of the first DMRS
 * symbol (what the caller then uses as pusch_vars->dmrs_symbol), or a negative error (-4: no DMRS symbol in the allocation or more than four). */
int32_t nrb200_chest_time_avg_dev(uint32_t fft_size, uint32_t nb_rx, uint32_t ch_stride, uint32_t start_symbol, uint32_t nr_of_symbols, uint32_t dmrs_symb_pos,
                                  uint32_t rb_size, int16_t *d_est, void *stream);
int32_t nrb200_chest_time_avg_host(uint32_t fft_size, uint32_t nb_rx, uint32_t start_symbol, uint32_t nr_of_symbols, uint32_t dmrs_symb_pos, uint32_t rb_size,
                                   int16_t *est);

/* ---- Part 8: the "offload" calling convention of the codec ABI -----------------------------------------------------------
 * OAI's second LDPC interface (ldpc_interface_offload, loaded as version "_t2" when --ldpc-offload-enable is given; reference
 * implementation nrLDPC_decoder_offload.c:1034-1145 in front of a T2 accelerator) uses the same four prototypes with different
 * semantics: the decoder receives the E raw int8 LLRs of ONE segment and does de-interleaving, rate recovery and HARQ combining
 * itself, keeping the soft buffers inside the library keyed by (ulsch_id, segment); the encoder returns E rate-matched, interleaved
 * bits.  libldpc_b200_t2.so exports LDPCinit/LDPCshutdown/LDPCdecoder/LDPCencoder with these semantics and forwards to the two entry
 * points below.  The arithmetic is this library's (the reference's arithmetic here lives in the accelerator): int16 HARQ accumulation as
 * nr_rate_matching_ldpc_rx, the bit-exact flooding decoder with parity-check stop, numMaxIter from the parameter block.
 * decode: p->{BG,Z,R,F,Qm,rv,E,numMaxIter,setCombIn}; llr = E int8; out = K/8 bytes (K = 22Z | 10Z); returns iterations, < 0 on error.
 * encode: impp->{BG,Zc,K,F,Qm,rv,E}; in = K/8 bytes; out = E bytes, one bit each. */
int32_t nrb200_ldpc_offload_init(void);      /* = LDPCinit of libldpc_b200.so under a name the shim can link to */
int32_t nrb200_ldpc_offload_release(void);   /* frees the library-owned soft buffers (LDPCshutdown of the _t2 module) */
int32_t nrb200_ldpc_offload_decode(const nrb200_ldpc_dec_params_t *p, uint8_t harq_pid, uint8_t ulsch_id, uint8_t r, const int8_t *llr, uint8_t *out);
int32_t nrb200_ldpc_offload_encode(const uint8_t *in, uint8_t *out, const nrb200_ldpc_enc_params_t *impp);

/* ---- Part 9: gNB PDSCH transmitter after the encoder -------------------------------------------------------------------
 * One launch replaces, for one code word on 1..4 layers (PT-RS and one wideband precoding matrix included: last fields), everything nr_generate_pdsch
 * (openair1/PHY/NR_TRANSPORT/nr_dlsch.c:56-583) does after nr_dlsch_encoding: nr_pdsch_codeword_scrambling (:160), nr_modulation (:175), nr_layer_mapping
 * (:192), DMRS generation and resource mapping (:236-478) and the copy into txdataF (:490-530).  Field names follow nfapi_nr_dl_tti_pdsch_pdu_rel15_t /
 * NR_DL_FRAME_PARMS / PHY_VARS_gNB.  f: the encoder's output, nrb200_pdsch_tx_num_bits() rate-matched and interleaved bits, one per byte (what
 * nrb200_ldpc_rm_tx_batch_* writes).  txdataF: [nb_tx][14][fft_size] c16 of the slot (txdataF_offset applied by the caller); only the allocation's REs of
 * the PDSCH symbols are written (antennas beyond the layers get zeros there).  DMRS ports 0..3 (type 1) / 0..5 (type 2) in CDM groups without data; the
 * amplitude quirks of the reference's resource mapping are reproduced.  Anything else returns -4. */
typedef struct nrb200_pdsch_tx_s {
  uint32_t fft_size, nb_tx;                 /* ofdm_symbol_size, nb_antennas_tx */
  uint32_t slot;                            /* slot in the frame (DMRS sequence) */
  uint32_t rb_start, bwp_start, rb_size, first_carrier_offset;
  uint32_t qam_mod_order, nrOfLayers;
  uint32_t start_symbol_index, nr_of_symbols, dl_dmrs_symb_pos, dmrs_config_type /* 0 = type 1 */, num_dmrs_cdm_grps_no_data;
  uint32_t dmrs_ports;                      /* bitmap; 0 = port 0 (DCI 1_0), layer l uses the l-th set bit (get_dmrs_port) */
  uint32_t scid, dl_dmrs_scrambling_id, data_scrambling_id, rnti;
  uint32_t amp;                             /* gNB->TX_AMP */
  uint32_t tx_stride;                       /* _dev: c16 between antennas of txdataF */
  uint32_t pm_idx;                          /* precodingAndBeamforming.prgs_list[0].pm_idx: 0 = identity (layer l -> antenna l, the others zeroed); > 0 = the
                                             * matrix below on every RB of the allocation (one PRG: the reference's nFAPI structure holds a single prgs_list entry),
                                             * nr_dlsch.c:536-590 with nr_layer_precoder_simd / nr_layer_precoder_cm; nb_tx <= 4 */
  int16_t pm_weights[4][4][2];              /* gNB_config.pmi_list.pmi_pdu[pm_idx - 1].weights[layer][antenna] {precoder_weight_Re, precoder_weight_Im} */
  uint32_t ptrs;                            /* 1: pduBitmap & 1 -- PT-RS (nr_dlsch.c:98-111, :287-352): on the symbols set_ptrs_symb_idx selects, the sub-carriers
                                             * is_ptrs_subcarrier selects (NR_REFSIG/ptrs_nr.c:53-129, `rnti` above) carry on EVERY layer the QPSK symbols of the first bits
                                             * of the symbol's DMRS Gold sequence, the data skip them, the symbol's scaling truncates, and the encoder delivers
                                             * harq->unav_res modulation symbols less per layer (nrb200_pdsch_tx_num_bits accounts for it) */
  uint32_t ptrs_time_density;               /* PTRSTimeDensity 0 1 2 (L_PTRS = 1 << value) */
  uint32_t ptrs_freq_density;               /* PTRSFreqDensity: K_PTRS 2 or 4 */
  uint32_t ptrs_re_offset;                  /* PTRSReOffset, used as k_RE_ref like the reference does (< 12) */
} nrb200_pdsch_tx_t;
uint32_t nrb200_pdsch_tx_num_bits(const nrb200_pdsch_tx_t *d);                 /* G = nb_re * Qm as nr_generate_pdsch derives it (minus the PT-RS REs), 0 if invalid */
int32_t nrb200_pdsch_tx_slot_dev(const nrb200_pdsch_tx_t *d, const uint8_t *d_f, int16_t *d_txdataF, void *stream);
/* host buffers; txdataF is contiguous [nb_tx][14][fft_size] and is read and written back (REs outside the allocation keep their values) */
int32_t nrb200_pdsch_tx_slot_host(const nrb200_pdsch_tx_t *d, const uint8_t *f, int16_t *txdataF);

/* Device in use / last CUDA error text (diagnostics; never NULL). */
int32_t nrb200_device_index(void);
const char *nrb200_last_error(void);
/* Kernel launches issued by this library since load (bench.py reports it as gpu_launches). */
uint64_t nrb200_launch_count(void);

/* ---- one process, several GPUs (OAI is one process; SURVEY 8e: code blocks are independent, no data-path collective) ----
 * The library keeps one context per visible device.  nrb200_set_device() selects the device the CALLING THREAD's following calls run on
 * (default: NRB200_DEVICE, else LOCAL_RANK, else 0); device pointers handed to the *_dev entry points must belong to it.
 * With NRB200_DEVICES=all|<n> in the environment the OAI-facing entry points spread their calls themselves: LDPCdecoder / LDPCencoder
 * round robin (the host owns the HARQ state in that convention), the offload convention by nrb200_sticky_device(ulsch_id, r) so that a
 * segment's library-owned soft buffer stays on one GPU across HARQ rounds. */
/* the constant tables of (BG, Z, R) as one byte blob (size returned; min(size, cap) bytes copied): what a multi-rank job broadcasts at init */
int32_t nrb200_ldpc_graph_blob(int BG, int Z, int R, void *out, uint32_t cap);
int32_t nrb200_device_count(void);
int32_t nrb200_set_device(int dev);
int32_t nrb200_sticky_device(uint32_t ulsch_id, uint32_t r, uint32_t n_dev);
/* nrb200_ldpc_decode_batch_host over devices 0 .. n_dev-1 of this process: contiguous shards, all devices' copies and kernels in flight together */
int32_t nrb200_ldpc_decode_batch_host_multi(const nrb200_ldpc_batch_desc_t *desc, const int8_t *llr, uint8_t *out, int32_t *iters, int32_t n_dev);
/* low-latency path statistics: kernel launches and code blocks so far (blocks / launches = callers combined per launch) */
void nrb200_ll_stats(uint64_t *launches, uint64_t *blocks);
/* where the low-latency calls spent their time, sums in nanoseconds: [0] staging (memcpy into the mapped row), [1] queue + kernel launch,
 * [2] waiting for the kernel's completion bytes, [3] copying the result out, [4] device-side time of the block (%globaltimer) */
void nrb200_ll_timing(uint64_t *out5);
/* debug: clock64() phase marks of the cluster decoder (64 per CTA of code block 0) when NRB200_CLUSTER_TIMERS=1; out512 = int64[8 * 64] */
int32_t nrb200_debug_cluster_marks(long long *out512);

#ifdef __cplusplus
}
#endif
#endif /* NRB200_LDPC_H */
