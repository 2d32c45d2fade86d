/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the OpenAirInterface NR LDPC hot path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may link or
 * load this.  The product (libldpc_b200.so) never does.
 *
 * Parity status: PINNED.  Every function below is differential-tested bit-exactly against the unmodified
 * reference compiled by oracle/build_ref.sh (oracle/_ref/libref_*.so) in tests/test_oracle_vs_reference.py,
 * and against the committed fixtures in tests/golden/ (generated from that compiled reference by
 * tools/gen_golden.py) on machines where /root/reference is absent.
 */
#ifndef NRB200_ORACLE_H
#define NRB200_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* reference-defect emulation, see nrb200_oracle.c (bit 0: AVX2 BG2 R15 generator defect) */
void orc_set_quirks(int q);

/* lifting-set index iLS (0..7) of Z, -1 if Z is not one of the 51 NR lifting sizes */
int orc_ils_of_z(int Z);
/* columns the decoder uses for rate selector R: BG1 13->68 23->35 89->27; BG2 15->52 13->32 23->17; -1 if invalid */
int orc_ncols_for_rate(int BG, int R);

/* MSB-first bitwise CRC, result left-aligned in 32 bits like crc_byte.c:148-312.
 * poly_id: 0=24A 1=24B 2=24C 3=16 4=12 5=11 6=8 7=6 */
uint32_t orc_crc(int poly_id, const uint8_t *data, uint32_t bitlen);
/* crc_byte.c:314-379 */
int orc_check_crc(const uint8_t *decoded_bytes, uint32_t n, int crc_type);

/* nrLDPC_decoder.c:172-881 flooding int8 min-sum.  out sized per outMode (BIT: ncols*Z/8).
 * use_crc=0 <=> check_crc==NULL.  abort_in != 0 models a peer segment having set decode_abort_t before the call.
 * Returns what LDPCdecoder returns. */
int orc_ldpc_decode(int BG, int Z, int R, int numMaxIter, int outMode, const int8_t *llr, int8_t *out, int use_crc,
                    uint32_t crc_len_bits, int crc_type, int abort_in);

/* ldpc_encoder_optim8segmulti.c:46-212 for one segment: in = K/8 packed bytes MSB-first, out = 66Z|50Z bytes of 0/1. */
int orc_ldpc_encode(int BG, int Z, int K, const uint8_t *in, uint8_t *out);

/* nr_segmentation.c:32-180; outputs as the reference; seg_out[r] must hold K/8 bytes when non-NULL. Returns Kb or -1. */
int orc_segmentation(const uint8_t *in, uint8_t **seg_out, unsigned B, unsigned *C, unsigned *K, unsigned *Zout, unsigned *F, int BG);

/* nr_rate_matching.c:424-505 */
int orc_rate_matching_tx(uint32_t Tbslbrm, int BG, int Z, const uint8_t *w, uint8_t *e, int C, uint32_t F, uint32_t Foffset, int rv,
                         uint32_t E);
/* nr_rate_matching.c:507-603 */
int orc_rate_matching_rx(uint32_t Tbslbrm, int BG, int Z, int16_t *w, const int16_t *soft, int C, int rv, int clear, uint32_t E,
                         uint32_t F, uint32_t Foffset);
/* nr_rate_matching.c:36-305 / 310-388 (f[i+j*Qm] = e[i*E/Qm + j]) */
void orc_interleave(uint32_t E, int Qm, const uint8_t *e, uint8_t *f);
void orc_deinterleave(uint32_t E, int Qm, int16_t *e, const int16_t *f);
/* nr_rate_matching.c:390-422 */
int orc_get_R_ldpc_decoder(int rv, int E, int BG, int Z, int *llrLen, int round);

/* Gold sequence words, nr_codeword_scrambling / _unscrambling, nr_modulation */
void orc_gold_words(uint32_t c_init, uint32_t n_words, uint32_t *out);
void orc_scramble(const uint8_t *in, uint32_t size, uint32_t q, uint32_t Nid, uint32_t n_RNTI, uint32_t *out);
void orc_unscramble_llr(int16_t *llr, uint32_t size, uint32_t q, uint32_t Nid, uint32_t n_RNTI);
void orc_modulate(const uint8_t *bits, uint32_t length, int Qm, int16_t *out);

/* nr_ulsch_llr_computation.c:45-312: single-layer max-log LLRs for Qm = 2, 4, 6, 8 */
void orc_ulsch_llr(int Qm, const int16_t *rxF, const int16_t *maga, const int16_t *magb, const int16_t *magc, int16_t *out, uint32_t nb_re);

/* Q15 DFT/IDFT of the OFDM sizes (nrb200_dft_oracle.c restates openair1/PHY/TOOLS/oai_dfts.c); interleaved {re,im} int16. */
int orc_dft(int N, int inverse, const int16_t *in, int16_t *out, int scale);

/* PUSCH channel estimation, DMRS type 1, frequency-domain interpolation (nrb200_chest_oracle.c) */
typedef struct {
  int32_t fft_size, nb_rx, slot, symbol, port, rb_start, bwp_start, rb_size, first_carrier_offset, scid, dmrs_scrambling_id;
  int32_t dmrs_type;    /* 0: DMRS configuration type 1, 1: type 2 (pusch_dmrs_type_t) */
  int32_t chest_freq;   /* gNB->chest_freq: 0 = frequency-domain interpolation, 1 = one average per PRB */
} orc_chest_t;
void orc_pusch_dmrs_pilots(const orc_chest_t *p, int16_t *pil);
int orc_pusch_channel_estimation(const orc_chest_t *p, const int16_t *rxdataF, int16_t *ul_ch_est, int32_t *out);
int orc_pdsch_channel_estimation(const orc_chest_t *p, const int16_t *rxdataF, int16_t *dl_ch_est);
int orc_chest_time_domain_avg(int N, int nb_rx, int num_symbols, int start_symbol, int dmrs_bitmap, int num_rbs, int16_t *est);

/* single-layer PUSCH inner receiver (nrb200_pusch_oracle.c) */
typedef struct {
  int32_t fft_size, nb_rx, rb_start, bwp_start, rb_size, first_carrier_offset, Qm, ul_dmrs_symb_pos, dmrs_config_type, num_dmrs_cdm_grps_no_data;
} orc_pusch_t;
int orc_pusch_nb_re(const orc_pusch_t *p, int symbol);
int orc_pusch_extract(const orc_pusch_t *p, int is_dmrs_symbol, const int16_t *rxF, const int16_t *ch, int16_t *rx_ext, int16_t *ch_ext);
int orc_pusch_log2_maxh(const orc_pusch_t *p, int meas_symbol, int ch_symbol, const int16_t *rxdataF, const int16_t *ch_est, int32_t *avg_out);
int orc_pdsch_rx_slot(const orc_pusch_t *p, int start_symbol, int nr_symbols, const int16_t *rxdataF, const int16_t *dl_ch_est, int16_t *llr, int32_t *log2_maxh_out);
/* PT-RS at the UE: on, L = PTRSTimeDensity (log2), K = PTRSFreqDensity, re_offset = PTRSReOffset (used as k_RE_ref), rnti, slot = nr_slot_rx, nscid, nid = scramblingID_dlsch */
typedef struct { int32_t on, L, K, re_offset, rnti, slot, nscid, nid; } orc_ptrs_t;
uint32_t orc_ptrs_symbols(int start_symbol, int duration, int L_ptrs, uint32_t dmrs_pos);
int orc_ptrs_process_slot(uint32_t dmrs, uint32_t ptrs, int16_t *est, int start, int nsym);
int orc_pdsch_rx_slot_ptrs(const orc_pusch_t *p, const orc_ptrs_t *t, int start_symbol, int nr_symbols, const int16_t *rxdataF, const int16_t *dl_ch_est, int16_t *llr,
                           int32_t *log2_maxh_out, int16_t *phase_out, int32_t *ptrs_re_out);
int orc_pdsch_rx_slot_nl(const orc_pusch_t *p, int NL, int start_symbol, int nr_symbols, const int16_t *rxdataF, const int16_t *dl_ch_est, int16_t *llr, int32_t *log2_maxh_out);
int orc_pdsch_rx_slot_2l(const orc_pusch_t *p, int start_symbol, int nr_symbols, const int16_t *rxdataF, const int16_t *dl_ch_est, int16_t *llr, int32_t *log2_maxh_out);
int orc_pusch_log2_maxh_2l(const orc_pusch_t *p, int meas_symbol, int ch_symbol, int max_ch, const int16_t *rxdataF, const int16_t *ch_est, int32_t *avg_out);
int orc_pusch_inner_rx_symbol_2l(const orc_pusch_t *p, int symbol, int ch_symbol, int shift, uint32_t nvar, const int16_t *rxdataF, const int16_t *ch_est,
                                 int16_t *llr, int16_t *comp_out);
int orc_pusch_inner_rx_symbol(const orc_pusch_t *p, int symbol, int ch_symbol, int output_shift, const int16_t *rxdataF, const int16_t *ch_est,
                              int16_t *llr, int16_t *comp_out);

/* slot-level OFDM front end (nrb200_ofdm_oracle.c) */
void orc_rotate_cpx_vector(const int16_t *x, int16_t ar, int16_t ai, int16_t *y, uint32_t N);
void orc_mult_cpx_vector(const int16_t *x1, const int16_t *x2, int16_t *y, uint32_t N);
void orc_symbol_rotation(int mu, double f0, int16_t *rot);
void orc_timeshift_rotation(int N, int sample_offset, int16_t *out);
void orc_ofdm_geometry(int N, int mu, int slot, uint32_t *prefix, uint32_t *cp_start, uint32_t *slot_start, uint32_t *frame_len);
void orc_ofdm_tx_slot(int N, int mu, int nb_rb, int slot, int nsymb, const int16_t *rot, int16_t *txdataF, int16_t *txdata);
void orc_ofdm_rx_slot(int N, int mu, int nb_rb, int slot, int divisor, int sample_offset, const int16_t *rot, const int16_t *rxdata, int16_t *rxdataF);

/* gNB PDSCH transmitter after the encoder (nrb200_pdschtx_oracle.c): bits (one per byte, G of them) -> txdataF [nb_tx][14][fft_size] c16; only the
 * allocation's REs of the PDSCH symbols are written.  Returns G (= the length nr_generate_pdsch derives) or < 0. */
typedef struct {
  int32_t fft_size, nb_tx, slot, rb_start, bwp_start, rb_size, first_carrier_offset, Qm, nrOfLayers, start_symbol, nr_of_symbols, dl_dmrs_symb_pos,
          dmrs_config_type, num_dmrs_cdm_grps_no_data, dmrs_ports, scid, dl_dmrs_scrambling_id, data_scrambling_id, rnti, amp;
  int32_t pm_idx;               /* 0: identity precoding; > 0: the wideband precoding matrix below (one PRG spanning the allocation) */
  int16_t pm_weights[4][4][2];  /* nfapi_nr_pm_pdu_t.weights[layer][antenna] {Re, Im} */
  int32_t ptrs_on, ptrs_L, ptrs_K, ptrs_re_offset;   /* PT-RS (pduBitmap & 1): PTRSTimeDensity (log2), PTRSFreqDensity, PTRSReOffset; the caller then supplies the reduced G */
} orc_pdsch_tx_t;
int orc_pdsch_tx_slot(const orc_pdsch_tx_t *p, const uint8_t *bits, int16_t *txdataF);

/* rfsimulator channel application (nrb200_rfsim_oracle.c): rxAddInput of radio/rfsimulator/apply_channelmod.c */
void orc_rfsim_rx_add_input(int nb_tx, int nb_rx, int channel_length, int channel_offset, double path_loss_dB, float noise_power_dB, const double *ch,
                            const int16_t *input_sig, int16_t *out, int rxAnt, int nbSamples, uint64_t TS, uint32_t CirSize, const double *noise);

/* gNB PRACH detector (nrb200_prach_oracle.c): rx_nr_prach of NR_TRANSPORT/nr_prach.c, unrestricted set */
int orc_db_fixed_times10(uint32_t x);
int orc_rx_nr_prach(int nb_rx, int short_sequence, int NCS, int prach_fmt, int mu, const int16_t *xu, const int16_t *rxsigF, int32_t *out3);

#ifdef __cplusplus
}
#endif
#endif
