/* TEST INFRASTRUCTURE ONLY (see nrb200_oracle.h).  CPU restatement of the single-layer PUSCH inner receiver of the reference
 * (openair1/PHY/NR_TRANSPORT/nr_ulsch_demodulation.c): nr_ulsch_extract_rbs :279-380, nr_ulsch_scale_channel :382-414,
 * get_nb_re_pusch :416-432, nr_ulsch_channel_level :434-466, nr_ulsch_channel_compensation :468-578 (rho == NULL), the log2_maxh rule
 * of nr_rx_pusch_tp :1595-1647 and the per-symbol composition inner_rx :1262-1384 followed by nr_ulsch_compute_llr.
 * Pinned against the compiled reference (oracle/_ref/libref_pusch.so) by tests/test_oracle_vs_reference.py. */
#include <limits.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "nrb200_oracle.h"

static inline int16_t sat16(int32_t v) { return v > 32767 ? 32767 : v < -32768 ? -32768 : (int16_t)v; }
static inline int32_t wrap32(int64_t v) { return (int32_t)(uint32_t)(uint64_t)v; }
static inline int16_t wrap16(int32_t v) { return (int16_t)(uint16_t)(uint32_t)v; }
static inline int16_t mulhrs16(int a, int b) { return wrap16(((a * b) + 0x4000) >> 15); }
static int log2_approx_(uint32_t x) { int l = 0; for (int i = 0; i < 31; i++) if (x & (1u << i)) l = i + 1; return l; }
static int factor2_(uint32_t x) { int i; for (i = 0; i < 31; i++) if (x & (1u << i)) break; return i; }

int orc_pusch_nb_re(const orc_pusch_t *p, int symbol)
{
  if ((p->ul_dmrs_symb_pos >> symbol) & 1) {
    if (p->dmrs_config_type == 0) return p->rb_size * (12 - p->num_dmrs_cdm_grps_no_data * 6);
    return p->rb_size * (12 - p->num_dmrs_cdm_grps_no_data * 4);
  }
  return p->rb_size * 12;
}

/* rxF: one antenna's symbol (N c16), ch: the estimates of the DMRS symbol in use, stored from index 0 for the first allocated RE.
 * Returns the number of REs written (the caller's buffers are zero-padded to buffer_length). */
int orc_pusch_extract(const orc_pusch_t *p, int is_dmrs_symbol, const int16_t *rxF, const int16_t *ch, int16_t *rx_ext, int16_t *ch_ext)
{
  const int N = p->fft_size, nb_re = 12 * p->rb_size;
  const int start_re = (p->first_carrier_offset + (p->rb_start + p->bwp_start) * 12) % N;
  int n = 0;
#define PUT(ri, ci) do { rx_ext[2 * n] = rxF[2 * (ri)]; rx_ext[2 * n + 1] = rxF[2 * (ri) + 1]; ch_ext[2 * n] = ch[2 * (ci)]; ch_ext[2 * n + 1] = ch[2 * (ci) + 1]; n++; } while (0)
  if (!is_dmrs_symbol) {
    for (int i = 0; i < nb_re; i++) PUT((start_re + i) % N, i);
  } else if (p->dmrs_config_type == 0) {              /* type 1: delta is hard-wired to 0, the odd REs carry data */
    if (start_re + nb_re < N) {
      for (int idx = 1; idx < nb_re; idx += 2) PUT(start_re + idx, idx);
    } else {
      const int neg = N - start_re, pos = nb_re - neg;
      int idx, idx2;
      for (idx = 1; idx < neg; idx += 2) PUT(start_re + idx, idx);
      idx2 = idx;
      for (idx = 1; idx < pos; idx += 2, idx2 += 2) PUT(idx, idx2);
    }
  } else {                                            /* type 2: REs 0,1 of every 6 are pilots */
    if (start_re + nb_re < N) {
      for (int idx = 0; idx < nb_re; idx++) { if (idx % 6 == 0 || idx % 6 == 1) continue; PUT(idx, idx); }   /* sic: no start_re (:345-351) */
    } else {
      const int neg = N - start_re, pos = nb_re - neg;
      int idx, idx2;
      for (idx = 0; idx < neg; idx++) { if (idx % 6 == 0 || idx % 6 == 1) continue; PUT(start_re + idx, idx); }
      idx2 = idx;
      for (idx = 0; idx < pos; idx++, idx2++) { if (idx % 6 == 0 || idx % 6 == 1) continue; PUT(idx, idx2); }
    }
  }
#undef PUT
  return n;
}

/* ch_symbol = DMRS symbol whose estimates are used; rxdataF [nb_rx][14 N] c16; ch_est [nb_rx][14 N] c16.  max_ch only matters for 2 layers. */
int orc_pusch_log2_maxh(const orc_pusch_t *p, int meas_symbol, int ch_symbol, const int16_t *rxdataF, const int16_t *ch_est, int32_t *avg_out)
{
  const int N = p->fft_size;
  const int len = (orc_pusch_nb_re(p, meas_symbol) + 15) & ~15;
  const int cap = (p->rb_size * 12 + 15) & ~15;        /* the extraction may write more than `len` REs (type-2 DMRS with 2 CDM groups) */
  int16_t *rx = calloc(2 * (size_t)cap, 2), *ch = calloc(2 * (size_t)cap, 2);
  int avgs = 0;
  const int x = factor2_(len), y = len >> x;
  for (int a = 0; a < p->nb_rx; a++) {
    memset(rx, 0, 4 * (size_t)cap); memset(ch, 0, 4 * (size_t)cap);
    orc_pusch_extract(p, (p->ul_dmrs_symb_pos >> meas_symbol) & 1, rxdataF + 2 * ((size_t)a * 14 + meas_symbol) * N,
                      ch_est + 2 * ((size_t)a * 14 + ch_symbol) * N, rx, ch);
    int32_t lane[4] = {0, 0, 0, 0};
    for (int i = 0; i < (len >> 2) * 4; i++) {
      /* nr_ulsch_scale_channel with shift_ch_ext = 0: mulhi by 8192 then << 3 */
      const int16_t r = wrap16((((int32_t)ch[2 * i] * 8192) >> 16) << 3), im = wrap16((((int32_t)ch[2 * i + 1] * 8192) >> 16) << 3);
      lane[i & 3] = wrap32((int64_t)lane[i & 3] + (wrap32((int64_t)r * r + (int64_t)im * im) >> x));
    }
    const int32_t avg = wrap32((int64_t)lane[0] + lane[1] + lane[2] + lane[3]) / y;
    if (avg_out) avg_out[a] = avg;
    if (avg > avgs) avgs = avg;
  }
  free(rx); free(ch);
  int l = (log2_approx_((uint32_t)avgs) >> 1) + 1 + log2_approx_((uint32_t)p->nb_rx >> 2);
  return l < 0 ? 0 : l;
}

/* ---- transform precoding (DFT-s-OFDM), one layer, Qm <= 6: inner_rx :1326-1336 runs nr_freq_equalization (NR_ESTIMATION/nr_freq_equalization.c:37-71, Qm > 2)
 * and nr_idft (:16-265) on the compensated symbol before the LLRs.
 *   equalisation: per group of 4 REs, amp = the FIRST int16 of the group's magnitude vector (clipped to 4095), comp = mullo(comp, 4096 / amp) >> 3 for all 8
 *   int16 of the group (amp == 0: the table entry is 0), magnitudes replaced by the constants 324 (16QAM) / 316, 158 (64QAM);
 *   nr_idft: conj -> four-way dft(DFT_<M>) with the data in lane 0 -> conj; M = 12: scale 0 then mulhi(9459) << 1; M = 1536 / 3072: idft() in place, no conj.
 * M = 768 and 2304 reach the single-transform dft768 / the broken dft2304 with four-way data (uninitialised lanes mix in): not reproducible, rejected here. */
int orc_dft4(int N, const int16_t *in, int16_t *out, int scale);
int orc_dft(int N, int inverse, const int16_t *in, int16_t *out, int scale);
static int g_transform_precoding;
void orc_pusch_set_transform_precoding(int on) { g_transform_precoding = on; }
static int16_t conj16(int16_t v) { return wrap16(-(int32_t)v); }                 /* sign_epi16 by -1: -32768 stays */
int orc_nr_idft(int16_t *z, int M)
{
  if (M == 1536 || M == 3072) {
    int16_t *t = malloc(4 * (size_t)M);
    const int rc = orc_dft(M, 1, z, t, 1);
    if (rc == 0) memcpy(z, t, 4 * (size_t)M);
    free(t);
    return rc;
  }
  if (M == 768 || M == 2304 || M < 12) return -1;
  int16_t *in = calloc(16 * (size_t)M, 1), *out = calloc(16 * (size_t)M, 1);
  for (int i = 0; i < M; i++) { in[8 * i] = z[2 * i]; in[8 * i + 1] = conj16(z[2 * i + 1]); }
  const int rc = orc_dft4(M, in, out, M == 12 ? 0 : 1);
  if (rc == 0)
    for (int i = 0; i < M; i++) {
      int16_t r = out[8 * i], q = out[8 * i + 1];
      if (M == 12) { r = wrap16(((r * 9459) >> 16) << 1); q = wrap16(((q * 9459) >> 16) << 1); }
      z[2 * i] = r; z[2 * i + 1] = conj16(q);
    }
  free(in); free(out);
  return rc;
}
static void freq_equalization(int16_t *comp, int16_t *ma, int16_t *mb, int M, int Qm)
{
  for (int g = 0; g < (M >> 2); g++) {
    int amp = ma[8 * g];
    if (amp > 4095) amp = 4095;
    const int inv = amp > 0 ? 4096 / amp : 0;                                    /* amp < 0 indexes before nr_inv_ch[] in the reference: undefined there */
    for (int k = 0; k < 8; k++) {
      comp[8 * g + k] = (int16_t)(wrap16(comp[8 * g + k] * inv) >> 3);
      if (Qm == 4) ma[8 * g + k] = 324;
      else if (Qm == 6) { ma[8 * g + k] = 316; mb[8 * g + k] = 158; }
    }
  }
}

/* One symbol of inner_rx, one layer.  Outputs: llr (valid_re * Qm int16), comp/maga/magb/magc (buffer_length c16 each, may be NULL). */
int orc_pusch_inner_rx_symbol(const orc_pusch_t *p, int symbol, int ch_symbol, int output_shift, const int16_t *rxdataF, const int16_t *ch_est,
                              int16_t *llr, int16_t *comp_out)
{
  const int N = p->fft_size, blen = (p->rb_size * 12 + 15) & ~15, Qm = p->Qm;
  const int is_dmrs = (p->ul_dmrs_symb_pos >> symbol) & 1;
  const int valid = orc_pusch_nb_re(p, symbol);
  int16_t *rx = malloc(4 * (size_t)blen), *ch = malloc(4 * (size_t)blen);
  int16_t *comp = calloc(4 * (size_t)blen, 1), *ma = calloc(4 * (size_t)blen, 1), *mb = calloc(4 * (size_t)blen, 1), *mc = calloc(4 * (size_t)blen, 1);
  const int ampa = Qm == 4 ? 20724 /* QAM16_n1 */ : Qm == 6 ? 20225 /* QAM64_n1 */ : Qm == 8 ? 20106 /* QAM256_n1 */ : 0;
  const int ampb = Qm == 6 ? 10112 /* QAM64_n2 */ : Qm == 8 ? 10053 /* QAM256_n2 */ : 0;
  const int ampc = Qm == 8 ? 5026 /* QAM256_n3 */ : 0;
  for (int a = 0; a < p->nb_rx; a++) {
    memset(rx, 0, 4 * (size_t)blen); memset(ch, 0, 4 * (size_t)blen);
    orc_pusch_extract(p, is_dmrs, rxdataF + 2 * ((size_t)a * 14 + symbol) * N, ch_est + 2 * ((size_t)a * 14 + ch_symbol) * N, rx, ch);
    for (int i = 0; i < (blen >> 3) * 8; i++) {
      const int32_t hr = ch[2 * i], hi = ch[2 * i + 1], yr = rx[2 * i], yi = rx[2 * i + 1];
      const int32_t nhi = wrap16(-hi);                                             /* sign_epi16(.., -1) keeps -32768 */
      const int16_t cr = sat16(wrap32((int64_t)hr * yr + (int64_t)hi * yi) >> output_shift);
      const int16_t ci = sat16(wrap32((int64_t)nhi * yr + (int64_t)hr * yi) >> output_shift);
      const int16_t m = sat16(wrap32((int64_t)hr * hr + (int64_t)hi * hi) >> output_shift);
      comp[2 * i] = wrap16(comp[2 * i] + cr); comp[2 * i + 1] = wrap16(comp[2 * i + 1] + ci);      /* MRC: add_epi16 wraps */
      if (Qm > 2) { const int16_t v = mulhrs16(m, ampa); ma[2 * i] = wrap16(ma[2 * i] + v); ma[2 * i + 1] = wrap16(ma[2 * i + 1] + v); }
      if (Qm > 4) { const int16_t v = mulhrs16(m, ampb); mb[2 * i] = wrap16(mb[2 * i] + v); mb[2 * i + 1] = wrap16(mb[2 * i + 1] + v); }
      if (Qm > 6) { const int16_t v = mulhrs16(m, ampc); mc[2 * i] = wrap16(mc[2 * i] + v); mc[2 * i + 1] = wrap16(mc[2 * i + 1] + v); }
    }
  }
  int rc = valid;
  if (g_transform_precoding && Qm <= 6) {
    if (Qm > 2) freq_equalization(comp, ma, mb, valid, Qm);
    if (orc_nr_idft(comp, valid) != 0) rc = -1;
  }
  if (rc >= 0) orc_ulsch_llr(Qm, comp, ma, mb, mc, llr, (uint32_t)valid);
  if (comp_out) memcpy(comp_out, comp, 4 * (size_t)blen);
  free(rx); free(ch); free(comp); free(ma); free(mb); free(mc);
  return rc;
}

/* ---- two layers, MMSE receiver (Qm >= 6): nr_ulsch_channel_compensation per layer + nr_ulsch_mmse_2layers (:870-1260) with its helpers
 * nr_ulsch_conjch0_mult_ch1 :651-695, nr_ulsch_construct_HhH_elements :761-868, nr_ulsch_det_HhH :579-644, nr_ulsch_comp_muli_sum :697-759.
 * ch_est: [2 * nb_rx][14 N] c16, index layer * nb_rx + rx.  llr: [2][valid * Qm].  comp_out (optional): [2][buffer_length] c16 after MMSE.
 * Returns the number of valid REs, or -1 for configurations the reference itself rejects (nb_rx not 2 or 4). */
static inline int32_t abs32w(int32_t v) { return v == INT32_MIN ? v : (v < 0 ? -v : v); }
static int log2a(uint32_t x) { return log2_approx_(x); }

/* ---- joint max-log ML detector for two layers, Qm < 6 (nr_ulsch_compute_ML_llr :2100-2130 -> nr_ulsch_qpsk_qpsk :375-525, nr_ulsch_qam16_qam16 :903-1135;
 * the 128-bit code path, `#define USE_128BIT` :39).  Everything is element-wise int16 arithmetic; one resource element at a time here. */
static inline int16_t mulhi16(int a, int b) { return (int16_t)((a * b) >> 16); }
static inline int16_t sll16(int a, int n) { return wrap16(a << n); }
static inline int16_t adds16(int a, int b) { return sat16(a + b); }
static inline int16_t subs16(int a, int b) { return sat16(a - b); }
static inline int16_t abs16(int a) { return wrap16(a < 0 ? -a : a); }          /* abs_epi16: -32768 stays */
static inline int16_t max16(int a, int b) { return (int16_t)(a > b ? a : b); }

/* y0: matched-filter output of the wanted layer, y1: of the other layer, rho: sum over rx of conj(h_wanted) h_other.  out: 2 LLRs (before the >> 4) */
static void ml_qpsk_qpsk(const int16_t *y0, const int16_t *y1, const int16_t *rho, int16_t *out)
{
  const int16_t y0r2 = sll16(mulhi16(y0[0], 23170), 1), y0i2 = sll16(mulhi16(y0[1], 23170), 1);
  const int16_t y1r2 = (int16_t)(y1[0] >> 1), y1i2 = (int16_t)(y1[1] >> 1);
  const int16_t rho_p = mulhi16(adds16(rho[0], rho[1]), 23170), rho_m = mulhi16(subs16(rho[0], rho[1]), 23170);
  const int16_t rpm = abs16(subs16(rho_p, y1r2)), imm = abs16(subs16(rho_m, y1i2)), rmm = abs16(subs16(rho_m, y1r2)), ipm = abs16(subs16(rho_p, y1i2));
  const int16_t rpp = abs16(adds16(rho_p, y1r2)), imp = abs16(adds16(rho_m, y1i2)), rmp = abs16(adds16(rho_m, y1r2)), ipp = abs16(adds16(rho_p, y1i2));
  const int16_t num_re_p = adds16(adds16(adds16(rpm, imm), y0r2), y0i2);
  const int16_t num_re_m = subs16(adds16(adds16(rmm, ipp), y0r2), y0i2);
  const int16_t den_re_p = adds16(subs16(adds16(rmp, ipm), y0r2), y0i2);
  const int16_t den_re_m = subs16(subs16(adds16(rpp, imp), y0r2), y0i2);
  /* second bit: num = {+,+}, {-,+}; den = {+,-}, {-,-} */
  out[0] = subs16(max16(num_re_p, num_re_m), max16(den_re_p, den_re_m));
  out[1] = subs16(max16(num_re_p, den_re_p), max16(num_re_m, den_re_m));
}

/* mag0 / mag1: the real part of ul_ch_mag of the wanted / the other layer.  out: 4 LLRs */
static void ml_qam16_qam16(const int16_t *y0, const int16_t *y1, int16_t mag_des, int16_t mag_int, const int16_t *rho, int16_t *out)
{
  enum { C10 = 20724 /* 1/sqrt(10) Q16 */, C10Q15 = 10362, C3 = 31086 /* 3/sqrt(10) Q15 */, CS = 25905 /* sqrt(10)/4 Q15 */, C9 = 23315 /* 9/(2 sqrt(10)) Q14 */ };
  const int16_t rr = rho[0], ri = rho[1], rpi = adds16(rr, ri), rmi = subs16(rr, ri);
  int16_t rs[8], psr[16], psi[16], y0s[8], bm[16];
  rs[0] = mulhi16(rpi, C10); rs[4] = mulhi16(rmi, C10);
  rs[3] = sll16(mulhi16(rpi, C3), 1); rs[7] = sll16(mulhi16(rmi, C3), 1);
  const int16_t x4 = mulhi16(rr, C10), x5 = sll16(mulhi16(ri, C3), 1), x6 = sll16(mulhi16(rr, C3), 1), x7 = mulhi16(ri, C10);
  rs[1] = adds16(x4, x5); rs[5] = subs16(x4, x5); rs[2] = adds16(x6, x7); rs[6] = subs16(x6, x7);
  for (int j = 0; j < 8; j++) psr[j] = abs16(subs16(rs[j], y1[0]));
  for (int j = 8; j < 16; j++) psr[j] = abs16(adds16(rs[(j - 4) & 7], y1[0]));
  static const uint8_t idx[16] = {4, 6, 5, 7, 0, 2, 1, 3, 0, 2, 1, 3, 4, 6, 5, 7};
  for (int k = 0; k < 16; k += 8)
    for (int j = k; j < k + 4; j++) { psi[j] = abs16(subs16(rs[idx[j]], y1[1])); psi[j + 4] = abs16(adds16(rs[idx[j + 4]], y1[1])); }
  const int16_t y0r1 = mulhi16(y0[0], C10), y0i1 = mulhi16(y0[1], C10), y0r3 = sll16(mulhi16(y0[0], C3), 1), y0i3 = sll16(mulhi16(y0[1], C3), 1);
  y0s[0] = adds16(y0r1, y0i1); y0s[4] = subs16(y0r1, y0i1); y0s[1] = adds16(y0r1, y0i3); y0s[5] = subs16(y0r1, y0i3);
  y0s[2] = adds16(y0r3, y0i1); y0s[6] = subs16(y0r3, y0i1); y0s[3] = adds16(y0r3, y0i3); y0s[7] = subs16(y0r3, y0i3);
  const int16_t ch10 = mulhi16(mag_des, C10Q15), ch2 = sll16(mulhi16(mag_des, CS), 1), ch910 = sll16(mulhi16(mag_des, C9), 2);
  const int16_t cc[4] = {ch10, ch2, ch2, ch910};
  for (int j = 0; j < 16; j++) {
    const int16_t ar = psr[j] < mag_int ? C10Q15 : C3, ai = psi[j] < mag_int ? C10Q15 : C3;                      /* interference_abs_epi16 :731 */
    const int16_t psa = adds16(sll16(mulhi16(psr[j], ar), 1), sll16(mulhi16(psi[j], ai), 1));                    /* prodsum_psi_a_epi16 :721 */
    const int16_t sq_r = sll16(mulhi16(sll16(mulhi16(sll16(mulhi16(ar, ar), 1), CS), 1), mag_int), 1);           /* square_a_epi16 :741 */
    const int16_t sq_i = sll16(mulhi16(sll16(mulhi16(sll16(mulhi16(ai, ai), 1), CS), 1), mag_int), 1);
    const int16_t t = subs16(psa, adds16(sq_r, sq_i));
    if (j < 8) bm[j] = subs16(adds16(t, y0s[j]), cc[j & 3]);
    else bm[j] = subs16(subs16(t, y0s[(j + 4) & 7]), cc[j & 3]);     /* j = 8..11 -> y0s[4..7], j = 12..15 -> y0s[0..3] */
  }
#define MX8(a, b, c, d, e, f, g, h) max16(max16(max16(bm[a], bm[b]), max16(bm[c], bm[d])), max16(max16(bm[e], bm[f]), max16(bm[g], bm[h])))
  out[0] = subs16(MX8(0, 1, 2, 3, 4, 5, 6, 7), MX8(8, 9, 10, 11, 12, 13, 14, 15));
  out[1] = subs16(MX8(0, 1, 3, 2, 8, 9, 10, 11), MX8(4, 5, 6, 7, 12, 13, 14, 15));
  out[2] = subs16(MX8(0, 1, 4, 5, 8, 9, 12, 13), MX8(2, 3, 6, 7, 10, 11, 14, 15));
  out[3] = subs16(MX8(0, 2, 4, 6, 8, 10, 12, 14), MX8(1, 3, 5, 7, 9, 11, 13, 15));
#undef MX8
}

int orc_pusch_inner_rx_symbol_2l(const orc_pusch_t *p, int symbol, int ch_symbol, int shift, uint32_t nvar, const int16_t *rxdataF, const int16_t *ch_est,
                                 int16_t *llr, int16_t *comp_out)
{
  const int N = p->fft_size, blen = (p->rb_size * 12 + 15) & ~15, Qm = p->Qm, nrx = p->nb_rx;
  if (Qm >= 6 && nrx != 2 && nrx != 4) return -1;
  const int is_dmrs = (p->ul_dmrs_symb_pos >> symbol) & 1;
  const int valid = orc_pusch_nb_re(p, symbol);
  const size_t B = 2 * (size_t)blen;
  int16_t *rx = calloc(B * nrx, 2), *ch = calloc(B * nrx * 2, 2), *comp = calloc(B * 2, 2), *mag[3];
  for (int i = 0; i < 3; i++) mag[i] = calloc(B * 2, 2);
  for (int a = 0; a < nrx; a++)
    for (int l = 0; l < 2; l++)
      orc_pusch_extract(p, is_dmrs, rxdataF + 2 * ((size_t)a * 14 + symbol) * N, ch_est + 2 * ((size_t)(l * nrx + a) * 14 + ch_symbol) * N, rx + B * a,
                        ch + B * (l * nrx + a));
  /* matched filter per layer (MRC sum wraps) */
  for (int l = 0; l < 2; l++)
    for (int a = 0; a < nrx; a++)
      for (int i = 0; i < (blen >> 3) * 8; i++) {
        const int32_t hr = ch[B * (l * nrx + a) + 2 * i], hi = ch[B * (l * nrx + a) + 2 * i + 1], yr = rx[B * a + 2 * i], yi = rx[B * a + 2 * i + 1];
        const int32_t nhi = wrap16(-hi);
        comp[B * l + 2 * i] = wrap16(comp[B * l + 2 * i] + sat16(wrap32((int64_t)hr * yr + (int64_t)hi * yi) >> shift));
        comp[B * l + 2 * i + 1] = wrap16(comp[B * l + 2 * i + 1] + sat16(wrap32((int64_t)nhi * yr + (int64_t)hr * yi) >> shift));
      }
  if (Qm < 6) {
    /* ML path (:1349-1361): rho[l][1-l] = saturating sum over rx of conj(h_l) h_(1-l) >> shift, ul_ch_maga[l] = wrapping sum over rx of
     * mulhrs(sat(|h_l|^2 >> shift), QAM16_n1) (nr_ulsch_channel_compensation :516-573); QPSK LLRs are shifted right by 4 afterwards (nr_ulsch_shift_llr) */
    /* x86 builds run the 256-bit loops (USE_128BIT is defined for aarch64 only, :37-39): `for (i = 0; i < length >> 3; i += 2)`, 16 REs per pass, so when
     * length mod 16 is 1..7 (12 rb_size = 4 mod 16, i.e. rb_size = 3 mod 4) the last REs of the symbol are never written.  The reference leaves whatever the
     * LLR buffer held; zeros here (the harness clears it), DESIGN.md defect 14. */
    const int covered = 16 * (((valid >> 3) + 1) >> 1);
    for (int i = 0; i < valid; i++) {
      if (i >= covered) {
        for (int l = 0; l < 2; l++) memset(llr + (size_t)l * valid * Qm + (size_t)i * Qm, 0, 2 * (size_t)Qm);
        continue;
      }
      int16_t rho[2][2] = {{0, 0}, {0, 0}}, mg[2] = {0, 0};
      for (int a = 0; a < nrx; a++)
        for (int l = 0; l < 2; l++) {
          const int16_t *h0 = ch + B * (l * nrx + a) + 2 * i, *h1 = ch + B * ((1 - l) * nrx + a) + 2 * i;
          const int32_t nai = wrap16(-h0[1]);
          rho[l][0] = sat16((int32_t)rho[l][0] + sat16(wrap32((int64_t)h0[0] * h1[0] + (int64_t)h0[1] * h1[1]) >> shift));
          rho[l][1] = sat16((int32_t)rho[l][1] + sat16(wrap32((int64_t)nai * h1[0] + (int64_t)h0[0] * h1[1]) >> shift));
          mg[l] = wrap16(mg[l] + mulhrs16(sat16(wrap32((int64_t)h0[0] * h0[0] + (int64_t)h0[1] * h0[1]) >> shift), 20724));
        }
      for (int l = 0; l < 2; l++) {
        int16_t *o = llr + (size_t)l * valid * Qm + (size_t)i * Qm;
        if (Qm == 2) { ml_qpsk_qpsk(comp + B * l + 2 * i, comp + B * (1 - l) + 2 * i, rho[l], o); o[0] >>= 4; o[1] >>= 4; }
        else ml_qam16_qam16(comp + B * l + 2 * i, comp + B * (1 - l) + 2 * i, mg[l], mg[1 - l], rho[l], o);
      }
    }
    if (comp_out) memcpy(comp_out, comp, B * 2 * 2);
    free(rx); free(ch); free(comp);
    for (int i = 0; i < 3; i++) free(mag[i]);
    return valid;
  }
  const int ampv[3] = {Qm == 4 ? 20724 : Qm == 6 ? 20225 : Qm == 8 ? 20106 : 0, Qm == 6 ? 10112 : Qm == 8 ? 10053 : 0, Qm == 8 ? 5026 : 0};
  const int nb_rb_0 = valid / 12 + ((valid % 12) ? 1 : 0);
  for (int g = 0; g < 3 * nb_rb_0; g++) {
    int32_t det[4], af[4][4][2];                                        /* af[k][00,01,10,11][re,im] */
    for (int k = 0; k < 4; k++) {
      const int i = 4 * g + k;
      int16_t s[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
      for (int a = 0; a < nrx; a++) {
        const int16_t *h0 = ch + B * (0 * nrx + a) + 2 * i, *h1 = ch + B * (1 * nrx + a) + 2 * i;
        const int16_t *pair[4][2] = {{h0, h0}, {h0, h1}, {h1, h0}, {h1, h1}};   /* conj(first) * second: 00, 01, 10, 11 */
        for (int e = 0; e < 4; e++) {
          const int32_t ar = pair[e][0][0], ai = pair[e][0][1], br = pair[e][1][0], bi = pair[e][1][1];
          const int32_t nai = wrap16(-ai);
          const int16_t re = sat16(wrap32((int64_t)ar * br + (int64_t)ai * bi) >> shift), im = sat16(wrap32((int64_t)nai * br + (int64_t)ar * bi) >> shift);
          if (a == 0) { s[e][0] = re; s[e][1] = im; } else { s[e][0] = sat16((int32_t)s[e][0] + re); s[e][1] = sat16((int32_t)s[e][1] + im); }
        }
      }
      for (int e = 0; e < 4; e++) { af[k][e][0] = s[e][0]; af[k][e][1] = s[e][1]; }
      if (nvar != 0)                                                     /* add_epi32 on the packed {re, im} word: carries into im */
        for (int e = 0; e < 4; e += 3) {
          uint32_t w = ((uint32_t)(uint16_t)af[k][e][0]) | ((uint32_t)(uint16_t)af[k][e][1] << 16);
          w += nvar;
          af[k][e][0] = (int16_t)(w & 0xFFFF); af[k][e][1] = (int16_t)(w >> 16);
        }
      const int32_t ad = wrap32((int64_t)af[k][0][0] * af[k][3][0] + (int64_t)wrap16(-af[k][0][1]) * af[k][3][1]);
      const int32_t bc = wrap32((int64_t)af[k][1][0] * af[k][2][0] + (int64_t)wrap16(-af[k][1][1]) * af[k][2][1]);
      det[k] = abs32w(wrap32((int64_t)ad - bc));
    }
    int32_t sum_det = 0;
    for (int k = 0; k < 4; k++) sum_det = wrap32((int64_t)sum_det + (det[k] >> 2));
    const int b = log2a((uint32_t)sum_det) - 8;
    for (int k = 0; k < 4; k++) {
      const int i = 4 * g + k;
      const int32_t dm = b > 0 ? det[k] >> b : (int32_t)((uint32_t)det[k] << (-b));
      const int16_t m = sat16(dm);
      for (int l = 0; l < 2; l++)
        for (int t = 0; t < 3; t++) {
          const int16_t v = wrap16(((m * ampv[t]) >> 16) << 1);
          mag[t][B * l + 2 * i] = v; mag[t][B * l + 2 * i + 1] = v;
        }
      /* x0 = comp0 * d - comp1 * b ; x1 = comp1 * a - comp0 * c */
      const int32_t c0r = comp[2 * i], c0i = comp[2 * i + 1], c1r = comp[B + 2 * i], c1i = comp[B + 2 * i + 1];
      const int32_t in[2][4][2] = {{{c0r, c0i}, {af[k][3][0], af[k][3][1]}, {c1r, c1i}, {af[k][1][0], af[k][1][1]}},
                                   {{c1r, c1i}, {af[k][0][0], af[k][0][1]}, {c0r, c0i}, {af[k][2][0], af[k][2][1]}}};
      for (int l = 0; l < 2; l++) {
        const int32_t xr = in[l][0][0], xi = in[l][0][1], yr = in[l][1][0], yi = in[l][1][1], wr = in[l][2][0], wi = in[l][2][1], zr = in[l][3][0], zi = in[l][3][1];
        int32_t re = wrap32((int64_t)wrap32((int64_t)xr * yr + (int64_t)wrap16(-xi) * yi) - wrap32((int64_t)wr * zr + (int64_t)wrap16(-wi) * zi));
        int32_t im = wrap32((int64_t)wrap32((int64_t)xi * yr + (int64_t)xr * yi) - wrap32((int64_t)wi * zr + (int64_t)wr * zi));
        if (b > 0) { re >>= b; im >>= b; } else { re = (int32_t)((uint32_t)re << (-b)); im = (int32_t)((uint32_t)im << (-b)); }
        comp[B * l + 2 * i] = sat16(re); comp[B * l + 2 * i + 1] = sat16(im);
      }
    }
  }
  for (int l = 0; l < 2; l++) orc_ulsch_llr(Qm, comp + B * l, mag[0] + B * l, mag[1] + B * l, mag[2] + B * l, llr + (size_t)l * valid * Qm, (uint32_t)valid);
  if (comp_out) memcpy(comp_out, comp, B * 2 * 2);
  free(rx); free(ch); free(comp);
  for (int i = 0; i < 3; i++) free(mag[i]);
  return valid;
}

/* log2_maxh for two layers with the MMSE receiver (Qm >= 6): the estimates of both layers are scaled by shift_ch_ext = log2_approx(max_ch >> 11)
 * (nr_ulsch_scale_channel :382-414) before the level, the rule is (log2_approx(max avg) >> 1) - 3, floored at 0 (:1640-1647).
 * ch_est [2 * nb_rx][14 N]; avg_out (optional) 2 * nb_rx values, index layer * nb_rx + rx. */
int orc_pusch_log2_maxh_2l(const orc_pusch_t *p, int meas_symbol, int ch_symbol, int max_ch, const int16_t *rxdataF, const int16_t *ch_est, int32_t *avg_out)
{
  const int N = p->fft_size, nrx = p->nb_rx;
  const int len = (orc_pusch_nb_re(p, meas_symbol) + 15) & ~15, cap = (p->rb_size * 12 + 15) & ~15;
  const int shift_ch_ext = log2_approx_((uint32_t)(max_ch >> 11));
  int b = 3, amp = 8192;
  if (shift_ch_ext > 3) { b = 0; amp = (int16_t)(amp >> (shift_ch_ext - 3)); if (amp == 0) amp = 1; } else b -= shift_ch_ext;
  int16_t *rx = calloc(2 * (size_t)cap, 2), *ch = calloc(2 * (size_t)cap, 2);
  const int x = factor2_(len), y = len >> x;
  int avgs = 0;
  for (int l = 0; l < 2; l++)
    for (int a = 0; a < nrx; a++) {
      memset(rx, 0, 4 * (size_t)cap); memset(ch, 0, 4 * (size_t)cap);
      orc_pusch_extract(p, (p->ul_dmrs_symb_pos >> meas_symbol) & 1, rxdataF + 2 * ((size_t)a * 14 + meas_symbol) * N,
                        ch_est + 2 * ((size_t)(l * nrx + a) * 14 + ch_symbol) * N, rx, ch);
      int32_t lane[4] = {0, 0, 0, 0};
      for (int i = 0; i < (len >> 2) * 4; i++) {
        const int16_t r = wrap16((((int32_t)ch[2 * i] * amp) >> 16) << b), im = wrap16((((int32_t)ch[2 * i + 1] * amp) >> 16) << b);
        lane[i & 3] = wrap32((int64_t)lane[i & 3] + (wrap32((int64_t)r * r + (int64_t)im * im) >> x));
      }
      const int32_t avg = wrap32((int64_t)lane[0] + lane[1] + lane[2] + lane[3]) / y;
      if (avg_out) avg_out[l * nrx + a] = avg;
      if (avg > avgs) avgs = avg;
    }
  free(rx); free(ch);
  const int l2 = (log2_approx_((uint32_t)avgs) >> 1) - (p->Qm >= 6 ? 3 : 0);   /* - 3 for the MMSE receiver only (:1640-1641) */
  return l2 < 0 ? 0 : l2;
}

/* ---- UE side: nr_rx_pdsch for one layer (NR_UE_TRANSPORT/nr_dlsch_demodulation.c:241-684 with nr_dlsch_extract_rbs :1182-1301, nr_dlsch_scale_channel
 * :1052-1101, nr_dlsch_channel_level :1104-1142, nr_dlsch_channel_compensation :737-1050, nr_dlsch_detection_mrc :1303-1368, nr_dlsch_llr :1909-1990,
 * get_valid_dmrs_idx_for_channel_est NR_REFSIG/dmrs_nr.c:321-340).  Differences to the gNB receiver: the estimates are scaled (mulhi 8192, << 3) BEFORE the
 * matched filter, every antenna's output is packed on its own and the antennas are combined with SATURATING adds, the magnitude thresholds use mulhi << 1,
 * and the LLRs of the whole slot are computed at the last symbol with the magnitudes of that LAST symbol (the per-symbol magnitude buffers are locals of
 * nr_rx_pdsch).  p->ul_dmrs_symb_pos = dlDmrsSymbPos, p->num_dmrs_cdm_grps_no_data = n_dmrs_cdm_groups.  Returns the number of LLRs written. */
static int valid_dmrs_idx(int pos, int symbol)
{
  if ((pos >> symbol) & 1) return symbol;
  for (int s = symbol; s >= 0; s--) if ((pos >> s) & 1) return s;
  for (int s = symbol; s < 14; s++) if ((pos >> s) & 1) return s;
  return -1;
}

/* ---- PT-RS at the UE (nr_pdsch_ptrs_processing, NR_UE_ESTIMATION/nr_dl_channel_estimation.c:1765-1907, called from nr_rx_pdsch :569-574 after the symbol's
 * compensation; NR_REFSIG/ptrs_nr.c).  Per PT-RS symbol the common phase error is estimated from the compensated PT-RS REs and those REs are squeezed out of
 * rxdataF_comp (nr_ptrs_cpe_estimation :181-263); at the slot's last symbol the estimates are interpolated over the symbols without PT-RS
 * (nr_ptrs_process_slot :281-337, only when PTRSTimeDensity > 0) and every non-DMRS symbol is rotated (rotate_cpx_vector over 12 * nb_rb REs) before the LLRs of
 * the slot are computed.  The magnitude thresholds are NOT squeezed: LLR j of a PT-RS symbol pairs the j-th data RE with threshold j of the last symbol.
 * Floating point: plain double arithmetic, no fused multiply-add (oracle/_ref is built with -mavx2 only; a -march=native build of the reference may contract
 * real * real + imag * imag and differ in the last bit -- DESIGN.md, defect 21).  Out-of-range double -> int16 conversions follow x86's cvttsd2si + truncation. */
static inline int16_t d2i16(double v)
{
  if (!(v > -2147483649.0 && v < 2147483648.0)) return 0;            /* cvttsd2si's "integer indefinite" 0x80000000: low half 0 */
  return (int16_t)(uint16_t)(uint32_t)(int32_t)v;
}
/* set_ptrs_symb_idx (ptrs_nr.c:53-86): L_ptrs = 1 << PTRSTimeDensity */
uint32_t orc_ptrs_symbols(int start_symbol, int duration, int L_ptrs, uint32_t dmrs_pos)
{
  uint32_t out = 0;
  int i = 0, l_ref = start_symbol;
  const int last = start_symbol + duration - 1;
  if (L_ptrs == 0) return 0;
  while (l_ref + i * L_ptrs <= last) {
    int is_dmrs = 0, l;
    const int lo = (l_ref + (i - 1) * L_ptrs + 1) > l_ref ? (l_ref + (i - 1) * L_ptrs + 1) : l_ref;
    for (l = l_ref + i * L_ptrs; l >= lo; l--) if ((dmrs_pos >> l) & 1) { is_dmrs = 1; break; }
    if (is_dmrs) { l_ref = l; i = 1; continue; }
    out |= 1u << (l_ref + i * L_ptrs);
    i++;
  }
  return out;
}
/* is_ptrs_subcarrier (ptrs_nr.c:107-129) with start_sc = 0 */
static int is_ptrs_sc(int k, int rnti, int K, int nb_rb, int k_re_ref)
{
  const int k_rb_ref = (nb_rb % K == 0) ? rnti % K : rnti % (nb_rb % K);
  return (k - k_re_ref - k_rb_ref * 12) % (K * 12) == 0;
}
static int next_bit(uint32_t mask, int from, int end) { for (int s = from; s < end; s++) if ((mask >> s) & 1) return s; return -1; }
static int next_est(uint32_t ptrs, uint32_t dmrs, int from, int end)
{
  const int np = next_bit(ptrs, from, end), nd = next_bit(dmrs, from, end);
  if (nd == -1) return np;
  if (np == -1) return nd;
  return np > nd ? nd : np;
}
static void slope_from(int start, int end, const int16_t *est, double *sl)
{
  const uint8_t distance = (uint8_t)(end - start);
  sl[0] = (double)(est[2 * end] - est[2 * start]) / distance;
  sl[1] = (double)(est[2 * end + 1] - est[2 * start + 1]) / distance;
}
static void est_from_slope(int16_t *est, const double *sl, int start, int end)
{
  for (uint8_t i = 1; i < (end - start); i++) {
    est[2 * (start + i)] = wrap16((int32_t)est[2 * start] + d2i16(i * sl[0]));
    est[2 * (start + i) + 1] = wrap16((int32_t)est[2 * start + 1] + d2i16(i * sl[1]));
  }
}
/* nr_ptrs_process_slot (ptrs_nr.c:281-337), its control flow kept literally (int8_t references, the initial leftRef = rightRef = 0) */
int orc_ptrs_process_slot(uint32_t dmrs, uint32_t ptrs, int16_t *est, int start, int nsym)
{
  double slope[2] = {0, 0};
  const uint8_t end = (uint8_t)(start + nsym);
  int8_t right = 0, left = 0, tmp = 0;
  for (uint8_t symb = (uint8_t)start; symb < end; symb++) {
    if (((ptrs >> symb) & 1) || ((dmrs >> symb) & 1)) {
      left = (int8_t)symb;
      right = (int8_t)next_est(ptrs, dmrs, symb + 1, end);
    } else {
      if (symb == start && left == -1 && right == -1) return -1;
      if (right != -1 && ((dmrs >> right) & 1)) {
        tmp = (int8_t)next_est(ptrs, dmrs, right + 1, end);
        if (tmp != -1) slope_from(right, tmp, est, slope);
        est_from_slope(est, slope, left, right);
        symb = (uint8_t)(right - 1);
      } else if (right != -1 && ((ptrs >> right) & 1)) {
        slope_from(left, right, est, slope);
        est_from_slope(est, slope, left, right);
        symb = (uint8_t)(right - 1);
      } else if (right == -1 && symb < end) {
        est_from_slope(est, slope, symb - 1, end);
        symb = end;
      } else return -1;
    }
  }
  return 0;
}
/* nr_ptrs_cpe_estimation (ptrs_nr.c:181-263) on one symbol's compensated REs (12 * nb_rb c16, squeezed in place); gold = the symbol's PDSCH DMRS sequence */
static int ptrs_cpe(const orc_ptrs_t *t, int nb_rb, int16_t *comp, const uint32_t *gold, int16_t *est)
{
  const int K = t->K, sc = (nb_rb + K - 1) / K;
  int32_t sr = 0, si = 0;
  int re_cnt = 0, cnt = 0;
  for (int re = 0; re < 12 * nb_rb; re++) {
    if (is_ptrs_sc(re, t->rnti, K, nb_rb, t->re_offset)) {
      /* nr_gen_ref_conj_symbols (nr_dmrs_rx.c:240-256): conjugated QPSK symbol of bits 2i, 2i+1 */
      const int b0 = (gold[(2 * re_cnt) >> 5] >> ((2 * re_cnt) & 31)) & 1, b1 = (gold[(2 * re_cnt + 1) >> 5] >> ((2 * re_cnt + 1) & 31)) & 1;
      const int32_t pr = b0 ? -23170 : 23170, pi = b1 ? 23170 : -23170, xr = comp[2 * re], xi = comp[2 * re + 1];
      /* mult_cpx_vector (cmult_vv.c:96-156), shift 15, packs_epi32 */
      if (re_cnt < sc) {
        sr += sat16(wrap32((int64_t)xr * pr + (int64_t)wrap16(-xi) * pi) >> 15);
        si += sat16(wrap32((int64_t)xi * pr + (int64_t)xr * pi) >> 15);
      }
      re_cnt++;
    } else { comp[2 * cnt] = comp[2 * re]; comp[2 * cnt + 1] = comp[2 * re + 1]; cnt++; }
  }
  double real = (double)sr, imag = (double)si;
  real /= sc; imag /= sc;
  const volatile double rr = real * real, ii = imag * imag;            /* volatile: no contraction into an fma whatever the flags */
  const double ab = sqrt(rr + ii);
  est[0] = d2i16((real / ab) * (1 << 15));
  est[1] = d2i16((-1) * (imag / ab) * (1 << 15));
  return re_cnt;
}

static int pdsch_rx_slot_impl(const orc_pusch_t *p, const orc_ptrs_t *t, int start_symbol, int nr_symbols, const int16_t *rxdataF, const int16_t *dl_ch_est, int16_t *llr,
                              int32_t *log2_maxh_out, int16_t *phase_out, int32_t *ptrs_re_out);
int orc_pdsch_rx_slot(const orc_pusch_t *p, int start_symbol, int nr_symbols, const int16_t *rxdataF, const int16_t *dl_ch_est, int16_t *llr, int32_t *log2_maxh_out)
{
  return pdsch_rx_slot_impl(p, NULL, start_symbol, nr_symbols, rxdataF, dl_ch_est, llr, log2_maxh_out, NULL, NULL);
}
/* phase_out: 14 {re, im} (ptrs_phase_per_slot[0]); ptrs_re_out: 14 (ptrs_re_per_slot[0]); both optional */
int orc_pdsch_rx_slot_ptrs(const orc_pusch_t *p, const orc_ptrs_t *t, int start_symbol, int nr_symbols, const int16_t *rxdataF, const int16_t *dl_ch_est, int16_t *llr,
                           int32_t *log2_maxh_out, int16_t *phase_out, int32_t *ptrs_re_out)
{
  return pdsch_rx_slot_impl(p, t, start_symbol, nr_symbols, rxdataF, dl_ch_est, llr, log2_maxh_out, phase_out, ptrs_re_out);
}

static int pdsch_rx_slot_impl(const orc_pusch_t *p, const orc_ptrs_t *t, int start_symbol, int nr_symbols, const int16_t *rxdataF, const int16_t *dl_ch_est, int16_t *llr,
                              int32_t *log2_maxh_out, int16_t *phase_out, int32_t *ptrs_re_out)
{
  const int N = p->fft_size, nrx = p->nb_rx, nb = p->rb_size, Qm = p->Qm, pos = p->ul_dmrs_symb_pos, type = p->dmrs_config_type, cdm = p->num_dmrs_cdm_grps_no_data;
  const int sz = (nb * 12 + 15) & ~15;
  const int start_re = (p->first_carrier_offset + (p->rb_start + p->bwp_start) * 12) % N;
  const int ampv[3] = {Qm == 4 ? 20724 : Qm == 6 ? 20225 : Qm == 8 ? 20106 : 0, Qm == 6 ? 10112 : Qm == 8 ? 10053 : 0, Qm == 8 ? 5026 : 0};
  int16_t *rx = malloc(4 * (size_t)sz * nrx), *ch = malloc(4 * (size_t)sz * nrx), *comp = calloc((size_t)14 * nb * 12 * 2, 2), *mag[3], *cm = malloc(4 * (size_t)sz);
  for (int t = 0; t < 3; t++) mag[t] = calloc(2 * (size_t)sz, 2);
  int valid[14] = {0}, log2_maxh = 0;
  int16_t phase[28] = {0};
  int32_t ptrs_re[14] = {0};
  uint32_t ptrs_pos = 0;
  int first_with_data = start_symbol;
  const int dmrs_data_re = type == 0 ? 12 - 6 * cdm : 12 - 4 * cdm;
  while (dmrs_data_re == 0 && ((pos >> first_with_data) & 1)) first_with_data++;
  for (int m = start_symbol; m < start_symbol + nr_symbols; m++) {
    const int pilots = (pos >> m) & 1, vd = valid_dmrs_idx(pos, m);
    memset(rx, 0, 4 * (size_t)sz * nrx); memset(ch, 0, 4 * (size_t)sz * nrx);
    for (int t = 0; t < 3; t++) memset(mag[t], 0, 4 * (size_t)sz);      /* the magnitude buffers are per-call locals */
    for (int a = 0; a < nrx; a++) {
      const int16_t *rxF = rxdataF + 2 * ((size_t)a * 14 + m) * N, *h = dl_ch_est + 2 * ((size_t)a * 14 + vd) * N;
      int16_t *re = rx + 2 * (size_t)sz * a, *ce = ch + 2 * (size_t)sz * a;
      int n = 0;
#define PUT2(ri, ci) do { re[2 * n] = rxF[2 * (ri)]; re[2 * n + 1] = rxF[2 * (ri) + 1]; ce[2 * n] = h[2 * (ci)]; ce[2 * n + 1] = h[2 * (ci) + 1]; n++; } while (0)
      if (!pilots) { for (int i = 0; i < 12 * nb; i++) PUT2((start_re + i) % N, i); }
      else if (type == 0) {
        if (cdm == 1) { int k = start_re; for (int j = 0; j < 6 * nb; j += 3) { PUT2(k + 1, 2 * j + 1); PUT2(k + 3, 2 * j + 3); PUT2(k + 5, 2 * j + 5); k += 6; if (k >= N) k -= N; } }
      } else {
        if (cdm == 1) { int k = start_re; for (int j = 0; j < 8 * nb; j += 4) { const int c0 = 6 * (j / 4); PUT2(k + 2, c0 + 2); PUT2(k + 3, c0 + 3); PUT2(k + 4, c0 + 4); PUT2(k + 5, c0 + 5); k += 6; if (k >= N) k -= N; } }
        else if (cdm == 2) { int k = start_re; for (int j = 0; j < 4 * nb; j += 2) { const int c0 = 6 * (j / 2); PUT2(k + 4, c0 + 4); PUT2(k + 5, c0 + 5); k += 6; if (k >= N) k -= N; } }
      }
#undef PUT2
    }
    const int len = pilots ? (type == 0 ? nb * (12 - 6 * cdm) : nb * (12 - 4 * cdm)) : nb * 12;
    const int nb_rb_0 = len / 12 + ((len % 12) ? 1 : 0), span = nb_rb_0 * 12;
    for (int a = 0; a < nrx; a++)                                       /* scale */
      for (int i = 0; i < 2 * span; i++) { int16_t *c = ch + 2 * (size_t)sz * a; c[i] = wrap16((((int32_t)c[i] * 8192) >> 16) << 3); }
    if (m == first_with_data) {                                         /* level -> log2_maxh */
      const int x = factor2_((uint32_t)len), y = len >> x;
      int avgs = 0;
      for (int a = 0; a < nrx; a++) {
        const int16_t *c = ch + 2 * (size_t)sz * a;
        int32_t lane[4] = {0, 0, 0, 0};
        for (int i = 0; i < span; i++) lane[i & 3] = wrap32((int64_t)lane[i & 3] + (wrap32((int64_t)c[2 * i] * c[2 * i] + (int64_t)c[2 * i + 1] * c[2 * i + 1]) >> x));
        const int avg = (int)(((int64_t)lane[0] + lane[1] + lane[2] + lane[3]) / y);
        if (avg > avgs) avgs = avg;
      }
      log2_maxh = (log2_approx_((uint32_t)avgs) / 2) + 1;
    }
    const int shift = log2_maxh;
    int16_t *out = comp + 2 * (size_t)m * nb * 12;
    for (int a = 0; a < nrx; a++) {                                     /* matched filter per antenna, then saturating MRC */
      const int16_t *c = ch + 2 * (size_t)sz * a, *y = rx + 2 * (size_t)sz * a;
      for (int i = 0; i < span; i++) {
        const int32_t hr = c[2 * i], hi = c[2 * i + 1], yr = y[2 * i], yi = y[2 * i + 1];
        const int16_t cr = sat16(wrap32((int64_t)hr * yr + (int64_t)hi * yi) >> shift), ci = sat16(wrap32((int64_t)wrap16(-hi) * yr + (int64_t)hr * yi) >> shift);
        const int16_t mm = sat16(wrap32((int64_t)hr * hr + (int64_t)hi * hi) >> shift);
        if (a == 0) { cm[2 * i] = cr; cm[2 * i + 1] = ci; } else { cm[2 * i] = sat16((int32_t)cm[2 * i] + cr); cm[2 * i + 1] = sat16((int32_t)cm[2 * i + 1] + ci); }
        if (Qm > 2)
          for (int t = 0; t < 3; t++) {
            const int16_t v = wrap16(((mm * ampv[t]) >> 16) << 1);
            if (a == 0) { mag[t][2 * i] = v; mag[t][2 * i + 1] = v; } else { mag[t][2 * i] = sat16((int32_t)mag[t][2 * i] + v); mag[t][2 * i + 1] = sat16((int32_t)mag[t][2 * i + 1] + v); }
          }
      }
    }
    /* rxdataF_comp of symbol m lives at m * nb_rb * 12; vectors beyond the last symbol's region are not modelled (span <= nb * 12) */
    memcpy(out, cm, 4 * (size_t)span);
    valid[m] = len;
    if (t && t->on) {                                                    /* nr_pdsch_ptrs_processing for receive antenna 0 (the plane the LLRs read) */
      ptrs_re[m] = 0; phase[2 * m + 1] = 0; phase[2 * m] = pilots ? 32767 : 0;
      if (m == start_symbol) ptrs_pos = orc_ptrs_symbols(start_symbol, nr_symbols, 1 << t->L, (uint32_t)pos);
      if ((ptrs_pos >> m) & 1) {
        uint32_t gold[20];
        const uint64_t x2tmp0 = ((uint64_t)(14 * t->slot + m + 1) * (((uint64_t)t->nid << 1) + 1)) << 17;     /* nr_gold_pdsch (nr_gold_ue.c:75-93) */
        orc_gold_words((uint32_t)((x2tmp0 + ((uint64_t)t->nid << 1) + t->nscid) % (1U << 31)), 20, gold);
        ptrs_re[m] = ptrs_cpe(t, nb, out, gold, phase + 2 * m);
      }
      valid[m] -= ptrs_re[m];
      if (m == start_symbol + nr_symbols - 1) {
        int ret = 0;
        if (t->L > 0) ret = orc_ptrs_process_slot((uint32_t)pos, ptrs_pos, phase, start_symbol, nr_symbols);
        for (int i = start_symbol; i < start_symbol + nr_symbols; i++)
          if (!((pos >> i) & 1) && ret == 0) orc_rotate_cpx_vector(comp + 2 * (size_t)i * nb * 12, phase[2 * i], phase[2 * i + 1], comp + 2 * (size_t)i * nb * 12, nb * 12);
      }
    }
  }
  if (phase_out) memcpy(phase_out, phase, sizeof(phase));
  if (ptrs_re_out) memcpy(ptrs_re_out, ptrs_re, sizeof(ptrs_re));
  /* LLRs of the whole slot with the magnitudes of the last symbol */
  size_t off = 0;
  for (int m = start_symbol; m < start_symbol + nr_symbols; m++) {
    orc_ulsch_llr(Qm, comp + 2 * (size_t)m * nb * 12, mag[0], mag[1], mag[2], llr + off, (uint32_t)valid[m]);
    off += (size_t)valid[m] * Qm;
  }
  if (log2_maxh_out) *log2_maxh_out = log2_maxh;
  free(rx); free(ch); free(comp); free(cm);
  for (int t = 0; t < 3; t++) free(mag[t]);
  return (int)off;
}

/* ---- UE side, two layers: nr_rx_pdsch with Nl = 2 (same file: nr_dlsch_channel_level_median :1144-1179, nr_dlsch_detection_mrc :1303-1368 per layer,
 * nr_zero_forcing_rx :1726-1869 with nr_conjch0_mult_ch1 :1679-1720, nr_matrix_inverse / nr_determin :1460-1506,1549-1610 (fixed-point branch),
 * nr_a_mult_b :1389-1427, nr_a_sum_b :1370-1383, nr_dlsch_layer_demapping :1871-1907).  dl_ch_est holds 2 * nb_rx planes, index layer * nb_rx + rx.
 * What differs from one layer: the level also takes the (max + min) / 2 "median" of the 4-RE power sums; after the per-layer matched filter + saturating
 * MRC, H^H H (2x2, each element packed per antenna and summed with saturating adds) is inverted as adjugate / determinant with products shifted by
 * log2_maxh - 2, the adjugate is applied to the two matched-filter outputs, and the QAM thresholds become determinant * QAM_amp (mulhi, << 1) -- again only
 * the LAST symbol's thresholds reach the LLR stage.  nb_rx >= 2 (the reference skips MRC and zero forcing when n_rx == 1). */
static inline int32_t sra32_(int32_t v, int s) { return ((unsigned)s & 0xFF) > 31 ? (v < 0 ? -1 : 0) : v >> (s & 0xFF); }
typedef struct { int16_t r, i; } c16o;
static inline c16o conj0_mult1(c16o a, c16o b, int s)   /* conj(a) * b >> s, packed (nr_conjch0_mult_ch1) */
{
  c16o o;
  o.r = sat16(sra32_(wrap32((int64_t)a.r * b.r + (int64_t)a.i * b.i), s));
  o.i = sat16(sra32_(wrap32((int64_t)wrap16(-a.i) * b.r + (int64_t)a.r * b.i), s));
  return o;
}
static inline c16o a_mult_b(c16o a, c16o b, int s)      /* a * b >> s, packed (nr_a_mult_b) */
{
  c16o o;
  o.r = sat16(sra32_(wrap32((int64_t)a.r * b.r + (int64_t)wrap16(-a.i) * b.i), s));
  o.i = sat16(sra32_(wrap32((int64_t)a.i * b.r + (int64_t)a.r * b.i), s));
  return o;
}
static inline c16o adds_c(c16o a, c16o b) { c16o o = {sat16((int32_t)a.r + b.r), sat16((int32_t)a.i + b.i)}; return o; }
static inline c16o neg_c(c16o a) { c16o o = {wrap16(-a.r), wrap16(-a.i)}; return o; }      /* sign_epi16 by -1: -32768 stays */
static void ue_src(const orc_pusch_t *p, int start_re, int pilots, int i, int *ri, int *ci)
{
  const int N = p->fft_size;
  if (!pilots) { *ri = (start_re + i) % N; *ci = i; return; }
  int per, first;
  if (p->dmrs_config_type == 0) { per = 3; first = 1; } else if (p->num_dmrs_cdm_grps_no_data == 1) { per = 4; first = 2; } else { per = 2; first = 4; }
  const int g = i / per, r = i - g * per;
  int k = start_re + 6 * g;
  while (k >= N) k -= N;
  const int o = p->dmrs_config_type == 0 ? 2 * r + first : r + first;
  *ri = k + o; *ci = 6 * g + o;
}

/* nr_determin (:1460-1506) on one resource element: Laplace expansion along column 0 exactly as written -- det = sum over rows rtx of a[0][rtx] * det(minor(rtx, 0))
 * with the sign (-1)^rtx handed DOWN the recursion (it is applied to the 1 x 1 leaves by nr_element_sign), every product nr_a_mult_b (>> shift0, packed) and every sum
 * nr_a_sum_b (saturating).  a[c][r] is indexed [column][row] like a44. */
static c16o det_re(int size, c16o a[4][4], int sign, int shift0)
{
  if (size == 1) return sign < 0 ? neg_c(a[0][0]) : a[0][0];
  c16o acc = {0, 0};
  for (int rtx = 0; rtx < size; rtx++) {
    c16o sub[4][4];
    int rr[3], cc[3], k = 0;
    for (int r = 0; r < size; r++) if (r != rtx) rr[k++] = r;
    k = 0;
    for (int c = 0; c < size; c++) if (c != 0) cc[k++] = c;
    for (int ri = 0; ri < size - 1; ri++) for (int ci = 0; ci < size - 1; ci++) sub[ci][ri] = a[cc[ci]][rr[ri]];
    const c16o prod = a_mult_b(a[0][rtx], det_re(size - 1, sub, ((rtx & 1) ? -1 : 1) * sign, shift0), shift0);
    acc = rtx == 0 ? prod : adds_c(acc, prod);
  }
  return acc;
}
/* nr_matrix_inverse, fixed-point branch (:1549-1610): inv[rtx][ctx] = det of the minor without row rtx and column ctx, sign (-1)^(rtx + ctx) */
static void inverse_re(int size, c16o a[4][4], c16o inv[4][4], int shift0)
{
  for (int rtx = 0; rtx < size; rtx++)
    for (int ctx = 0; ctx < size; ctx++) {
      c16o sub[4][4];
      int rr[3], cc[3], k = 0;
      for (int r = 0; r < size; r++) if (r != rtx) rr[k++] = r;
      k = 0;
      for (int c = 0; c < size; c++) if (c != ctx) cc[k++] = c;
      for (int ri = 0; ri < size - 1; ri++) for (int ci = 0; ci < size - 1; ci++) sub[ci][ri] = a[cc[ci]][rr[ri]];
      inv[rtx][ctx] = det_re(size - 1, sub, ((rtx & 1) ? -1 : 1) * ((ctx & 1) ? -1 : 1), shift0);
    }
}

/* Nl = 2, 3 or 4 layers: the reference's code is generic in n_tx (nr_zero_forcing_rx "for 2, 3, and 4 Tx layers", :528) */
int orc_pdsch_rx_slot_nl(const orc_pusch_t *p, int NL, int start_symbol, int nr_symbols, const int16_t *rxdataF, const int16_t *dl_ch_est, int16_t *llr, int32_t *log2_maxh_out)
{
  if (NL < 2 || NL > 4) return -1;
  const int N = p->fft_size, nrx = p->nb_rx, nb = p->rb_size, Qm = p->Qm, pos = p->ul_dmrs_symb_pos, type = p->dmrs_config_type, cdm = p->num_dmrs_cdm_grps_no_data;
  const int sz = (nb * 12 + 15) & ~15;
  const int start_re = (p->first_carrier_offset + (p->rb_start + p->bwp_start) * 12) % N;
  const int ampv[3] = {Qm == 4 ? 20724 : Qm == 6 ? 20225 : Qm == 8 ? 20106 : 0, Qm == 6 ? 10112 : Qm == 8 ? 10053 : 0, Qm == 8 ? 5026 : 0};
  if (nrx < 2) return -1;
  c16o *rx = malloc(sizeof(c16o) * (size_t)sz * nrx), *ch = malloc(sizeof(c16o) * (size_t)sz * nrx * NL);
  c16o *comp[4];
  int16_t *mag[3], *lay[4];
  for (int l = 0; l < NL; l++) { comp[l] = calloc((size_t)14 * nb * 12 + sz, sizeof(c16o)); lay[l] = calloc((size_t)14 * nb * 12 * Qm + 64, 2); }
  for (int t = 0; t < 3; t++) mag[t] = calloc(2 * (size_t)sz, 2);
  int valid[14] = {0}, log2_maxh = 0;
  int first_with_data = start_symbol;
  const int dmrs_data_re = type == 0 ? 12 - 6 * cdm : 12 - 4 * cdm;
  while (dmrs_data_re == 0 && ((pos >> first_with_data) & 1)) first_with_data++;
  for (int m = start_symbol; m < start_symbol + nr_symbols; m++) {
    const int pilots = (pos >> m) & 1, vd = valid_dmrs_idx(pos, m);
    const int len = pilots ? (type == 0 ? nb * (12 - 6 * cdm) : nb * (12 - 4 * cdm)) : nb * 12;
    const int nb_rb_0 = len / 12 + ((len % 12) ? 1 : 0), span = nb_rb_0 * 12;
    memset(rx, 0, sizeof(c16o) * (size_t)sz * nrx); memset(ch, 0, sizeof(c16o) * (size_t)sz * nrx * NL);
    for (int t = 0; t < 3; t++) memset(mag[t], 0, 4 * (size_t)sz);
    for (int i = 0; i < len; i++) {
      int ri, ci;
      ue_src(p, start_re, pilots, i, &ri, &ci);
      for (int a = 0; a < nrx; a++) {
        const int16_t *rxF = rxdataF + 2 * ((size_t)a * 14 + m) * N;
        rx[(size_t)a * sz + i].r = rxF[2 * ri]; rx[(size_t)a * sz + i].i = rxF[2 * ri + 1];
        for (int l = 0; l < NL; l++) {
          const int16_t *h = dl_ch_est + 2 * ((size_t)(l * nrx + a) * 14 + vd) * N;
          ch[(size_t)(l * nrx + a) * sz + i].r = h[2 * ci]; ch[(size_t)(l * nrx + a) * sz + i].i = h[2 * ci + 1];
        }
      }
    }
    for (int q = 0; q < NL * nrx; q++)                                  /* nr_dlsch_scale_channel */
      for (int i = 0; i < span; i++) {
        c16o *c = ch + (size_t)q * sz + i;
        c->r = wrap16((((int32_t)c->r * 8192) >> 16) << 3); c->i = wrap16((((int32_t)c->i * 8192) >> 16) << 3);
      }
    if (m == first_with_data) {                                         /* level + median -> log2_maxh (:433-452) */
      const int x = factor2_((uint32_t)len), y = len >> x;
      int avgs = 0;
      int32_t avg[4 * 8];
      for (int q = 0; q < NL * nrx; q++) {
        const c16o *c = ch + (size_t)q * sz;
        int32_t lane[4] = {0, 0, 0, 0};
        for (int i = 0; i < span; i++) lane[i & 3] = wrap32((int64_t)lane[i & 3] + (wrap32((int64_t)c[i].r * c[i].r + (int64_t)c[i].i * c[i].i) >> x));
        avg[q] = (int32_t)(((int64_t)lane[0] + lane[1] + lane[2] + lane[3]) / y);
        if (avg[q] > avgs) avgs = avg[q];
      }
      for (int q = 0; q < NL * nrx; q++) {
        const c16o *c = ch + (size_t)q * sz;
        int64_t mx = avg[q], mn = avg[q];
        for (int v = 0; v < (len >> 2); v++) {
          int64_t s = 0;
          for (int j = 0; j < 4; j++) s += wrap32((int64_t)c[4 * v + j].r * c[4 * v + j].r + (int64_t)c[4 * v + j].i * c[4 * v + j].i) >> 2;
          if (s > mx) mx = s;
          if (s < mn) mn = s;
        }
        const int32_t med = (int32_t)((mx + mn) >> 1);
        if (med > avgs) avgs = med;
      }
      log2_maxh = (log2_approx_((uint32_t)avgs) / 2) + 1;
    }
    const int shift = log2_maxh, shift0 = shift - 2;
    for (int i = 0; i < span; i++) {
      c16o mf[4], E[4][4];                                              /* E[c][r] = sum_a conj(H[r][a]) H[c][a] */
      for (int l = 0; l < NL; l++)
        for (int a = 0; a < nrx; a++) {
          const c16o v = conj0_mult1(ch[(size_t)(l * nrx + a) * sz + i], rx[(size_t)a * sz + i], shift);
          mf[l] = a == 0 ? v : adds_c(mf[l], v);
        }
      for (int r = 0; r < NL; r++)
        for (int c = 0; c < NL; c++)
          for (int a = 0; a < nrx; a++) {
            const c16o v = conj0_mult1(ch[(size_t)(r * nrx + a) * sz + i], ch[(size_t)(c * nrx + a) * sz + i], shift);
            E[c][r] = a == 0 ? v : adds_c(E[c][r], v);
          }
      /* determinant and adjugate (2 x 2: a44[0][0] * a44[1][1] + a44[0][1] * (-a44[1][0]); inv[r][c] = (-1)^(r+c) a44[1-c][1-r]); output layer r = sum over c of
       * inv[c][r] * mf[c], accumulated from zero with saturating adds (:1786-1814) */
      const c16o det = det_re(NL, E, +1, shift0);
      c16o inv[4][4];
      inverse_re(NL, E, inv, shift0);
      for (int r = 0; r < NL; r++) {
        c16o acc = {0, 0};
        for (int c = 0; c < NL; c++) acc = adds_c(acc, a_mult_b(inv[c][r], mf[c], shift0));
        comp[r][(size_t)m * nb * 12 + i] = acc;
      }
      if (Qm > 2)
        for (int t = 0; t < 3; t++) { const int16_t v = wrap16(((det.r * ampv[t]) >> 16) << 1); mag[t][2 * i] = v; mag[t][2 * i + 1] = v; }
    }
    valid[m] = len;
  }
  size_t off = 0;
  for (int m = start_symbol; m < start_symbol + nr_symbols; m++) {
    for (int l = 0; l < NL; l++) orc_ulsch_llr(Qm, (const int16_t *)(comp[l] + (size_t)m * nb * 12), mag[0], mag[1], mag[2], lay[l] + off, (uint32_t)valid[m]);
    off += (size_t)valid[m] * Qm;
  }
  const size_t n_re = off / Qm;
  for (size_t i = 0; i < n_re; i++)
    for (int l = 0; l < NL; l++)
      for (int q = 0; q < Qm; q++) llr[NL * Qm * i + l * Qm + q] = lay[l][i * Qm + q];
  if (log2_maxh_out) *log2_maxh_out = log2_maxh;
  free(rx); free(ch);
  for (int l = 0; l < NL; l++) { free(comp[l]); free(lay[l]); }
  for (int t = 0; t < 3; t++) free(mag[t]);
  return (int)(NL * off);
}

int orc_pdsch_rx_slot_2l(const orc_pusch_t *p, int start_symbol, int nr_symbols, const int16_t *rxdataF, const int16_t *dl_ch_est, int16_t *llr, int32_t *log2_maxh_out)
{
  return orc_pdsch_rx_slot_nl(p, 2, start_symbol, nr_symbols, rxdataF, dl_ch_est, llr, log2_maxh_out);
}

/* flat entry points of the two ML kernels for unit tests: one resource element */
void orc_ml_qpsk_qpsk(const int16_t *y0, const int16_t *y1, const int16_t *rho, int16_t *out2) { ml_qpsk_qpsk(y0, y1, rho, out2); out2[0] >>= 4; out2[1] >>= 4; }
void orc_ml_qam16_qam16(const int16_t *y0, const int16_t *y1, int mag_des, int mag_int, const int16_t *rho, int16_t *out4)
{
  ml_qam16_qam16(y0, y1, (int16_t)mag_des, (int16_t)mag_int, rho, out4);
}
