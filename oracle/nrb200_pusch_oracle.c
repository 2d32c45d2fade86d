/* TEST INFRASTRUCTURE ONLY (see nrb200_oracle.h).  CPU restatement of the single-layer PUSCH inner receiver of the reference
 * (openair1/PHY/NR_TRANSPORT/nr_ulsch_demodulation.c): nr_ulsch_extract_rbs :279-380, nr_ulsch_scale_channel :382-414,
 * get_nb_re_pusch :416-432, nr_ulsch_channel_level :434-466, nr_ulsch_channel_compensation :468-578 (rho == NULL), the log2_maxh rule
 * of nr_rx_pusch_tp :1595-1647 and the per-symbol composition inner_rx :1262-1384 followed by nr_ulsch_compute_llr.
 * Pinned against the compiled reference (oracle/_ref/libref_pusch.so) by tests/test_oracle_vs_reference.py. */
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
  orc_ulsch_llr(Qm, comp, ma, mb, mc, llr, (uint32_t)valid);
  if (comp_out) memcpy(comp_out, comp, 4 * (size_t)blen);
  free(rx); free(ch); free(comp); free(ma); free(mb); free(mc);
  return valid;
}
