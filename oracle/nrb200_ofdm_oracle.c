/* TEST INFRASTRUCTURE ONLY (see nrb200_oracle.h).  CPU restatement of the slot-level OFDM front end of the reference:
 *   apply_nr_rotation_TX + nr_normal_prefix_mod/PHY_ofdm_mod   openair1/PHY/MODULATION/ofdm_mod.c:67-127, 130-281, 337-376
 *   nr_slot_fep_ul + apply_nr_rotation_RX                      openair1/PHY/MODULATION/slot_fep_nr.c:223-332
 *   rotate_cpx_vector (AVX2 branch) / multadd_cpx_vector       openair1/PHY/TOOLS/cmult_sv.c:77-145, cmult_vv.c:158-213
 *   init_symbol_rotation / init_timeshift_rotation             openair1/PHY/MODULATION/nr_modulation.c:586-660
 * Pinned against the compiled reference (oracle/_ref/libref_ofdm.so) by tests/test_oracle_vs_reference.py. */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "nrb200_oracle.h"

static inline int16_t sat16(int32_t v) { return v > 32767 ? 32767 : v < -32768 ? -32768 : (int16_t)v; }
static inline int32_t wrap32(int64_t v) { return (int32_t)(uint32_t)(uint64_t)v; }

/* y = x * alpha >> 15.  The vector body (8 elements at a time) packs with saturation, the scalar tail (N % 8 elements) casts. */
void orc_rotate_cpx_vector(const int16_t *x, int16_t ar, int16_t ai, int16_t *y, uint32_t N)
{
  const int16_t nai = (int16_t)(uint16_t)(0u - (uint16_t)ai);
  const uint32_t body = (N / 8) * 8;
  for (uint32_t n = 0; n < N; n++) {
    const int32_t xr = x[2 * n], xi = x[2 * n + 1];
    if (n < body) {
      const int32_t re = wrap32((int64_t)xr * ar + (int64_t)xi * nai) >> 15;   /* madd_epi16 with {r, -i} */
      const int32_t im = wrap32((int64_t)xr * ai + (int64_t)xi * ar) >> 15;    /* madd_epi16 with {i, r}  */
      y[2 * n] = sat16(re); y[2 * n + 1] = sat16(im);
    } else {
      y[2 * n] = (int16_t)(wrap32((int64_t)xr * ar - (int64_t)xi * ai) >> 15);
      y[2 * n + 1] = (int16_t)(wrap32((int64_t)xr * ai + (int64_t)xi * ar) >> 15);
    }
  }
}

/* multadd_cpx_vector(x1, x2, y, zero_flag = 1, N, 15): y = x1 * x2 >> 15 (plain product despite the name), 4 * (N >> 2) elements */
void orc_mult_cpx_vector(const int16_t *x1, const int16_t *x2, int16_t *y, uint32_t N)
{
  for (uint32_t n = 0; n < (N >> 2) * 4; n++) {
    const int32_t ar = x1[2 * n], ai = x1[2 * n + 1], br = x2[2 * n], bi = x2[2 * n + 1];
    const int32_t nai = (int16_t)(uint16_t)(0u - (uint16_t)ai);                 /* sign_epi16(x, -1) wraps -32768 */
    y[2 * n] = sat16(wrap32((int64_t)ar * br + (int64_t)nai * bi) >> 15);
    y[2 * n + 1] = sat16(wrap32((int64_t)ai * br + (int64_t)ar * bi) >> 15);
  }
}

/* carrier-phase pre-compensation per symbol of a subframe: 14 << mu {re,im} pairs */
void orc_symbol_rotation(int mu, double f0, int16_t *rot)
{
  const int nsymb = 14 << mu;
  const double Tc = (1 / 480e3 / 4096);
  const double Nu = 2048 * 64 * (1 / (float)(1 << mu));
  const double Ncp0 = 16 * 64 + (144 * 64 * (1 / (float)(1 << mu)));
  const double Ncp1 = (144 * 64 * (1 / (float)(1 << mu)));
  double tl = 0.0;
  for (int l = 0; l < nsymb; l++) {
    const double Ncp = (l == 0 || l == (7 * (1 << mu))) ? Ncp0 : Ncp1;
    const double poff = 2 * M_PI * (tl + (Ncp * Tc)) * f0;
    rot[2 * l] = (int16_t)floor(cos(poff) * 32767);
    rot[2 * l + 1] = (int16_t)floor(sin(-poff) * 32767);
    tl += (Nu + Ncp) * Tc;
  }
}

void orc_timeshift_rotation(int N, int sample_offset, int16_t *out)
{
  for (int i = 0; i < N; i++) {
    const double poff = -i * 2.0 * M_PI * sample_offset / N;
    out[2 * i] = (int16_t)round(cos(poff) * 32767);
    out[2 * i + 1] = (int16_t)round(sin(-poff) * 32767);
  }
}

/* Slot geometry.  prefix[l] = CP length of symbol l, cp_start[l] = first CP sample of symbol l relative to the slot start,
 * *slot_start = first sample of the slot in the frame, *frame_len = samples per frame. */
void orc_ofdm_geometry(int N, int mu, int slot, uint32_t *prefix, uint32_t *cp_start, uint32_t *slot_start, uint32_t *frame_len)
{
  const uint32_t p = N / 128 * 9, p0 = N / 128 * (9 + (1 << mu));
  uint32_t pos = 0;
  for (int l = 0; l < 14; l++) {
    prefix[l] = ((slot * 14 + l) % (7 << mu)) ? p : p0;
    cp_start[l] = pos;
    pos += prefix[l] + N;
  }
  const uint32_t slotN0 = (p + N) * 14, slot0 = p0 + 13 * p + 14 * N;
  const uint32_t subframe = (p0 + N) * 2 + (p + N) * (14 * (1 << mu) - 2);
  uint32_t s = 0;
  for (int i = 0; i < slot; i++) s += mu == 0 ? subframe : ((i % ((1 << mu) / 2)) ? slotN0 : slot0);
  *slot_start = s;
  *frame_len = 10 * subframe;
}

static void rot_ranges(int N, int nb_rb, uint32_t start[2], uint32_t *len)
{
  const uint32_t fco = N - nb_rb * 6;
  if (nb_rb & 1) { *len = (nb_rb + 1) * 6; start[0] = 0; start[1] = fco - 6; }
  else { *len = nb_rb * 6; start[0] = 0; start[1] = fco; }
}

/* rot = the 14 << mu entries of symbol_rotation (or NULL: no rotation).  txdataF (nsymb * N c16) is rotated in place like the reference. */
void orc_ofdm_tx_slot(int N, int mu, int nb_rb, int slot, int nsymb, const int16_t *rot, int16_t *txdataF, int16_t *txdata)
{
  uint32_t prefix[14], cp_start[14], ss, fl, st[2], len;
  orc_ofdm_geometry(N, mu, slot, prefix, cp_start, &ss, &fl);
  rot_ranges(N, nb_rb, st, &len);
  const int symb_offset = (slot % (1 << mu)) * 14;
  if (rot)
    for (int l = 0; l < nsymb; l++)
      for (int h = 0; h < 2; h++) {
        int16_t *p = txdataF + 2 * ((size_t)l * N + st[h]);
        orc_rotate_cpx_vector(p, rot[2 * (symb_offset + l)], rot[2 * (symb_offset + l) + 1], p, len);
      }
  if (mu == 0) nsymb = 14;                                   /* nr_normal_prefix_mod ignores nsymb for numerology 0 */
  int16_t *tmp = (int16_t *)malloc((size_t)N * 4);
  for (int l = 0; l < nsymb; l++) {
    orc_dft(N, 1, txdataF + 2 * (size_t)l * N, tmp, 1);
    int16_t *o = txdata + 2 * (size_t)cp_start[l];
    memcpy(o + 2 * prefix[l], tmp, (size_t)N * 4);
    memcpy(o, tmp + 2 * (N - prefix[l]), (size_t)prefix[l] * 4);
  }
  free(tmp);
}

/* rxdata = one frame (frame_len c16, ring); sample_offset = timing advance offset (N_TA_offset); rot = UL symbol_rotation or NULL */
void orc_ofdm_rx_slot(int N, int mu, int nb_rb, int slot, int divisor, int sample_offset, const int16_t *rot, const int16_t *rxdata, int16_t *rxdataF)
{
  uint32_t prefix[14], cp_start[14], ss, fl, st[2], len;
  orc_ofdm_geometry(N, mu, slot, prefix, cp_start, &ss, &fl);
  rot_ranges(N, nb_rb, st, &len);
  const uint32_t p = N / 128 * 9;
  int16_t *tmp = (int16_t *)malloc((size_t)N * 4), *ts = (int16_t *)malloc((size_t)N * 4);
  orc_timeshift_rotation(N, p / divisor, ts);
  const int symb_offset = (slot % (1 << mu)) * 14;
  for (int l = 0; l < 14; l++) {
    /* FFT window: start of the useful part minus 1/divisor of the (short) CP */
    const int64_t off = (int64_t)ss + cp_start[l] + prefix[l] - p / divisor - sample_offset;
    for (int i = 0; i < N; i++) {
      const int64_t k = ((off + i) % (int64_t)fl + fl) % fl;
      tmp[2 * i] = rxdata[2 * k]; tmp[2 * i + 1] = rxdata[2 * k + 1];
    }
    int16_t *F = rxdataF + 2 * (size_t)l * N;
    orc_dft(N, 0, tmp, F, 1);
    if (rot) {
      const int16_t rr = rot[2 * (symb_offset + l)], ri = (int16_t)(uint16_t)(0u - (uint16_t)rot[2 * (symb_offset + l) + 1]);
      for (int h = 0; h < 2; h++) orc_rotate_cpx_vector(F + 2 * st[h], rr, ri, F + 2 * st[h], len);
      for (int h = 0; h < 2; h++) orc_mult_cpx_vector(F + 2 * st[h], ts + 2 * st[h], F + 2 * st[h], len);
    }
  }
  free(tmp); free(ts);
}
