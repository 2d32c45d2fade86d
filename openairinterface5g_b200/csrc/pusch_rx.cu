// Single-layer PUSCH inner receiver: resource extraction, MRC channel compensation, QAM magnitude thresholds, max-log LLRs and
// (optionally) descrambling in ONE launch per slot, plus the channel-level measurement that fixes the compensation shift.
// Reference: openair1/PHY/NR_TRANSPORT/nr_ulsch_demodulation.c -- nr_ulsch_extract_rbs :279-380, nr_ulsch_scale_channel :382-414,
// get_nb_re_pusch :416-432, nr_ulsch_channel_level :434-466, nr_ulsch_channel_compensation :468-578, inner_rx :1262-1384,
// log2_maxh rule and llr_offset bookkeeping of nr_rx_pusch_tp :1595-1700, unscrambling in nr_pusch_symbol_processing :1430-1433.
// The reference walks symbol by symbol through four intermediate buffers (rxFext, chFext, rxdataF_comp, the three magnitude planes);
// here one thread owns one resource element of one symbol: it gathers y and h of every rx antenna straight from rxdataF /
// ul_ch_estimates (12 B per antenna and RE in, 2 * Qm B out), so nothing but the LLRs is ever written.  HBM bound.
#include "nrb200_ctx.h"
#include "gold_seq.cuh"
#include "../../include/nrb200_ldpc.h"
#include <climits>
#include <cstring>
#include <algorithm>

namespace nrb200 {

struct PuschGeom {
  int N, nb_rx, start_re, nb_re, Qm, dmrs_type, shift_from_dev, shift;
  unsigned rx_stride, ch_stride;
  int n_sym;                       // symbols with at least one valid RE
  int sym[14], ch_sym[14], is_dmrs[14], valid[14];
  unsigned llr_off[14];
  unsigned unscramble, c_init;
  int nl;                          // layers (1 | 2)
  int ue, cdm;                     // 1: the UE's PDSCH receiver (nr_rx_pdsch); cdm = n_dmrs_cdm_groups
  int last_is_dmrs, last_ch_sym, last_span;   // UE: the symbol whose magnitude buffers survive until the LLRs are computed
  unsigned nvar;                   // noise variance added to the diagonal of H^H H (2 layers)
  int lvl_amp, lvl_b;              // nr_ulsch_scale_channel constants of the level measurement
  const int *est_state;            // optional (device): the channel estimator's state, 18 int32 per port; replaces nvar / lvl_amp / lvl_b (est_scalars)
  int est_ports, est_div;          // ports in est_state; nr_of_symbols * nrOfLayers * nb_rx
  int tp_direct;                   // transform precoding with M = 1536 / 3072: plain [symbol][M] layout for idft(), no conjugation
};

__device__ __forceinline__ int p_sat16(int v) { return max(-32768, min(32767, v)); }
__device__ __forceinline__ int p_wrap16(int v) { return (int)(short)v; }
__device__ __forceinline__ int p_lo(unsigned w) { return (int)(short)(w & 0xFFFFu); }
__device__ __forceinline__ int p_hi(unsigned w) { return (int)(short)(w >> 16); }
__device__ __forceinline__ int p_abs16w(int v) { return v == -32768 ? -32768 : abs(v); }
__device__ __forceinline__ int p_subs16(int a, int b) { return max(-32768, min(32767, a - b)); }
__device__ __forceinline__ int p_mulhrs(int a, int b) { return p_wrap16((a * b + 0x4000) >> 15); }

// nvar and the nr_ulsch_scale_channel constants, from the descriptor or -- stream ordered behind the estimator, no host round trip -- from the
// estimator's device state: max_ch = max over ports, nvar = sum over ports / (symbols * layers * antennas) (nr_ulsch_demodulation.c:1470-1524),
// shift_ch_ext = log2_approx(max_ch >> 11) (:382-432)
__device__ __forceinline__ void est_scalars(const PuschGeom &G, unsigned &nvar, int &amp, int &b)
{
  nvar = G.nvar; amp = G.lvl_amp; b = G.lvl_b;
  if (G.est_state == nullptr) return;
  int mx = 0;
  unsigned long long nv = 0;
  for (int p = 0; p < G.est_ports; p++) { mx = max(mx, __ldg(G.est_state + 18 * p)); nv += (unsigned)__ldg(G.est_state + 18 * p + 1); }
  nvar = (unsigned)(nv / (unsigned long long)G.est_div);
  const unsigned v = (unsigned)mx >> 11;
  const int sce = v ? 32 - __clz(v) : 0;
  b = 3; amp = 8192;
  if (sce > 3) { b = 0; amp = (short)(amp >> (sce - 3)); if (amp == 0) amp = 1; } else b -= sce;
}

// i-th extracted RE of a symbol -> (sub-carrier in the symbol, index into the channel estimates), exactly the reference's loops
// (including the type-2 branch that forgets start_re when the allocation does not wrap, :345-351)
__device__ __forceinline__ void re_source(const PuschGeom &G, int is_dmrs, int i, int &rx_idx, int &ch_idx)
{
  const int N = G.N, s = G.start_re;
  if (!is_dmrs) { rx_idx = s + i; if (rx_idx >= N) rx_idx -= N; ch_idx = i; return; }
  const bool nowrap = s + G.nb_re < N;
  const int neg = N - s;
  if (G.dmrs_type == 0) {
    const int idx = 2 * i + 1;
    if (nowrap || idx < neg) { rx_idx = s + idx; ch_idx = idx; return; }
    const int n1 = neg >> 1;                     // odd indices below neg
    const int j = i - n1;
    rx_idx = 2 * j + 1; ch_idx = (neg | 1) + 2 * j;
    return;
  }
  int idx = 6 * (i >> 2) + 2 + (i & 3);
  if (nowrap) { rx_idx = idx; ch_idx = idx; return; }
  const int c1 = 4 * (neg / 6) + max(0, neg % 6 - 2);
  if (i < c1) { rx_idx = s + idx; ch_idx = idx; return; }
  const int j = i - c1;
  idx = 6 * (j >> 2) + 2 + (j & 3);
  rx_idx = idx; ch_idx = neg + idx;
}

// Transform precoding (DFT-s-OFDM, one layer, Qm <= 6; inner_rx :1326-1336): between the compensation and the LLRs the reference equalises the symbol
// (nr_freq_equalization, Qm > 2: every group of 4 REs is multiplied by 4096 / amp of the group's FIRST magnitude and shifted by 3, the thresholds become
// constants) and takes an M-point transform across the symbol's REs (nr_idft: conj -> four-way dft(DFT_M) with the data in lane 0 -> conj; M = 12 has its
// own scaling, M = 1536 / 3072 go through idft() directly).  The kernel runs twice around the library's own batched transform:
//   MODE 1: extraction + MRC + equalisation + conj, written into the transform's input layout -- symbol k is lane (k & 3) of four-way call (k >> 2), so
//           four symbols share one four-way transform instead of using one lane of four (the lanes are independent);
//   MODE 2: reads the transform's output, undoes the layout, conj, LLRs with the constant thresholds, descrambling.
// MODE 0 is the ordinary receiver.
__device__ __forceinline__ size_t tp_slot(const PuschGeom &G, int k, int i)
{
  return G.tp_direct ? (size_t)k * G.nb_re + i : (size_t)(k >> 2) * 4 * G.nb_re + 4 * (size_t)i + (k & 3);
}

template <int QM, int MODE>
__global__ void __launch_bounds__(256) pusch_rx_kernel(PuschGeom G, const GoldTables *__restrict__ T, const int *__restrict__ d_shift, const unsigned *__restrict__ rxF,
                                                       const unsigned *__restrict__ ch, short *__restrict__ llr, unsigned *__restrict__ tp)
{
  __shared__ uint32_t s_gold[(256 * QM) / 32 + 2];
  const int k = blockIdx.y, symbol = G.sym[k], valid = G.valid[k], is_dmrs = G.is_dmrs[k];
  const int i0 = blockIdx.x * 256, i = i0 + threadIdx.x;
  if (i0 >= valid) return;
  const unsigned bit0 = G.llr_off[k] + (unsigned)i0 * QM;          // first LLR (= scrambling bit) index handled by this CTA
  if (MODE != 1 && G.unscramble) {
    const unsigned w0 = bit0 >> 5, nw = ((bit0 + 256u * QM + 31u) >> 5) - w0;
    if (threadIdx.x < nw) s_gold[threadIdx.x] = gold_word(T, G.c_init, w0 + threadIdx.x);
    __syncthreads();
  }
  if (i >= valid) return;
  constexpr int ampa = QM == 4 ? 20724 : QM == 6 ? 20225 : QM == 8 ? 20106 : 0;     // QAM16_n1 / QAM64_n1 / QAM256_n1 (impl_defs_top.h:205-222)
  constexpr int ampb = QM == 6 ? 10112 : QM == 8 ? 10053 : 0;
  constexpr int ampc = QM == 8 ? 5026 : 0;
  int cr = 0, ci = 0, ma = 0, mb = 0, mc = 0;
  if (MODE == 2) {
    const unsigned z = tp[tp_slot(G, k, i)];
    cr = p_lo(z); ci = p_hi(z);
    if (G.nb_re == 12) { cr = p_wrap16(((cr * 9459) >> 16) << 1); ci = p_wrap16(((ci * 9459) >> 16) << 1); }   // DFT_12 is called unscaled: mulhi(9459) << 1 (:41-48)
    if (!G.tp_direct) ci = p_wrap16(-ci);                                              // conjugate output (:254-257)
    ma = QM == 4 ? 324 : 316; mb = 158;                                                 // 512 * 2 / sqrt(10); 512 * 4 / sqrt(42), 512 * 2 / sqrt(42)
  }
  const int shift = MODE == 2 ? 0 : (G.shift_from_dev ? *d_shift : G.shift);
  int rx_idx = 0, ch_idx = 0;
  if (MODE != 2) re_source(G, is_dmrs, i, rx_idx, ch_idx);
  for (int a = 0; MODE != 2 && a < G.nb_rx; a++) {
    const unsigned y = __ldg(rxF + (size_t)a * G.rx_stride + (size_t)symbol * G.N + rx_idx);
    const unsigned h = __ldg(ch + (size_t)a * G.ch_stride + (size_t)G.ch_sym[k] * G.N + ch_idx);
    const int hr = p_lo(h), hi = p_hi(h), yr = p_lo(y), yi = p_hi(y), nhi = p_wrap16(-hi);
    // madd_epi16 wraps in 32 bits, srai, packs_epi32 saturates; the MRC sum over antennas is add_epi16 (wraps)
    cr = p_wrap16(cr + p_sat16(((int)((unsigned)(hr * yr) + (unsigned)(hi * yi))) >> shift));
    ci = p_wrap16(ci + p_sat16(((int)((unsigned)(nhi * yr) + (unsigned)(hr * yi))) >> shift));
    if (QM > 2) {
      const int m = p_sat16(((int)((unsigned)(hr * hr) + (unsigned)(hi * hi))) >> shift);
      ma = p_wrap16(ma + p_mulhrs(m, ampa));
      if (QM > 4) mb = p_wrap16(mb + p_mulhrs(m, ampb));
      if (QM > 6) mc = p_wrap16(mc + p_mulhrs(m, ampc));
    }
  }
  if (MODE == 1) {
    if (QM > 2) {
      // the group's first magnitude: RE (i & ~3) sits in lane (lane & ~3) of this warp (valid is a multiple of 4, so groups are never split)
      int amp = __shfl_sync(__activemask(), ma, (threadIdx.x & 31) & ~3);
      amp = min(amp, 4095);
      const int inv = amp > 0 ? 4096 / amp : 0;                                         // nr_inv_ch[0] is 0; a negative amp is out of bounds in the reference
      cr = p_wrap16(cr * inv) >> 3; ci = p_wrap16(ci * inv) >> 3;                       // mullo_epi16, srai 3
    }
    if (!G.tp_direct) ci = p_wrap16(-ci);                                               // conjugate input (:27-31)
    tp[tp_slot(G, k, i)] = ((unsigned)cr & 0xFFFFu) | ((unsigned)ci << 16);
    return;
  }
  int o[8];
  if (QM == 2) { o[0] = cr >> 3; o[1] = ci >> 3; }
  else {
    o[0] = cr; o[1] = ci;
    o[2] = p_subs16(ma, p_abs16w(cr)); o[3] = p_subs16(ma, p_abs16w(ci));
    if (QM > 4) { o[4] = p_subs16(mb, p_abs16w(o[2])); o[5] = p_subs16(mb, p_abs16w(o[3])); }
    if (QM > 6) { o[6] = p_subs16(mc, p_abs16w(o[4])); o[7] = p_subs16(mc, p_abs16w(o[5])); }
  }
  const unsigned b = G.llr_off[k] + (unsigned)i * QM;
  if (G.unscramble) {
    const unsigned rel = b - ((bit0 >> 5) << 5);
#pragma unroll
    for (int m = 0; m < QM; m++) {
      const unsigned r = rel + m;
      if ((s_gold[r >> 5] >> (r & 31u)) & 1u) o[m] = p_wrap16(-o[m]);           // llr * s with s = -1: -32768 stays (int16 wrap)
    }
  }
  // Qm int16 per RE; b * 2 bytes is 4-byte aligned because Qm is even
  unsigned *dst = reinterpret_cast<unsigned *>(llr + b);
#pragma unroll
  for (int m = 0; m < QM / 2; m++) dst[m] = ((unsigned)o[2 * m] & 0xFFFFu) | ((unsigned)o[2 * m + 1] << 16);
}

// ---- two layers, MMSE (Qm >= 6): matched filter per layer, H^H H + nvar I from the same estimates, determinant, 2x2 adjugate, per-layer LLRs,
// layer de-mapping and descrambling (nr_ulsch_mmse_2layers :870-1260 and helpers, nr_pusch_symbol_processing :1420-1433).  One thread per RE;
// the scaling exponent b is shared by groups of 4 consecutive REs (one SSE vector in the reference): 4-lane shuffle sum.
__device__ __forceinline__ int p_log2_approx(unsigned v) { return v ? 32 - __clz(v) : 0; }      // bit length; the reference scans bits 0..30 only
__device__ __forceinline__ int p_mad(int ar, int br, int ai, int bi) { return (int)((unsigned)(ar * br) + (unsigned)(ai * bi)); }

template <int QM>
__global__ void __launch_bounds__(256) pusch_rx2_kernel(PuschGeom G, const GoldTables *__restrict__ T, const int *__restrict__ d_shift,
                                                        const unsigned *__restrict__ rxF, const unsigned *__restrict__ ch, short *__restrict__ llr)
{
  __shared__ uint32_t s_gold[(256 * 2 * QM) / 32 + 2];
  const int k = blockIdx.y, symbol = G.sym[k], valid = G.valid[k], is_dmrs = G.is_dmrs[k];
  const int i0 = blockIdx.x * 256, i = i0 + threadIdx.x;
  if (i0 >= valid) return;
  const unsigned bit0 = 2u * G.llr_off[k] + (unsigned)i0 * 2 * QM;      // both layers interleaved per RE after de-mapping
  if (G.unscramble) {
    const unsigned w0 = bit0 >> 5, nw = ((bit0 + 256u * 2 * QM + 31u) >> 5) - w0;
    if (threadIdx.x < nw) s_gold[threadIdx.x] = gold_word(T, G.c_init, w0 + threadIdx.x);
    __syncthreads();
  }
  const int shift = G.shift_from_dev ? *d_shift : G.shift;
  // number of REs the extraction writes (beyond it the reference's buffers hold zeros)
  const int n_ext = !is_dmrs ? G.nb_re : G.dmrs_type == 0 ? G.nb_re / 2 : (G.nb_re / 6) * 4;
  int c0r = 0, c0i = 0, c1r = 0, c1i = 0;
  int af[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};                       // 00, 01, 10, 11
  if (i < n_ext) {
    int rx_idx, ch_idx;
    re_source(G, is_dmrs, i, rx_idx, ch_idx);
    for (int a = 0; a < G.nb_rx; a++) {
      const unsigned y = __ldg(rxF + (size_t)a * G.rx_stride + (size_t)symbol * G.N + rx_idx);
      const unsigned h0 = __ldg(ch + (size_t)a * G.ch_stride + (size_t)G.ch_sym[k] * G.N + ch_idx);
      const unsigned h1 = __ldg(ch + (size_t)(G.nb_rx + a) * G.ch_stride + (size_t)G.ch_sym[k] * G.N + ch_idx);
      const int yr = p_lo(y), yi = p_hi(y), h0r = p_lo(h0), h0i = p_hi(h0), h1r = p_lo(h1), h1i = p_hi(h1);
      c0r = p_wrap16(c0r + p_sat16(p_mad(h0r, yr, h0i, yi) >> shift)); c0i = p_wrap16(c0i + p_sat16(p_mad(p_wrap16(-h0i), yr, h0r, yi) >> shift));
      c1r = p_wrap16(c1r + p_sat16(p_mad(h1r, yr, h1i, yi) >> shift)); c1i = p_wrap16(c1i + p_sat16(p_mad(p_wrap16(-h1i), yr, h1r, yi) >> shift));
      const int hr[2] = {h0r, h1r}, hi[2] = {h0i, h1i};
#pragma unroll
      for (int e = 0; e < 4; e++) {                                      // conj(h_first) * h_second, first = e >> 1, second = e & 1
        const int ar = hr[e >> 1], ai = hi[e >> 1], br = hr[e & 1], bi = hi[e & 1];
        const int re = p_sat16(p_mad(ar, br, ai, bi) >> shift), im = p_sat16(p_mad(p_wrap16(-ai), br, ar, bi) >> shift);
        if (a == 0) { af[e][0] = re; af[e][1] = im; } else { af[e][0] = p_sat16(af[e][0] + re); af[e][1] = p_sat16(af[e][1] + im); }
      }
    }
  }
  unsigned nvar; int lvl_amp_unused, lvl_b_unused;
  est_scalars(G, nvar, lvl_amp_unused, lvl_b_unused);
  if (nvar) {                                                            // add_epi32 on the packed {re, im} word (carries into im)
#pragma unroll
    for (int e = 0; e < 4; e += 3) {
      const unsigned w = (((unsigned)af[e][0] & 0xFFFFu) | ((unsigned)af[e][1] << 16)) + nvar;
      af[e][0] = p_lo(w); af[e][1] = p_hi(w);
    }
  }
  const int ad = p_mad(af[0][0], af[3][0], p_wrap16(-af[0][1]), af[3][1]);
  const int bc = p_mad(af[1][0], af[2][0], p_wrap16(-af[1][1]), af[2][1]);
  int det = (int)((unsigned)ad - (unsigned)bc);
  det = det == INT_MIN ? det : abs(det);
  int sum = det >> 2;
  sum = (int)((unsigned)sum + (unsigned)__shfl_xor_sync(0xffffffffu, sum, 1));
  sum = (int)((unsigned)sum + (unsigned)__shfl_xor_sync(0xffffffffu, sum, 2));
  if (i >= valid) return;
  const int b = p_log2_approx((unsigned)sum & 0x7FFFFFFFu) - 8;
  const int m = p_sat16(b > 0 ? det >> b : (int)((unsigned)det << (-b)));
  constexpr int ampa = QM == 6 ? 20225 : 20106, ampb = QM == 6 ? 10112 : 10053, ampc = QM == 8 ? 5026 : 0;
  const int ma = p_wrap16(((m * ampa) >> 16) << 1), mb = p_wrap16(((m * ampb) >> 16) << 1), mc = p_wrap16(((m * ampc) >> 16) << 1);
  // x0 = comp0 * d - comp1 * b ; x1 = comp1 * a - comp0 * c  (complex products, 32-bit wrap, shift by the group's exponent, pack)
  int xr[2], xi[2];
  {
    const int in[2][4][2] = {{{c0r, c0i}, {af[3][0], af[3][1]}, {c1r, c1i}, {af[1][0], af[1][1]}}, {{c1r, c1i}, {af[0][0], af[0][1]}, {c0r, c0i}, {af[2][0], af[2][1]}}};
#pragma unroll
    for (int l = 0; l < 2; l++) {
      int re = (int)((unsigned)p_mad(in[l][0][0], in[l][1][0], p_wrap16(-in[l][0][1]), in[l][1][1]) - (unsigned)p_mad(in[l][2][0], in[l][3][0], p_wrap16(-in[l][2][1]), in[l][3][1]));
      int im = (int)((unsigned)p_mad(in[l][0][1], in[l][1][0], in[l][0][0], in[l][1][1]) - (unsigned)p_mad(in[l][2][1], in[l][3][0], in[l][2][0], in[l][3][1]));
      if (b > 0) { re >>= b; im >>= b; } else { re = (int)((unsigned)re << (-b)); im = (int)((unsigned)im << (-b)); }
      xr[l] = p_sat16(re); xi[l] = p_sat16(im);
    }
  }
  const unsigned bb = 2u * G.llr_off[k] + (unsigned)i * 2 * QM;
  const unsigned rel = bb - ((bit0 >> 5) << 5);
#pragma unroll
  for (int l = 0; l < 2; l++) {
    int o[8];
    o[0] = xr[l]; o[1] = xi[l];
    o[2] = p_subs16(ma, p_abs16w(o[0])); o[3] = p_subs16(ma, p_abs16w(o[1]));
    o[4] = p_subs16(mb, p_abs16w(o[2])); o[5] = p_subs16(mb, p_abs16w(o[3]));
    if (QM > 6) { o[6] = p_subs16(mc, p_abs16w(o[4])); o[7] = p_subs16(mc, p_abs16w(o[5])); }
    if (G.unscramble) {
#pragma unroll
      for (int mm = 0; mm < QM; mm++) {
        const unsigned r = rel + l * QM + mm;
        if ((s_gold[r >> 5] >> (r & 31u)) & 1u) o[mm] = p_wrap16(-o[mm]);
      }
    }
    unsigned *dst = reinterpret_cast<unsigned *>(llr + bb + l * QM);
#pragma unroll
    for (int mm = 0; mm < QM / 2; mm++) dst[mm] = ((unsigned)o[2 * mm] & 0xFFFFu) | ((unsigned)o[2 * mm + 1] << 16);
  }
}

// ---- two layers, joint max-log ML detector (Qm < 6): nr_ulsch_compute_ML_llr (nr_ulsch_llr_computation.c:2100-2130) -> nr_ulsch_qpsk_qpsk (:375-525) and
// nr_ulsch_qam16_qam16 (:903-1135) on the matched-filter outputs, rho[l][1-l] (saturating sum over rx of conj(h_l) h_(1-l)) and, for 16QAM, ul_ch_maga of both
// layers (wrapping sum over rx of mulhrs(sat(|h|^2 >> shift), QAM16_n1)) as nr_ulsch_channel_compensation (:468-575) leaves them; QPSK LLRs >> 4
// (nr_ulsch_shift_llr :2064-2098); layer de-mapping and descrambling as in the MMSE kernel.  Element-wise int16 arithmetic, one thread per RE.  The x86
// build of the reference walks 16 REs per pass for floor(valid / 8) / 2 passes (rounded up): REs beyond that keep what the LLR buffer held -- zeros here.
__device__ __forceinline__ int m_mulhi(int a, int b) { return (a * b) >> 16; }
__device__ __forceinline__ int m_sll(int a, int n) { return p_wrap16(a << n); }
__device__ __forceinline__ int m_adds(int a, int b) { return p_sat16(a + b); }
__device__ __forceinline__ int m_subs(int a, int b) { return p_sat16(a - b); }
__device__ __forceinline__ int m_abs(int a) { return p_abs16w(a); }   // abs_epi16: -32768 stays.  NOT (short)(a < 0 ? -a : a): nvcc 12.9 folds that cast into a saturating conversion

__device__ __forceinline__ void ml_qpsk_qpsk(int y0r, int y0i, int y1r, int y1i, int rhor, int rhoi, int (&o)[4])
{
  const int y0r2 = m_sll(m_mulhi(y0r, 23170), 1), y0i2 = m_sll(m_mulhi(y0i, 23170), 1), y1r2 = y1r >> 1, y1i2 = y1i >> 1;
  const int rho_p = m_mulhi(m_adds(rhor, rhoi), 23170), rho_m = m_mulhi(m_subs(rhor, rhoi), 23170);
  const int rpm = m_abs(m_subs(rho_p, y1r2)), imm = m_abs(m_subs(rho_m, y1i2)), rmm = m_abs(m_subs(rho_m, y1r2)), ipm = m_abs(m_subs(rho_p, y1i2));
  const int rpp = m_abs(m_adds(rho_p, y1r2)), imp = m_abs(m_adds(rho_m, y1i2)), rmp = m_abs(m_adds(rho_m, y1r2)), ipp = m_abs(m_adds(rho_p, y1i2));
  const int pp = m_adds(m_adds(m_adds(rpm, imm), y0r2), y0i2);      // x = (+, +)
  const int pm = m_subs(m_adds(m_adds(rmm, ipp), y0r2), y0i2);      // x = (+, -)
  const int mp = m_adds(m_subs(m_adds(rmp, ipm), y0r2), y0i2);      // x = (-, +)
  const int mm = m_subs(m_subs(m_adds(rpp, imp), y0r2), y0i2);      // x = (-, -)
  o[0] = m_subs(max(pp, pm), max(mp, mm)) >> 4;
  o[1] = m_subs(max(pp, mp), max(pm, mm)) >> 4;
}

__device__ __forceinline__ void ml_qam16_qam16(int y0r, int y0i, int y1r, int y1i, int mag_des, int mag_int, int rr, int ri, int (&o)[4])
{
  constexpr int C10 = 20724, C10Q15 = 10362, C3 = 31086, CS = 25905, C9 = 23315;
  const int rpi = m_adds(rr, ri), rmi = m_subs(rr, ri);
  int rs[8], y0s[8], bm[16];
  rs[0] = m_mulhi(rpi, C10); rs[4] = m_mulhi(rmi, C10); rs[3] = m_sll(m_mulhi(rpi, C3), 1); rs[7] = m_sll(m_mulhi(rmi, C3), 1);
  const int x4 = m_mulhi(rr, C10), x5 = m_sll(m_mulhi(ri, C3), 1), x6 = m_sll(m_mulhi(rr, C3), 1), x7 = m_mulhi(ri, C10);
  rs[1] = m_adds(x4, x5); rs[5] = m_subs(x4, x5); rs[2] = m_adds(x6, x7); rs[6] = m_subs(x6, x7);
  const int y0r1 = m_mulhi(y0r, C10), y0i1 = m_mulhi(y0i, C10), y0r3 = m_sll(m_mulhi(y0r, C3), 1), y0i3 = m_sll(m_mulhi(y0i, C3), 1);
  y0s[0] = m_adds(y0r1, y0i1); y0s[4] = m_subs(y0r1, y0i1); y0s[1] = m_adds(y0r1, y0i3); y0s[5] = m_subs(y0r1, y0i3);
  y0s[2] = m_adds(y0r3, y0i1); y0s[6] = m_subs(y0r3, y0i1); y0s[3] = m_adds(y0r3, y0i3); y0s[7] = m_subs(y0r3, y0i3);
  const int cc[4] = {m_mulhi(mag_des, C10Q15), m_sll(m_mulhi(mag_des, CS), 1), m_sll(m_mulhi(mag_des, CS), 1), m_sll(m_mulhi(mag_des, C9), 2)};
  // the two possible values of square_a_epi16 (:741) per component: a = 1/sqrt(10) or 3/sqrt(10)
  const int sq1 = m_sll(m_mulhi(m_sll(m_mulhi(m_sll(m_mulhi(C10Q15, C10Q15), 1), CS), 1), mag_int), 1);
  const int sq3 = m_sll(m_mulhi(m_sll(m_mulhi(m_sll(m_mulhi(C3, C3), 1), CS), 1), mag_int), 1);
  constexpr int idx[16] = {4, 6, 5, 7, 0, 2, 1, 3, 0, 2, 1, 3, 4, 6, 5, 7};
#pragma unroll
  for (int j = 0; j < 16; j++) {
    const int psr = j < 8 ? m_abs(m_subs(rs[j], y1r)) : m_abs(m_adds(rs[(j - 4) & 7], y1r));
    const int psi = (j & 4) ? m_abs(m_adds(rs[idx[j]], y1i)) : m_abs(m_subs(rs[idx[j]], y1i));
    const bool lr = psr < mag_int, li = psi < mag_int;
    const int psa = m_adds(m_sll(m_mulhi(psr, lr ? C10Q15 : C3), 1), m_sll(m_mulhi(psi, li ? C10Q15 : C3), 1));
    const int t = m_subs(psa, m_adds(lr ? sq1 : sq3, li ? sq1 : sq3));
    bm[j] = j < 8 ? m_subs(m_adds(t, y0s[j]), cc[j & 3]) : m_subs(m_subs(t, y0s[(j + 4) & 7]), cc[j & 3]);
  }
#define NRB200_MX8(a, b, c, d, e, f, g, h) max(max(max(bm[a], bm[b]), max(bm[c], bm[d])), max(max(bm[e], bm[f]), max(bm[g], bm[h])))
  o[0] = m_subs(NRB200_MX8(0, 1, 2, 3, 4, 5, 6, 7), NRB200_MX8(8, 9, 10, 11, 12, 13, 14, 15));
  o[1] = m_subs(NRB200_MX8(0, 1, 3, 2, 8, 9, 10, 11), NRB200_MX8(4, 5, 6, 7, 12, 13, 14, 15));
  o[2] = m_subs(NRB200_MX8(0, 1, 4, 5, 8, 9, 12, 13), NRB200_MX8(2, 3, 6, 7, 10, 11, 14, 15));
  o[3] = m_subs(NRB200_MX8(0, 2, 4, 6, 8, 10, 12, 14), NRB200_MX8(1, 3, 5, 7, 9, 11, 13, 15));
#undef NRB200_MX8
}

template <int QM>
__global__ void __launch_bounds__(256) pusch_rx2ml_kernel(PuschGeom G, const GoldTables *__restrict__ T, const int *__restrict__ d_shift,
                                                          const unsigned *__restrict__ rxF, const unsigned *__restrict__ ch, short *__restrict__ llr)
{
  __shared__ uint32_t s_gold[(256 * 2 * QM) / 32 + 2];
  const int k = blockIdx.y, symbol = G.sym[k], valid = G.valid[k], is_dmrs = G.is_dmrs[k];
  const int i0 = blockIdx.x * 256, i = i0 + threadIdx.x;
  if (i0 >= valid) return;
  const unsigned bit0 = 2u * G.llr_off[k] + (unsigned)i0 * 2 * QM;
  if (G.unscramble) {
    const unsigned w0 = bit0 >> 5, nw = ((bit0 + 256u * 2 * QM + 31u) >> 5) - w0;
    if (threadIdx.x < nw) s_gold[threadIdx.x] = gold_word(T, G.c_init, w0 + threadIdx.x);
    __syncthreads();
  }
  if (i >= valid) return;
  const int shift = G.shift_from_dev ? *d_shift : G.shift;
  const int covered = 16 * (((valid >> 3) + 1) >> 1);
  int o[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
  if (i < covered) {
    const int n_ext = !is_dmrs ? G.nb_re : G.dmrs_type == 0 ? G.nb_re / 2 : (G.nb_re / 6) * 4;
    int c[2][2] = {{0, 0}, {0, 0}}, rho[2][2] = {{0, 0}, {0, 0}}, mg[2] = {0, 0};
    if (i < n_ext) {
      int rx_idx, ch_idx;
      re_source(G, is_dmrs, i, rx_idx, ch_idx);
      for (int a = 0; a < G.nb_rx; a++) {
        const unsigned y = __ldg(rxF + (size_t)a * G.rx_stride + (size_t)symbol * G.N + rx_idx);
        const unsigned hw[2] = {__ldg(ch + (size_t)a * G.ch_stride + (size_t)G.ch_sym[k] * G.N + ch_idx),
                                __ldg(ch + (size_t)(G.nb_rx + a) * G.ch_stride + (size_t)G.ch_sym[k] * G.N + ch_idx)};
        const int yr = p_lo(y), yi = p_hi(y);
#pragma unroll
        for (int l = 0; l < 2; l++) {
          const int hr = p_lo(hw[l]), hi = p_hi(hw[l]), gr = p_lo(hw[1 - l]), gi = p_hi(hw[1 - l]), nhi = p_wrap16(-hi);
          c[l][0] = p_wrap16(c[l][0] + p_sat16(p_mad(hr, yr, hi, yi) >> shift));
          c[l][1] = p_wrap16(c[l][1] + p_sat16(p_mad(nhi, yr, hr, yi) >> shift));
          rho[l][0] = p_sat16(rho[l][0] + p_sat16(p_mad(hr, gr, hi, gi) >> shift));
          rho[l][1] = p_sat16(rho[l][1] + p_sat16(p_mad(nhi, gr, hr, gi) >> shift));
          if (QM == 4) mg[l] = p_wrap16(mg[l] + p_wrap16((p_sat16(p_mad(hr, hr, hi, hi) >> shift) * 20724 + 0x4000) >> 15));
        }
      }
    }
#pragma unroll
    for (int l = 0; l < 2; l++) {
      if (QM == 2) ml_qpsk_qpsk(c[l][0], c[l][1], c[1 - l][0], c[1 - l][1], rho[l][0], rho[l][1], o[l]);
      else ml_qam16_qam16(c[l][0], c[l][1], c[1 - l][0], c[1 - l][1], mg[l], mg[1 - l], rho[l][0], rho[l][1], o[l]);
    }
  }
  const unsigned bb = 2u * G.llr_off[k] + (unsigned)i * 2 * QM;
  const unsigned rel = bb - ((bit0 >> 5) << 5);
#pragma unroll
  for (int l = 0; l < 2; l++) {
    if (G.unscramble) {
#pragma unroll
      for (int mm = 0; mm < QM; mm++) {
        const unsigned r = rel + l * QM + mm;
        if ((s_gold[r >> 5] >> (r & 31u)) & 1u) o[l][mm] = p_wrap16(-o[l][mm]);
      }
    }
    unsigned *dst = reinterpret_cast<unsigned *>(llr + bb + l * QM);
#pragma unroll
    for (int mm = 0; mm < QM / 2; mm++) dst[mm] = ((unsigned)o[l][2 * mm] & 0xFFFFu) | ((unsigned)o[l][2 * mm + 1] << 16);
  }
}

// ---- UE side, one layer: nr_rx_pdsch (NR_UE_TRANSPORT/nr_dlsch_demodulation.c:241-684).  Same structure as pusch_rx_kernel with the UE's arithmetic:
// estimates scaled (mulhi 8192, << 3) before the matched filter, per-antenna outputs packed and combined with SATURATING adds, thresholds mulhi << 1, and --
// because the reference computes the slot's LLRs after the last symbol with that call's local magnitude buffers -- the thresholds of EVERY symbol come from
// the LAST symbol's extraction (zero beyond its span).
__device__ __forceinline__ void ue_source(const PuschGeom &G, int is_dmrs, int i, int &rx_idx, int &ch_idx)
{
  const int N = G.N, s = G.start_re;
  if (!is_dmrs) { rx_idx = s + i; if (rx_idx >= N) rx_idx -= N; ch_idx = i; return; }
  int per, first;                                  // data REs per 6-RE group and the first of them (nr_dlsch_extract_rbs :1233-1297)
  if (G.dmrs_type == 0) { per = 3; first = 1; } else if (G.cdm == 1) { per = 4; first = 2; } else { per = 2; first = 4; }
  const int g = i / per, r = i - g * per;
  int k = s + 6 * g;
  if (k >= N) k -= N;
  const int o = G.dmrs_type == 0 ? 2 * r + first : r + first;
  rx_idx = k + o; ch_idx = 6 * g + o;
}
__device__ __forceinline__ unsigned ue_scale(unsigned h)
{
  return ((unsigned)(unsigned short)p_wrap16(((p_lo(h) * 8192) >> 16) << 3)) | ((unsigned)(unsigned short)p_wrap16(((p_hi(h) * 8192) >> 16) << 3) << 16);
}

// ---- PT-RS at the UE (nr_pdsch_ptrs_processing, NR_UE_ESTIMATION/nr_dl_channel_estimation.c:1765-1907; NR_REFSIG/ptrs_nr.c), one layer.
// The reference estimates a common phase error per PT-RS symbol from the compensated PT-RS REs, squeezes those REs out of rxdataF_comp, interpolates the
// estimates over the other symbols at the slot's last symbol, rotates every non-DMRS symbol and only then computes the slot's LLRs.  Here:
//   pdsch_ptrs_kernel  one CTA per symbol: matched filter + MRC of the symbol's PT-RS REs only (nb_rb / K of them), product with the conjugated QPSK
//                      pilot (Gold sequence of the symbol's PDSCH DMRS), exact integer sum across the CTA, thread 0 normalises in IEEE double arithmetic without
//                      fused multiply-adds (what oracle/_ref is built with).  32 words of state: raw estimates in [16..29].
//   pdsch_rx_kernel    PTRS = true: one thread per CTA runs nr_ptrs_process_slot's interpolation on the 14 raw estimates; output index i of a PT-RS symbol is mapped past the PT-RS REs before it, the MRC output is rotated by the symbol's
//                      phase (AVX2 body / scalar tail of rotate_cpx_vector by buffer position), thresholds stay at index i like the reference's unsqueezed
//                      magnitude buffers.  No intermediate buffer, still one pass over the slot.
struct PtrsGeom {
  int on, L, K12, q0, n;           // L = PTRSTimeDensity (log2); 12 K; first PT-RS RE of a symbol; PT-RS REs per PT-RS symbol
  int start, nsym;
  unsigned pos, dmrs_pos;          // PT-RS symbols (set_ptrs_symb_idx), DMRS symbols
  unsigned cinit[14];              // nr_gold_pdsch's c_init per symbol
  int ch_sym[14];                  // symbol of the estimates per symbol
  unsigned *state;                 // device, 32 words: [0..13] phase {re, im} packed, [14] nr_ptrs_process_slot's return value, [16..29] raw per-symbol estimates
};
// data RE i of a PT-RS symbol -> RE of the allocation (the PT-RS REs q0 + j * K12 skipped)
__device__ __forceinline__ int ptrs_unsqueeze(const PtrsGeom &T, int i)
{
  if (i < T.q0) return i;
  return i + min(T.n, (i - T.q0) / (T.K12 - 1) + 1);
}
__device__ __forceinline__ int p_d2i16(double v)                       // x86 cvttsd2si + truncation to 16 bits ("integer indefinite" has a zero low half)
{
  if (!(v > -2147483649.0 && v < 2147483648.0)) return 0;
  return p_wrap16((int)v);
}
// matched filter + saturating MRC of RE `re` of non-DMRS symbol `symbol` (the one-layer arithmetic of pdsch_rx_kernel)
__device__ __forceinline__ void ue_mrc(const PuschGeom &G, const unsigned *__restrict__ rxF, const unsigned *__restrict__ ch, int symbol, int chs, int re, int shift,
                                       int &cr, int &ci)
{
  int rx_idx, ch_idx;
  ue_source(G, 0, re, rx_idx, ch_idx);
  cr = 0; ci = 0;
  for (int a = 0; a < G.nb_rx; a++) {
    const unsigned y = __ldg(rxF + (size_t)a * G.rx_stride + (size_t)symbol * G.N + rx_idx);
    const unsigned h = ue_scale(__ldg(ch + (size_t)a * G.ch_stride + (size_t)chs * G.N + ch_idx));
    const int hr = p_lo(h), hi = p_hi(h), yr = p_lo(y), yi = p_hi(y);
    const int r = p_sat16(((int)((unsigned)(hr * yr) + (unsigned)(hi * yi))) >> shift), im = p_sat16(((int)((unsigned)(p_wrap16(-hi) * yr) + (unsigned)(hr * yi))) >> shift);
    if (a == 0) { cr = r; ci = im; } else { cr = p_sat16(cr + r); ci = p_sat16(ci + im); }
  }
}
__device__ __forceinline__ int ptrs_next(unsigned mask, int from, int end) { for (int s = from; s < end; s++) if ((mask >> s) & 1u) return s; return -1; }
__device__ __forceinline__ int ptrs_next_est(unsigned ptrs, unsigned dmrs, int from, int end)
{
  const int np = ptrs_next(ptrs, from, end), nd = ptrs_next(dmrs, from, end);
  if (nd == -1) return np;
  if (np == -1) return nd;
  return np > nd ? nd : np;
}
__device__ void ptrs_slope(int start, int end, const short *est, double *sl)
{
  const double distance = (double)(unsigned char)(end - start);
  sl[0] = (double)((int)est[2 * end] - (int)est[2 * start]) / distance;
  sl[1] = (double)((int)est[2 * end + 1] - (int)est[2 * start + 1]) / distance;
}
__device__ void ptrs_from_slope(short *est, const double *sl, int start, int end)
{
  for (int i = 1; i < end - start; i++) {
    est[2 * (start + i)] = (short)p_wrap16((int)est[2 * start] + p_d2i16(__dmul_rn((double)i, sl[0])));
    est[2 * (start + i) + 1] = (short)p_wrap16((int)est[2 * start + 1] + p_d2i16(__dmul_rn((double)i, sl[1])));
  }
}
// nr_ptrs_process_slot (ptrs_nr.c:281-337), control flow kept as written (8-bit symbol counters, leftRef = rightRef = 0 to begin with)
__device__ int ptrs_process_slot(unsigned dmrs, unsigned ptrs, short *est, int start, int nsym)
{
  double slope[2] = {0.0, 0.0};
  const int end = start + nsym;
  int right = 0, left = 0;
  for (int symb = start; symb < end; symb++) {
    if (((ptrs >> symb) & 1u) || ((dmrs >> symb) & 1u)) { left = symb; right = ptrs_next_est(ptrs, dmrs, symb + 1, end); continue; }
    if (symb == start && left == -1 && right == -1) return -1;
    if (right != -1 && ((dmrs >> right) & 1u)) {
      const int tmp = ptrs_next_est(ptrs, dmrs, right + 1, end);
      if (tmp != -1) ptrs_slope(right, tmp, est, slope);
      ptrs_from_slope(est, slope, left, right);
      symb = right - 1;
    } else if (right != -1 && ((ptrs >> right) & 1u)) {
      ptrs_slope(left, right, est, slope);
      ptrs_from_slope(est, slope, left, right);
      symb = right - 1;
    } else if (right == -1) {
      ptrs_from_slope(est, slope, symb - 1, end);
      symb = end;
    } else return -1;
  }
  return 0;
}
// One CTA per symbol of the slot (grid 14): the symbol's PT-RS REs are spread over the CTA's threads, the integer sums are reduced exactly, thread 0 normalises.
// state[16 + m] = the symbol's RAW estimate (DMRS symbols: 32767 + 0 j, others 0); the interpolation over the slot is redone by every CTA of the receiver kernel
// (a few hundred instructions of one thread) instead of a serial tail here, and CTA (0, 0) of the receiver publishes the final phasors in state[0..13], state[14].
constexpr int kPtrsTpb = 160;
__global__ void __launch_bounds__(kPtrsTpb) pdsch_ptrs_kernel(PuschGeom G, PtrsGeom T, const GoldTables *__restrict__ GT, const int *__restrict__ d_shift,
                                                              const unsigned *__restrict__ rxF, const unsigned *__restrict__ ch)
{
  __shared__ unsigned s_gw[12];
  __shared__ int s_sum[2][kPtrsTpb / 32];
  const int m = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool in_alloc = m >= T.start && m < T.start + T.nsym;
  if (!(in_alloc && ((T.pos >> m) & 1u))) {
    if (threadIdx.x == 0) T.state[16 + m] = (in_alloc && ((T.dmrs_pos >> m) & 1u)) ? 32767u : 0u;
    return;
  }
  const int shift = G.shift_from_dev ? *d_shift : G.shift;
  // the pilots are the first 2 n bits of the symbol's PDSCH-DMRS Gold sequence: at most 9 words for 138 PT-RS REs
  if (threadIdx.x < 12 && (int)threadIdx.x <= (2 * T.n - 1) >> 5) s_gw[threadIdx.x] = gold_word(GT, T.cinit[m], threadIdx.x);
  __syncthreads();
  int sr = 0, si = 0;
  for (int j = threadIdx.x; j < T.n; j += kPtrsTpb) {
    int cr, ci;
    ue_mrc(G, rxF, ch, m, T.ch_sym[m], T.q0 + j * T.K12, shift, cr, ci);
    const unsigned gw = s_gw[j >> 4];
    const int b0 = (gw >> ((2 * j) & 31)) & 1u, b1 = (gw >> ((2 * j + 1) & 31)) & 1u;
    const int pr = b0 ? -23170 : 23170, pi = b1 ? 23170 : -23170;          // nr_gen_ref_conj_symbols: conjugated QPSK (nr_dmrs_rx.c:54-55, :240-256)
    sr += p_sat16(((int)((unsigned)(cr * pr) + (unsigned)(p_wrap16(-ci) * pi))) >> 15);   // mult_cpx_vector, shift 15, packs (cmult_vv.c:96-156)
    si += p_sat16(((int)((unsigned)(ci * pr) + (unsigned)(cr * pi))) >> 15);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { sr += __shfl_xor_sync(0xffffffffu, sr, o); si += __shfl_xor_sync(0xffffffffu, si, o); }
  if (lane == 0) { s_sum[0][warp] = sr; s_sum[1][warp] = si; }
  __syncthreads();
  if (threadIdx.x == 0) {
    sr = 0; si = 0;
    for (int w = 0; w < kPtrsTpb / 32; w++) { sr += s_sum[0][w]; si += s_sum[1][w]; }
    const double sc = (double)T.n;
    const double real = (double)sr / sc, imag = (double)si / sc;
    const double ab = sqrt(__dadd_rn(__dmul_rn(real, real), __dmul_rn(imag, imag)));
    const int er = p_d2i16(__dmul_rn(real / ab, 32768.0)), ei = p_d2i16(__dmul_rn(-(imag / ab), 32768.0));
    T.state[16 + m] = ((unsigned)er & 0xFFFFu) | ((unsigned)ei << 16);
  }
}
// the slot's phasors from the raw per-symbol estimates: nr_ptrs_process_slot when PTRSTimeDensity > 0.  Returns its status; est = 14 {re, im}.
__device__ __forceinline__ int ptrs_finish(const PtrsGeom &T, short *est)
{
  for (int q = 0; q < 14; q++) { const unsigned v = T.state[16 + q]; est[2 * q] = (short)(v & 0xFFFFu); est[2 * q + 1] = (short)(v >> 16); }
  return T.L > 0 ? ptrs_process_slot(T.dmrs_pos, T.pos, est, T.start, T.nsym) : 0;
}

template <int QM, bool PTRS = false>
__global__ void __launch_bounds__(256) pdsch_rx_kernel(PuschGeom G, const GoldTables *__restrict__ T, const int *__restrict__ d_shift, const unsigned *__restrict__ rxF,
                                                       const unsigned *__restrict__ ch, short *__restrict__ llr, PtrsGeom PT)
{
  __shared__ uint32_t s_gold[(256 * QM) / 32 + 2];
  __shared__ unsigned s_phase;
  __shared__ int s_ptrs_ret;
  const int k = blockIdx.y, symbol = G.sym[k], valid = G.valid[k], is_dmrs = G.is_dmrs[k];
  const int i0 = blockIdx.x * 256, i = i0 + threadIdx.x;
  if (i0 >= valid) return;
  const unsigned bit0 = G.llr_off[k] + (unsigned)i0 * QM;
  if (PTRS && threadIdx.x == 255) {
    // every CTA redoes the slot's interpolation from the 14 raw estimates (cheap, and no serial tail in the estimator kernel); CTA (0, 0) publishes the result
    short est[28];
    const int ret = ptrs_finish(PT, est);
    s_phase = ((unsigned)(unsigned short)est[2 * symbol]) | ((unsigned)(unsigned short)est[2 * symbol + 1] << 16);
    s_ptrs_ret = ret;
    if (blockIdx.x == 0 && blockIdx.y == 0) {
      for (int q = 0; q < 14; q++) PT.state[q] = ((unsigned)(unsigned short)est[2 * q]) | ((unsigned)(unsigned short)est[2 * q + 1] << 16);
      PT.state[14] = (unsigned)ret;
    }
  }
  if (G.unscramble) {
    const unsigned w0 = bit0 >> 5, nw = ((bit0 + 256u * QM + 31u) >> 5) - w0;
    if (threadIdx.x < nw) s_gold[threadIdx.x] = gold_word(T, G.c_init, w0 + threadIdx.x);
  }
  if (G.unscramble || PTRS) __syncthreads();
  if (i >= valid) return;
  const int shift = G.shift_from_dev ? *d_shift : G.shift;
  int rx_idx, ch_idx, mch_idx, dummy;
  ue_source(G, is_dmrs, (PTRS && ((PT.pos >> symbol) & 1u)) ? ptrs_unsqueeze(PT, i) : i, rx_idx, ch_idx);
  ue_source(G, G.last_is_dmrs, i, dummy, mch_idx);
  const bool mag_ok = i < G.last_span;
  constexpr int ampa = QM == 4 ? 20724 : QM == 6 ? 20225 : QM == 8 ? 20106 : 0, ampb = QM == 6 ? 10112 : QM == 8 ? 10053 : 0, ampc = QM == 8 ? 5026 : 0;
  int cr = 0, ci = 0, ma = 0, mb = 0, mc = 0;
  for (int a = 0; a < G.nb_rx; a++) {
    const unsigned y = __ldg(rxF + (size_t)a * G.rx_stride + (size_t)symbol * G.N + rx_idx);
    const unsigned h = ue_scale(__ldg(ch + (size_t)a * G.ch_stride + (size_t)G.ch_sym[k] * G.N + ch_idx));
    const int hr = p_lo(h), hi = p_hi(h), yr = p_lo(y), yi = p_hi(y);
    const int r = p_sat16(((int)((unsigned)(hr * yr) + (unsigned)(hi * yi))) >> shift), im = p_sat16(((int)((unsigned)(p_wrap16(-hi) * yr) + (unsigned)(hr * yi))) >> shift);
    if (a == 0) { cr = r; ci = im; } else { cr = p_sat16(cr + r); ci = p_sat16(ci + im); }
    if (QM > 2 && mag_ok) {
      const unsigned hm = ue_scale(__ldg(ch + (size_t)a * G.ch_stride + (size_t)G.last_ch_sym * G.N + mch_idx));
      const int m = p_sat16(((int)((unsigned)(p_lo(hm) * p_lo(hm)) + (unsigned)(p_hi(hm) * p_hi(hm)))) >> shift);
      const int va = p_wrap16(((m * ampa) >> 16) << 1), vb = p_wrap16(((m * ampb) >> 16) << 1), vc = p_wrap16(((m * ampc) >> 16) << 1);
      if (a == 0) { ma = va; mb = vb; mc = vc; } else { ma = p_sat16(ma + va); mb = p_sat16(mb + vb); mc = p_sat16(mc + vc); }
    }
  }
  if (PTRS && !is_dmrs && s_ptrs_ret == 0) {
    // rotate_cpx_vector(rxdataF_comp of the symbol, phase, 12 * nb_rb, 15) (cmult_sv.c:77-145): madd + packs for whole groups of 8 REs, c16mulShift for the rest
    const unsigned ph = s_phase;
    const int ar = p_lo(ph), ai = p_hi(ph);
    int xr, xi;
    if (i < (G.nb_re & ~7)) {
      xr = p_sat16(((int)((unsigned)(cr * ar) + (unsigned)(ci * p_wrap16(-ai)))) >> 15);
      xi = p_sat16(((int)((unsigned)(cr * ai) + (unsigned)(ci * ar))) >> 15);
    } else {
      xr = p_wrap16(((int)((unsigned)(cr * ar) - (unsigned)(ci * ai))) >> 15);
      xi = p_wrap16(((int)((unsigned)(cr * ai) + (unsigned)(ci * ar))) >> 15);
    }
    cr = xr; ci = xi;
  }
  int o[8];
  if (QM == 2) { o[0] = cr >> 3; o[1] = ci >> 3; }
  else {
    o[0] = cr; o[1] = ci;
    o[2] = p_subs16(ma, p_abs16w(cr)); o[3] = p_subs16(ma, p_abs16w(ci));
    if (QM > 4) { o[4] = p_subs16(mb, p_abs16w(o[2])); o[5] = p_subs16(mb, p_abs16w(o[3])); }
    if (QM > 6) { o[6] = p_subs16(mc, p_abs16w(o[4])); o[7] = p_subs16(mc, p_abs16w(o[5])); }
  }
  const unsigned b = G.llr_off[k] + (unsigned)i * QM;
  if (G.unscramble) {
    const unsigned rel = b - ((bit0 >> 5) << 5);
#pragma unroll
    for (int m = 0; m < QM; m++) { const unsigned r = rel + m; if ((s_gold[r >> 5] >> (r & 31u)) & 1u) o[m] = p_wrap16(-o[m]); }
  }
  unsigned *dst = reinterpret_cast<unsigned *>(llr + b);
#pragma unroll
  for (int m = 0; m < QM / 2; m++) dst[m] = ((unsigned)o[2 * m] & 0xFFFFu) | ((unsigned)o[2 * m + 1] << 16);
}

// ---- UE side, two layers: matched filter + saturating MRC per layer, then nr_zero_forcing_rx (:1726-1869) per resource element: H^H H element by element
// (each antenna's product packed, antennas combined with saturating adds), determinant and adjugate with products shifted by log2_maxh - 2, the adjugate applied
// to the two matched-filter outputs, thresholds = determinant * QAM_amp (mulhi, << 1) -- of the LAST symbol, like the one-layer path -- and the LLRs written
// layer de-mapped (nr_dlsch_layer_demapping :1871-1907) and descrambled.  Estimates: plane layer * nb_rx + rx.
struct C16 { int r, i; };
__device__ __forceinline__ int p_sra(int v, int s) { return ((unsigned)s & 0xFFu) > 31u ? (v >> 31) : (v >> (s & 0xFF)); }
__device__ __forceinline__ C16 c_unpack(unsigned w) { return C16{p_lo(w), p_hi(w)}; }
__device__ __forceinline__ C16 c_conj0_mult1(C16 a, C16 b, int s)     // conj(a) * b >> s, packed (nr_conjch0_mult_ch1, nr_dlsch_channel_compensation)
{
  return C16{p_sat16(p_sra((int)((unsigned)(a.r * b.r) + (unsigned)(a.i * b.i)), s)), p_sat16(p_sra((int)((unsigned)(p_wrap16(-a.i) * b.r) + (unsigned)(a.r * b.i)), s))};
}
__device__ __forceinline__ C16 c_mult(C16 a, C16 b, int s)            // a * b >> s, packed (nr_a_mult_b)
{
  return C16{p_sat16(p_sra((int)((unsigned)(a.r * b.r) + (unsigned)(p_wrap16(-a.i) * b.i)), s)), p_sat16(p_sra((int)((unsigned)(a.i * b.r) + (unsigned)(a.r * b.i)), s))};
}
__device__ __forceinline__ C16 c_adds(C16 a, C16 b) { return C16{p_sat16(a.r + b.r), p_sat16(a.i + b.i)}; }
__device__ __forceinline__ C16 c_neg(C16 a) { return C16{p_wrap16(-a.r), p_wrap16(-a.i)}; }
// E[c][r] = sum over rx antennas of conj(H[r][a]) H[c][a] >> shift, H read (and scaled) from plane l * nb_rx + a at (symbol chs, index ci)
__device__ __forceinline__ void ue_gram(const PuschGeom &G, const unsigned *__restrict__ ch, int chs, int ci, int shift, C16 (&H)[2][4], C16 (&E)[2][2])
{
#pragma unroll
  for (int l = 0; l < 2; l++)
#pragma unroll
    for (int a = 0; a < 4; a++)
      if (a < G.nb_rx) H[l][a] = c_unpack(ue_scale(__ldg(ch + (size_t)(l * G.nb_rx + a) * G.ch_stride + (size_t)chs * G.N + ci)));
#pragma unroll
  for (int r = 0; r < 2; r++)
#pragma unroll
    for (int c = 0; c < 2; c++) {
      C16 acc = c_conj0_mult1(H[r][0], H[c][0], shift);
#pragma unroll
      for (int a = 1; a < 4; a++)
        if (a < G.nb_rx) acc = c_adds(acc, c_conj0_mult1(H[r][a], H[c][a], shift));
      E[c][r] = acc;
    }
}
__device__ __forceinline__ C16 ue_det(const C16 (&E)[2][2], int shift0) { return c_adds(c_mult(E[0][0], E[1][1], shift0), c_mult(E[0][1], c_neg(E[1][0]), shift0)); }

template <int QM>
__global__ void __launch_bounds__(256) pdsch_rx2_kernel(PuschGeom G, const GoldTables *__restrict__ T, const int *__restrict__ d_shift, const unsigned *__restrict__ rxF,
                                                        const unsigned *__restrict__ ch, short *__restrict__ llr)
{
  __shared__ uint32_t s_gold[(512 * QM) / 32 + 2];
  const int k = blockIdx.y, symbol = G.sym[k], valid = G.valid[k], is_dmrs = G.is_dmrs[k];
  const int i0 = blockIdx.x * 256, i = i0 + threadIdx.x;
  if (i0 >= valid) return;
  const unsigned bit0 = 2u * (G.llr_off[k] + (unsigned)i0 * QM);
  if (G.unscramble) {
    const unsigned w0 = bit0 >> 5, nw = ((bit0 + 512u * QM + 31u) >> 5) - w0;
    if (threadIdx.x < nw) s_gold[threadIdx.x] = gold_word(T, G.c_init, w0 + threadIdx.x);
    __syncthreads();
  }
  if (i >= valid) return;
  const int shift = G.shift_from_dev ? *d_shift : G.shift, shift0 = shift - 2;
  int rx_idx, ch_idx, mch_idx, dummy;
  ue_source(G, is_dmrs, i, rx_idx, ch_idx);
  ue_source(G, G.last_is_dmrs, i, dummy, mch_idx);
  C16 H[2][4], E[2][2], mf[2], out[2];
  ue_gram(G, ch, G.ch_sym[k], ch_idx, shift, H, E);
#pragma unroll
  for (int a = 0; a < 4; a++)
    if (a < G.nb_rx) {
      const C16 y = c_unpack(__ldg(rxF + (size_t)a * G.rx_stride + (size_t)symbol * G.N + rx_idx));
#pragma unroll
      for (int l = 0; l < 2; l++) { const C16 v = c_conj0_mult1(H[l][a], y, shift); mf[l] = a == 0 ? v : c_adds(mf[l], v); }
    }
  // adjugate: inv[r][c] = (-1)^(r + c) E[1 - c][1 - r]; layer r = sum over c of inv[c][r] * mf[c], accumulated from zero with saturating adds
  for (int r = 0; r < 2; r++) {
    C16 acc = c_adds(C16{0, 0}, c_mult(r == 0 ? E[1][1] : c_neg(E[0][1]), mf[0], shift0));
    out[r] = c_adds(acc, c_mult(r == 0 ? c_neg(E[1][0]) : E[0][0], mf[1], shift0));
  }
  int ma = 0, mb = 0, mc = 0;
  if (QM > 2 && i < G.last_span) {
    constexpr int ampa = QM == 4 ? 20724 : QM == 6 ? 20225 : QM == 8 ? 20106 : 0, ampb = QM == 6 ? 10112 : QM == 8 ? 10053 : 0, ampc = QM == 8 ? 5026 : 0;
    C16 Hm[2][4], Em[2][2];
    ue_gram(G, ch, G.last_ch_sym, mch_idx, shift, Hm, Em);
    const int det = ue_det(Em, shift0).r;
    ma = p_wrap16(((det * ampa) >> 16) << 1); mb = p_wrap16(((det * ampb) >> 16) << 1); mc = p_wrap16(((det * ampc) >> 16) << 1);
  }
  const unsigned b = 2u * (G.llr_off[k] + (unsigned)i * QM);
#pragma unroll
  for (int l = 0; l < 2; l++) {
    const int cr = out[l].r, ci = out[l].i;
    int o[8];
    if (QM == 2) { o[0] = cr >> 3; o[1] = ci >> 3; }
    else {
      o[0] = cr; o[1] = ci;
      o[2] = p_subs16(ma, p_abs16w(cr)); o[3] = p_subs16(ma, p_abs16w(ci));
      if (QM > 4) { o[4] = p_subs16(mb, p_abs16w(o[2])); o[5] = p_subs16(mb, p_abs16w(o[3])); }
      if (QM > 6) { o[6] = p_subs16(mc, p_abs16w(o[4])); o[7] = p_subs16(mc, p_abs16w(o[5])); }
    }
    const unsigned bl = b + (unsigned)l * QM;
    if (G.unscramble) {
      const unsigned rel = bl - ((bit0 >> 5) << 5);
#pragma unroll
      for (int m = 0; m < QM; m++) { const unsigned r = rel + m; if ((s_gold[r >> 5] >> (r & 31u)) & 1u) o[m] = p_wrap16(-o[m]); }
    }
    unsigned *dst = reinterpret_cast<unsigned *>(llr + bl);
#pragma unroll
    for (int m = 0; m < QM / 2; m++) dst[m] = ((unsigned)o[2 * m] & 0xFFFFu) | ((unsigned)o[2 * m + 1] << 16);
  }
}

// ---- UE side, three and four layers: the reference's receiver is generic in n_tx (nr_zero_forcing_rx "for 2, 3, and 4 Tx layers", :528; nr_determin :1460-1506 and
// nr_matrix_inverse :1549-1610 recurse over minors).  Same structure as pdsch_rx2_kernel with the Laplace expansion unrolled at compile time: det = sum over rows r of
// a[0][r] * det(minor(r, 0)), the sign (-1)^r handed down to the 1 x 1 leaves (nr_element_sign), every product nr_a_mult_b (>> shift0, packed), every sum saturating;
// inv[r][c] = det(minor(r, c)) with sign (-1)^(r + c); layer r = sum over c of inv[c][r] * mf[c].  a[c][r] is indexed [column][row] like the reference's a44.
template <int S> struct UeMat { C16 a[S][S]; };
template <int S>
__device__ __forceinline__ C16 ue_det_n(const UeMat<S> &A, int sign, int shift0)
{
  if constexpr (S == 1) return sign < 0 ? c_neg(A.a[0][0]) : A.a[0][0];
  else {
    C16 acc = C16{0, 0};
#pragma unroll
    for (int rtx = 0; rtx < S; rtx++) {
      UeMat<S - 1> sub;
#pragma unroll
      for (int ri = 0; ri < S - 1; ri++)
#pragma unroll
        for (int ci = 0; ci < S - 1; ci++) sub.a[ci][ri] = A.a[ci + 1][ri < rtx ? ri : ri + 1];
      const C16 prod = c_mult(A.a[0][rtx], ue_det_n<S - 1>(sub, ((rtx & 1) ? -1 : 1) * sign, shift0), shift0);
      acc = rtx == 0 ? prod : c_adds(acc, prod);
    }
    return acc;
  }
}
template <int S>
__device__ __forceinline__ C16 ue_cofactor(const UeMat<S> &A, int rtx, int ctx, int shift0)
{
  UeMat<S - 1> sub;
#pragma unroll
  for (int ri = 0; ri < S - 1; ri++)
#pragma unroll
    for (int ci = 0; ci < S - 1; ci++) sub.a[ci][ri] = A.a[ci < ctx ? ci : ci + 1][ri < rtx ? ri : ri + 1];
  return ue_det_n<S - 1>(sub, ((rtx & 1) ? -1 : 1) * ((ctx & 1) ? -1 : 1), shift0);
}
// E.a[c][r] = sum over rx antennas of conj(H[r][a]) H[c][a] >> shift (saturating sum of the packed per-antenna products)
template <int NL>
__device__ __forceinline__ void ue_gram_n(const PuschGeom &G, const unsigned *__restrict__ ch, int chs, int ci, int shift, C16 (&H)[NL][4], UeMat<NL> &E)
{
#pragma unroll
  for (int l = 0; l < NL; l++)
#pragma unroll
    for (int a = 0; a < 4; a++)
      if (a < G.nb_rx) H[l][a] = c_unpack(ue_scale(__ldg(ch + (size_t)(l * G.nb_rx + a) * G.ch_stride + (size_t)chs * G.N + ci)));
#pragma unroll
  for (int r = 0; r < NL; r++)
#pragma unroll
    for (int c = 0; c < NL; c++) {
      C16 acc = c_conj0_mult1(H[r][0], H[c][0], shift);
#pragma unroll
      for (int a = 1; a < 4; a++)
        if (a < G.nb_rx) acc = c_adds(acc, c_conj0_mult1(H[r][a], H[c][a], shift));
      E.a[c][r] = acc;
    }
}

template <int QM, int NL>
__global__ void __launch_bounds__(128) pdsch_rxn_kernel(PuschGeom G, const GoldTables *__restrict__ T, const int *__restrict__ d_shift, const unsigned *__restrict__ rxF,
                                                        const unsigned *__restrict__ ch, short *__restrict__ llr)
{
  constexpr int TPB = 128;
  __shared__ uint32_t s_gold[(TPB * NL * QM) / 32 + 2];
  const int k = blockIdx.y, symbol = G.sym[k], valid = G.valid[k], is_dmrs = G.is_dmrs[k];
  const int i0 = blockIdx.x * TPB, i = i0 + threadIdx.x;
  if (i0 >= valid) return;
  const unsigned bit0 = (unsigned)NL * (G.llr_off[k] + (unsigned)i0 * QM);
  if (G.unscramble) {
    const unsigned w0 = bit0 >> 5, nw = ((bit0 + (unsigned)(TPB * NL * QM) + 31u) >> 5) - w0;
    for (unsigned w = threadIdx.x; w < nw; w += TPB) s_gold[w] = gold_word(T, G.c_init, w0 + w);
    __syncthreads();
  }
  if (i >= valid) return;
  const int shift = G.shift_from_dev ? *d_shift : G.shift, shift0 = shift - 2;
  int rx_idx, ch_idx, mch_idx, dummy;
  ue_source(G, is_dmrs, i, rx_idx, ch_idx);
  ue_source(G, G.last_is_dmrs, i, dummy, mch_idx);
  C16 H[NL][4], mf[NL], out[NL];
  UeMat<NL> E;
  ue_gram_n<NL>(G, ch, G.ch_sym[k], ch_idx, shift, H, E);
#pragma unroll
  for (int a = 0; a < 4; a++)
    if (a < G.nb_rx) {
      const C16 y = c_unpack(__ldg(rxF + (size_t)a * G.rx_stride + (size_t)symbol * G.N + rx_idx));
#pragma unroll
      for (int l = 0; l < NL; l++) { const C16 v = c_conj0_mult1(H[l][a], y, shift); mf[l] = a == 0 ? v : c_adds(mf[l], v); }
    }
#pragma unroll
  for (int r = 0; r < NL; r++) {
    C16 acc = C16{0, 0};
#pragma unroll
    for (int c = 0; c < NL; c++) acc = c_adds(acc, c_mult(ue_cofactor<NL>(E, c, r, shift0), mf[c], shift0));     // inv[c][r] * mf[c]
    out[r] = acc;
  }
  int ma = 0, mb = 0, mc = 0;
  if (QM > 2 && i < G.last_span) {
    constexpr int ampa = QM == 4 ? 20724 : QM == 6 ? 20225 : QM == 8 ? 20106 : 0, ampb = QM == 6 ? 10112 : QM == 8 ? 10053 : 0, ampc = QM == 8 ? 5026 : 0;
    C16 Hm[NL][4];
    UeMat<NL> Em;
    ue_gram_n<NL>(G, ch, G.last_ch_sym, mch_idx, shift, Hm, Em);
    const int det = ue_det_n<NL>(Em, +1, shift0).r;
    ma = p_wrap16(((det * ampa) >> 16) << 1); mb = p_wrap16(((det * ampb) >> 16) << 1); mc = p_wrap16(((det * ampc) >> 16) << 1);
  }
  const unsigned b = (unsigned)NL * (G.llr_off[k] + (unsigned)i * QM);
#pragma unroll
  for (int l = 0; l < NL; l++) {
    const int cr = out[l].r, ci = out[l].i;
    int o[8];
    if (QM == 2) { o[0] = cr >> 3; o[1] = ci >> 3; }
    else {
      o[0] = cr; o[1] = ci;
      o[2] = p_subs16(ma, p_abs16w(cr)); o[3] = p_subs16(ma, p_abs16w(ci));
      if (QM > 4) { o[4] = p_subs16(mb, p_abs16w(o[2])); o[5] = p_subs16(mb, p_abs16w(o[3])); }
      if (QM > 6) { o[6] = p_subs16(mc, p_abs16w(o[4])); o[7] = p_subs16(mc, p_abs16w(o[5])); }
    }
    const unsigned bl = b + (unsigned)l * QM;
    if (G.unscramble) {
      const unsigned rel = bl - ((bit0 >> 5) << 5);
#pragma unroll
      for (int m = 0; m < QM; m++) { const unsigned r = rel + m; if ((s_gold[r >> 5] >> (r & 31u)) & 1u) o[m] = p_wrap16(-o[m]); }
    }
    unsigned *dst = reinterpret_cast<unsigned *>(llr + bl);
#pragma unroll
    for (int m = 0; m < QM / 2; m++) dst[m] = ((unsigned)o[2 * m] & 0xFFFFu) | ((unsigned)o[2 * m + 1] << 16);
  }
}

// UE: nr_dlsch_scale_channel + nr_dlsch_channel_level on the first symbol with data, log2_maxh = log2_approx(max avg) / 2 + 1 (:433-452)
// Two layers: one CTA per (layer, antenna) plane, and nr_dlsch_channel_level_median (:1144-1179) on top: (max + min) / 2 of the 4-RE power sums, both
// seeded with the plane's average; the plane's entry in d_out is then max(average, median), which is all the log2_maxh rule uses.
__global__ void __launch_bounds__(256) pdsch_level_kernel(PuschGeom G, int meas_k, const unsigned *__restrict__ ch, int *__restrict__ d_out, unsigned *__restrict__ d_count)
{
  __shared__ unsigned s_lane[4][64];
  __shared__ long long s_mx[256], s_mn[256];
  const int a = blockIdx.x, is_dmrs = G.is_dmrs[meas_k], len = G.valid[meas_k];
  int x = 0;
  while (x < 31 && !((len >> x) & 1)) x++;
  const int y = len >> x, span = (len / 12 + ((len % 12) ? 1 : 0)) * 12;
  const int n_ext = !is_dmrs ? G.nb_re : G.dmrs_type == 0 ? (G.cdm == 1 ? G.nb_re / 2 : 0) : (G.cdm == 1 ? (G.nb_re / 6) * 4 : G.cdm == 2 ? (G.nb_re / 6) * 2 : 0);
  // four 32-bit lanes like the SSE accumulator: lane = i & 3; thread t owns lane t & 3 (stride 256 keeps the lane)
  unsigned acc = 0;
  for (int i = threadIdx.x; i < min(n_ext, span); i += blockDim.x) {
    int rx_idx, ch_idx;
    ue_source(G, is_dmrs, i, rx_idx, ch_idx);
    const unsigned h = ue_scale(__ldg(ch + (size_t)a * G.ch_stride + (size_t)G.ch_sym[meas_k] * G.N + ch_idx));
    acc += (unsigned)(((int)((unsigned)(p_lo(h) * p_lo(h)) + (unsigned)(p_hi(h) * p_hi(h)))) >> x);
  }
  s_lane[threadIdx.x & 3][threadIdx.x >> 2] = acc;
  __syncthreads();
  if (threadIdx.x < 4) {
    unsigned s = 0;
    for (int j = 0; j < 64; j++) s += s_lane[threadIdx.x][j];
    s_lane[threadIdx.x][0] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const long long tot = (long long)(int)s_lane[0][0] + (int)s_lane[1][0] + (int)s_lane[2][0] + (int)s_lane[3][0];
    s_lane[0][1] = (unsigned)(int)(tot / y);
  }
  __syncthreads();
  const int avg = (int)s_lane[0][1];
  int lvl = avg;
  if (G.nl > 1) {
    long long mx = avg, mn = avg;
    for (int v = threadIdx.x; v < (len >> 2); v += blockDim.x) {
      long long sum = 0;
      for (int j = 0; j < 4; j++) {
        int rx_idx, ch_idx;
        ue_source(G, is_dmrs, 4 * v + j, rx_idx, ch_idx);
        const unsigned h = ue_scale(__ldg(ch + (size_t)a * G.ch_stride + (size_t)G.ch_sym[meas_k] * G.N + ch_idx));
        sum += ((int)((unsigned)(p_lo(h) * p_lo(h)) + (unsigned)(p_hi(h) * p_hi(h)))) >> 2;
      }
      mx = max(mx, sum); mn = min(mn, sum);
    }
    s_mx[threadIdx.x] = mx; s_mn[threadIdx.x] = mn;
    __syncthreads();
    for (int st = 128; st > 0; st >>= 1) {
      if (threadIdx.x < st) { s_mx[threadIdx.x] = max(s_mx[threadIdx.x], s_mx[threadIdx.x + st]); s_mn[threadIdx.x] = min(s_mn[threadIdx.x], s_mn[threadIdx.x + st]); }
      __syncthreads();
    }
    lvl = max(avg, (int)((s_mx[0] + s_mn[0]) >> 1));
  }
  if (threadIdx.x == 0) {
    // up to 4 layers x 4 antennas = 16 planes: the levels go through the launch's counter block (d_count[1 + plane]); d_out[0..7] mirrors the first eight
    reinterpret_cast<int *>(d_count)[1 + a] = lvl;
    if (a < 8) d_out[a] = lvl;
    __threadfence();
    if (atomicAdd(d_count, 1u) == (unsigned)(G.nb_rx * G.nl) - 1) {
      int avgs = 0;
      for (int q = 0; q < G.nb_rx * G.nl; q++) avgs = max(avgs, ((volatile int *)d_count)[1 + q]);
      const unsigned v = (unsigned)avgs & 0x7FFFFFFFu;
      d_out[8] = ((v ? 32 - __clz(v) : 0) / 2) + 1;
      *d_count = 0;
    }
  }
}

// nr_ulsch_scale_channel (shift_ch_ext = 0) + nr_ulsch_channel_level on the measurement symbol, one CTA per rx antenna, then the
// log2_maxh rule for one layer.  avg[a] and the final shift are left in d_out[0..nb_rx) and d_out[8].
__global__ void __launch_bounds__(256) pusch_level_kernel(PuschGeom G, int meas_k, int len, const unsigned *__restrict__ ch, int *__restrict__ d_out,
                                                          unsigned *__restrict__ d_count)
{
  __shared__ unsigned s_sum[256];
  const int a = blockIdx.x, is_dmrs = G.is_dmrs[meas_k];          // a = layer * nb_rx + rx
  int x = 0;
  while (x < 31 && !((len >> x) & 1)) x++;                       // factor2(len)
  const int y = len >> x;
  // number of REs the extraction writes for this symbol (everything beyond stays zero and adds nothing)
  const int n_ext = !is_dmrs ? G.nb_re : G.dmrs_type == 0 ? G.nb_re / 2 : (G.nb_re / 6) * 4;
  unsigned nvar_unused; int lvl_amp, lvl_b;
  est_scalars(G, nvar_unused, lvl_amp, lvl_b);
  unsigned acc = 0;
  for (int i = threadIdx.x; i < min(n_ext, len & ~3); i += blockDim.x) {
    int rx_idx, ch_idx;
    re_source(G, is_dmrs, i, rx_idx, ch_idx);
    const unsigned h = __ldg(ch + (size_t)a * G.ch_stride + (size_t)G.ch_sym[meas_k] * G.N + ch_idx);
    const int r = p_wrap16(((p_lo(h) * lvl_amp) >> 16) << lvl_b), im = p_wrap16(((p_hi(h) * lvl_amp) >> 16) << lvl_b);   // mulhi by ch_amp, slli b
    acc += (unsigned)(((int)((unsigned)(r * r) + (unsigned)(im * im))) >> x);
  }
  s_sum[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) { if (threadIdx.x < s) s_sum[threadIdx.x] += s_sum[threadIdx.x + s]; __syncthreads(); }
  if (threadIdx.x == 0) {
    d_out[a] = (int)s_sum[0] / y;
    __threadfence();
    if (atomicAdd(d_count, 1u) == (unsigned)(G.nb_rx * G.nl) - 1) {   // last (layer, antenna) pair: combine
      int avgs = 0;
      for (int k = 0; k < G.nb_rx * G.nl; k++) avgs = max(avgs, ((volatile int *)d_out)[k]);
      auto l2 = [](unsigned v) { return v ? 32 - __clz(v) : 0; };  // log2_approx: bit length (values < 2^31)
      int l = G.nl == 2 ? (l2((unsigned)avgs) >> 1) - (G.Qm >= 6 ? 3 : 0) : (l2((unsigned)avgs) >> 1) + 1 + l2((unsigned)G.nb_rx >> 2);   // - 3: MMSE only (:1640)
      d_out[8] = l < 0 ? 0 : l;
      *d_count = 0;
    }
  }
}

static int nb_re_symbol(const nrb200_pusch_rx_t &d, int symbol)
{
  if ((d.ul_dmrs_symb_pos >> symbol) & 1) return d.rb_size * (12 - d.num_dmrs_cdm_grps_no_data * (d.dmrs_config_type == 0 ? 6 : 4));
  return d.rb_size * 12;
}

// set_ptrs_symb_idx (ptrs_nr.c:53-86)
static unsigned ptrs_symbol_mask(int start_symbol, int duration, int L_ptrs, unsigned dmrs_pos)
{
  unsigned out = 0;
  int i = 0, l_ref = start_symbol;
  const int last = start_symbol + duration - 1;
  while (l_ref + i * L_ptrs <= last) {
    int is_dmrs = 0, l;
    const int lo = std::max(l_ref + (i - 1) * L_ptrs + 1, l_ref);
    for (l = l_ref + i * L_ptrs; l >= lo; l--) if ((dmrs_pos >> l) & 1u) { is_dmrs = 1; break; }
    if (is_dmrs) { l_ref = l; i = 1; continue; }
    out |= 1u << (l_ref + i * L_ptrs);
    i++;
  }
  return out;
}
// PT-RS geometry of a descriptor (one layer at the UE only): 0 = no PT-RS, 1 = filled, < 0 = not a configuration the library reproduces
static int make_ptrs(const nrb200_pusch_rx_t &d, PtrsGeom *T)
{
  std::memset(T, 0, sizeof(*T));
  if (!d.ptrs) return 0;
  const int nl = d.nrOfLayers == 0 ? 1 : (int)d.nrOfLayers;
  // nr_pdsch_ptrs_processing squeezes and rotates plane [0][aarx] only: with two layers the second layer would be left as it is (not reproduced);
  // the gNB side has no input -> output function (defect 19)
  if (!d.pdsch_ue || nl != 1) return -4;
  const int K = (int)d.ptrs_freq_density, nb = (int)d.rb_size;
  if ((K != 2 && K != 4) || d.ptrs_time_density > 2 || d.ptrs_re_offset >= 12 || d.ptrs_nscid > 1 || d.ptrs_slot >= 160) return -4;
  const int k_rb_ref = (nb % K == 0) ? (int)(d.rnti & 0xFFFFu) % K : (int)(d.rnti & 0xFFFFu) % (nb % K);          // is_ptrs_subcarrier (ptrs_nr.c:107-129)
  T->on = 1; T->L = (int)d.ptrs_time_density; T->K12 = 12 * K; T->q0 = (int)d.ptrs_re_offset + 12 * k_rb_ref;
  T->n = T->q0 < 12 * nb ? (12 * nb - 1 - T->q0) / T->K12 + 1 : 0;
  if (T->n < 1) return -4;
  T->start = (int)d.start_symbol_index; T->nsym = (int)d.nr_of_symbols; T->dmrs_pos = d.ul_dmrs_symb_pos;
  T->pos = ptrs_symbol_mask(T->start, T->nsym, 1 << T->L, d.ul_dmrs_symb_pos);
  const uint64_t nid = d.ptrs_dmrs_scrambling_id & 0xFFFFu;
  for (int l = 0; l < 14; l++) {                                          // nr_gold_pdsch (nr_gold_ue.c:75-93)
    const uint64_t x2tmp0 = ((uint64_t)(14 * d.ptrs_slot + l + 1) * ((nid << 1) + 1)) << 17;
    T->cinit[l] = (uint32_t)((x2tmp0 + (nid << 1) + d.ptrs_nscid) % (1ull << 31));
    int chs = -1;                                                         // get_valid_dmrs_idx_for_channel_est
    for (int q = l; q >= 0 && chs < 0; q--) if ((d.ul_dmrs_symb_pos >> q) & 1) chs = q;
    for (int q = l; q < 14 && chs < 0; q++) if ((d.ul_dmrs_symb_pos >> q) & 1) chs = q;
    T->ch_sym[l] = chs < 0 ? 0 : chs;
  }
  T->state = reinterpret_cast<unsigned *>((uintptr_t)d.d_ptrs_state);
  return 1;
}

static int make_geom(const nrb200_pusch_rx_t &d, PuschGeom *G, uint32_t *total_llr)
{
  const int Qm = d.qam_mod_order;
  const int nl = d.nrOfLayers == 0 ? 1 : (int)d.nrOfLayers;
  PtrsGeom PT;
  if (make_ptrs(d, &PT) < 0) return -4;
  if (nl > (d.pdsch_ue ? 4 : 2) || (nl == 2 && !d.pdsch_ue && Qm >= 6 && d.nb_rx != 2 && d.nb_rx != 4)) return -4;   // 2 layers: MMSE receiver (Qm >= 6; 2 or 4 rx like the reference), joint ML below
  if ((Qm != 2 && Qm != 4 && Qm != 6 && Qm != 8) || d.nb_rx < 1 || d.nb_rx > 8 || d.rb_size < 1 || d.fft_size < 12 * d.rb_size ||
      d.start_symbol_index + d.nr_of_symbols > 14 || d.dmrs_config_type > 1 || (d.log2_maxh > 31 && d.log2_maxh != 0xFFFFFFFFu))
    return -4;
  G->N = d.fft_size; G->nb_rx = d.nb_rx; G->nb_re = 12 * d.rb_size; G->Qm = Qm; G->dmrs_type = d.dmrs_config_type;
  G->start_re = (d.first_carrier_offset + (d.rb_start + d.bwp_start) * 12) % d.fft_size;
  G->shift = d.log2_maxh; G->shift_from_dev = 0;
  G->rx_stride = d.rx_stride; G->ch_stride = d.ch_stride;
  G->unscramble = d.unscramble; G->c_init = (d.rnti << 15) + d.data_scrambling_id;
  G->nl = nl; G->nvar = d.noise_var;
  G->est_state = nullptr; G->est_ports = 0; G->est_div = 1;
  if (d.d_est_state != 0) {
    if (nl != 2 || d.pdsch_ue || d.est_state_ports < 1 || d.est_state_ports > 2) return -4;
    G->est_state = reinterpret_cast<const int *>((uintptr_t)d.d_est_state);
    G->est_ports = (int)d.est_state_ports; G->est_div = (int)(d.nr_of_symbols * nl * d.nb_rx);
  }
  G->ue = d.pdsch_ue ? 1 : 0; G->cdm = d.num_dmrs_cdm_grps_no_data; G->tp_direct = 0;
  if (G->ue && (d.nb_rx > 4 || (nl >= 2 && d.nb_rx < 2))) return -4;     // the reference applies neither MRC nor zero forcing with one rx antenna
  {
    // nr_ulsch_scale_channel: shift_ch_ext = log2_approx(max_ch >> 11) for 2 layers, 0 for one
    int sce = 0;
    if (nl > 1) { const unsigned v = d.max_ch >> 11; sce = v ? 32 - __builtin_clz(v) : 0; }
    int b = 3, amp = 8192;
    if (sce > 3) { b = 0; amp = (short)(amp >> (sce - 3)); if (amp == 0) amp = 1; } else b -= sce;
    G->lvl_amp = amp; G->lvl_b = b;
  }
  int first_dmrs = -1;
  for (uint32_t s = d.start_symbol_index; s < d.start_symbol_index + d.nr_of_symbols; s++)
    if ((d.ul_dmrs_symb_pos >> s) & 1) { first_dmrs = s; break; }
  if (first_dmrs < 0) return -4;
  G->n_sym = 0;
  unsigned off = 0;
  int cur = first_dmrs;
  for (uint32_t s = d.start_symbol_index; s < d.start_symbol_index + d.nr_of_symbols; s++) {
    const int dm = (d.ul_dmrs_symb_pos >> s) & 1;
    if (dm) cur = s;                                            // nr_pusch_symbol_processing :1398-1404
    int chs = cur;
    if (G->ue) {                                                // get_valid_dmrs_idx_for_channel_est (dmrs_nr.c:321-340): this, previous, else next DMRS symbol
      chs = -1;
      for (int q = (int)s; q >= 0 && chs < 0; q--) if ((d.ul_dmrs_symb_pos >> q) & 1) chs = q;
      for (int q = (int)s; q < 14 && chs < 0; q++) if ((d.ul_dmrs_symb_pos >> q) & 1) chs = q;
    }
    const int v0 = nb_re_symbol(d, s);                                   // what the symbol's extraction / magnitude buffers span
    int v = v0;
    if (PT.on && ((PT.pos >> s) & 1u)) v -= PT.n;                     // dl_valid_re[symbol] -= ptrs_re_per_slot[0][symbol] (nr_dlsch_demodulation.c:573)
    if (v > 0) {
      const int k = G->n_sym++;
      G->sym[k] = s; G->ch_sym[k] = chs; G->is_dmrs[k] = dm; G->valid[k] = v; G->llr_off[k] = off;
    }
    off += (unsigned)v * Qm;
    if (s == d.start_symbol_index + d.nr_of_symbols - 1) { G->last_is_dmrs = dm; G->last_ch_sym = chs; G->last_span = v0; }   // thresholds exist for the last symbol's EXTRACTED REs only: the reference's vectors run on to whole PRBs, but over zero-padded estimates (threshold 0, like beyond the span)
  }
  if (total_llr) *total_llr = off * (unsigned)nl;
  return G->n_sym > 0 ? 0 : -4;
}

int pusch_ptrs_layout(const nrb200_pusch_rx_t &d, uint32_t *mask, uint32_t *n_re)
{
  PtrsGeom T;
  const int rc = make_ptrs(d, &T);
  if (rc < 0) return rc;
  if (mask) *mask = T.on ? T.pos : 0u;
  if (n_re) *n_re = T.on ? (uint32_t)T.n : 0u;
  return 0;
}

uint32_t pusch_num_llr(const nrb200_pusch_rx_t &d)
{
  PuschGeom G;
  uint32_t n = 0;
  return make_geom(d, &G, &n) == 0 ? n : 0;
}

int launch_pusch_level(const nrb200_pusch_rx_t &d, const int16_t *ch, int32_t *d_out9, uint32_t *d_count, cudaStream_t st)
{
  PuschGeom G;
  int rc = make_geom(d, &G, nullptr);
  if (rc) return rc;
  if (d.ptrs) {                                                   // the level is measured on the symbol as extracted, PT-RS REs included (nr_dlsch_demodulation.c:436-470)
    nrb200_pusch_rx_t e = d;
    e.ptrs = 0;
    if ((rc = make_geom(e, &G, nullptr)) != 0) return rc;
  }
  if (G.ue) {
    pdsch_level_kernel<<<G.nb_rx * G.nl, 256, 0, st>>>(G, 0, (const unsigned *)ch, d_out9, d_count);
    ctx().launches++;
    NRB200_CUDA_OK(cudaGetLastError(), "pdsch_level launch");
    return 0;
  }
  const int len = (G.valid[0] + 15) & ~15;                      // first symbol with data (:1601-1611)
  pusch_level_kernel<<<G.nb_rx * G.nl, 256, 0, st>>>(G, 0, len, (const unsigned *)ch, d_out9, d_count);
  ctx().launches++;
  NRB200_CUDA_OK(cudaGetLastError(), "pusch_level launch");
  return 0;
}

// nr_idft's sizes (nr_ulsch_demodulation.c:38-246) without 768 and 2304: there the reference hands four-way data to the single-transform dft768 / to
// dft2304 (which combines uninitialised stack), so its own output is not reproducible
bool pusch_tp_supported(int M)
{
  static const int sizes[] = {12, 24, 36, 48, 60, 72, 96, 108, 120, 144, 180, 192, 216, 240, 288, 300, 324, 360, 384, 432, 480, 540, 576, 600, 648, 720, 864, 900, 960, 972,
                              1080, 1152, 1200, 1296, 1440, 1500, 1536, 1620, 1728, 1800, 1920, 1944, 2160, 2400, 2592, 2700, 2880, 2916, 3000, 3072, 3240};
  for (int v : sizes) if (v == M) return true;
  return false;
}
size_t pusch_tp_scratch_bytes(const nrb200_pusch_rx_t &d) { return (size_t)128 * 12 * d.rb_size; }   // two planes of 16 M c16 (input and output of the transforms)

int dft_batch_internal(int N, int inverse, uint32_t n, const int16_t *d_in, int16_t *d_out, int scale, cudaStream_t st);   // dfts_internal.cu

int launch_pusch_rx(const nrb200_pusch_rx_t &d, const int16_t *rxF, const int16_t *ch, const int32_t *d_shift, int16_t *llr, cudaStream_t st)
{
  PuschGeom G;
  int rc = make_geom(d, &G, nullptr);
  if (rc) return rc;
  if (d.unscramble && scramble_mod_init() != 0) return -5;
  G.shift_from_dev = d_shift != nullptr;
  const GoldTables *T = d.unscramble ? gold_tables_dev() : nullptr;
  int vmax = 0;
  for (int k = 0; k < G.n_sym; k++) vmax = std::max(vmax, G.valid[k]);
  const dim3 grid((((vmax + 3) & ~3) + 255) / 256, G.n_sym);
  const unsigned *R = (const unsigned *)rxF, *C = (const unsigned *)ch;
  if (G.ue && G.nl > 2) {
    const dim3 gridn((((vmax + 3) & ~3) + 127) / 128, G.n_sym);
#define NRB200_RXN(QM_) do { if (G.nl == 3) pdsch_rxn_kernel<QM_, 3><<<gridn, 128, 0, st>>>(G, T, d_shift, R, C, llr); \
                             else pdsch_rxn_kernel<QM_, 4><<<gridn, 128, 0, st>>>(G, T, d_shift, R, C, llr); } while (0)
    switch (G.Qm) { case 2: NRB200_RXN(2); break; case 4: NRB200_RXN(4); break; case 6: NRB200_RXN(6); break; default: NRB200_RXN(8); break; }
#undef NRB200_RXN
  } else if (G.ue && G.nl == 2) {
    switch (G.Qm) {
      case 2: pdsch_rx2_kernel<2><<<grid, 256, 0, st>>>(G, T, d_shift, R, C, llr); break;
      case 4: pdsch_rx2_kernel<4><<<grid, 256, 0, st>>>(G, T, d_shift, R, C, llr); break;
      case 6: pdsch_rx2_kernel<6><<<grid, 256, 0, st>>>(G, T, d_shift, R, C, llr); break;
      default: pdsch_rx2_kernel<8><<<grid, 256, 0, st>>>(G, T, d_shift, R, C, llr); break;
    }
  } else if (G.ue) {
    PtrsGeom PT;
    const int pt = make_ptrs(d, &PT);
    if (pt < 0) return pt;
    if (pt == 1) {
      if (PT.state == nullptr || scramble_mod_init() != 0) return PT.state == nullptr ? -4 : -5;
      pdsch_ptrs_kernel<<<14, kPtrsTpb, 0, st>>>(G, PT, gold_tables_dev(), d_shift, R, C);
      NRB200_CUDA_OK(cudaGetLastError(), "pdsch_ptrs launch");
      ctx().launches++;
      switch (G.Qm) {
        case 2: pdsch_rx_kernel<2, true><<<grid, 256, 0, st>>>(G, T, d_shift, R, C, llr, PT); break;
        case 4: pdsch_rx_kernel<4, true><<<grid, 256, 0, st>>>(G, T, d_shift, R, C, llr, PT); break;
        case 6: pdsch_rx_kernel<6, true><<<grid, 256, 0, st>>>(G, T, d_shift, R, C, llr, PT); break;
        default: pdsch_rx_kernel<8, true><<<grid, 256, 0, st>>>(G, T, d_shift, R, C, llr, PT); break;
      }
    } else {
      switch (G.Qm) {
        case 2: pdsch_rx_kernel<2><<<grid, 256, 0, st>>>(G, T, d_shift, R, C, llr, PT); break;
        case 4: pdsch_rx_kernel<4><<<grid, 256, 0, st>>>(G, T, d_shift, R, C, llr, PT); break;
        case 6: pdsch_rx_kernel<6><<<grid, 256, 0, st>>>(G, T, d_shift, R, C, llr, PT); break;
        default: pdsch_rx_kernel<8><<<grid, 256, 0, st>>>(G, T, d_shift, R, C, llr, PT); break;
      }
    }
  } else if (G.nl == 2) {
    if (G.Qm == 2) pusch_rx2ml_kernel<2><<<grid, 256, 0, st>>>(G, T, d_shift, R, C, llr);
    else if (G.Qm == 4) pusch_rx2ml_kernel<4><<<grid, 256, 0, st>>>(G, T, d_shift, R, C, llr);
    else if (G.Qm == 6) pusch_rx2_kernel<6><<<grid, 256, 0, st>>>(G, T, d_shift, R, C, llr);
    else pusch_rx2_kernel<8><<<grid, 256, 0, st>>>(G, T, d_shift, R, C, llr);
  } else if (d.transform_precoding && G.Qm <= 6) {
    // inner_rx applies the equalisation / nr_idft step to one layer and Qm <= 6 only (:1326); 256QAM and two layers take the ordinary path like the reference
    const int M = G.nb_re;
    if (!pusch_tp_supported(M) || d.d_tp_scratch == 0) return -4;
    for (int k = 0; k < G.n_sym; k++) if (G.valid[k] != M) return -4;           // data on a DMRS symbol: nr_idft has no such size
    G.tp_direct = (M == 1536 || M == 3072) ? 1 : 0;
    unsigned *tin = reinterpret_cast<unsigned *>((uintptr_t)d.d_tp_scratch), *tout = tin + 16 * (size_t)M;
    switch (G.Qm) {
      case 2: pusch_rx_kernel<2, 1><<<grid, 256, 0, st>>>(G, T, d_shift, R, C, llr, tin); break;
      case 4: pusch_rx_kernel<4, 1><<<grid, 256, 0, st>>>(G, T, d_shift, R, C, llr, tin); break;
      default: pusch_rx_kernel<6, 1><<<grid, 256, 0, st>>>(G, T, d_shift, R, C, llr, tin); break;
    }
    NRB200_CUDA_OK(cudaGetLastError(), "pusch_rx (transform precoding, stage 1) launch");
    rc = G.tp_direct ? dft_batch_internal(M, 1, (uint32_t)G.n_sym, (const int16_t *)tin, (int16_t *)tout, 1, st)
                     : dft_batch_internal(M, 0, (uint32_t)((G.n_sym + 3) / 4), (const int16_t *)tin, (int16_t *)tout, M == 12 ? 0 : 1, st);
    if (rc) return rc;
    switch (G.Qm) {
      case 2: pusch_rx_kernel<2, 2><<<grid, 256, 0, st>>>(G, T, d_shift, R, C, llr, tout); break;
      case 4: pusch_rx_kernel<4, 2><<<grid, 256, 0, st>>>(G, T, d_shift, R, C, llr, tout); break;
      default: pusch_rx_kernel<6, 2><<<grid, 256, 0, st>>>(G, T, d_shift, R, C, llr, tout); break;
    }
    ctx().launches += 2;
  } else {
    switch (G.Qm) {
      case 2: pusch_rx_kernel<2, 0><<<grid, 256, 0, st>>>(G, T, d_shift, R, C, llr, nullptr); break;
      case 4: pusch_rx_kernel<4, 0><<<grid, 256, 0, st>>>(G, T, d_shift, R, C, llr, nullptr); break;
      case 6: pusch_rx_kernel<6, 0><<<grid, 256, 0, st>>>(G, T, d_shift, R, C, llr, nullptr); break;
      default: pusch_rx_kernel<8, 0><<<grid, 256, 0, st>>>(G, T, d_shift, R, C, llr, nullptr); break;
    }
  }
  ctx().launches++;
  NRB200_CUDA_OK(cudaGetLastError(), "pusch_rx launch");
  return 0;
}

}  // namespace nrb200
