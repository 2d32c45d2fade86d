// PUSCH channel estimation.  Main path: DMRS configuration type 1, frequency-domain interpolation (the reference's default, chest_freq == 0);
// the other three branches of the reference function (type 2, and chest_freq == 1 for both types) are the "variants" further down, and so are the UE's
// (nr_pdsch_channel_estimation's NFAPI_NR_DMRS_TYPE2_linear_interp / TYPE1_average_prb / TYPE2_average_prb, pdsch_ue = 1).
// Reference: nr_pusch_channel_estimation (openair1/PHY/NR_ESTIMATION/nr_ul_channel_estimation.c:67-243, 483-487) with nr_gold_pusch /
// nr_pusch_dmrs_rx (NR_REFSIG/nr_gold.c:99-116, nr_dmrs_rx.c:44-116), nr_est_delay / get_delay_idx / init_delay_table
// (common/utils/nr/nr_common.c:906-990), c16multaddVectRealComplex and filt16_ul_* (tools_defs.h:266-297, filt16a_32.h:242-249).
// The reference walks the pilots serially and overlap-adds a 16-tap window per pilot into the estimate.  Here:
//   chest_ls_kernel      one thread per pilot pair: DMRS bits from the Gold jump-ahead tables, LS estimate held over 4 REs, max_ch
//   (IDFT)               the library's own Q15 transform kernel on the zero-padded LS estimates of all antennas (nr_est_delay / freq2time)
//   chest_peak_kernel    peak of |h(t)|^2 >> 1 per antenna, then the reference's running maximum ACROSS antennas (delay_t is only reset per call)
//   chest_interp_kernel  one thread per output sub-carrier GATHERS the <= 8 pilot windows that cover it, in pilot order, with the same saturating
//                        adds; then reverts the delay and accumulates the noise estimate.
// All four are stream ordered; nothing returns to the host.
#include "nrb200_ctx.h"
#include "gold_seq.cuh"
#include "../../include/nrb200_ldpc.h"
#include <cmath>
#include <map>

namespace nrb200 {

int dft_batch_internal(int N, int inverse, uint32_t n, const int16_t *d_in, int16_t *d_out, int scale, cudaStream_t st);   // dfts_internal.cu

struct ChestGeom {
  int N, nb_rx, symbol, nb, np, k0, dmrs_offset;
  int ue;                                   // 1: the UE's PDSCH estimator (least-squares step of NFAPI_NR_DMRS_TYPE1_linear_interp)
  int n_ports, delta[2], wsign_odd[2];      // ports handled by one call (DMRS ports p, p+1 share the Gold sequence, differ in delta / w_f)
  unsigned rx_stride, ch_stride, x2;
  int type2, chest_freq;                    // DMRS configuration type 2 / one average per PRB (the variants below)
  int nushift, tail;                        // variants: (p >> 1) & 1 added to the symbol POINTER; c16 readable beyond the symbol's N (0 on the slot's last symbol)
  const unsigned *lowpapr;                  // transform precoding: the 6 * nb c16 low-PAPR type-1 sequence (device); pilots = its conjugate, w = +1, index from 0
};
constexpr int kChestState = 18;             // int32 per port, see chest_ls_kernel

__device__ __forceinline__ int c_sat16(int v) { return max(-32768, min(32767, v)); }
__device__ __forceinline__ int c_wrap16(int v) { return (int)(short)v; }
__device__ __forceinline__ int c_lo(unsigned w) { return (int)(short)(w & 0xFFFFu); }
__device__ __forceinline__ int c_hi(unsigned w) { return (int)(short)(w >> 16); }
__device__ __forceinline__ unsigned c_pk(int r, int i) { return ((unsigned)r & 0xFFFFu) | ((unsigned)i << 16); }
__device__ __forceinline__ int c_mulhrs(int a, int b) { return c_wrap16((a * b + 0x4000) >> 15); }
// c16mulShift(a, b, 8): truncating casts
__device__ __forceinline__ unsigned c_mul8(unsigned a, unsigned b)
{
  const int ar = c_lo(a), ai = c_hi(a), br = c_lo(b), bi = c_hi(b);
  return c_pk(c_wrap16((ar * br - ai * bi) >> 8), c_wrap16((ar * bi + ai * br) >> 8));
}

// per port: state[0] = max_ch, [1] = nvar, [2] = est_delay (after the last antenna), [3] = delay_max_pos, [4] = delay_max_val, [5] = CTA completion counter,
// [6..7] = 64-bit noise accumulator, [8 + a] = est_delay seen by antenna a (diagnostic).  raw (scratch): per (port, antenna) {peak value, peak position}.
__global__ void __launch_bounds__(256) chest_ls_kernel(ChestGeom G, const GoldTables *__restrict__ T, const unsigned *__restrict__ rxF, unsigned *__restrict__ ls,
                                                       int *__restrict__ state)
{
  __shared__ uint32_t s_gold[40];
  const int pa = blockIdx.y, port = pa / G.nb_rx, a = pa - port * G.nb_rx, n0 = blockIdx.x * 256, n = n0 + threadIdx.x;
  state += port * kChestState;
  const unsigned bit0 = 2u * (unsigned)(G.dmrs_offset + 2 * n0);            // first DMRS bit this CTA needs (2 bits per pilot, 2 pilots per thread)
  const unsigned w0 = bit0 >> 5;
  if (n0 < 3 * G.nb) {
    if (threadIdx.x < 34) s_gold[threadIdx.x] = gold_word(T, G.x2, w0 + threadIdx.x);
  }
  __syncthreads();
  unsigned *dst = ls + (size_t)pa * G.N;
  if (4 * n >= G.N) return;
  unsigned v = 0;
  if (n < 3 * G.nb) {
    const unsigned *rx = rxF + (size_t)a * G.rx_stride + (size_t)G.symbol * G.N;
    int cr = 0, ci = 0;
#pragma unroll
    for (int kl = 0; kl < 2; kl++) {
      const int i = G.dmrs_offset + 2 * n + kl;                            // pilot index in the sequence
      int pr, pi;
      if (G.lowpapr != nullptr) {                                          // nr_pusch_lowpaprtype1_dmrs_rx(p = 1000, re_offset = 0): conj, int16 wrap
        const unsigned sq = __ldg(G.lowpapr + 2 * n + kl);
        pr = c_lo(sq); pi = c_wrap16(-c_hi(sq));
      } else {
        const unsigned r0 = 2u * (unsigned)i - (w0 << 5);
        const int b0 = (s_gold[r0 >> 5] >> (r0 & 31u)) & 1u, b1 = (s_gold[(r0 + 1) >> 5] >> ((r0 + 1) & 31u)) & 1u;
        const int w = (i & 1) ? G.wsign_odd[port] : 1;
        // conj of the QPSK symbol (nr_rx_mod_table): re = +A for b0 = 0, im = -A for b1 = 0, negated when w = -1
        pr = w * (b0 ? -23170 : 23170); pi = w * (b1 ? 23170 : -23170);
      }
      int re = G.k0 + (n << 2) + (kl << 1);
      if (!G.ue) { re += G.delta[port]; re %= G.N; }
      else { re %= G.N; re += G.delta[port]; }                             // UE: the comb offset moves the symbol pointer, it does not wrap (:1689)
      const unsigned y = __ldg(rx + re);
      if (!G.ue) {
        cr += (pr * c_lo(y) - pi * c_hi(y)) >> 16;
        ci += (pr * c_hi(y) + pi * c_lo(y)) >> 16;
      } else {                                                             // c16mulShift / c16maddShift: int16 accumulation, >> 15 per product
        cr = c_wrap16(((pr * c_lo(y) - pi * c_hi(y)) >> 15) + cr);
        ci = c_wrap16(((pr * c_hi(y) + pi * c_lo(y)) >> 15) + ci);
      }
    }
    if (G.ue) { cr >>= 1; ci >>= 1; }                                      // c16Shift(ch, 1)
    const int m = max(abs(cr), abs(ci));
    if (m > 0 && !G.ue) atomicMax(state, m);
    v = c_pk(c_wrap16(cr), c_wrap16(ci));
  }
  reinterpret_cast<uint4 *>(dst)[n] = make_uint4(v, v, v, v);
}

// peak of |h(t)|^2 >> 1 of one (port, antenna): {value, first position}.  The reference's running maximum across antennas is applied by the consumers.
__global__ void __launch_bounds__(256) chest_peak_kernel(ChestGeom G, const unsigned *__restrict__ tim, int *__restrict__ raw)
{
  __shared__ int s_val[256], s_pos[256];
  const unsigned *t = tim + (size_t)blockIdx.x * G.N;
  int bv = -1, bp = 0;
  for (int i = threadIdx.x; i < G.N; i += blockDim.x) {
    const unsigned w = t[i];
    const int v = (int)(((unsigned)(c_lo(w) * c_lo(w) + c_hi(w) * c_hi(w))) >> 1);
    if (v > bv) { bv = v; bp = i; }                                       // ascending i per thread: keeps the first maximum
  }
  s_val[threadIdx.x] = bv; s_pos[threadIdx.x] = bp;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      const int ov = s_val[threadIdx.x + s], op = s_pos[threadIdx.x + s];
      if (ov > s_val[threadIdx.x] || (ov == s_val[threadIdx.x] && op < s_pos[threadIdx.x])) { s_val[threadIdx.x] = ov; s_pos[threadIdx.x] = op; }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) { raw[2 * blockIdx.x] = s_val[0]; raw[2 * blockIdx.x + 1] = s_pos[0]; }
}

// nr_est_delay's state after antennas 0..a of one port: delay_t is reset once per call, so the maximum runs ACROSS antennas (strict >: an earlier
// antenna's peak wins ties) and the wrap to negative delays is applied after every antenna.
__device__ __forceinline__ void running_delay(const ChestGeom &G, const int *__restrict__ raw_port, int a, int &max_pos, int &max_val)
{
  max_pos = 0; max_val = 0;
  for (int k = 0; k <= a; k++) {
    if (raw_port[2 * k] > max_val) { max_val = raw_port[2 * k]; max_pos = raw_port[2 * k + 1]; }
    if (max_pos > G.N / 2) max_pos -= G.N;
  }
}

__device__ __forceinline__ int filt_tap(int pc, int np, int t)
{
  if (pc == 0) return t < 8 ? 4096 : 0;                                      // filt16_ul_p0
  if (pc <= 2) return t < 4 ? 4096 : t < 12 ? 2048 : 0;                      // filt16_ul_p1p2
  if (pc == np - 1) return t < 4 ? 4096 : t < 8 ? 8192 : 0;                  // filt16_ul_last
  return 2048;                                                               // filt16_ul_middle
}

__global__ void __launch_bounds__(256) chest_interp_kernel(ChestGeom G, const unsigned *__restrict__ ls, const unsigned *__restrict__ dtab /* [41][N] */,
                                                           const int *__restrict__ raw, unsigned *__restrict__ est, int *__restrict__ state)
{
  const int pa = blockIdx.y, port = pa / G.nb_rx, a = pa - port * G.nb_rx, k = blockIdx.x * 256 + threadIdx.x;
  state += port * kChestState;
  const int nre = 12 * G.nb, kmax = min(nre + 8, G.N);
  const unsigned *l = ls + (size_t)pa * G.N;
  int ed, mv;
  running_delay(G, raw + 2 * port * G.nb_rx, a, ed, mv);
  const unsigned *tb = dtab + (size_t)min(max(20 + ed, 0), 40) * G.N, *ti = dtab + (size_t)min(max(20 - ed, 0), 40) * G.N;
  unsigned long long noise = 0;
  if (k < G.N) {
    unsigned out = 0;
    if (k < kmax) {
      int yr = 0, yi = 0;
      const int bq = k >> 2;
      for (int b = max(0, bq - 3); b <= bq; b++) {
        const int p_lo = b == 0 ? 0 : 2 * b + 3, p_hi = min(2 * b + 4, G.np - 1);
        for (int pc = p_lo; pc <= p_hi; pc++) {
          const int f = filt_tap(pc, G.np, k - 4 * b);
          if (f == 0) continue;
          const int si = (G.ue && G.type2) ? (pc / 3) * 6 : 2 * pc;          // UE type 2: three "pilots" share a CDM pair's value (:1499-1501)
          const unsigned c = c_mul8(l[si], __ldg(tb + si));                  // delay-compensated LS estimate of pilot pc
          const int mr = c_mulhrs(c_lo(c), f), mi = c_mulhrs(c_hi(c), f);
          yr = c_sat16(yr + c_sat16(2 * mr)); yi = c_sat16(yi + c_sat16(2 * mi));
        }
      }
      out = c_pk(yr, yi);
      if (k < nre) {
        out = c_mul8(out, __ldg(ti + k));                                    // revert the delay
        const unsigned lv = l[k];
        const int dr = c_wrap16(c_lo(lv) - c_lo(out)), di = c_wrap16(c_hi(lv) - c_hi(out));
        noise = G.ue ? 0u : (unsigned)(dr * dr + di * di);
      }
    }
    est[(size_t)pa * G.ch_stride + (size_t)G.symbol * G.N + k] = out;        // the whole symbol is rewritten (memset in the reference)
  }
  // block reduction of the noise power
  __shared__ unsigned long long s_n[256];
  s_n[threadIdx.x] = noise;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) { if (threadIdx.x < s) s_n[threadIdx.x] += s_n[threadIdx.x + s]; __syncthreads(); }
  if (threadIdx.x == 0) {
    if (s_n[0]) atomicAdd(reinterpret_cast<unsigned long long *>(state + 6), s_n[0]);
    if (blockIdx.x == 0) state[8 + a] = ed;
    __threadfence();
    // the last CTA of this port publishes nvar and the final delay_t
    if (atomicAdd(reinterpret_cast<unsigned *>(state + 5), 1u) == gridDim.x * (unsigned)G.nb_rx - 1) {
      __threadfence();
      const unsigned long long n = *reinterpret_cast<volatile unsigned long long *>(state + 6);
      state[1] = (int)(unsigned)(n / (unsigned long long)(12 * G.nb * G.nb_rx));
      int fp, fv;
      running_delay(G, raw + 2 * port * G.nb_rx, G.nb_rx - 1, fp, fv);
      state[2] = fp; state[3] = fp; state[4] = fv;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------------------------
// Variants of nr_pusch_channel_estimation: DMRS type 2 with chest_freq == 0 (nr_ul_channel_estimation.c:258-283) and one average per PRB,
// chest_freq == 1 (:285-343 type 1, :343-460 type 2; NO_INTERP is defined to 1 in that file, so a PRB's 12 REs take the average itself).
// The reference's quirks are part of the arithmetic and are reproduced (oracle/nrb200_chest_oracle.c lists them): nushift = (p >> 1) & 1 moves
// the symbol pointer for both DMRS types; type 2 accumulates the first four REs of every CDM group ACROSS antennas; type 2 + chest_freq == 1
// uses the first PRB's third pilot twice and needs slot % 4 == 0.

// conj of DMRS symbol i of the sequence (nr_rx_mod_table / nr_rx_nmod_table) as {re, im}
__device__ __forceinline__ void dmrs_conj(const GoldTables *__restrict__ T, unsigned x2, int i, int wsign_odd, int &pr, int &pi)
{
  const unsigned b = 2u * (unsigned)i;
  const uint32_t w = gold_word(T, x2, b >> 5);                             // bits 2i, 2i+1 never straddle a word
  const int b0 = (w >> (b & 31u)) & 1u, b1 = (w >> ((b & 31u) + 1u)) & 1u;
  const int s = (i & 1) ? wsign_odd : 1;
  pr = s * (b0 ? -23170 : 23170); pi = s * (b1 ? 23170 : -23170);
}
__device__ __forceinline__ unsigned rx_at(const ChestGeom &G, const unsigned *__restrict__ rx, int idx)
{
  return idx < G.N + G.tail ? __ldg(rx + idx) : 0u;                        // beyond the slot's last symbol: not ours to read
}

// type 2, chest_freq == 0: one thread per CDM pair, antennas in sequence (the running saturating sum of the first four REs)
__global__ void __launch_bounds__(128) chest_t2_ls_kernel(ChestGeom G, const GoldTables *__restrict__ T, const unsigned *__restrict__ rxF, unsigned *__restrict__ ls,
                                                          int *__restrict__ state)
{
  const int m = blockIdx.x * 128 + threadIdx.x;                            // pair index: pilots 2m, 2m + 1, sub-carriers 6m .. 6m + 5
  unsigned long long noise = 0;
  if (m < 2 * G.nb) {
    int p0r, p0i, p1r, p1i;
    dmrs_conj(T, G.x2, G.dmrs_offset + 2 * m, G.wsign_odd[0], p0r, p0i);
    dmrs_conj(T, G.x2, G.dmrs_offset + 2 * m + 1, G.wsign_odd[0], p1r, p1i);
    const int re0 = (G.k0 + 6 * m) % G.N + G.nushift, re1 = (G.k0 + 6 * m + 1) % G.N + G.nushift;
    int accr = 0, acci = 0, mx = 0;
    for (int a = 0; a < G.nb_rx; a++) {
      const unsigned *rx = rxF + (size_t)a * G.rx_stride + (size_t)G.symbol * G.N;
      const unsigned y0 = rx_at(G, rx, re0), y1 = rx_at(G, rx, re1);
      const int c0r = c_wrap16((p0r * c_lo(y0) - p0i * c_hi(y0)) >> 15), c0i = c_wrap16((p0r * c_hi(y0) + p0i * c_lo(y0)) >> 15);
      const int c1r = c_wrap16((p1r * c_lo(y1) - p1i * c_hi(y1)) >> 15), c1i = c_wrap16((p1r * c_hi(y1) + p1i * c_lo(y1)) >> 15);
      const int cr = c_wrap16((c0r + c1r) >> 1), ci = c_wrap16((c0i + c1i) >> 1);
      mx = max(mx, max(abs(cr), abs(ci)));
      accr = c_sat16(accr + c_wrap16((cr >> 2) << 2)); acci = c_sat16(acci + c_wrap16((ci >> 2) << 2));   // mulhi_s1_int16(ch, 16384), adds_epi16
      unsigned *dst = ls + (size_t)a * G.N + 6 * m;
      const unsigned va = c_pk(accr, acci), vc = c_pk(cr, ci);
      dst[0] = va; dst[1] = va; dst[2] = va; dst[3] = va; dst[4] = vc; dst[5] = vc;
      const int dr = c_wrap16(c0r - cr), di = c_wrap16(c0i - ci);
      noise += (unsigned)(dr * dr + di * di);
    }
    if (mx > 0) atomicMax(state, mx);
  }
  __shared__ unsigned long long s_n[128];
  s_n[threadIdx.x] = noise;
  __syncthreads();
  for (int s = 64; s > 0; s >>= 1) { if (threadIdx.x < s) s_n[threadIdx.x] += s_n[threadIdx.x + s]; __syncthreads(); }
  if (threadIdx.x == 0 && s_n[0]) atomicAdd(reinterpret_cast<unsigned long long *>(state + 6), s_n[0]);
}

// UE, type 2, chest_freq == 0 (NFAPI_NR_DMRS_TYPE2_linear_interp, nr_dl_channel_estimation.c:1463-1490): one thread per CDM pair and antenna,
// the pair's average held over its 6 sub-carriers; the interpolation that follows is the type 1 one (chest_interp_kernel)
__global__ void __launch_bounds__(128) chest_ue_t2_ls_kernel(ChestGeom G, const GoldTables *__restrict__ T, const unsigned *__restrict__ rxF, unsigned *__restrict__ ls)
{
  const int m = blockIdx.x * 128 + threadIdx.x, a = blockIdx.y;
  if (m >= 2 * G.nb) return;
  int p0r, p0i, p1r, p1i;
  dmrs_conj(T, G.x2, G.dmrs_offset + 2 * m, G.wsign_odd[0], p0r, p0i);
  dmrs_conj(T, G.x2, G.dmrs_offset + 2 * m + 1, G.wsign_odd[0], p1r, p1i);
  const unsigned *rx = rxF + (size_t)a * G.rx_stride + (size_t)G.symbol * G.N;
  const unsigned y0 = rx_at(G, rx, (G.k0 + 6 * m) % G.N + G.nushift), y1 = rx_at(G, rx, (G.k0 + 6 * m + 1) % G.N + G.nushift);
  const int lr = c_wrap16((p0r * c_lo(y0) - p0i * c_hi(y0)) >> 15), li = c_wrap16((p0r * c_hi(y0) + p0i * c_lo(y0)) >> 15);
  const int rr = c_wrap16((p1r * c_lo(y1) - p1i * c_hi(y1)) >> 15), ri = c_wrap16((p1r * c_hi(y1) + p1i * c_lo(y1)) >> 15);
  const unsigned v = c_pk(c_wrap16((lr + rr) >> 1), c_wrap16((li + ri) >> 1));
  unsigned *dst = ls + (size_t)a * G.N + 6 * m;
#pragma unroll
  for (int k = 0; k < 6; k++) dst[k] = v;
}

// type 2, chest_freq == 0: ul_ch[n] = c16mulShift(ls[n], delay_table[get_delay_idx(-est_delay)][n % 6], 8); the rest of the symbol is cleared
__global__ void __launch_bounds__(256) chest_t2_apply_kernel(ChestGeom G, const unsigned *__restrict__ ls, const unsigned *__restrict__ dtab, const int *__restrict__ raw,
                                                             unsigned *__restrict__ est, int *__restrict__ state)
{
  const int a = blockIdx.y, k = blockIdx.x * 256 + threadIdx.x;
  int ed, mv;
  running_delay(G, raw, a, ed, mv);
  if (k < G.N) {
    const unsigned *ti = dtab + (size_t)min(max(20 - ed, 0), 40) * G.N;
    est[(size_t)a * G.ch_stride + (size_t)G.symbol * G.N + k] = k < 12 * G.nb ? c_mul8(ls[(size_t)a * G.N + k], __ldg(ti + k % 6)) : 0u;
  }
  if (k == 0) {
    state[8 + a] = ed;
    if (a == G.nb_rx - 1) {                                                 // chest_t2_ls_kernel has completed (stream order): publish nvar and delay_t
      const unsigned long long n = *reinterpret_cast<const unsigned long long *>(state + 6);
      state[1] = (int)(unsigned)(n / (unsigned long long)(2 * G.nb * G.nb_rx));
      state[2] = ed; state[3] = ed; state[4] = mv;
    }
  }
}

// chest_freq == 1: every sub-carrier of PRB j takes the PRB's average; no delay estimation, no noise estimate
__global__ void __launch_bounds__(256) chest_avg_kernel(ChestGeom G, const GoldTables *__restrict__ T, const unsigned *__restrict__ rxF, unsigned *__restrict__ est,
                                                        int *__restrict__ state)
{
  const int a = blockIdx.y, j = blockIdx.x * 256 + threadIdx.x;            // one thread per PRB; PRBs beyond the allocation clear their 12 sub-carriers
  if (12 * j >= G.N) return;
  unsigned v = 0;
  if (j < G.nb) {
    const unsigned *rx = rxF + (size_t)a * G.rx_stride + (size_t)G.symbol * G.N;
    const int cnt = G.type2 ? 4 : 6;
    int sr = 0, si = 0;
    for (int i = 0; i < cnt; i++) {
      int pidx, re;
      if (!G.type2) { pidx = 6 * j + i; re = (G.k0 + 12 * j + 2 * i) % G.N; }
      else if (!G.ue) { pidx = j == 0 ? min(i, 2) : 4 * j - 1 + i; re = (G.k0 + 12 * j + (i & 1) + 6 * (i >> 1)) % G.N; }
      else { pidx = 4 * j + i; re = (j == 0 ? G.k0 + i : G.k0 + 4 + 20 * (j - 1) + 5 * i) % G.N; }   // the UE's walk (:1541-1575): 4 consecutive, then 5 apart
      int pr, pi;
      if (G.lowpapr != nullptr) { const unsigned sq = __ldg(G.lowpapr + pidx); pr = c_lo(sq); pi = c_wrap16(-c_hi(sq)); }
      else dmrs_conj(T, G.x2, G.dmrs_offset + pidx, G.wsign_odd[0], pr, pi);
      const unsigned y = rx_at(G, rx, re + G.nushift);
      sr += (pr * c_lo(y) - pi * c_hi(y)) >> 15;
      si += (pr * c_hi(y) + pi * c_lo(y)) >> 15;
    }
    const int cr = c_wrap16(sr / cnt), ci = c_wrap16(si / cnt);
    if (!G.ue && j > 0 && j < G.nb - 1) { const int m = max(abs(cr), abs(ci)); if (m > 0) atomicMax(state, m); }
    v = c_pk(cr, ci);
  }
  unsigned *dst = est + (size_t)a * G.ch_stride + (size_t)G.symbol * G.N + 12 * j;
  for (int k = 0; k < 12 && 12 * j + k < G.N; k++) dst[k] = v;
}

// nr_chest_time_domain_avg (NR_REFSIG/dmrs_nr.c:343-417): the slot's DMRS-symbol estimates summed (adds_epi16) into the first DMRS symbol over the
// first 12 * num_rbs entries of the symbol and divided by their number (>> 1, / 3 towards zero, >> 2).  One thread per antenna and entry; the
// reference makes one pass over the symbol per additional DMRS symbol plus one for the division.
__global__ void __launch_bounds__(256) chest_time_avg_kernel(int N, unsigned plane_stride, int first, unsigned later_mask, int ndmrs, int n, unsigned *__restrict__ est)
{
  const int k = blockIdx.x * 256 + threadIdx.x;
  if (k >= n) return;
  unsigned *plane = est + (size_t)blockIdx.y * plane_stride;
  const unsigned v = plane[(size_t)first * N + k];
  int r = c_lo(v), i = c_hi(v);
  for (int s = first + 1; s < 14; s++)
    if ((later_mask >> s) & 1u) {
      const unsigned x = plane[(size_t)s * N + k];
      r = c_sat16(r + c_lo(x)); i = c_sat16(i + c_hi(x));
    }
  if (ndmrs == 2) { r >>= 1; i >>= 1; }
  else if (ndmrs == 4) { r >>= 2; i >>= 2; }
  else if (ndmrs == 3) { r /= 3; i /= 3; }
  plane[(size_t)first * N + k] = c_pk(r, i);
}

// returns the first DMRS symbol (>= 0) or a negative error
int launch_chest_time_avg(uint32_t N, uint32_t nb_rx, uint32_t ch_stride, uint32_t start_symbol, uint32_t nr_of_symbols, uint32_t dmrs_symb_pos, uint32_t rb_size,
                          int16_t *d_est, cudaStream_t st)
{
  const uint32_t total = start_symbol + nr_of_symbols;
  if (total > 14 || nb_rx < 1 || 12 * rb_size > N || rb_size < 1) return -4;
  int ndmrs = 0, first = -1;
  for (uint32_t s = 0; s < total; s++) ndmrs += (dmrs_symb_pos >> s) & 1u;          // get_dmrs_symbols_in_slot counts from symbol 0
  for (uint32_t s = start_symbol; s < total; s++) if ((dmrs_symb_pos >> s) & 1u) { first = (int)s; break; }
  if (first < 0 || ndmrs < 1 || ndmrs > 4) return -4;                                // AssertFatal in the reference
  const unsigned later = dmrs_symb_pos & ((1u << total) - 1u) & ~((2u << first) - 1u);
  const int n = 12 * (int)rb_size;
  chest_time_avg_kernel<<<dim3((n + 255) / 256, nb_rx), 256, 0, st>>>((int)N, ch_stride, first, later, ndmrs, n, (unsigned *)d_est);
  ctx().launches += 1;
  NRB200_CUDA_OK(cudaGetLastError(), "chest_time_avg launch");
  return first;
}

// fp->delay_table (init_delay_table): round(256 e^{j 2 pi k d / N}) for d = -20..20, built once per N
static const unsigned *delay_table_dev(int N)
{
  static std::mutex mu;
  static std::map<int, unsigned *> cache;
  std::lock_guard<std::mutex> lk(mu);
  const int key = N | (ctx().dev << 20);                 // one table per size and device
  auto it = cache.find(key);
  if (it != cache.end()) return it->second;
  std::vector<unsigned> h((size_t)41 * N);
  for (int d = -20; d <= 20; d++)
    for (int k = 0; k < N; k++) {
      const double ang = 2.0 * M_PI * k * d / N;
      const short r = (short)std::round(256 * std::cos(ang)), i = (short)std::round(256 * std::sin(ang));
      h[(size_t)(20 + d) * N + k] = ((unsigned)(unsigned short)r) | ((unsigned)(unsigned short)i << 16);
    }
  unsigned *dptr = nullptr;
  if (cudaMalloc(&dptr, h.size() * 4) != cudaSuccess) return nullptr;
  cudaMemcpy(dptr, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  cache[key] = dptr;
  return dptr;
}

static int chest_geom(const nrb200_pusch_chest_t &d, ChestGeom *G)
{
  if (d.nb_rx < 1 || d.nb_rx > 8 || d.symbol > 13 || d.port > ((d.pdsch_ue && d.dmrs_config_type) ? 5u : 3u) || d.rb_size < 1 || 12 * d.rb_size > d.fft_size || d.scid > 1 || (d.fft_size & 3)) return -4;
  G->N = d.fft_size; G->nb_rx = d.nb_rx; G->symbol = d.symbol; G->nb = d.rb_size; G->np = 6 * d.rb_size;
  G->k0 = ((d.rb_start + d.bwp_start) * 12 + d.first_carrier_offset) % d.fft_size;
  G->ue = d.pdsch_ue ? 1 : 0;
  G->n_ports = d.n_ports == 0 ? 1 : (int)d.n_ports;
  if (G->n_ports > 2 || d.port + G->n_ports > ((d.pdsch_ue && d.dmrs_config_type) ? 6u : 4u)) return -4;
  for (int q = 0; q < G->n_ports; q++) {
    const unsigned pp = d.port + q;
    G->delta[q] = (pp >> 1) & 1;                                            // delta1[p]
    G->wsign_odd[q] = (pp & 1) ? -1 : 1;                                    // wf1[p][1]
  }
  G->type2 = d.dmrs_config_type ? 1 : 0; G->chest_freq = d.chest_freq ? 1 : 0;
  G->lowpapr = nullptr;
  if (d.transform_precoding) {                                              // NR_SC_FDMA supports type 1 DMRS only (:123); the sequence is the caller's
    if (d.dmrs_config_type || d.pdsch_ue || d.lowpapr_seq == 0) return -4;
    G->lowpapr = reinterpret_cast<const unsigned *>((uintptr_t)d.lowpapr_seq);
  }
  G->nushift = (d.port >> 1) & 1;
  G->tail = d.symbol < 13 ? 4 : 0;
  if (G->type2 || G->chest_freq) {
    // variants: gNB estimator, one port per call; the pointer shift needs an even first sub-carrier to stay inside the symbol (always the case
    // for an even first_carrier_offset); PRB averages need two PRBs; type 2 averages need slot % 4 == 0 (see the kernels' header)
    if (G->n_ports != 1 || d.dmrs_config_type > 1 || d.chest_freq > 1) return -4;
    if (G->chest_freq && (d.rb_size < 2 || (G->type2 && !G->ue && (d.slot & 3)))) return -4;
    G->np = (G->type2 && !G->ue ? 4 : 6) * d.rb_size;                       // the UE's type 2 branch feeds the type 1 interpolation: 6 "pilots" per PRB
    if (G->ue && G->type2) { static const int delta2[6] = {0, 0, 2, 2, 4, 4}; G->nushift = delta2[d.port]; }   // get_delta(p, NFAPI_NR_DMRS_TYPE2)
  }
  G->dmrs_offset = ((d.bwp_start + d.rb_start) * 12) / (G->type2 ? 3 : 2);
  G->rx_stride = d.rx_stride; G->ch_stride = d.ch_stride;
  const unsigned long long nid = d.ul_dmrs_scrambling_id;
  const unsigned long long t = ((1ULL << 17) * (unsigned long long)(14 * d.slot + d.symbol + 1) * ((nid << 1) + 1) + ((nid << 1) + d.scid));
  G->x2 = (unsigned)(t % (1ULL << 31));
  return 0;
}

// Host-side DMRS generation (nr_gold_pusch + nr_pusch_dmrs_rx, serial word recurrence like the reference): 6 * rb_size conjugated pilots {re, im}.
// The estimator kernel derives the same bits from the jump-ahead tables; this entry point serves transmit-side synthesis and tests.
int pusch_dmrs_pilots_host(const nrb200_pusch_chest_t &d, int16_t *pil)
{
  ChestGeom G;
  int rc = chest_geom(d, &G);
  if (rc) return rc;
  if (d.transform_precoding) {                                               // here lowpapr_seq is a HOST address
    const int16_t *sq = reinterpret_cast<const int16_t *>((uintptr_t)d.lowpapr_seq);
    for (int k = 0; k < 6 * G.nb; k++) { pil[2 * k] = sq[2 * k]; pil[2 * k + 1] = (int16_t)-sq[2 * k + 1]; }
    return 0;
  }
  uint32_t x1 = 1u + (1u << 31), x2 = G.x2;
  x2 = x2 ^ ((x2 ^ (x2 >> 1) ^ (x2 >> 2) ^ (x2 >> 3)) << 31);
  auto step = [&]() {
    x1 = (x1 >> 1) ^ (x1 >> 4); x1 = x1 ^ (x1 << 31) ^ (x1 << 28);
    x2 = (x2 >> 1) ^ (x2 >> 2) ^ (x2 >> 3) ^ (x2 >> 4); x2 = x2 ^ (x2 << 31) ^ (x2 << 30) ^ (x2 << 29) ^ (x2 << 28);
  };
  for (int n = 1; n < 50; n++) step();
  const int last = G.dmrs_offset + (G.type2 ? 4 : 6) * G.nb;                 // 6 (type 1) or 4 (type 2) pilots per PRB
  std::vector<uint32_t> g((size_t)(2 * last + 31) / 32 + 1);
  for (auto &w : g) { step(); w = x1 ^ x2; }
  for (int i = G.dmrs_offset; i < last; i++) {
    const int b0 = (g[(2 * i) >> 5] >> ((2 * i) & 31)) & 1, b1 = (g[(2 * i + 1) >> 5] >> ((2 * i + 1) & 31)) & 1;
    const int w = (i & 1) ? G.wsign_odd[0] : 1;
    pil[2 * (i - G.dmrs_offset)] = (int16_t)(w * (b0 ? -23170 : 23170));
    pil[2 * (i - G.dmrs_offset) + 1] = (int16_t)(w * (b1 ? 23170 : -23170));
  }
  return 0;
}

// r_{u,v}(n) of TS 38.211 5.2.2 with the reference's arithmetic (ul_ref_seq_nr.c:55-196): length 30 in closed form (5.2.2.2), 36 and longer as the cyclic
// extension of the Zadoff-Chu sequence of the largest prime below the length (5.2.2.1); double precision, floor.
int lowpapr_sequence_host(uint32_t u, uint32_t v, uint32_t n_re, uint32_t scaling, int16_t *seq)
{
  if (u > 29 || v > 1 || !seq || (n_re != 30 && n_re < 36) || n_re > 12 * 275) return -4;
  if (n_re == 30) {
    for (uint32_t n = 0; n < n_re; n++) {
      const double x = -(M_PI * (u + 1) * (n + 1) * (n + 2)) / (double)31;
      seq[2 * n] = (int16_t)std::floor(scaling * std::cos(x)); seq[2 * n + 1] = (int16_t)std::floor(scaling * std::sin(x));
    }
    return 0;
  }
  uint32_t nzc = n_re - 1;
  for (;; nzc--) { bool pr = true; for (uint32_t q = 2; q * q <= nzc; q++) if (nzc % q == 0) { pr = false; break; } if (pr) break; }
  const double qb = nzc * (u + 1) / (double)31;
  const unsigned q = (((int)std::floor(2 * qb)) & 1) == 0 ? (unsigned)((int)std::floor(qb + .5) - (int)v) : (unsigned)((int)std::floor(qb + .5) + (int)v);
  for (uint32_t n = 0; n < n_re; n++) {
    const unsigned m = n % nzc;
    const double x = (double)q * m * (m + 1) / nzc;
    seq[2 * n] = (int16_t)std::floor(scaling * std::cos(M_PI * x));
    seq[2 * n + 1] = (int16_t)-(int16_t)std::floor(scaling * std::sin(M_PI * x));
  }
  return 0;
}

size_t pusch_chest_scratch_bytes(const nrb200_pusch_chest_t &d) { return (size_t)2 * 2 * d.nb_rx * d.fft_size * 4 + 256; }   // LS + time planes for up to 2 ports, raw peaks

// d_scratch: pusch_chest_scratch_bytes(); d_state: 18 int32 per port (see chest_ls_kernel), zeroed here
// buf_symbol >= 0: the buffers hold the DMRS symbol at that index (the host entry point stages a one-symbol slot); the DMRS sequence always uses d.symbol
int launch_pusch_chest(const nrb200_pusch_chest_t &d, const int16_t *rxF, int16_t *est, void *d_scratch, int32_t *d_state, cudaStream_t st, int buf_symbol,
                       int tail_override)
{
  ChestGeom G;
  int rc = chest_geom(d, &G);
  if (rc) return rc;
  if (buf_symbol >= 0) G.symbol = buf_symbol;
  if (scramble_mod_init() != 0) return -5;
  const unsigned *dtab = delay_table_dev(G.N);
  if (!dtab) return -5;
  const int npa = G.n_ports * G.nb_rx;
  unsigned *ls = (unsigned *)d_scratch, *tim = ls + (size_t)npa * G.N;
  int *raw = (int *)(tim + (size_t)npa * G.N);
  NRB200_CUDA_OK(cudaMemsetAsync(d_state, 0, (size_t)G.n_ports * kChestState * 4, st), "chest memset");
  if (buf_symbol >= 0) G.tail = tail_override;
  if (G.chest_freq) {
    chest_avg_kernel<<<dim3((G.N / 12 + 1 + 255) / 256, G.nb_rx), 256, 0, st>>>(G, gold_tables_dev(), (const unsigned *)rxF, (unsigned *)est, d_state);
    ctx().launches += 1;
    NRB200_CUDA_OK(cudaGetLastError(), "chest_avg launch");
    return 0;
  }
  if (G.type2 && G.ue) {
    NRB200_CUDA_OK(cudaMemsetAsync(ls, 0, (size_t)G.nb_rx * G.N * 4, st), "chest ls memset");
    chest_ue_t2_ls_kernel<<<dim3((2 * G.nb + 127) / 128, G.nb_rx), 128, 0, st>>>(G, gold_tables_dev(), (const unsigned *)rxF, ls);
    NRB200_CUDA_OK(cudaGetLastError(), "chest_ue_t2_ls launch");
    if ((rc = dft_batch_internal(G.N, 1, G.nb_rx, (const int16_t *)ls, (int16_t *)tim, 1, st)) != 0) return rc;
    chest_peak_kernel<<<G.nb_rx, 256, 0, st>>>(G, tim, raw);
    chest_interp_kernel<<<dim3((G.N + 255) / 256, G.nb_rx), 256, 0, st>>>(G, ls, dtab, raw, (unsigned *)est, d_state);
    ctx().launches += 4;
    NRB200_CUDA_OK(cudaGetLastError(), "chest_ue_t2 launch");
    return 0;
  }
  if (G.type2) {
    NRB200_CUDA_OK(cudaMemsetAsync(ls, 0, (size_t)G.nb_rx * G.N * 4, st), "chest ls memset");
    chest_t2_ls_kernel<<<(2 * G.nb + 127) / 128, 128, 0, st>>>(G, gold_tables_dev(), (const unsigned *)rxF, ls, d_state);
    NRB200_CUDA_OK(cudaGetLastError(), "chest_t2_ls launch");
    if ((rc = dft_batch_internal(G.N, 1, G.nb_rx, (const int16_t *)ls, (int16_t *)tim, 1, st)) != 0) return rc;
    chest_peak_kernel<<<G.nb_rx, 256, 0, st>>>(G, tim, raw);
    chest_t2_apply_kernel<<<dim3((G.N + 255) / 256, G.nb_rx), 256, 0, st>>>(G, ls, dtab, raw, (unsigned *)est, d_state);
    ctx().launches += 4;
    NRB200_CUDA_OK(cudaGetLastError(), "chest_t2 launch");
    return 0;
  }
  chest_ls_kernel<<<dim3((G.N / 4 + 255) / 256, npa), 256, 0, st>>>(G, gold_tables_dev(), (const unsigned *)rxF, ls, d_state);
  NRB200_CUDA_OK(cudaGetLastError(), "chest_ls launch");
  if ((rc = dft_batch_internal(G.N, 1, npa, (const int16_t *)ls, (int16_t *)tim, 1, st)) != 0) return rc;
  chest_peak_kernel<<<npa, 256, 0, st>>>(G, tim, raw);
  chest_interp_kernel<<<dim3((G.N + 255) / 256, npa), 256, 0, st>>>(G, ls, dtab, raw, (unsigned *)est, d_state);
  ctx().launches += 4;
  NRB200_CUDA_OK(cudaGetLastError(), "chest launch");
  return 0;
}

}  // namespace nrb200
