// libdfts_b200.so -- batched bit-exact Q15 DFT/IDFT kernels + the OAI `dft`/`idft` plug-in ABI (include/nrb200_dfts.h).
//
// Arithmetic restated from the reference openair1/PHY/TOOLS/oai_dfts.c: N = r3 * r2 * N4 with r3 in {1,3}, r2 in {1,2},
// N4 = 16 * 4^D, decimation in time: 16-point saturating kernel (:1190-1458), radix-4 levels (16-bit saturating butterfly with
// hand-rounded tables for 64 and forward 256 :803-948, 32-bit butterfly + wrapping add of x0 otherwise :633-721), optional
// radix-2 level (:390-476) and radix-3 level (:477-540); per-level scaling >>3 (64), >>1 (256+), mulhrs 1/sqrt2, 1/sqrt3.
// Here the recursion is unrolled into in-place passes over one shared-memory buffer: leaves are written in digit-reversed order
// so every later butterfly reads and writes the same 4 (2, 3) addresses.  One CTA handles `tpb` transforms of N points.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>
#include <cuda_runtime.h>
#include "../../include/nrb200_dfts.h"
#include "nr_dft_tables.h"

// dfts_internal.cu compiles this file a second time into libldpc_b200.so (the PUSCH delay estimator needs an IDFT): there the C ABI stays hidden.
#ifdef NRB200_DFTS_INTERNAL
#define NRB200_EXPORT extern "C" __attribute__((visibility("hidden")))
#else
#define NRB200_EXPORT extern "C" __attribute__((visibility("default")))
#endif

namespace {

struct cx { int r, i; };

__device__ __forceinline__ int sat16(int v) { return max(-32768, min(32767, v)); }
__device__ __forceinline__ int wrap16(int v) { return (int)(short)v; }
__device__ __forceinline__ int neg16(int v) { return v == -32768 ? -32768 : -v; }
__device__ __forceinline__ cx mjn(cx a) { return {a.i, neg16(a.r)}; }
// adds_epi16 / subs_epi16 on values held sign-extended in 32-bit registers: add-then-min is one VIADDMNMX, the max a VIMNMX.  (Written as max(min(a + b))
// the compiler recognises a 16-bit saturating add and expands it into a seven-instruction overflow test.)
__device__ __forceinline__ int sadd16(int a, int b) { return max(__viaddmin_s32(a, b, 32767), -32768); }
__device__ __forceinline__ cx sadd(cx a, cx b) { return {sadd16(a.r, b.r), sadd16(a.i, b.i)}; }
__device__ __forceinline__ cx ssub(cx a, cx b) { return {sadd16(a.r, -b.r), sadd16(a.i, -b.i)}; }
// The butterflies are bound by the integer ALU pipe (SHF / VIMNMX / PRMT / LOP3: 87 % busy in ncu, FMA pipe 15 %).  Issuing the arithmetic shifts as IMAD.HI
// (x * 2^(32-n) >> 32) to move them to the FMA pipe was measured and is SLOWER (49.3 -> 47.0 M IDFT-4096/s for the >> 15 alone, 40.4 M with the sign
// extensions too): IMAD.HI does not issue at the IMAD rate.  Plain shifts stay.
__device__ __forceinline__ int sra15(unsigned v) { return ((int)v) >> 15; }
__device__ __forceinline__ cx unpack(unsigned w) { return {(int)(short)(w & 0xFFFFu), (int)(short)(w >> 16)}; }
// both components are int16 values: cvt.pack.sat (one I2IP) packs them, the saturation never acts
__device__ __forceinline__ unsigned pack(cx a)
{
  unsigned r;
  asm("cvt.pack.sat.s16.s32 %0, %1, %2;" : "=r"(r) : "r"(a.i), "r"(a.r));
  return r;
}
// packed_cmult2: (a . ta, a . tb) >> 15, packs
// twiddles are {re, im} int16 pairs at even offsets of the table blob: one 32-bit load each
__device__ __forceinline__ cx ldtw(const short *p) { return unpack(__ldg(reinterpret_cast<const unsigned *>(p))); }
__device__ __forceinline__ cx cmult2(cx a, const short *ta, const short *tb)
{
  const cx A = ldtw(ta), B = ldtw(tb);
  return {sat16(sra15((unsigned)(a.r * A.r) + (unsigned)(a.i * A.i))), sat16(sra15((unsigned)(a.r * B.r) + (unsigned)(a.i * B.i)))};
}
__device__ __forceinline__ int mulhrs(int a, int b) { return wrap16((a * b + 0x4000) >> 15); }

__device__ __forceinline__ void bfly4_sat(cx x0, cx a1, cx a2, cx a3, bool inv, cx &y0, cx &y1, cx &y2, cx &y3)
{
  cx x02 = sadd(x0, a2), x13 = sadd(a1, a3);
  y0 = sadd(x02, x13);
  y2 = ssub(x02, x13);
  x02 = ssub(x0, a2);
  x13 = ssub(mjn(a1), mjn(a3));
  const cx ya = sadd(x02, x13), yb = ssub(x02, x13);
  y1 = inv ? yb : ya;
  y3 = inv ? ya : yb;
}
// 32-bit product by w (forward) or conj(w) (inverse); unsigned arithmetic wraps like the reference's epi32 adds
__device__ __forceinline__ void cm32(cx x, int wr, int wi, bool inv, unsigned &re, unsigned &im)
{
  if (inv) { re = (unsigned)(x.r * wr) + (unsigned)(x.i * wi); im = (unsigned)(x.i * wr) - (unsigned)(x.r * wi); }
  else { re = (unsigned)(x.r * wr) - (unsigned)(x.i * wi); im = (unsigned)(x.r * wi) + (unsigned)(x.i * wr); }
}
__device__ __forceinline__ cx pk32(unsigned r, unsigned i) { return {sat16(sra15(r)), sat16(sra15(i))}; }

__device__ __forceinline__ void bfly4_32(cx x0, cx x1, cx x2, cx x3, const short *w1, const short *w2, const short *w3, bool inv, cx &y0, cx &y1,
                                         cx &y2, cx &y3)
{
  unsigned x1r, x1i, x2r, x2i, x3r, x3i;
  const cx W1 = ldtw(w1), W2 = ldtw(w2), W3 = ldtw(w3);
  cm32(x1, W1.r, W1.i, inv, x1r, x1i);
  cm32(x2, W2.r, W2.i, inv, x2r, x2i);
  cm32(x3, W3.r, W3.i, inv, x3r, x3i);
  const cx d0 = pk32(x1r + x2r + x3r, x1i + x2i + x3i);
  const cx da = pk32(x1i - (x2r + x3i), (x3r - x2i) - x1r);
  const cx d2 = pk32((x2r - x3r) - x1r, (x2i - x3i) - x1i);
  const cx db = pk32((x3i - x2r) - x1i, x1r - (x2i + x3r));
  y0 = {wrap16(x0.r + d0.r), wrap16(x0.i + d0.i)};
  y2 = {wrap16(x0.r + d2.r), wrap16(x0.i + d2.i)};
  const cx oa = {wrap16(x0.r + da.r), wrap16(x0.i + da.i)}, ob = {wrap16(x0.r + db.r), wrap16(x0.i + db.i)};
  y1 = inv ? ob : oa;
  y3 = inv ? oa : ob;
}

// The same butterfly on PACKED {re, im} int16 words (the 4096-point kernel): x0 is only ever added with 16-bit wrap-around, so it stays packed
// (one VIADD.16x2 per output instead of two adds + two sign extensions), and the four 32-bit sums are shifted, saturated and packed by one I2IP each
// (cvt.pack.sat.s16.s32 = packs_epi32).  Identical results to bfly4_32.
__device__ __forceinline__ unsigned pk32p(unsigned r, unsigned i)
{
  unsigned o;
  asm("cvt.pack.sat.s16.s32 %0, %1, %2;" : "=r"(o) : "r"(sra15(i)), "r"(sra15(r)));
  return o;
}
__device__ __forceinline__ unsigned sra1_16x2(unsigned p) { return ((p >> 1) & 0x7FFF7FFFu) | (p & 0x80008000u); }   // >> 1 of both int16 lanes
__device__ __forceinline__ void bfly4_32p(unsigned x0, unsigned p1, unsigned p2, unsigned p3, const short *w1, const short *w2, const short *w3, bool inv, bool half,
                                          unsigned &y0, unsigned &y1, unsigned &y2, unsigned &y3)
{
  unsigned x1r, x1i, x2r, x2i, x3r, x3i;
  const cx W1 = ldtw(w1), W2 = ldtw(w2), W3 = ldtw(w3);
  cm32(unpack(p1), W1.r, W1.i, inv, x1r, x1i);
  cm32(unpack(p2), W2.r, W2.i, inv, x2r, x2i);
  cm32(unpack(p3), W3.r, W3.i, inv, x3r, x3i);
  const unsigned d0 = pk32p(x1r + x2r + x3r, x1i + x2i + x3i);
  const unsigned da = pk32p(x1i - (x2r + x3i), (x3r - x2i) - x1r);
  const unsigned d2 = pk32p((x2r - x3r) - x1r, (x2i - x3i) - x1i);
  const unsigned db = pk32p((x3i - x2r) - x1i, x1r - (x2i + x3r));
  y0 = __vadd2(x0, d0);
  y2 = __vadd2(x0, d2);
  const unsigned oa = __vadd2(x0, da), ob = __vadd2(x0, db);
  y1 = inv ? ob : oa;
  y3 = inv ? oa : ob;
  if (half) { y0 = sra1_16x2(y0); y1 = sra1_16x2(y1); y2 = sra1_16x2(y2); y3 = sra1_16x2(y3); }
}

// The same butterfly with the complex products on the dot-product unit (4096-point kernel).  A twiddle is stored as two words of bytes
//   B_re = {lo(a), lo(b), hi(a), hi(b)},  B_im = {lo(c), lo(d), hi(c), hi(d)}      with re = xr a + xi b, im = xr c + xi d
// (forward: a = wr, b = -wi, c = wi, d = wr; inverse: a = wr, b = wi, c = -wi, d = wr), lo = the unsigned low byte, hi = the signed high byte of the int16.
// Then re = 256 * dp2a.hi.s32.s32(x, B_re) + dp2a.lo.s32.u32(x, B_re) with x the PACKED sample: two IDP + one IMAD on the FMA pipe and no unpacking of the
// sample or the twiddle on the ALU pipe, which is the pipe that limits the kernel.  All arithmetic is modulo 2^32 like the reference's madd_epi16.
__device__ __forceinline__ unsigned dp_prod(unsigned x, unsigned B)
{
  int lo, hi;
  asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(lo) : "r"(x), "r"(B), "r"(0));
  asm("dp2a.hi.s32.s32 %0, %1, %2, %3;" : "=r"(hi) : "r"(x), "r"(B), "r"(0));
  return (unsigned)(hi * 256 + lo);
}
// packed_cmult2 on the dot-product unit: T = {B_re from table a, B_im from table b}; (x . ta, x . tb) >> 15, saturated
__device__ __forceinline__ cx cmult2dp(unsigned x, const uint2 *t)
{
  const uint2 T = __ldg(t);
  return {sat16(sra15(dp_prod(x, T.x))), sat16(sra15(dp_prod(x, T.y)))};
}
__device__ __forceinline__ void bfly4_32dp(unsigned x0, unsigned p1, unsigned p2, unsigned p3, const uint2 *d1, const uint2 *d2, const uint2 *d3, bool inv, bool half,
                                           unsigned &y0, unsigned &y1, unsigned &y2, unsigned &y3)
{
  const uint2 T1 = __ldg(d1), T2 = __ldg(d2), T3 = __ldg(d3);
  const unsigned x1r = dp_prod(p1, T1.x), x1i = dp_prod(p1, T1.y), x2r = dp_prod(p2, T2.x), x2i = dp_prod(p2, T2.y), x3r = dp_prod(p3, T3.x), x3i = dp_prod(p3, T3.y);
  const unsigned d0 = pk32p(x1r + x2r + x3r, x1i + x2i + x3i);
  const unsigned da = pk32p(x1i - (x2r + x3i), (x3r - x2i) - x1r);
  const unsigned dd2 = pk32p((x2r - x3r) - x1r, (x2i - x3i) - x1i);
  const unsigned db = pk32p((x3i - x2r) - x1i, x1r - (x2i + x3r));
  y0 = __vadd2(x0, d0);
  y2 = __vadd2(x0, dd2);
  const unsigned oa = __vadd2(x0, da), ob = __vadd2(x0, db);
  y1 = inv ? ob : oa;
  y3 = inv ? oa : ob;
  if (half) { y0 = sra1_16x2(y0); y1 = sra1_16x2(y1); y2 = sra1_16x2(y2); y3 = sra1_16x2(y3); }
}

// Device-resident twiddle blob (int16): the hand-rounded tables of nr_dft_tables.h followed by the generated ones.
struct TwOffsets {
  int tw16, tw16a, tw16b, tw16c, tw64, tw64a, tw64b, tw64c, tw128, tw128a, tw128b, tw256, tw256a, tw256b, tw512;
  int rad4_1024, rad4_4096, rad2_2048, rad2_8192, rad3[4];   // rad3: 768, 1536, 3072, 6144 (twa then twb)
  int dp256i, dp1024[2], dp4096[2];                           // byte-arranged copies for the dot-product butterflies (dp_tables below): [0] forward, [1] inverse
  int dps16[2], dps64[2], dps256f;                            // the same for the saturating levels' table pairs (packed_cmult2: re from table a, im from table b)
};

struct DftPlan {
  int N, r3, r2, N4, D;       // N = r3 * r2 * N4, N4 = 16 * 4^D
  int inverse, scale;
  int tpb;                    // transforms per CTA
  int top4_tw;                // offset of the generated radix-4 table of the largest 4^k level (-1 if none)
  int rad2_tw, rad3_tw;       // offsets, -1 if unused
};

// one radix-4 butterfly of the level that builds size-4M transforms from size-M ones (k = position inside the sub-transform), with the level's
// scaling.  The arithmetic per level is the reference's: 16-bit saturating butterfly with hand-rounded tables for 64 and forward 256, 32-bit
// butterfly with generated twiddles otherwise.
template <bool INV>
__device__ __forceinline__ void r4_level(const TwOffsets &O, const short *__restrict__ tw, int M, int k, bool do_scale, cx &x0, cx &x1, cx &x2, cx &x3)
{
  constexpr bool inv = INV;
  const int size = 4 * M;
  cx y0, y1, y2, y3;
  if (size == 64) {
    const short *ta = tw + (inv ? O.tw64 : O.tw64a), *tb = tw + (inv ? O.tw64c : O.tw64b);
    bfly4_sat(x0, cmult2(x1, ta + 2 * k, tb + 2 * k), cmult2(x2, ta + 32 + 2 * k, tb + 32 + 2 * k), cmult2(x3, ta + 64 + 2 * k, tb + 64 + 2 * k), inv, y0, y1, y2, y3);
  } else if (size == 256 && !inv) {
    const short *ta = tw + O.tw256a, *tb = tw + O.tw256b;
    bfly4_sat(x0, cmult2(x1, ta + 2 * k, tb + 2 * k), cmult2(x2, ta + 128 + 2 * k, tb + 128 + 2 * k), cmult2(x3, ta + 256 + 2 * k, tb + 256 + 2 * k), false, y0, y1, y2, y3);
  } else {
    const short *t4 = tw + (size == 256 ? O.tw256 : size == 1024 ? O.rad4_1024 : O.rad4_4096);
    bfly4_32(x0, x1, x2, x3, t4 + 2 * k, t4 + 2 * M + 2 * k, t4 + 4 * M + 2 * k, inv, y0, y1, y2, y3);
  }
  if (do_scale) {
    const int sh = size == 64 ? 3 : 1;
    y0.r >>= sh; y0.i >>= sh; y1.r >>= sh; y1.i >>= sh; y2.r >>= sh; y2.i >>= sh; y3.r >>= sh; y3.i >>= sh;
  }
  x0 = y0; x1 = y1; x2 = y2; x3 = y3;
}

// Slot-level OFDM front end fused into the transform's load / store phases (include/nrb200_dfts.h Part 3):
//   mode 1 (TX): apply_nr_rotation_TX on load, IDFT, cyclic-prefix insertion on store      (ofdm_mod.c:130-281, 337-376)
//   mode 2 (RX): FFT-window gather from the frame ring on load, DFT, apply_nr_rotation_RX on store (slot_fep_nr.c:223-332)
struct SlotIO {
  int n_symb, rotate;
  unsigned f_stride, t_stride, t_ring;
  unsigned t_off[14], prefix[14];
  unsigned r_start[2], r_len;               // the two sub-carrier ranges the reference rotates
  short rot[14][2];
  const unsigned *timeshift;                // RX: fp->timeshift_symbol_rotation (N c16)
};

// rotate_cpx_vector (cmult_sv.c:77-145, AVX2 branch): 8-element vector body saturates, the scalar tail truncates
__device__ __forceinline__ unsigned rotate_c16(unsigned w, int ar, int ai, bool body)
{
  const cx x = unpack(w);
  if (body) {
    const int nai = wrap16(-ai);
    return pack(cx{sat16(sra15((unsigned)(x.r * ar) + (unsigned)(x.i * nai))), sat16(sra15((unsigned)(x.r * ai) + (unsigned)(x.i * ar)))});
  }
  return pack(cx{wrap16(sra15((unsigned)(x.r * ar) - (unsigned)(x.i * ai))), wrap16(sra15((unsigned)(x.r * ai) + (unsigned)(x.i * ar)))});
}
// multadd_cpx_vector(zero_flag = 1) (cmult_vv.c:158-213): plain product >> 15, saturating
__device__ __forceinline__ unsigned mult_c16(unsigned w, unsigned t)
{
  const cx a = unpack(w), b = unpack(t);
  const int nai = wrap16(-a.i);
  return pack(cx{sat16(sra15((unsigned)(a.r * b.r) + (unsigned)(nai * b.i))), sat16(sra15((unsigned)(a.i * b.r) + (unsigned)(a.r * b.i)))});
}
// position of sub-carrier i inside the rotated ranges: -1 outside, else offset j within its range
__device__ __forceinline__ int range_pos(const SlotIO &S, unsigned i)
{
  if (i - S.r_start[0] < S.r_len) return (int)(i - S.r_start[0]);
  if (i - S.r_start[1] < S.r_len) return (int)(i - S.r_start[1]);
  return -1;
}

#ifndef NRB200_DFT_MINBLOCKS
#define NRB200_DFT_MINBLOCKS 4      /* 64 registers: 15.8 M idft4096/s against 12.9 M at 2 blocks (95 registers), measured */
#endif
template <int MODE, bool INV>
__global__ void __launch_bounds__(256, NRB200_DFT_MINBLOCKS) dft_kernel(DftPlan P, TwOffsets O, const short *__restrict__ tw, const unsigned *__restrict__ in,
                                                  unsigned *__restrict__ out, unsigned n, SlotIO S)
{
  extern __shared__ unsigned sm[];          // [tpb][N] natural-order input copy, then [tpb][N] work buffer
  const int N = P.N, tpb = P.tpb;
  unsigned *xin = sm, *buf = sm + tpb * N;
  constexpr bool inv = INV;                 // compile time: the butterflies select their outputs without SEL / ISETP
  const unsigned t0 = blockIdx.x * tpb;
  const int nt = min((unsigned)tpb, n - t0);
  if (MODE == 0) {
    for (int i = threadIdx.x; i < nt * N; i += blockDim.x) xin[i] = in[(size_t)t0 * N + i];
  } else {
    for (int w = threadIdx.x; w < nt * N; w += blockDim.x) {
      const unsigned tr = w / N, i = w - tr * N, t = t0 + tr, ant = t / S.n_symb, l = t - ant * S.n_symb;
      if (MODE == 1) {
        unsigned v = in[(size_t)ant * S.f_stride + (size_t)l * N + i];
        if (S.rotate) {
          const int j = range_pos(S, i);
          if (j >= 0) v = rotate_c16(v, S.rot[l][0], S.rot[l][1], (unsigned)j < (S.r_len & ~7u));
        }
        xin[w] = v;
      } else {
        unsigned k = S.t_off[l] + i;
        if (S.t_ring) k %= S.t_ring;
        xin[w] = in[(size_t)ant * S.t_stride + k];
      }
    }
  }
  __syncthreads();
  // ---- leaves: 16-point kernels, output slot = digit-reversed leaf id so that all later passes are in place
  const int T = P.r3 * P.r2, leaves4 = P.N4 >> 4, D = P.D;
  const short *ta16 = tw + (inv ? O.tw16 : O.tw16a), *tb16 = tw + (inv ? O.tw16c : O.tw16b);
  for (int w = threadIdx.x; w < nt * T * leaves4; w += blockDim.x) {
    const int tr = w / (T * leaves4), rem = w - tr * (T * leaves4);
    const int t = rem / leaves4, s = rem - t * leaves4;
    const int q3 = t / P.r2, q2 = t - q3 * P.r2;
    int mo = 0;                                            // m' = d1 + 4 d2 + ...; s = dD + 4 d(D-1) + ... + 4^(D-1) d1
    for (int l = 0, ss = s; l < D; l++, ss >>= 2) mo = (mo << 2) | (ss & 3);
    const unsigned *x = xin + tr * N;
    cx v[16];
#pragma unroll
    for (int m = 0; m < 16; m++) v[m] = unpack(x[q3 + P.r3 * (q2 + P.r2 * (mo + (m << (2 * D))))]);
    cx A[4][4];
#pragma unroll
    for (int c = 0; c < 4; c++) bfly4_sat(v[c], v[4 + c], v[8 + c], v[12 + c], inv, A[0][c], A[1][c], A[2][c], A[3][c]);
    unsigned yv[16];
#pragma unroll
    for (int k1 = 0; k1 < 4; k1++) {
      cx b1 = cmult2(A[k1][1], ta16 + 2 * k1, tb16 + 2 * k1), b2 = cmult2(A[k1][2], ta16 + 8 + 2 * k1, tb16 + 8 + 2 * k1),
         b3 = cmult2(A[k1][3], ta16 + 16 + 2 * k1, tb16 + 16 + 2 * k1);
      cx y0, y1, y2, y3;
      bfly4_sat(A[k1][0], b1, b2, b3, inv, y0, y1, y2, y3);
      yv[k1] = pack(y0); yv[k1 + 4] = pack(y1); yv[k1 + 8] = pack(y2); yv[k1 + 12] = pack(y3);
    }
    uint4 *y4 = reinterpret_cast<uint4 *>(buf + tr * N + t * P.N4 + s * 16);       // one leaf = 64 contiguous bytes: four 128-bit stores
#pragma unroll
    for (int q = 0; q < 4; q++) y4[q] = make_uint4(yv[4 * q], yv[4 * q + 1], yv[4 * q + 2], yv[4 * q + 3]);
  }
  __syncthreads();
  // ---- radix-4 levels, in place.  Two consecutive levels (M -> 4M -> 16M) are fused: a thread keeps the 16 values {k + M a + 4M b} in registers, does the
  // four size-4M butterflies (over a, twiddle index k) and then the four size-16M butterflies (over b, twiddle index k + M a) -- the same operations in the
  // same order as two separate passes, with one shared-memory round trip and one barrier instead of two.
  for (int l = 1, M = 16; l <= D;) {
    if (l + 1 <= D) {
      const bool last2 = (l + 1 == D) && T == 1;
      const bool scale2 = last2 ? (P.scale != 0) : true;
      const int per = N >> 4;
      for (int w = threadIdx.x; w < nt * per; w += blockDim.x) {
        const int tr = w / per, rem = w - tr * per;
        const int g = rem / M, k = rem - g * M;
        unsigned *p = buf + tr * N + g * 16 * M + k;
        cx v[4][4];
#pragma unroll
        for (int b = 0; b < 4; b++)
#pragma unroll
          for (int a = 0; a < 4; a++) v[a][b] = unpack(p[M * a + 4 * M * b]);
#pragma unroll
        for (int b = 0; b < 4; b++) r4_level<INV>(O, tw, M, k, true, v[0][b], v[1][b], v[2][b], v[3][b]);
#pragma unroll
        for (int a = 0; a < 4; a++) r4_level<INV>(O, tw, 4 * M, k + M * a, scale2, v[a][0], v[a][1], v[a][2], v[a][3]);
#pragma unroll
        for (int b = 0; b < 4; b++)
#pragma unroll
          for (int a = 0; a < 4; a++) p[M * a + 4 * M * b] = pack(v[a][b]);
      }
      __syncthreads();
      l += 2; M <<= 4;
    } else {
      const bool last_overall = (l == D) && T == 1;
      const bool do_scale = last_overall ? (P.scale != 0) : true;
      const int nb = nt * (N >> 2);
      for (int w = threadIdx.x; w < nb; w += blockDim.x) {
        const int tr = w / (N >> 2), rem = w - tr * (N >> 2);
        const int g = rem / M, k = rem - g * M;              // g enumerates (t, s') groups: base = g * 4M
        unsigned *p = buf + tr * N + g * 4 * M + k;
        cx x0 = unpack(p[0]), x1 = unpack(p[M]), x2 = unpack(p[2 * M]), x3 = unpack(p[3 * M]);
        r4_level<INV>(O, tw, M, k, do_scale, x0, x1, x2, x3);
        p[0] = pack(x0); p[M] = pack(x1); p[2 * M] = pack(x2); p[3 * M] = pack(x3);
      }
      __syncthreads();
      l += 1; M <<= 2;
    }
  }
  // ---- radix-2 level
  if (P.r2 == 2) {
    const int M = P.N4, size = 2 * M;
    const bool do_scale = P.r3 == 1 ? (P.scale != 0) : true;
    for (int w = threadIdx.x; w < nt * (N >> 1); w += blockDim.x) {
      const int tr = w / (N >> 1), rem = w - tr * (N >> 1);
      const int g = rem / M, k = rem - g * M;
      unsigned *p = buf + tr * N + g * size + k;
      const cx x0 = unpack(p[0]), x1 = unpack(p[M]);
      cx y0, y1;
      if (size == 128 && !inv) {
        const cx t = cmult2(x1, tw + O.tw128a + 2 * k, tw + O.tw128b + 2 * k);
        y0 = sadd(x0, t); y1 = ssub(x0, t);
      } else {
        const short *t2 = tw + (size == 128 ? O.tw128 : size == 512 ? O.tw512 : P.rad2_tw) + 2 * k;
        unsigned br, bi;
        const cx T2 = ldtw(t2);
        cm32(x1, T2.r, T2.i, inv, br, bi);
        const unsigned ar = (unsigned)(x0.r * 32767), ai = (unsigned)(x0.i * 32767);
        y0 = pk32(ar + br, ai + bi); y1 = pk32(ar - br, ai - bi);
      }
      if (do_scale) { y0.r = mulhrs(y0.r, 23170); y0.i = mulhrs(y0.i, 23170); y1.r = mulhrs(y1.r, 23170); y1.i = mulhrs(y1.i, 23170); }
      p[0] = pack(y0); p[M] = pack(y1);
    }
    __syncthreads();
  }
  // ---- radix-3 level
  if (P.r3 == 3) {
    const int M = N / 3;
    const short *twa = tw + P.rad3_tw, *twb = twa + 2 * M;
    for (int w = threadIdx.x; w < nt * M; w += blockDim.x) {
      const int tr = w / M, k = w - tr * M;
      unsigned *p = buf + tr * N + k;
      const cx x0 = unpack(p[0]);
      unsigned r, i, r2, i2;
      const cx TA = ldtw(twa + 2 * k), TB = ldtw(twb + 2 * k);
      cm32(unpack(p[M]), TA.r, TA.i, inv, r, i);
      const cx x1 = pk32(r, i);
      cm32(unpack(p[2 * M]), TB.r, TB.i, inv, r, i);
      const cx x2 = pk32(r, i);
      cx y0 = sadd(x0, sadd(x1, x2));
      cm32(x1, -16384, -28378, inv, r, i); cm32(x2, -16384, 28378, inv, r2, i2);
      cx y1 = sadd(x0, pk32(r + r2, i + i2));
      cm32(x1, -16384, 28378, inv, r, i); cm32(x2, -16384, -28378, inv, r2, i2);
      cx y2 = sadd(x0, pk32(r + r2, i + i2));
      if (P.scale == 1) {
        y0.r = mulhrs(y0.r, 18919); y0.i = mulhrs(y0.i, 18919); y1.r = mulhrs(y1.r, 18919); y1.i = mulhrs(y1.i, 18919);
        y2.r = mulhrs(y2.r, 18919); y2.i = mulhrs(y2.i, 18919);
      }
      p[0] = pack(y0); p[M] = pack(y1); p[2 * M] = pack(y2);
    }
    __syncthreads();
  }
  if (MODE == 0) {
    for (int i = threadIdx.x; i < nt * N; i += blockDim.x) out[(size_t)t0 * N + i] = buf[i];
  } else {
    for (int w = threadIdx.x; w < nt * N; w += blockDim.x) {
      const unsigned tr = w / N, i = w - tr * N, t = t0 + tr, ant = t / S.n_symb, l = t - ant * S.n_symb;
      unsigned v = buf[w];
      if (MODE == 1) {
        unsigned *o = out + (size_t)ant * S.t_stride + S.t_off[l];
        const unsigned cp = S.prefix[l];
        o[cp + i] = v;
        if (i >= (unsigned)N - cp) o[i - ((unsigned)N - cp)] = v;        // cyclic prefix = the last cp samples
      } else {
        if (S.rotate) {
          const int j = range_pos(S, i);
          if (j >= 0) {
            v = rotate_c16(v, S.rot[l][0], wrap16(-S.rot[l][1]), (unsigned)j < (S.r_len & ~7u));
            if ((unsigned)j < (S.r_len & ~3u)) v = mult_c16(v, __ldg(S.timeshift + i));
          }
        }
        out[(size_t)ant * S.f_stride + (size_t)l * N + i] = v;
      }
    }
  }
}

#include "dft4096_tma.cuh"

// ------------------------------------------------------------------------------------------------ four-way DFT-s-OFDM family (12 ... 3240)
// oai_dfts.c:4352-7706: entry points that work on 128-bit vectors holding four independent transforms (element n of transform l is c16 number 4 n + l).
// N = 12 * R[L-1] * ... * R[0], decimation in time at every level: M-point transforms of x[m + R n], then per k < M one radix-R butterfly (bfly{2,3,4,5}_tw1 for
// k = 0, bfly{2,3,4,5} with generated twiddles otherwise) writing y[k + q M], then mulhrs by the level's constant when that level is called with scale 1.
// Here: one CTA per call (4 N c16 in shared memory, lanes interleaved as in memory, so consecutive threads touch consecutive words), the 12-point kernels
// read their digit-reversed inputs straight from global memory, every later level runs in place.
struct SmallPlan {
  int N, L;
  int R[5], norm[5], apply[5], tw[5];   // level 0 = top; tw: blob offset (shorts) of (R-1) planes of (M-1) {re, im} pairs, plane p-1, entry k-1
};

__device__ __forceinline__ cx pcm(cx x, cx w) { unsigned r, i; cm32(x, w.r, w.i, false, r, i); return pk32(r, i); }   // packed_cmult
__device__ __forceinline__ void bfly3_fwd(cx x0, cx x1, cx x2, cx &y0, cx &y1, cx &y2)
{
  unsigned r, i, r2, i2;
  y0 = sadd(x0, sadd(x1, x2));
  cm32(x1, -16384, -28378, false, r, i); cm32(x2, -16384, 28378, false, r2, i2);
  y1 = sadd(x0, pk32(r + r2, i + i2));
  cm32(x1, -16384, 28378, false, r, i); cm32(x2, -16384, -28378, false, r2, i2);
  y2 = sadd(x0, pk32(r + r2, i + i2));
}
__device__ __forceinline__ void bfly5_fwd(cx x0, const cx (&x)[4], cx (&y)[5])
{
  const int W[4][2] = {{10126, -31163}, {-26509, -19260}, {-26510, 19260}, {10126, 31163}};   // W15 .. W45 (oai_dfts.c:324-327)
  const int sel[4][4] = {{0, 1, 2, 3}, {1, 3, 0, 2}, {2, 0, 3, 1}, {3, 2, 1, 0}};
  y[0] = sadd(x0, sadd(x[0], sadd(x[1], sadd(x[2], x[3]))));
#pragma unroll
  for (int q = 0; q < 4; q++) {
    unsigned r = 0, i = 0;
#pragma unroll
    for (int p = 0; p < 4; p++) { unsigned a, b; cm32(x[p], W[sel[q][p]][0], W[sel[q][p]][1], false, a, b); r += a; i += b; }
    y[q + 1] = sadd(x0, pk32(r, i));
  }
}

__global__ void __launch_bounds__(256)
dft_small_kernel(SmallPlan P, const short *__restrict__ tw, const unsigned *__restrict__ in, unsigned *__restrict__ out, unsigned n_calls)
{
  extern __shared__ unsigned sm[];
  const int N = P.N, L = P.L;
  for (unsigned call = blockIdx.x; call < n_calls; call += gridDim.x) {
    const unsigned *src = in + (size_t)call * 4 * N;
    // ---- 12-point kernels (dft12f :4365-4470): three bfly4_tw1, then bfly3_tw1 and three bfly3 with the hand-entered W12 constants
    for (int w = threadIdx.x; w < 4 * (N / 12); w += blockDim.x) {
      const int l = w & 3, b = w >> 2;
      int stride = 1, first = 0;                      // input index of element n12 of block b: first + stride * n12
      {
        int bb = b;
        // digits of b, innermost level last: n = m0 + R0 (m1 + R1 (... + R[L-1] n12))
        int mult = 1, digs[5];
        for (int j = L - 1; j >= 0; j--) { digs[j] = bb % P.R[j]; bb /= P.R[j]; }
        for (int j = 0; j < L; j++) { first += digs[j] * mult; mult *= P.R[j]; }
        stride = mult;
      }
      cx x[12], t[12], y[12];
#pragma unroll
      for (int i = 0; i < 12; i++) x[i] = unpack(__ldg(src + 4 * (first + stride * i) + l));
#pragma unroll
      for (int c = 0; c < 3; c++) bfly4_sat(x[c], x[c + 3], x[c + 6], x[c + 9], false, t[c], t[c + 3], t[c + 6], t[c + 9]);
      bfly3_fwd(t[0], t[1], t[2], y[0], y[4], y[8]);
      bfly3_fwd(t[3], pcm(t[4], {28377, -16383}), pcm(t[5], {16383, -28377}), y[1], y[5], y[9]);
      bfly3_fwd(t[6], pcm(t[7], {16383, -28377}), pcm(t[8], {-16383, -28377}), y[2], y[6], y[10]);
      bfly3_fwd(t[9], pcm(t[10], {0, -32767}), pcm(t[11], {-32767, 0}), y[3], y[7], y[11]);
#pragma unroll
      for (int k = 0; k < 12; k++) sm[(b * 12 + k) * 4 + l] = pack(y[k]);
    }
    __syncthreads();
    int M = 12;
    for (int j = L - 1; j >= 0; j--) {
      const int R = P.R[j], NJ = R * M, norm = P.norm[j];
      const bool sc = P.apply[j] != 0;
      const short *t0 = tw + P.tw[j];
      for (int w = threadIdx.x; w < 4 * (N / R); w += blockDim.x) {
        const int l = w & 3, t = w >> 2, g = t / M, k = t - g * M;
        unsigned *base = sm + (size_t)(g * NJ + k) * 4 + l;
        cx v[5], o[5];
#pragma unroll
        for (int p = 0; p < 5; p++) if (p < R) v[p] = unpack(base[p * M * 4]);
        cx W[4];
        if (k) {
#pragma unroll
          for (int p = 1; p < 5; p++) if (p < R) W[p - 1] = ldtw(t0 + 2 * ((p - 1) * (M - 1) + (k - 1)));
        }
        if (R == 2) {
          if (!k) { o[0] = sadd(v[0], v[1]); o[1] = ssub(v[0], v[1]); }
          else {
            unsigned br, bi;
            const unsigned ar = (unsigned)(v[0].r * 32767), ai = (unsigned)(v[0].i * 32767);
            cm32(v[1], W[0].r, W[0].i, false, br, bi);
            o[0] = pk32(ar + br, ai + bi); o[1] = pk32(ar - br, ai - bi);
          }
        } else if (R == 3) {
          if (!k) bfly3_fwd(v[0], v[1], v[2], o[0], o[1], o[2]);
          else bfly3_fwd(v[0], pcm(v[1], W[0]), pcm(v[2], W[1]), o[0], o[1], o[2]);
        } else if (R == 4) {
          if (!k) bfly4_sat(v[0], v[1], v[2], v[3], false, o[0], o[1], o[2], o[3]);
          else {
            unsigned x1r, x1i, x2r, x2i, x3r, x3i;
            cm32(v[1], W[0].r, W[0].i, false, x1r, x1i);
            cm32(v[2], W[1].r, W[1].i, false, x2r, x2i);
            cm32(v[3], W[2].r, W[2].i, false, x3r, x3i);
            const cx d0 = pk32(x1r + x2r + x3r, x1i + x2i + x3i), da = pk32(x1i - (x2r + x3i), (x3r - x2i) - x1r);
            const cx d2 = pk32((x2r - x3r) - x1r, (x2i - x3i) - x1i), db = pk32((x3i - x2r) - x1i, x1r - (x2i + x3r));
            o[0] = {wrap16(v[0].r + d0.r), wrap16(v[0].i + d0.i)}; o[1] = {wrap16(v[0].r + da.r), wrap16(v[0].i + da.i)};
            o[2] = {wrap16(v[0].r + d2.r), wrap16(v[0].i + d2.i)}; o[3] = {wrap16(v[0].r + db.r), wrap16(v[0].i + db.i)};
          }
        } else {
          cx xx[4];
#pragma unroll
          for (int p = 0; p < 4; p++) xx[p] = k ? pcm(v[p + 1], W[p]) : v[p + 1];
          bfly5_fwd(v[0], xx, o);
        }
#pragma unroll
        for (int q = 0; q < 5; q++)
          if (q < R) {
            cx r = o[q];
            if (sc) { r.r = mulhrs(r.r, norm); r.i = mulhrs(r.i, norm); }
            base[q * M * 4] = pack(r);
          }
      }
      __syncthreads();
      M = NJ;
    }
    unsigned *dst = out + (size_t)call * 4 * N;
    for (int i = threadIdx.x; i < 4 * N; i += blockDim.x) dst[i] = sm[i];
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------ sizes above 8192 (12288 ... 98304)
// oai_dfts.c:2846-3138, :3614-4350: radix-3 / radix-4 / radix-2 levels on top of a transform that fits in shared memory (4096, 6144 or 8192 points).  Here:
// one gather pass that puts every base transform's decimated input in a row, the batched shared-memory kernel over the rows, then one in-place pass over
// global memory per top level (a transform of 98304 points is 384 KB: L2 resident).
struct BigPlan {
  int N, B, T, L;                 // N = T * B, T = R[0] * ... * R[L-1], level 0 = top
  int R[3], sval[3], tw[3];       // radix, the scale argument that level's function receives, blob offset of its (R-1) planes of NJ/R twiddles (k from 0)
  int base_scale;
};

__global__ void big_gather_kernel(BigPlan P, const unsigned *__restrict__ in, unsigned *__restrict__ rows, unsigned n_calls)
{
  const size_t total = (size_t)n_calls * P.N;
  for (size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x; w < total; w += (size_t)gridDim.x * blockDim.x) {
    const unsigned call = (unsigned)(w / P.N), rem = (unsigned)(w - (size_t)call * P.N);
    const unsigned srow = rem / P.B, n = rem - srow * P.B;
    unsigned first = 0, mult = 1, ss = srow, digs[3];
    for (int j = P.L - 1; j >= 0; j--) { digs[j] = ss % P.R[j]; ss /= P.R[j]; }       // row id = ((m0 R1 + m1) R2 + m2); input index m0 + R0 (m1 + R1 (m2 + R2 n))
    for (int j = 0; j < P.L; j++) { first += digs[j] * mult; mult *= P.R[j]; }
    rows[w] = in[(size_t)call * P.N + first + (size_t)P.T * n];
  }
}

template <bool INV>
__global__ void big_combine_kernel(int N, int NJ, int R, int sval, const short *__restrict__ tw, const unsigned *src, unsigned *dst, unsigned n_calls)
{
  constexpr bool inv = INV;
  const int M = NJ / R;
  const size_t total = (size_t)n_calls * (N / R);
  for (size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x; w < total; w += (size_t)gridDim.x * blockDim.x) {
    const unsigned call = (unsigned)(w / (N / R)), rem = (unsigned)(w - (size_t)call * (N / R));
    const unsigned g = rem / M, k = rem - g * M;
    const size_t base = (size_t)call * N + (size_t)g * NJ + k;
    if (R == 3) {                                                          // bfly3 / ibfly3 (:477-560), twiddles from init_rad3
      const cx x0 = unpack(src[base]), a = unpack(src[base + M]), b = unpack(src[base + 2 * M]);
      const cx T1 = ldtw(tw + 2 * k), T2 = ldtw(tw + 2 * (M + k));
      unsigned r, i, r2, i2;
      cm32(a, T1.r, T1.i, inv, r, i);
      const cx x1 = pk32(r, i);
      cm32(b, T2.r, T2.i, inv, r, i);
      const cx x2 = pk32(r, i);
      cx y0 = sadd(x0, sadd(x1, x2));
      cm32(x1, -16384, -28378, inv, r, i); cm32(x2, -16384, 28378, inv, r2, i2);
      cx y1 = sadd(x0, pk32(r + r2, i + i2));
      cm32(x1, -16384, 28378, inv, r, i); cm32(x2, -16384, -28378, inv, r2, i2);
      cx y2 = sadd(x0, pk32(r + r2, i + i2));
      if (sval == 1) {
        y0 = {mulhrs(y0.r, 18919), mulhrs(y0.i, 18919)}; y1 = {mulhrs(y1.r, 18919), mulhrs(y1.i, 18919)}; y2 = {mulhrs(y2.r, 18919), mulhrs(y2.i, 18919)};
      }
      dst[base] = pack(y0); dst[base + M] = pack(y1); dst[base + 2 * M] = pack(y2);
    } else if (R == 4) {                                                   // bfly4_256 / ibfly4_256 (:633-721), twiddles from init_rad4
      const cx x0 = unpack(src[base]), x1 = unpack(src[base + M]), x2 = unpack(src[base + 2 * M]), x3 = unpack(src[base + 3 * M]);
      cx y0, y1, y2, y3;
      bfly4_32(x0, x1, x2, x3, tw + 2 * k, tw + 2 * (M + k), tw + 2 * (2 * M + k), inv, y0, y1, y2, y3);
      if (sval > 0) {
        const int sh = NJ == 65536 ? sval : 1;
        y0.r >>= sh; y0.i >>= sh; y1.r >>= sh; y1.i >>= sh; y2.r >>= sh; y2.i >>= sh; y3.r >>= sh; y3.i >>= sh;
      }
      dst[base] = pack(y0); dst[base + M] = pack(y1); dst[base + 2 * M] = pack(y2); dst[base + 3 * M] = pack(y3);
    } else {                                                               // bfly2_256 / ibfly2_256 (:390-476), twiddles from init_rad2
      const cx x0 = unpack(src[base]), x1 = unpack(src[base + M]);
      const cx T2 = ldtw(tw + 2 * k);
      unsigned br, bi;
      cm32(x1, T2.r, T2.i, inv, br, bi);
      const unsigned ar = (unsigned)(x0.r * 32767), ai = (unsigned)(x0.i * 32767);
      cx y0 = pk32(ar + br, ai + bi), y1 = pk32(ar - br, ai - bi);
      if (sval > 0) { y0 = {mulhrs(y0.r, 23170), mulhrs(y0.i, 23170)}; y1 = {mulhrs(y1.r, 23170), mulhrs(y1.i, 23170)}; }
      dst[base] = pack(y0); dst[base + M] = pack(y1);
    }
  }
}

// ------------------------------------------------------------------------------------------------ host side
// the four-way family: N = R x M, whether the M-point transforms are called with scale 1, and the mulhrs constant applied at this level when ITS scale is 1
// (dft_norm_table oai_dfts.c:352-368 for 24 ... 300, 1/sqrt(R) in Q15 above; read off every dftN :4680-7706).  768 is dft768p (:6330), reachable only through
// 2304: the reference's dft2304 (:7288) calls the single-transform dft768 on four-way data and then combines uninitialised stack, so its result is not a function
// of its input; this library returns the "768 x 3" transform its comment describes (defect 11).
// sizes above 8192: the radix of the top level (the rest is the next smaller size).  9216 and 73728 are AssertFatal("Need to do this") in the reference;
// dft32768/idft32768 overrun their stack buffers (:2966-3004: 256 x 64 input vectors read from a 4096-vector array) and crash, dft98304/idft98304 call them,
// idft65536 reads its second and third twiddle planes at twice the right offset, the third beyond the table (:4200-4202): for these four sizes this library
// computes the transform the code intends (same level arithmetic as the sizes that work), which cannot be pinned against the reference.
struct BigSize { int N, R; };
const BigSize kBig[] = {{12288, 3}, {16384, 4}, {18432, 3}, {24576, 3}, {32768, 2}, {36864, 3}, {49152, 3}, {65536, 4}, {98304, 3}};
constexpr int kNumBig = (int)(sizeof(kBig) / sizeof(kBig[0]));
int big_index(int N) { for (int i = 0; i < kNumBig; i++) if (kBig[i].N == N) return i; return -1; }

struct SmallSize { int N, R, M, subscale, norm; };
const SmallSize kSmall[] = {
    {24, 2, 12, 0, 6689}, {36, 3, 12, 0, 5461}, {48, 4, 12, 0, 4729}, {60, 5, 12, 0, 4230}, {72, 2, 36, 1, 23170}, {96, 2, 48, 0, 3344}, {108, 3, 36, 0, 3153},
    {120, 2, 60, 0, 2991}, {144, 3, 48, 1, 18918}, {180, 3, 60, 1, 18918}, {192, 4, 48, 1, 16384}, {216, 3, 72, 1, 18918}, {240, 4, 60, 1, 16384},
    {288, 3, 96, 1, 18918}, {300, 5, 60, 1, 14654}, {324, 3, 108, 1, 18918}, {360, 3, 120, 1, 18918}, {384, 4, 96, 1, 16384}, {432, 4, 108, 1, 16384},
    {480, 4, 120, 1, 16384}, {540, 3, 180, 1, 18918}, {576, 3, 192, 1, 18918}, {600, 2, 300, 1, 23170}, {648, 3, 216, 1, 18918}, {720, 4, 180, 1, 16384},
    {768, 4, 192, 1, 16384}, {864, 3, 288, 1, 18918}, {900, 3, 300, 1, 18918}, {960, 4, 240, 1, 16384}, {972, 3, 324, 1, 18918}, {1080, 3, 360, 1, 18918},
    {1152, 4, 288, 1, 16384}, {1200, 4, 300, 1, 16384}, {1296, 3, 432, 1, 18918}, {1440, 3, 480, 1, 18918}, {1500, 5, 300, 1, 14654}, {1620, 3, 540, 1, 18918},
    {1728, 3, 576, 1, 18918}, {1800, 3, 600, 1, 18918}, {1920, 4, 480, 1, 16384}, {1944, 3, 648, 1, 18918}, {2160, 3, 720, 1, 18918}, {2304, 3, 768, 1, 18918},
    {2400, 4, 600, 1, 16384}, {2592, 3, 864, 1, 18918}, {2700, 3, 900, 1, 18918}, {2880, 3, 960, 1, 18918}, {2916, 3, 972, 1, 18918}, {3000, 5, 600, 1, 14654},
    {3240, 3, 1080, 1, 18918}};
constexpr int kNumSmall = (int)(sizeof(kSmall) / sizeof(kSmall[0]));
int small_index(int N) { for (int i = 0; i < kNumSmall; i++) if (kSmall[i].N == N) return i; return -1; }
// sizes the `dft` table serves with the four-way entry points (768 is the OFDM transform there)
bool is_fourway(int N) { return N == 12 || (N != 768 && small_index(N) >= 0); }

struct DftCtx {
  std::recursive_mutex mu;
  bool inited = false;
  int dev = 0, sm_count = 148;
  short *d_tw = nullptr;
  TwOffsets off;
  int big_tw[9];                          // blob offsets of the top-level twiddles of kBig[i]
  int small_tw[51];                       // blob offset of the level twiddles of four-way size kSmall[i].N (-1: none)
  cudaStream_t stream = nullptr;
  void *d_in = nullptr, *d_out = nullptr, *h_in = nullptr, *h_out = nullptr;
  size_t cap = 0;
  std::atomic<uint64_t> launches{0};
  std::string last_error;
};
// one context per device; the calling thread's current device selects it (nrb200_dfts_set_device; inside libldpc_b200.so the thread's
// nrb200_set_device selection; default NRB200_DEVICE / LOCAL_RANK / 0)
#ifdef NRB200_DFTS_INTERNAL
}  // namespace
namespace nrb200 { int current_device(); }
namespace {
int cur_ddev() { return nrb200::current_device(); }
#else
thread_local int tls_ddev = -1;
int cur_ddev()
{
  if (tls_ddev < 0) {
    int n = 0, want = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); n = 0; }
    if (const char *s = getenv("NRB200_DEVICE")) want = atoi(s);
    else if (const char *s2 = getenv("LOCAL_RANK")) want = n > 0 ? atoi(s2) % n : 0;
    if (want < 0 || want >= (n > 0 ? n : 1) || want >= 16) want = 0;
    tls_ddev = want;
  }
  return tls_ddev;
}
#endif
DftCtx &dctx() { static DftCtx c[16]; return c[cur_ddev()]; }

short rnd16(double v) { return (short)std::round(v); }

int dft_init()
{
  DftCtx &c = dctx();
  std::lock_guard<std::recursive_mutex> lk(c.mu);
  if (c.inited) return 0;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) { c.last_error = "no CUDA device"; return -1; }
  const int want = cur_ddev();
  if (want < 0 || want >= n) { c.last_error = "no such CUDA device"; return -1; }
  if (cudaSetDevice(want) != cudaSuccess) { c.last_error = "cudaSetDevice failed"; return -1; }
  c.dev = want;
  std::vector<short> blob;
  auto put = [&](const int16_t *p, int len) { int o = (int)blob.size(); blob.insert(blob.end(), p, p + len); return o; };
  TwOffsets &O = c.off;
  O.tw16 = put(NRB200_TW16, 24); O.tw16a = put(NRB200_TW16A, 24); O.tw16b = put(NRB200_TW16B, 24); O.tw16c = put(NRB200_TW16C, 24);
  O.tw64 = put(NRB200_TW64, 96); O.tw64a = put(NRB200_TW64A, 96); O.tw64b = put(NRB200_TW64B, 96); O.tw64c = put(NRB200_TW64C, 96);
  O.tw128 = put(NRB200_TW128, 128); O.tw128a = put(NRB200_TW128A, 128); O.tw128b = put(NRB200_TW128B, 128);
  O.tw256 = put(NRB200_TW256, 384); O.tw256a = put(NRB200_TW256A, 384); O.tw256b = put(NRB200_TW256B, 384);
  O.tw512 = put(NRB200_TW512, 512);
  auto rad4 = [&](int N) {   // init_rad4 (oai_dfts.c:7709-7722): three planes W^k, W^2k, W^3k of N/4 entries
    int o = (int)blob.size();
    for (int p = 1; p <= 3; p++)
      for (int i = 0; i < N / 4; i++) { blob.push_back(rnd16(32767.0 * cos(2 * M_PI * p * i / N))); blob.push_back((short)-rnd16(32767.0 * sin(2 * M_PI * p * i / N))); }
    return o;
  };
  auto rad2 = [&](int N) {   // init_rad2 (:7747-7756)
    int o = (int)blob.size();
    for (int i = 0; i < N / 2; i++) { blob.push_back(rnd16(32767.0 * cos(2 * M_PI * i / N))); blob.push_back((short)-rnd16(32767.0 * sin(2 * M_PI * i / N))); }
    return o;
  };
  auto rad3 = [&](int N) {   // init_rad3 (:7772-7782): twa then twb
    int o = (int)blob.size();
    for (int p = 1; p <= 2; p++)
      for (int i = 0; i < N / 3; i++) { blob.push_back(rnd16(32767.0 * cos(2 * M_PI * p * i / N))); blob.push_back((short)-rnd16(32767.0 * sin(2 * M_PI * p * i / N))); }
    return o;
  };
  O.rad4_1024 = rad4(1024); O.rad4_4096 = rad4(4096); O.rad2_2048 = rad2(2048); O.rad2_8192 = rad2(8192);
  // dp_tables: byte-arranged copies of a twiddle table for bfly4_32dp (see there), 4 shorts per twiddle, 8-byte aligned
  auto dp_table = [&](int src, int count, bool inverse) {
    while (blob.size() % 4) blob.push_back(0);
    const int o = (int)blob.size();
    for (int i = 0; i < count; i++) {
      const int wr = blob[src + 2 * i], wi = blob[src + 2 * i + 1];
      const int a = wr, b = inverse ? wi : -wi, cc = inverse ? -wi : wi, d = wr;
      auto lo = [](int v) { return v & 0xFF; };
      auto hi = [](int v) { return (v >> 8) & 0xFF; };
      blob.push_back((short)(lo(a) | (lo(b) << 8))); blob.push_back((short)(hi(a) | (hi(b) << 8)));
      blob.push_back((short)(lo(cc) | (lo(d) << 8))); blob.push_back((short)(hi(cc) | (hi(d) << 8)));
    }
    return o;
  };
  auto dp_pair = [&](int srca, int srcb, int count) {       // packed_cmult2's two tables: re = xr a.r + xi a.i, im = xr b.r + xi b.i
    while (blob.size() % 4) blob.push_back(0);
    const int o = (int)blob.size();
    for (int i = 0; i < count; i++) {
      const int a = blob[srca + 2 * i], b = blob[srca + 2 * i + 1], cc = blob[srcb + 2 * i], d = blob[srcb + 2 * i + 1];
      auto lo = [](int v) { return v & 0xFF; };
      auto hi = [](int v) { return (v >> 8) & 0xFF; };
      blob.push_back((short)(lo(a) | (lo(b) << 8))); blob.push_back((short)(hi(a) | (hi(b) << 8)));
      blob.push_back((short)(lo(cc) | (lo(d) << 8))); blob.push_back((short)(hi(cc) | (hi(d) << 8)));
    }
    return o;
  };
  O.dps16[0] = dp_pair(O.tw16a, O.tw16b, 12); O.dps16[1] = dp_pair(O.tw16, O.tw16c, 12);
  O.dps64[0] = dp_pair(O.tw64a, O.tw64b, 48); O.dps64[1] = dp_pair(O.tw64, O.tw64c, 48);
  O.dps256f = dp_pair(O.tw256a, O.tw256b, 192);
  O.dp256i = dp_table(O.tw256, 192, true);
  for (int dir = 0; dir < 2; dir++) { O.dp1024[dir] = dp_table(O.rad4_1024, 768, dir == 1); O.dp4096[dir] = dp_table(O.rad4_4096, 3072, dir == 1); }
  const int r3n[4] = {768, 1536, 3072, 6144};
  for (int i = 0; i < 4; i++) O.rad3[i] = rad3(r3n[i]);
  for (int i = 0; i < kNumBig; i++) c.big_tw[i] = kBig[i].R == 3 ? rad3(kBig[i].N) : kBig[i].R == 4 ? rad4(kBig[i].N) : rad2(kBig[i].N);
  for (int i = 0; i < kNumSmall; i++) {   // init_rad{2,3,4,5}_rep (:7725-7828): entries k = 1 .. M-1 of W^(p k), p = 1 .. R-1, one copy instead of four
    const int N = kSmall[i].N, R = kSmall[i].R, M = kSmall[i].M;
    c.small_tw[i] = (int)blob.size();
    for (int p = 1; p < R; p++)
      for (int k = 1; k < M; k++) { blob.push_back(rnd16(32767.0 * cos(2 * M_PI * p * k / N))); blob.push_back((short)-rnd16(32767.0 * sin(2 * M_PI * p * k / N))); }
  }
  if (cudaMalloc(&c.d_tw, blob.size() * sizeof(short)) != cudaSuccess) { c.last_error = "cudaMalloc twiddles"; return -1; }
  cudaMemcpy(c.d_tw, blob.data(), blob.size() * sizeof(short), cudaMemcpyHostToDevice);
  if (cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking) != cudaSuccess) { c.last_error = "stream"; return -1; }
  cudaFuncSetAttribute(dft_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 8192 * 4);
  cudaFuncSetAttribute(dft_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 8192 * 4);
  cudaFuncSetAttribute(dft_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 8192 * 4);
  cudaFuncSetAttribute(dft_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 8192 * 4);
  cudaFuncSetAttribute(dft4096_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dft4096::kSmemBytes);
  cudaFuncSetAttribute(dft4096_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dft4096::kSmemBytes);
  cudaFuncSetAttribute(dft4096_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dft4096::kSmemBytes);
  cudaFuncSetAttribute(dft4096_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dft4096::kSmemBytes);
  {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, c.dev);
    c.sm_count = prop.multiProcessorCount;
  }
  cudaFuncSetAttribute(dft_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 3240 * 4);
  c.inited = true;
  return 0;
}

bool make_plan(int N, int inverse, int scale, DftPlan *P)
{
  const DftCtx &c = dctx();
  int r3 = (N % 3 == 0) ? 3 : 1, rest = N / r3, r2 = 1, D = -1;
  for (int d = 1, n4 = 64; d <= 4; d++, n4 *= 4) {
    if (rest == n4) { D = d; break; }
    if (rest == 2 * n4) { D = d; r2 = 2; break; }
  }
  if (D < 0 || N > 8192) return false;
  P->N = N; P->r3 = r3; P->r2 = r2; P->N4 = rest / r2; P->D = D; P->inverse = inverse; P->scale = scale;
  P->tpb = N >= 2048 ? 1 : 2048 / N;
  P->top4_tw = -1;
  P->rad2_tw = r2 == 2 ? (2 * P->N4 == 2048 ? c.off.rad2_2048 : 2 * P->N4 == 8192 ? c.off.rad2_8192 : -1) : -1;
  P->rad3_tw = r3 == 3 ? c.off.rad3[N == 768 ? 0 : N == 1536 ? 1 : N == 3072 ? 2 : 3] : -1;
  return true;
}

int launch_dft_small(int N, uint32_t n_calls, const int16_t *d_in, int16_t *d_out, int scale, cudaStream_t st);
int launch_dft_big(int N, int inverse, uint32_t n_calls, const int16_t *d_in, int16_t *d_out, int scale, cudaStream_t st);
bool is_fourway(int N);

// the 4096-point bulk-copy kernel needs 16-byte aligned sample arrays (NRB200_DFT_TMA=0 keeps the generic kernel, for A/B runs)
static bool use_tma4096(int N, const void *a, const void *b)
{
  static const bool on = []() { const char *e = getenv("NRB200_DFT_TMA"); return !(e && atoi(e) == 0); }();
  return on && N == 4096 && ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0;
}

int launch_dft(int N, int inverse, uint32_t n, const int16_t *d_in, int16_t *d_out, int scale, cudaStream_t st)
{
  if (is_fourway(N)) return inverse ? -4 : launch_dft_small(N, n, d_in, d_out, scale, st);
  if (N > 8192) return launch_dft_big(N, inverse, n, d_in, d_out, scale, st);
  DftPlan P;
  if (!make_plan(N, inverse, scale, &P)) return -4;
  if (n == 0) return 0;
  DftCtx &c = dctx();
  const unsigned grid = (n + P.tpb - 1) / P.tpb;
  const size_t smem = (size_t)2 * P.tpb * N * 4;
  if (use_tma4096(N, d_in, d_out)) {   // persistent, bulk-copy fed, conflict-free layouts (dft4096_tma.cuh)
    const unsigned g4 = std::min<unsigned>(n, 3u * (unsigned)c.sm_count);
    if (inverse) dft4096_kernel<0, true><<<g4, 256, dft4096::kSmemBytes, st>>>(P, c.off, c.d_tw, (const unsigned *)d_in, (unsigned *)d_out, n, SlotIO{});
    else dft4096_kernel<0, false><<<g4, 256, dft4096::kSmemBytes, st>>>(P, c.off, c.d_tw, (const unsigned *)d_in, (unsigned *)d_out, n, SlotIO{});
  } else if (inverse) dft_kernel<0, true><<<grid, 256, smem, st>>>(P, c.off, c.d_tw, (const unsigned *)d_in, (unsigned *)d_out, n, SlotIO{});
  else dft_kernel<0, false><<<grid, 256, smem, st>>>(P, c.off, c.d_tw, (const unsigned *)d_in, (unsigned *)d_out, n, SlotIO{});
  c.launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { std::lock_guard<std::recursive_mutex> lk(c.mu); c.last_error = std::string("dft launch: ") + cudaGetErrorString(e); return -2; }
  return 0;
}

bool make_big_plan(int N, int inverse, int scale, BigPlan *P)
{
  const DftCtx &c = dctx();
  if (big_index(N) < 0) return false;
  if (N == 65536 && !inverse) return false;                 // the reference has no dft65536
  P->N = N; P->L = 0; P->T = 1;
  int n = N, sv = scale;
  while (n > 8192) {
    const int i = big_index(n);
    if (i < 0 || P->L >= 3) return false;
    const int R = kBig[i].R;
    P->R[P->L] = R; P->sval[P->L] = sv; P->tw[P->L] = c.big_tw[i];
    sv = (n == 12288 || n == 18432) ? sv : 1;                 // dft12288 / dft18432 hand their scale argument down (:3641, :3755), every other level passes 1
    n /= R; P->T *= R; P->L++;
  }
  for (int j = P->L; j < 3; j++) { P->R[j] = 1; P->sval[j] = 0; P->tw[j] = 0; }
  P->B = n; P->base_scale = sv;
  return true;
}

int launch_dft(int N, int inverse, uint32_t n, const int16_t *d_in, int16_t *d_out, int scale, cudaStream_t st);

int launch_dft_big(int N, int inverse, uint32_t n_calls, const int16_t *d_in, int16_t *d_out, int scale, cudaStream_t st)
{
  BigPlan P;
  if (!make_big_plan(N, inverse, scale, &P)) return -4;
  if (n_calls == 0) return 0;
  DftCtx &c = dctx();
  unsigned *rows = nullptr;
  const size_t total = (size_t)n_calls * N;
  if (cudaMallocAsync((void **)&rows, total * 4, st) != cudaSuccess) { std::lock_guard<std::recursive_mutex> lk(c.mu); c.last_error = "dft_big scratch"; return -5; }
  const unsigned grid = (unsigned)std::min<size_t>((total + 255) / 256, 148 * 16);
  big_gather_kernel<<<grid, 256, 0, st>>>(P, (const unsigned *)d_in, rows, n_calls);
  c.launches++;
  int rc = launch_dft(P.B, inverse, n_calls * P.T, (const int16_t *)rows, d_out, P.base_scale, st);
  int M = P.B;
  for (int j = P.L - 1; j >= 0 && rc == 0; j--) {
    const int NJ = M * P.R[j];
    const size_t work = total / P.R[j];
    const unsigned g2 = (unsigned)std::min<size_t>((work + 255) / 256, 148 * 16);
    if (inverse) big_combine_kernel<true><<<g2, 256, 0, st>>>(N, NJ, P.R[j], P.sval[j], c.d_tw + P.tw[j], (const unsigned *)d_out, (unsigned *)d_out, n_calls);
    else big_combine_kernel<false><<<g2, 256, 0, st>>>(N, NJ, P.R[j], P.sval[j], c.d_tw + P.tw[j], (const unsigned *)d_out, (unsigned *)d_out, n_calls);
    c.launches++;
    M = NJ;
  }
  cudaFreeAsync(rows, st);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { std::lock_guard<std::recursive_mutex> lk(c.mu); c.last_error = std::string("dft_big launch: ") + cudaGetErrorString(e); return -2; }
  return rc;
}

bool make_small_plan(int N, int scale, SmallPlan *P)
{
  const DftCtx &c = dctx();
  if (!is_fourway(N)) return false;
  P->N = N; P->L = 0;
  int n = N, apply = scale == 1;
  while (n != 12) {
    const int i = small_index(n);
    if (i < 0 || P->L >= 5) return false;
    P->R[P->L] = kSmall[i].R; P->norm[P->L] = kSmall[i].norm; P->apply[P->L] = apply; P->tw[P->L] = c.small_tw[i];
    apply = kSmall[i].subscale; n = kSmall[i].M; P->L++;
  }
  for (int j = P->L; j < 5; j++) { P->R[j] = 1; P->norm[j] = 0; P->apply[j] = 0; P->tw[j] = 0; }
  return true;
}

// n_calls x (4 N c16 in, 4 N c16 out)
int launch_dft_small(int N, uint32_t n_calls, const int16_t *d_in, int16_t *d_out, int scale, cudaStream_t st)
{
  SmallPlan P;
  if (!make_small_plan(N, scale, &P)) return -4;
  if (n_calls == 0) return 0;
  DftCtx &c = dctx();
  dft_small_kernel<<<n_calls, 256, (size_t)4 * N * 4, st>>>(P, c.d_tw, (const unsigned *)d_in, (unsigned *)d_out, n_calls);
  c.launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { std::lock_guard<std::recursive_mutex> lk(c.mu); c.last_error = std::string("dft_small launch: ") + cudaGetErrorString(e); return -2; }
  return 0;
}

const int kDftSizes[] = {12, 24, 36, 48, 60, 64, 72, 96, 108, 120, 128, 144, 180, 192, 216, 240, 256, 288, 300, 324, 360, 384, 432, 480, 512, 540, 576, 600,
                         648, 720, 768, 864, 900, 960, 972, 1024, 1080, 1152, 1200, 1296, 1440, 1500, 1536, 1620, 1728, 1800, 1920, 1944, 2048, 2160,
                         2304, 2400, 2592, 2700, 2880, 2916, 3000, 3072, 3240, 4096, 6144, 8192, 9216, 12288, 18432, 24576, 36864, 49152, 73728, 98304};
const int kIdftSizes[] = {64, 128, 256, 512, 768, 1024, 1536, 2048, 3072, 4096, 6144, 8192, 9216, 12288, 16384, 18432, 24576, 32768, 36864, 49152, 65536,
                          73728, 98304};

void one_transform(bool inverse, uint8_t sizeidx, int16_t *in, int16_t *out, unsigned char scale)
{
  const int N = nrb200_dft_size_of_index(inverse, sizeidx);
  if (N < 0 || !nrb200_dft_supported(N) || dft_init() != 0) {
    fprintf(stderr, "[nrb200 dfts] %s size index %d (N=%d) is not available in libdfts_b200.so: %s\n", inverse ? "idft" : "dft", sizeidx, N,
            dctx().inited ? "size not implemented yet" : dctx().last_error.c_str());
    abort();   // fail loudly: this library has no CPU fallback
  }
  if (nrb200_dft_batch_host(N, inverse, 1, in, out, scale) != 0) {
    fprintf(stderr, "[nrb200 dfts] transform failed: %s\n", dctx().last_error.c_str());
    abort();
  }
}

}  // namespace

NRB200_EXPORT int32_t nrb200_dft_size_of_index(int inverse, int sizeidx)
{
  if (sizeidx < 0) return -1;
  if (inverse) return sizeidx < (int)(sizeof(kIdftSizes) / sizeof(int)) ? kIdftSizes[sizeidx] : -1;
  return sizeidx < (int)(sizeof(kDftSizes) / sizeof(int)) ? kDftSizes[sizeidx] : -1;
}

NRB200_EXPORT int32_t nrb200_dft_supported(int N)
{
  DftPlan P;
  int r3 = (N % 3 == 0) ? 3 : 1, rest = N / r3;
  (void)P;
  if (N <= 8192)
    for (int n4 = 64; n4 <= 4096; n4 *= 4) if (rest == n4 || rest == 2 * n4) return 1;
  return (is_fourway(N) || big_index(N) >= 0) ? 1 : 0;
}

NRB200_EXPORT int dfts_autoinit(void) { return dft_init(); }

NRB200_EXPORT int32_t nrb200_dft_batch_dev(int N, int inverse, uint32_t n, const int16_t *d_in, int16_t *d_out, int scale, void *stream)
{
  if (dft_init() != 0) return -1;
  cudaSetDevice(dctx().dev);
  return launch_dft(N, inverse, n, d_in, d_out, scale, (cudaStream_t)stream);
}

NRB200_EXPORT int32_t nrb200_dft_batch_host(int N, int inverse, uint32_t n, const int16_t *in, int16_t *out, int scale)
{
  if (dft_init() != 0) return -1;
  if (!nrb200_dft_supported(N)) return -4;
  if (n == 0) return 0;
  DftCtx &c = dctx();
  std::lock_guard<std::recursive_mutex> lk(c.mu);   // one staging buffer: host-buffer calls are serialised (the batched entry point is the fast path)
  cudaSetDevice(c.dev);
  const bool four = is_fourway(N);
  if (four && inverse) return -4;
  const size_t bytes = (size_t)n * N * 4 * (four ? 4 : 1);   // the four-way sizes move 4 N c16 per call
  if (bytes > c.cap) {
    if (c.d_in) { cudaFree(c.d_in); cudaFree(c.d_out); cudaFreeHost(c.h_in); cudaFreeHost(c.h_out); }
    c.cap = bytes + bytes / 4 + 65536;
    if (cudaMalloc(&c.d_in, c.cap) != cudaSuccess || cudaMalloc(&c.d_out, c.cap) != cudaSuccess || cudaHostAlloc(&c.h_in, c.cap, 0) != cudaSuccess ||
        cudaHostAlloc(&c.h_out, c.cap, 0) != cudaSuccess) { c.cap = 0; c.last_error = "staging alloc"; return -5; }
  }
  std::memcpy(c.h_in, in, bytes);
  cudaMemcpyAsync(c.d_in, c.h_in, bytes, cudaMemcpyHostToDevice, c.stream);
  if (!four && N > 8192) {
    const int rc = launch_dft_big(N, inverse, n, (const int16_t *)c.d_in, (int16_t *)c.d_out, scale, c.stream);
    if (rc != 0) return rc;
    c.launches--;   // counted inside
  } else if (four) {
    SmallPlan SP;
    if (!make_small_plan(N, scale, &SP)) return -4;
    dft_small_kernel<<<n, 256, (size_t)4 * N * 4, c.stream>>>(SP, c.d_tw, (const unsigned *)c.d_in, (unsigned *)c.d_out, n);
  } else {
    DftPlan P;
    if (!make_plan(N, inverse, scale, &P)) return -4;
    const unsigned grid = (n + P.tpb - 1) / P.tpb;
    if (inverse) dft_kernel<0, true><<<grid, 256, (size_t)2 * P.tpb * N * 4, c.stream>>>(P, c.off, c.d_tw, (const unsigned *)c.d_in, (unsigned *)c.d_out, n, SlotIO{});
    else dft_kernel<0, false><<<grid, 256, (size_t)2 * P.tpb * N * 4, c.stream>>>(P, c.off, c.d_tw, (const unsigned *)c.d_in, (unsigned *)c.d_out, n, SlotIO{});
  }
  c.launches++;
  cudaMemcpyAsync(c.h_out, c.d_out, bytes, cudaMemcpyDeviceToHost, c.stream);
  cudaError_t e = cudaStreamSynchronize(c.stream);
  if (e != cudaSuccess) { c.last_error = std::string("dft: ") + cudaGetErrorString(e); return -2; }
  std::memcpy(out, c.h_out, bytes);
  return 0;
}

// ------------------------------------------------------------------------------------------------ slot-level OFDM
namespace {
bool slot_io(const nrb200_ofdm_slot_t *d, bool rx, const unsigned *d_timeshift, SlotIO *S)
{
  if (!d || d->n_symb < 1 || d->n_symb > 14 || d->n_ant < 1) return false;
  const unsigned N = d->fft_size;
  S->n_symb = (int)d->n_symb; S->rotate = d->rotate ? 1 : 0;
  S->f_stride = d->f_stride; S->t_stride = d->t_stride; S->t_ring = rx ? d->t_ring : 0;
  for (unsigned l = 0; l < 14; l++) {
    S->t_off[l] = d->t_off[l]; S->prefix[l] = d->prefix[l]; S->rot[l][0] = d->rot[l][0]; S->rot[l][1] = d->rot[l][1];
    if (!rx && l < d->n_symb && d->prefix[l] > N) return false;
  }
  // the two ranges of apply_nr_rotation_TX/RX (odd nb_rb: one extra half PRB on both sides)
  const unsigned odd = d->nb_rb & 1u;
  S->r_len = (d->nb_rb + odd) * 6;
  S->r_start[0] = 0; S->r_start[1] = d->first_carrier_offset - (odd ? 6 : 0);
  if (S->rotate && (S->r_len > N || S->r_start[1] + S->r_len > N)) return false;
  S->timeshift = d_timeshift;
  if (rx && S->rotate && !d_timeshift) return false;
  return true;
}

int launch_slot(const nrb200_ofdm_slot_t *d, bool rx, const void *d_in, void *d_out, const unsigned *d_timeshift, cudaStream_t st)
{
  DftPlan P;
  SlotIO S;
  if (!make_plan((int)d->fft_size, rx ? 0 : 1, 1, &P)) return -4;
  if (!slot_io(d, rx, d_timeshift, &S)) return -4;
  DftCtx &c = dctx();
  const unsigned n = d->n_symb * d->n_ant, grid = (n + P.tpb - 1) / P.tpb;
  const size_t smem = (size_t)2 * P.tpb * P.N * 4;
  // bulk copies carry the frequency-domain side (always whole, contiguous symbols) when its base and antenna stride are 16-byte multiples; the
  // time-domain side is checked per symbol inside the kernel (timing offsets are arbitrary sample counts)
  if (use_tma4096((int)d->fft_size, rx ? d_out : d_in, nullptr) && (S.f_stride & 3u) == 0u) {
    const unsigned g4 = std::min<unsigned>(n, 3u * (unsigned)c.sm_count);
    if (rx) dft4096_kernel<2, false><<<g4, 256, dft4096::kSmemBytes, st>>>(P, c.off, c.d_tw, (const unsigned *)d_in, (unsigned *)d_out, n, S);
    else dft4096_kernel<1, true><<<g4, 256, dft4096::kSmemBytes, st>>>(P, c.off, c.d_tw, (const unsigned *)d_in, (unsigned *)d_out, n, S);
  } else if (rx) dft_kernel<2, false><<<grid, 256, smem, st>>>(P, c.off, c.d_tw, (const unsigned *)d_in, (unsigned *)d_out, n, S);
  else dft_kernel<1, true><<<grid, 256, smem, st>>>(P, c.off, c.d_tw, (const unsigned *)d_in, (unsigned *)d_out, n, S);
  c.launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { std::lock_guard<std::recursive_mutex> lk(c.mu); c.last_error = std::string("ofdm slot launch: ") + cudaGetErrorString(e); return -2; }
  return 0;
}

bool stage_reserve(DftCtx &c, size_t bytes)
{
  if (bytes <= c.cap) return true;
  if (c.d_in) { cudaFree(c.d_in); cudaFree(c.d_out); cudaFreeHost(c.h_in); cudaFreeHost(c.h_out); }
  c.cap = bytes + bytes / 4 + 65536;
  if (cudaMalloc(&c.d_in, c.cap) != cudaSuccess || cudaMalloc(&c.d_out, c.cap) != cudaSuccess || cudaHostAlloc(&c.h_in, c.cap, 0) != cudaSuccess ||
      cudaHostAlloc(&c.h_out, c.cap, 0) != cudaSuccess) { c.cap = 0; c.d_in = nullptr; c.last_error = "staging alloc"; return false; }
  return true;
}
}  // namespace

NRB200_EXPORT int32_t nrb200_ofdm_mod_slot_dev(const nrb200_ofdm_slot_t *d, const int16_t *d_txdataF, int16_t *d_txdata, void *stream)
{
  if (dft_init() != 0) return -1;
  cudaSetDevice(dctx().dev);
  return launch_slot(d, false, d_txdataF, d_txdata, nullptr, (cudaStream_t)stream);
}

NRB200_EXPORT int32_t nrb200_ofdm_demod_slot_dev(const nrb200_ofdm_slot_t *d, const int16_t *d_rxdata, const int16_t *d_timeshift, int16_t *d_rxdataF,
                                                 void *stream)
{
  if (dft_init() != 0) return -1;
  cudaSetDevice(dctx().dev);
  return launch_slot(d, true, d_rxdata, d_rxdataF, (const unsigned *)d_timeshift, (cudaStream_t)stream);
}

NRB200_EXPORT int32_t nrb200_ofdm_mod_slot_host(const nrb200_ofdm_slot_t *d, const int16_t *const *txdataF, int16_t *const *txdata)
{
  if (dft_init() != 0) return -1;
  if (!d || d->n_symb < 1 || d->n_symb > 14 || d->n_ant < 1 || !nrb200_dft_supported((int)d->fft_size)) return -4;
  DftCtx &c = dctx();
  std::lock_guard<std::recursive_mutex> lk(c.mu);
  cudaSetDevice(c.dev);
  const unsigned N = d->fft_size, ns = d->n_symb, na = d->n_ant;
  // time-domain span written by this call, relative to txdata[a]
  unsigned lo = d->t_off[0], hi = 0;
  for (unsigned l = 0; l < ns; l++) { lo = std::min(lo, d->t_off[l]); hi = std::max(hi, d->t_off[l] + d->prefix[l] + N); }
  const size_t fin = (size_t)ns * N, tout = hi - lo;
  if (!stage_reserve(c, 4 * std::max(fin, tout) * na)) return -5;
  for (unsigned a = 0; a < na; a++) std::memcpy((uint8_t *)c.h_in + 4 * fin * a, txdataF[a], 4 * fin);
  cudaMemcpyAsync(c.d_in, c.h_in, 4 * fin * na, cudaMemcpyHostToDevice, c.stream);
  nrb200_ofdm_slot_t e = *d;
  e.f_stride = (uint32_t)fin; e.t_stride = (uint32_t)tout;
  for (unsigned l = 0; l < ns; l++) e.t_off[l] -= lo;
  const int rc = launch_slot(&e, false, c.d_in, c.d_out, nullptr, c.stream);
  if (rc != 0) return rc;
  cudaMemcpyAsync(c.h_out, c.d_out, 4 * tout * na, cudaMemcpyDeviceToHost, c.stream);
  cudaError_t er = cudaStreamSynchronize(c.stream);
  if (er != cudaSuccess) { c.last_error = std::string("ofdm mod: ") + cudaGetErrorString(er); return -2; }
  for (unsigned a = 0; a < na; a++) {
    // only the samples this call produced are written back (symbols may be non-contiguous when n_symb < 14)
    for (unsigned l = 0; l < ns; l++)
      std::memcpy(txdata[a] + 2 * (size_t)d->t_off[l], (uint8_t *)c.h_out + 4 * (tout * a + e.t_off[l]), 4 * (size_t)(d->prefix[l] + N));
  }
  return 0;
}

NRB200_EXPORT int32_t nrb200_ofdm_demod_slot_host(const nrb200_ofdm_slot_t *d, const int16_t *const *rxdata, const int16_t *timeshift,
                                                  int16_t *const *rxdataF)
{
  if (dft_init() != 0) return -1;
  if (!d || d->n_symb < 1 || d->n_symb > 14 || d->n_ant < 1 || !nrb200_dft_supported((int)d->fft_size)) return -4;
  if (d->rotate && !timeshift) return -4;
  DftCtx &c = dctx();
  std::lock_guard<std::recursive_mutex> lk(c.mu);
  cudaSetDevice(c.dev);
  const unsigned N = d->fft_size, ns = d->n_symb, na = d->n_ant, ring = d->t_ring;
  // Only the FFT windows travel: window l of antenna a lands at (a * ns + l) * N in the staging buffer (the ring wrap is resolved here).
  const size_t win = (size_t)ns * N;
  if (!stage_reserve(c, 4 * (win * na + N))) return -5;
  for (unsigned a = 0; a < na; a++)
    for (unsigned l = 0; l < ns; l++) {
      uint8_t *dst = (uint8_t *)c.h_in + 4 * (win * a + (size_t)l * N);
      const size_t k = ring ? d->t_off[l] % ring : d->t_off[l];
      const size_t first = ring ? std::min<size_t>(N, ring - k) : N;
      std::memcpy(dst, rxdata[a] + 2 * k, 4 * first);
      if (first < N) std::memcpy(dst + 4 * first, rxdata[a], 4 * (N - first));
    }
  if (d->rotate) std::memcpy((uint8_t *)c.h_in + 4 * win * na, timeshift, 4 * (size_t)N);
  cudaMemcpyAsync(c.d_in, c.h_in, 4 * (win * na + N), cudaMemcpyHostToDevice, c.stream);
  nrb200_ofdm_slot_t e = *d;
  e.t_ring = 0; e.t_stride = (uint32_t)win; e.f_stride = (uint32_t)win;
  for (unsigned l = 0; l < ns; l++) e.t_off[l] = l * N;
  const int rc = launch_slot(&e, true, c.d_in, c.d_out, (const unsigned *)c.d_in + win * na, c.stream);
  if (rc != 0) return rc;
  cudaMemcpyAsync(c.h_out, c.d_out, 4 * win * na, cudaMemcpyDeviceToHost, c.stream);
  cudaError_t er = cudaStreamSynchronize(c.stream);
  if (er != cudaSuccess) { c.last_error = std::string("ofdm demod: ") + cudaGetErrorString(er); return -2; }
  for (unsigned a = 0; a < na; a++) std::memcpy(rxdataF[a], (uint8_t *)c.h_out + 4 * win * a, 4 * win);
  return 0;
}

NRB200_EXPORT void dft(uint8_t sizeidx, int16_t *sigF, int16_t *sig, unsigned char scale_flag) { one_transform(false, sizeidx, sigF, sig, scale_flag); }
NRB200_EXPORT void idft(uint8_t sizeidx, int16_t *sigF, int16_t *sig, unsigned char scale_flag) { one_transform(true, sizeidx, sigF, sig, scale_flag); }
#ifndef NRB200_DFTS_INTERNAL
// one process, several GPUs: selects the device the calling thread's following calls run on (see include/nrb200_dfts.h)
NRB200_EXPORT int32_t nrb200_dfts_set_device(int dev)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || dev < 0 || dev >= n || dev >= 16) return -1;
  tls_ddev = dev;
  return 0;
}
#endif
NRB200_EXPORT const char *nrb200_dfts_last_error(void) { return dctx().last_error.c_str(); }
NRB200_EXPORT uint64_t nrb200_dfts_launch_count(void) { return dctx().launches.load(); }

#ifdef NRB200_DFTS_INTERNAL
namespace nrb200 {
int dft_batch_internal(int N, int inverse, uint32_t n, const int16_t *d_in, int16_t *d_out, int scale, cudaStream_t st)
{
  if (dft_init() != 0) return -1;
  return launch_dft(N, inverse, n, d_in, d_out, scale, st);
}
}  // namespace nrb200
#endif
