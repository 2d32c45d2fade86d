/*
 * TEST INFRASTRUCTURE ONLY -- scalar C restatement of the reference's Q15 fixed-point DFT/IDFT (openair1/PHY/TOOLS/oai_dfts.c)
 * for the OFDM sizes 64, 128, 256, 512, 768, 1024, 1536, 2048, 3072, 4096, 6144, 8192, written as the recursive
 * decimation-in-time factorisation the reference uses (radix-3 on top of radix-2 on top of radix-4 on top of a 16-point kernel)
 * in natural index order, with the reference's quantisation points: which stages saturate (adds_epi16/packs_epi32), which wrap
 * (add_epi16 in bfly4_256), where products are shifted (>>15) and which scaling each level applies.
 * Parity status: PINNED bit-exactly against oracle/_ref/libref_dfts.so (tests/test_oracle_vs_reference.py, tests/golden/dft.npz).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "../openairinterface5g_b200/csrc/nr_dft_tables.h"

typedef struct { int32_t r, i; } cx;   /* values always within int16 range between stages */

static inline int32_t sat16(int64_t v) { return v > 32767 ? 32767 : v < -32768 ? -32768 : (int32_t)v; }
static inline int32_t wrap16(int32_t v) { return (int16_t)(uint16_t)v; }
static inline int32_t wrap32(int64_t v) { return (int32_t)(uint32_t)(uint64_t)v; }
static inline int32_t neg16(int32_t v) { return v == -32768 ? -32768 : -v; }   /* sign_epi16(x, -1) */
static inline cx mjn(cx a) { cx o = {a.i, neg16(a.r)}; return o; }             /* "x_flip": -j * x  (oai_dfts.c:1254-1257) */
static inline cx sadd(cx a, cx b) { cx o = {sat16((int64_t)a.r + b.r), sat16((int64_t)a.i + b.i)}; return o; }
static inline cx ssub(cx a, cx b) { cx o = {sat16((int64_t)a.r - b.r), sat16((int64_t)a.i - b.i)}; return o; }
/* packed_cmult2 (oai_dfts.c:303-317): re = a . ta, im = a . tb, >>15, packs */
static inline cx cmult2(cx a, const int16_t *ta, const int16_t *tb)
{
  cx o = {sat16(wrap32((int64_t)a.r * ta[0] + (int64_t)a.i * ta[1]) >> 15), sat16(wrap32((int64_t)a.r * tb[0] + (int64_t)a.i * tb[1]) >> 15)};
  return o;
}
static inline int32_t mulhrs(int32_t a, int32_t b) { return wrap16((a * b + 0x4000) >> 15); }   /* mulhrs_epi16 */

/* 16-bit saturating radix-4 butterfly: dft16 stages and bfly4_16_256 / ibfly4_16_256 (oai_dfts.c:803-948) */
static void bfly4_sat(cx x0, cx a1, cx a2, cx a3, int inverse, cx *y0, cx *y1, cx *y2, cx *y3)
{
  cx x02 = sadd(x0, a2), x13 = sadd(a1, a3);
  *y0 = sadd(x02, x13);
  *y2 = ssub(x02, x13);
  x02 = ssub(x0, a2);
  x13 = ssub(mjn(a1), mjn(a3));
  cx ya = sadd(x02, x13), yb = ssub(x02, x13);
  if (inverse) { *y1 = yb; *y3 = ya; } else { *y1 = ya; *y3 = yb; }
}

/* cmult / cmultc (oai_dfts.c:140-240): 32-bit complex product by w or conj(w), no shift */
static inline void cm32(cx x, int32_t wr, int32_t wi, int inverse, int64_t *re, int64_t *im)
{
  if (inverse) { *re = (int64_t)x.r * wr + (int64_t)x.i * wi; *im = (int64_t)x.i * wr - (int64_t)x.r * wi; }
  else { *re = (int64_t)x.r * wr - (int64_t)x.i * wi; *im = (int64_t)x.r * wi + (int64_t)x.i * wr; }
}
static inline cx pk32(int64_t r, int64_t i) { cx o = {sat16(wrap32(r) >> 15), sat16(wrap32(i) >> 15)}; return o; }   /* cpack */

/* bfly4_256 / ibfly4_256 (oai_dfts.c:633-721): 32-bit sums, >>15, packs, then NON-saturating add of x0 */
static void bfly4_32(cx x0, cx x1, cx x2, cx x3, const int32_t *w /* wr1 wi1 wr2 wi2 wr3 wi3 */, int inverse, cx *y0, cx *y1, cx *y2, cx *y3)
{
  int64_t x1r, x1i, x2r, x2i, x3r, x3i;
  cm32(x1, w[0], w[1], inverse, &x1r, &x1i);
  cm32(x2, w[2], w[3], inverse, &x2r, &x2i);
  cm32(x3, w[4], w[5], inverse, &x3r, &x3i);
  cx d0 = pk32(x1r + x2r + x3r, x1i + x2i + x3i);
  cx da = pk32(x1i - (x2r + x3i), (x3r - x2i) - x1r);
  cx d2 = pk32((x2r - x3r) - x1r, (x2i - x3i) - x1i);
  cx db = pk32((x3i - x2r) - x1i, x1r - (x2i + x3r));
  cx o0 = {wrap16(x0.r + d0.r), wrap16(x0.i + d0.i)}, oa = {wrap16(x0.r + da.r), wrap16(x0.i + da.i)};
  cx o2 = {wrap16(x0.r + d2.r), wrap16(x0.i + d2.i)}, ob = {wrap16(x0.r + db.r), wrap16(x0.i + db.i)};
  *y0 = o0; *y2 = o2;
  if (inverse) { *y1 = ob; *y3 = oa; } else { *y1 = oa; *y3 = ob; }
}

static int32_t rnd(double v) { return (int32_t)(int16_t)round(v); }   /* (int16_t)round(32767.0*cos(..)) as in init_rad4/2/3 */

/* 16-point kernel, natural order in and out (dft16_simd256 oai_dfts.c:1190-1297, idft16_simd256 :1346-1458) */
static void dft16(const cx *x, int stride, cx *y, int inverse)
{
  const int16_t *ta = inverse ? NRB200_TW16 : NRB200_TW16A, *tb = inverse ? NRB200_TW16C : NRB200_TW16B;
  cx A[4][4];   /* [k1][c] */
  for (int c = 0; c < 4; c++)
    bfly4_sat(x[(0 * 4 + c) * stride], x[(1 * 4 + c) * stride], x[(2 * 4 + c) * stride], x[(3 * 4 + c) * stride], inverse, &A[0][c], &A[1][c], &A[2][c], &A[3][c]);
  for (int k1 = 0; k1 < 4; k1++) {
    cx b[4];
    b[0] = A[k1][0];
    for (int c = 1; c < 4; c++) b[c] = cmult2(A[k1][c], ta + 8 * (c - 1) + 2 * k1, tb + 8 * (c - 1) + 2 * k1);   /* W16^(c*k1), applied for k1 = 0 too */
    bfly4_sat(b[0], b[1], b[2], b[3], inverse, &y[k1], &y[k1 + 4], &y[k1 + 8], &y[k1 + 12]);
  }
}

static void dft_pow4(const cx *x, int stride, cx *y, int N, int inverse, int scale);

static void dft64(const cx *x, int stride, cx *y, int inverse, int scale)
{
  const int16_t *ta = inverse ? NRB200_TW64 : NRB200_TW64A, *tb = inverse ? NRB200_TW64C : NRB200_TW64B;
  cx Y[4][16];
  for (int q = 0; q < 4; q++) dft16(x + q * stride, 4 * stride, Y[q], inverse);
  for (int k = 0; k < 16; k++) {
    cx a1 = cmult2(Y[1][k], ta + 2 * k, tb + 2 * k), a2 = cmult2(Y[2][k], ta + 32 + 2 * k, tb + 32 + 2 * k), a3 = cmult2(Y[3][k], ta + 64 + 2 * k, tb + 64 + 2 * k);
    bfly4_sat(Y[0][k], a1, a2, a3, inverse, &y[k], &y[k + 16], &y[k + 32], &y[k + 48]);
  }
  if (scale) for (int k = 0; k < 64; k++) { y[k].r >>= 3; y[k].i >>= 3; }   /* oai_dfts.c:1654-1663 */
}

static void dft_pow4(const cx *x, int stride, cx *y, int N, int inverse, int scale)
{
  if (N == 64) { dft64(x, stride, y, inverse, scale); return; }
  const int M = N / 4;
  cx *Y = malloc(sizeof(cx) * (size_t)N);
  for (int q = 0; q < 4; q++) dft_pow4(x + q * stride, 4 * stride, Y + q * M, M, inverse, 1);
  for (int k = 0; k < M; k++) {
    if (N == 256 && !inverse) {   /* dft256 uses the 16-bit saturating butterfly with tw256a/b (oai_dfts.c:1915-1976) */
      cx a1 = cmult2(Y[M + k], NRB200_TW256A + 2 * k, NRB200_TW256B + 2 * k);
      cx a2 = cmult2(Y[2 * M + k], NRB200_TW256A + 128 + 2 * k, NRB200_TW256B + 128 + 2 * k);
      cx a3 = cmult2(Y[3 * M + k], NRB200_TW256A + 256 + 2 * k, NRB200_TW256B + 256 + 2 * k);
      bfly4_sat(Y[k], a1, a2, a3, 0, &y[k], &y[k + M], &y[k + 2 * M], &y[k + 3 * M]);
    } else {                      /* idft256 and every size >= 1024: 32-bit butterfly (oai_dfts.c:2006-2052, 2260-2310, 2553-2663) */
      int32_t w[6];
      for (int p = 1; p <= 3; p++) {
        if (N == 256) { w[2 * p - 2] = NRB200_TW256[128 * (p - 1) + 2 * k]; w[2 * p - 1] = NRB200_TW256[128 * (p - 1) + 2 * k + 1]; }
        else { w[2 * p - 2] = rnd(32767.0 * cos(2 * M_PI * p * k / N)); w[2 * p - 1] = -rnd(32767.0 * sin(2 * M_PI * p * k / N)); }   /* init_rad4 :7709-7722 */
      }
      bfly4_32(Y[k], Y[M + k], Y[2 * M + k], Y[3 * M + k], w, inverse, &y[k], &y[k + M], &y[k + 2 * M], &y[k + 3 * M]);
    }
  }
  if (scale > 0) { const int sh = N == 65536 ? scale : 1;   /* idft65536 shifts by `scale` itself (oai_dfts.c:4206-4226), every other level by 1 */
    for (int k = 0; k < N; k++) { y[k].r >>= sh; y[k].i >>= sh; } }
  free(Y);
}

/* radix-2 on top: 128, 512, 2048, 8192, 32768 (oai_dfts.c:1775-1900, 2093-2255, 2372-2550, 2665-2842, 2958-3138) */
static void dft_pow2(const cx *x, int stride, cx *y, int N, int inverse, int scale)
{
  if (N == 64 || N == 256 || N == 1024 || N == 4096 || N == 16384 || N == 65536) { dft_pow4(x, stride, y, N, inverse, scale); return; }
  const int M = N / 2;
  cx *Y = malloc(sizeof(cx) * (size_t)N);
  dft_pow4(x, 2 * stride, Y, M, inverse, 1);
  dft_pow4(x + stride, 2 * stride, Y + M, M, inverse, 1);
  for (int k = 0; k < M; k++) {
    if (N == 128 && !inverse) {   /* bfly2_16_256 with tw128a/b: 16-bit product, saturating add/sub */
      cx t = cmult2(Y[M + k], NRB200_TW128A + 2 * k, NRB200_TW128B + 2 * k);
      y[k] = sadd(Y[k], t); y[k + M] = ssub(Y[k], t);
    } else {                      /* bfly2_256 / ibfly2_256: x0*(32767+0j) +- x1*w in 32 bit, >>15, packs */
      int32_t wr, wi;
      if (N == 128) { wr = NRB200_TW128[2 * k]; wi = NRB200_TW128[2 * k + 1]; }
      else if (N == 512) { wr = NRB200_TW512[2 * k]; wi = NRB200_TW512[2 * k + 1]; }
      else { wr = rnd(32767.0 * cos(2 * M_PI * k / N)); wi = -rnd(32767.0 * sin(2 * M_PI * k / N)); }   /* init_rad2 :7747-7756 */
      int64_t br, bi, ar = (int64_t)Y[k].r * 32767, ai = (int64_t)Y[k].i * 32767;
      cm32(Y[M + k], wr, wi, inverse, &br, &bi);
      y[k] = pk32(ar + br, ai + bi); y[k + M] = pk32(ar - br, ai - bi);
    }
  }
  if (scale) for (int k = 0; k < N; k++) { y[k].r = mulhrs(y[k].r, 23170); y[k].i = mulhrs(y[k].i, 23170); }   /* ONE_OVER_SQRT2_Q15 */
  free(Y);
}

/* radix-3 on top: 768, 1536, 3072, 6144 and the large sizes 12288, 18432, 24576, 36864, 49152, 98304 (bfly3/ibfly3 oai_dfts.c:477-540, drivers :3140-3600,
 * :3614-4350).  The M-point transforms are called with scale 1, except by 12288 and 18432, which hand their own scale argument down (:3641-3643, :3755-3757). */
static void dft_3x(const cx *x, int stride, cx *y, int N, int inverse, int scale)
{
  const int M = N / 3, sub = (N == 12288 || N == 18432) ? scale : 1;
  cx *Y = malloc(sizeof(cx) * (size_t)N);
  for (int q = 0; q < 3; q++) {
    if (M % 3 == 0) dft_3x(x + q * stride, 3 * stride, Y + q * M, M, inverse, sub);
    else dft_pow2(x + q * stride, 3 * stride, Y + q * M, M, inverse, sub);
  }
  for (int k = 0; k < M; k++) {
    int64_t r, i, r2, i2;
    cm32(Y[M + k], rnd(32767.0 * cos(2 * M_PI * k / N)), -rnd(32767.0 * sin(2 * M_PI * k / N)), inverse, &r, &i);            /* init_rad3 :7772-7782 */
    cx x1 = pk32(r, i);
    cm32(Y[2 * M + k], rnd(32767.0 * cos(2 * M_PI * 2 * k / N)), -rnd(32767.0 * sin(2 * M_PI * 2 * k / N)), inverse, &r, &i);
    cx x2 = pk32(r, i);
    y[k] = sadd(Y[k], sadd(x1, x2));
    cm32(x1, -16384, -28378, inverse, &r, &i); cm32(x2, -16384, 28378, inverse, &r2, &i2);     /* W13, W23 (:321-322) */
    y[k + M] = sadd(Y[k], pk32(r + r2, i + i2));
    cm32(x1, -16384, 28378, inverse, &r, &i); cm32(x2, -16384, -28378, inverse, &r2, &i2);
    y[k + 2 * M] = sadd(Y[k], pk32(r + r2, i + i2));
  }
  if (scale == 1) for (int k = 0; k < N; k++) { y[k].r = mulhrs(y[k].r, 18919); y[k].i = mulhrs(y[k].i, 18919); }   /* ONE_OVER_SQRT3_Q15, `scale==1` */
  free(Y);
}

/* ------------------------------------------------------------------------------------------------------------------------------------------------
 * The DFT-s-OFDM family 12 ... 3240 (PUSCH transform precoding; oai_dfts.c:4352-7706).  These entry points work on 128-bit vectors holding FOUR independent
 * transforms ("4-way parallel DFTS (i.e. 4 DFTS with interleaved input/output)", :4352): element n of transform l is c16 number 4 n + l, in and out.
 * Every size is N = R x M, decimation in time: M-point transforms of x[m], x[m + R], ... (m < R) in natural order, then for every k < M one radix-R butterfly
 *   k = 0: bfly{2,3,4,5}_tw1 (no twiddles: saturating 16-bit sums for R = 2, 4; 32-bit W products for R = 3, 5)
 *   k > 0: bfly{2,3,4,5} with twiddles (int16)round(32767 cos/sin(2 pi p k / N)), p = 1 .. R-1 (init_rad{2,3,4,5}_rep :7725-7828)
 * writing y[k + q M], q < R, followed by mulhrs with a per-size constant when scale_flag == 1.  The 12-point kernel (dft12f :4365-4470) is three bfly4_tw1 and
 * four bfly3 with hand-entered constants and never scales.  Sub-transforms are called with scale 1 except 96 -> 48, 108 -> 36 and 120 -> 60 (scale 0).
 * dft2304 (:7288) calls the single-transform dft768 on the four-way data and combines uninitialised stack: not reproducible; orc_dft4 returns the transform
 * the function's comment ("768 x 3") describes, built on the four-way 768 (dft768p :6330). */
typedef struct { int R, M, subscale, norm; } small_t;
static int small_plan(int N, small_t *p)
{
  static const int tab[][5] = {
    {24, 2, 12, 0, 6689}, {36, 3, 12, 0, 5461}, {48, 4, 12, 0, 4729}, {60, 5, 12, 0, 4230}, {72, 2, 36, 1, 23170}, {96, 2, 48, 0, 3344}, {108, 3, 36, 0, 3153},
    {120, 2, 60, 0, 2991}, {144, 3, 48, 1, 18918}, {180, 3, 60, 1, 18918}, {192, 4, 48, 1, 16384}, {216, 3, 72, 1, 18918}, {240, 4, 60, 1, 16384},
    {288, 3, 96, 1, 18918}, {300, 5, 60, 1, 14654}, {324, 3, 108, 1, 18918}, {360, 3, 120, 1, 18918}, {384, 4, 96, 1, 16384}, {432, 4, 108, 1, 16384},
    {480, 4, 120, 1, 16384}, {540, 3, 180, 1, 18918}, {576, 3, 192, 1, 18918}, {600, 2, 300, 1, 23170}, {648, 3, 216, 1, 18918}, {720, 4, 180, 1, 16384},
    {768, 4, 192, 1, 16384}, {864, 3, 288, 1, 18918}, {900, 3, 300, 1, 18918}, {960, 4, 240, 1, 16384}, {972, 3, 324, 1, 18918}, {1080, 3, 360, 1, 18918},
    {1152, 4, 288, 1, 16384}, {1200, 4, 300, 1, 16384}, {1296, 3, 432, 1, 18918}, {1440, 3, 480, 1, 18918}, {1500, 5, 300, 1, 14654}, {1620, 3, 540, 1, 18918},
    {1728, 3, 576, 1, 18918}, {1800, 3, 600, 1, 18918}, {1920, 4, 480, 1, 16384}, {1944, 3, 648, 1, 18918}, {2160, 3, 720, 1, 18918}, {2304, 3, 768, 1, 18918},
    {2400, 4, 600, 1, 16384}, {2592, 3, 864, 1, 18918}, {2700, 3, 900, 1, 18918}, {2880, 3, 960, 1, 18918}, {2916, 3, 972, 1, 18918}, {3000, 5, 600, 1, 14654},
    {3240, 3, 1080, 1, 18918}};
  for (unsigned i = 0; i < sizeof(tab) / sizeof(tab[0]); i++)
    if (tab[i][0] == N) { p->R = tab[i][1]; p->M = tab[i][2]; p->subscale = tab[i][3]; p->norm = tab[i][4]; return 1; }
  return 0;
}

/* packed_cmult: cmult then cpack (oai_dfts.c:246-256) */
static inline cx pcm(cx x, int32_t wr, int32_t wi) { int64_t r, i; cm32(x, wr, wi, 0, &r, &i); return pk32(r, i); }
/* sum_p x_p * W_p in 32 bit, cpack, saturating add of x0: the y1.. outputs of bfly3 / bfly5 (:477-500, :949-1000) */
static inline cx wsum(cx x0, const cx *x, const int16_t (*W)[2], const int *sel, int n)
{
  int64_t r = 0, i = 0;
  for (int p = 0; p < n; p++) { int64_t a, b; cm32(x[p], W[sel[p]][0], W[sel[p]][1], 0, &a, &b); r += a; i += b; }
  return sadd(x0, pk32(r, i));
}
static const int16_t W3C[2][2] = {{-16384, -28378}, {-16384, 28378}};                                     /* W13, W23 (:321-322) */
static const int16_t W5C[4][2] = {{10126, -31163}, {-26509, -19260}, {-26510, 19260}, {10126, 31163}};    /* W15 .. W45 (:324-327) */

static void bfly3_fwd(cx x0, cx x1, cx x2, cx *y0, cx *y1, cx *y2)   /* x1, x2 already multiplied by their twiddles (or taken as they are for k = 0) */
{
  const cx x[2] = {x1, x2};
  static const int s1[2] = {0, 1}, s2[2] = {1, 0};
  *y0 = sadd(x0, sadd(x1, x2));
  *y1 = wsum(x0, x, W3C, s1, 2);
  *y2 = wsum(x0, x, W3C, s2, 2);
}
static void bfly5_fwd(cx x0, const cx *x /* 4 */, cx *y /* 5 */)
{
  static const int s[4][4] = {{0, 1, 2, 3}, {1, 3, 0, 2}, {2, 0, 3, 1}, {3, 2, 1, 0}};
  y[0] = sadd(x0, sadd(x[0], sadd(x[1], sadd(x[2], x[3]))));
  for (int q = 0; q < 4; q++) y[q + 1] = wsum(x0, x, W5C, s[q], 4);
}

static void dft12_lane(const cx *x, int stride, cx *y)
{
  static const int16_t W12[5][2] = {{28377, -16383}, {16383, -28377}, {0, -32767}, {-16383, -28377}, {-32767, 0}};   /* W1, W2, W3, W4, W6 (:4353-4357) */
  cx t[12];
  for (int c = 0; c < 3; c++)
    bfly4_sat(x[c * stride], x[(c + 3) * stride], x[(c + 6) * stride], x[(c + 9) * stride], 0, &t[c], &t[c + 3], &t[c + 6], &t[c + 9]);
  bfly3_fwd(t[0], t[1], t[2], &y[0], &y[4], &y[8]);
  bfly3_fwd(t[3], pcm(t[4], W12[0][0], W12[0][1]), pcm(t[5], W12[1][0], W12[1][1]), &y[1], &y[5], &y[9]);
  bfly3_fwd(t[6], pcm(t[7], W12[1][0], W12[1][1]), pcm(t[8], W12[3][0], W12[3][1]), &y[2], &y[6], &y[10]);
  bfly3_fwd(t[9], pcm(t[10], W12[2][0], W12[2][1]), pcm(t[11], W12[4][0], W12[4][1]), &y[3], &y[7], &y[11]);
}

static void dft_small(const cx *x, int stride, cx *y, int N, int scale)
{
  if (N == 12) { dft12_lane(x, stride, y); return; }
  small_t P = {0, 0, 0, 0};
  small_plan(N, &P);
  const int R = P.R, M = P.M;
  cx *Y = malloc(sizeof(cx) * (size_t)N);
  for (int m = 0; m < R; m++) dft_small(x + m * stride, R * stride, Y + m * M, M, P.subscale);
  for (int k = 0; k < M; k++) {
    cx in[5], out[5];
    int32_t w[8];
    for (int p = 0; p < R; p++) in[p] = Y[p * M + k];
    for (int p = 1; p < R; p++) { w[2 * p - 2] = rnd(32767.0 * cos(2 * M_PI * p * k / N)); w[2 * p - 1] = -rnd(32767.0 * sin(2 * M_PI * p * k / N)); }
    if (R == 2) {
      if (k == 0) { out[0] = sadd(in[0], in[1]); out[1] = ssub(in[0], in[1]); }                /* bfly2_tw1 (:438-443) */
      else {                                                                                     /* bfly2 (:362-390) */
        int64_t br, bi, ar = (int64_t)in[0].r * 32767, ai = (int64_t)in[0].i * 32767;
        cm32(in[1], w[0], w[1], 0, &br, &bi);
        out[0] = pk32(ar + br, ai + bi); out[1] = pk32(ar - br, ai - bi);
      }
    } else if (R == 3) {
      if (k == 0) bfly3_fwd(in[0], in[1], in[2], &out[0], &out[1], &out[2]);
      else bfly3_fwd(in[0], pcm(in[1], w[0], w[1]), pcm(in[2], w[2], w[3]), &out[0], &out[1], &out[2]);
    } else if (R == 4) {
      if (k == 0) bfly4_sat(in[0], in[1], in[2], in[3], 0, &out[0], &out[1], &out[2], &out[3]);   /* bfly4_tw1 (:709-745) */
      else bfly4_32(in[0], in[1], in[2], in[3], w, 0, &out[0], &out[1], &out[2], &out[3]);        /* bfly4 (:584-632) */
    } else {
      cx t[4];
      for (int p = 1; p < 5; p++) t[p - 1] = k == 0 ? in[p] : pcm(in[p], w[2 * p - 2], w[2 * p - 1]);
      bfly5_fwd(in[0], t, out);
    }
    for (int q = 0; q < R; q++) y[k + q * M] = out[q];
  }
  if (scale == 1) for (int k = 0; k < N; k++) { y[k].r = mulhrs(y[k].r, P.norm); y[k].i = mulhrs(y[k].i, P.norm); }
  free(Y);
}

/* dft(DFT_<N>, in, out, scale_flag) for the four-way sizes: in/out hold 4 N c16, transform l at c16 positions 4 n + l.  768 here is dft768p. */
int orc_dft4(int N, const int16_t *in, int16_t *out, int scale)
{
  small_t P = {0, 0, 0, 0};
  if (N != 12 && !small_plan(N, &P)) return -1;
  cx *x = malloc(sizeof(cx) * (size_t)N), *y = malloc(sizeof(cx) * (size_t)N);
  for (int l = 0; l < 4; l++) {
    for (int n = 0; n < N; n++) { x[n].r = in[2 * (4 * n + l)]; x[n].i = in[2 * (4 * n + l) + 1]; }
    dft_small(x, 1, y, N, scale);
    for (int n = 0; n < N; n++) { out[2 * (4 * n + l)] = (int16_t)y[n].r; out[2 * (4 * n + l) + 1] = (int16_t)y[n].i; }
  }
  free(x); free(y);
  return 0;
}

/* in/out: interleaved {re, im} int16 like the reference's dft()/idft() (tools_defs.h:514-521); returns 0, -1 for an unsupported size */
int orc_dft(int N, int inverse, const int16_t *in, int16_t *out, int scale)
{
  /* 9216 and 73728 are AssertFatal("Need to do this") in the reference (:3601-3610, :4240-4250); 65536 exists only as idft65536 */
  if (!(N == 64 || N == 128 || N == 256 || N == 512 || N == 768 || N == 1024 || N == 1536 || N == 2048 || N == 3072 || N == 4096 || N == 6144 || N == 8192 ||
        N == 12288 || N == 16384 || N == 18432 || N == 24576 || N == 32768 || N == 36864 || N == 49152 || N == 65536 || N == 98304)) return -1;
  cx *x = malloc(sizeof(cx) * (size_t)N), *y = malloc(sizeof(cx) * (size_t)N);
  for (int n = 0; n < N; n++) { x[n].r = in[2 * n]; x[n].i = in[2 * n + 1]; }
  if (N % 3 == 0) dft_3x(x, 1, y, N, inverse, scale);
  else dft_pow2(x, 1, y, N, inverse, scale);
  for (int n = 0; n < N; n++) { out[2 * n] = (int16_t)y[n].r; out[2 * n + 1] = (int16_t)y[n].i; }
  free(x); free(y);
  return 0;
}
