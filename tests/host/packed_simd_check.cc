// CPU check of the packed decoder's byte-SIMD identities (openairinterface5g_b200/csrc/ldpc_packed_simd.cuh) against the scalar
// definition of the reference's check-node update (nrLDPC_cnProc.h:388-877 on inputs formed as in nrLDPC_bnProc.h:325).
// Test infrastructure: compiled with g++ by tests/test_packed_simd_host.py, never part of the product.
#define NRB200_HOST_EMUL
#include "ldpc_packed_simd.cuh"
#include <cstdio>
#include <cstdlib>
#include <random>

using namespace nrb200;

static int fails = 0;
#define CHECK(c, ...) do { if (!(c)) { if (fails < 10) { printf("FAIL %s:%d: ", __FILE__, __LINE__); printf(__VA_ARGS__); printf("\n"); } fails++; } } while (0)

static int clampi(int x, int lo, int hi) { return x < lo ? lo : (x > hi ? hi : x); }
static int qref(int ap, int rp) { return clampi((ap - 128) - (rp - 128), -128, 127); }   // subs_epi8(A, R)

int main()
{
  std::mt19937 rng(12345);
  const uint32_t one = 1u, mone = 0xFFFFFFFFu;
  // 1) cn_input, exhaustive per byte position with random neighbours
  for (int pos = 0; pos < 4; pos++)
    for (int ap = 0; ap < 256; ap++)
      for (int rp = 0; rp < 256; rp++) {
        const uint32_t na = rng(), nr = rng();
        const uint32_t aw = (na & ~(0xFFu << (8 * pos))) | ((uint32_t)ap << (8 * pos));
        const uint32_t ro = (nr & ~(0xFFu << (8 * pos))) | ((uint32_t)rp << (8 * pos));
        uint32_t mag, qsm;
        cn_input(aw, ro, mone, mag, qsm);
        const int q = qref(ap, rp), m = std::abs(q) > 127 ? 127 : std::abs(q);
        const int gm = (mag >> (8 * pos)) & 0xFF, gq = (qsm >> (8 * pos)) & 0xFF;
        CHECK(gm == m, "mag pos %d A' %d R' %d: %d != %d", pos, ap, rp, gm, m);
        CHECK((gq & 0x7F) == m, "qsm magnitude");
        if (q != 0) CHECK((gq >> 7) == (q < 0), "sign pos %d A' %d R' %d", pos, ap, rp);
      }
  // 2) whole rows, every degree the base graphs have, value ranges that provoke zeros, ties and the clip
  const int degs[] = {2, 3, 4, 5, 6, 7, 8, 9, 10, 19};
  const int spans[] = {1, 2, 6, 40, 128};
  for (int D : degs)
    for (int span : spans)
      for (int trial = 0; trial < 20000; trial++) {
        uint32_t aw[19], ro[19], q[19], rn[19];
        int Q[19][4];
        for (int j = 0; j < D; j++) {
          aw[j] = ro[j] = 0;
          for (int b = 0; b < 4; b++) {
            int ap, rp;
            if (span == 128) { ap = rng() & 255; rp = 1 + rng() % 255; }   // R in -127..127
            else { ap = 128 + (int)(rng() % (2 * span + 1)) - span; rp = 128 + (int)(rng() % (2 * span + 1)) - span; }
            if ((rng() & 63) == 0) ap = (rng() & 1) ? 0 : 255;
            aw[j] |= (uint32_t)ap << (8 * b); ro[j] |= (uint32_t)rp << (8 * b);
            Q[j][b] = qref(ap, rp);
          }
        }
        // the kernel's sequence (cn_row): inputs in pairs, three-input XORs for the sign product
        uint32_t sgn = 0u, qprev = 0u;
        TwoMin tm = twomin_init();
        for (int j = 0; j < D; j++) {
          uint32_t mag;
          cn_input(aw[j], ro[j], mone, mag, q[j]);
          if (j & 1) sgn = lop3<kLutXor3>(sgn, qprev, q[j]);
          qprev = q[j];
          twomin(mag, tm, one, mone);
        }
        if (D & 1) sgn ^= qprev;
        const uint32_t p1 = twomin_min1(tm, mone) | kH, p2 = twomin_min2(tm, mone) | kH;
        for (int j = 0; j < D; j++) rn[j] = make_r(q[j], tm.n1, p1, p2, sgn, one, mone);
        if (D >= 4) {   // the cluster decoder's split rows: two halves tracked separately, then merged either way round (twomin_merge)
          const int h = (D + 1) / 2;
          TwoMin ta = twomin_init(), tb = twomin_init();
          for (int j = 0; j < D; j++) { uint32_t mag, qq; cn_input(aw[j], ro[j], mone, mag, qq); twomin(mag, j < h ? ta : tb, one, mone); }
          TwoMin ma = ta, mb = tb;
          twomin_merge(ma, tb, one, mone);
          twomin_merge(mb, ta, one, mone);
          CHECK(twomin_min1(ma, mone) == twomin_min1(tm, mone) && twomin_min2(ma, mone) == twomin_min2(tm, mone), "merge a<-b D %d", D);
          CHECK(twomin_min1(mb, mone) == twomin_min1(tm, mone) && twomin_min2(mb, mone) == twomin_min2(tm, mone), "merge b<-a D %d", D);
          CHECK(ma.n1 == tm.n1 && mb.n1 == tm.n1, "merged n1 D %d", D);
        }
        {   // the tracked minima themselves
          for (int b = 0; b < 4; b++) {
            int m1 = 127, m2 = 127;
            for (int k = 0; k < D; k++) { int a = std::abs(Q[k][b]); if (a > 127) a = 127; if (a < m1) { m2 = m1; m1 = a; } else if (a < m2) m2 = a; }
            CHECK((int)((twomin_min1(tm, mone) >> (8 * b)) & 0xFF) == m1 && (int)((twomin_min2(tm, mone) >> (8 * b)) & 0xFF) == m2, "two minima D %d", D);
          }
        }
        for (int j = 0; j < D; j++)
          for (int b = 0; b < 4; b++) {
            int mn = 127, sg = 1;
            for (int k = 0; k < D; k++) {
              if (k == j) continue;
              const int a = std::abs(Q[k][b]);
              if (a < mn) mn = a;
              sg *= Q[k][b] < 0 ? -1 : (Q[k][b] == 0 ? 0 : 1);
            }
            const int want = 128 + sg * mn, got = (rn[j] >> (8 * b)) & 0xFF;
            CHECK(got == want, "row D %d span %d edge %d byte %d: %d != %d", D, span, j, b, got, want);
          }
      }
  // 3) per-byte negate identity behind make_r, exhaustive over one byte with random neighbours in 128..255
  for (int pos = 0; pos < 4; pos++)
    for (int v = 128; v < 256; v++) {
      uint32_t x = (rng() | kH);
      x = (x & ~(0xFFu << (8 * pos))) | ((uint32_t)v << (8 * pos));
      const uint32_t n = add_fma(x, kNegC, mone);
      for (int b = 0; b < 4; b++) CHECK(((n >> (8 * b)) & 0xFF) == 256 - ((x >> (8 * b)) & 0xFF), "negate");
    }
  if (fails) { printf("%d failures\n", fails); return 1; }
  printf("packed_simd_check OK\n");
  return 0;
}
