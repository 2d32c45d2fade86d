// Byte-SIMD-in-word arithmetic of the packed min-sum decoder (four lifts per 32-bit register, offset binary = value + 128).
//
// Device build: inline PTX pins the instruction selection (one LOP3 per 3-input boolean, one PRMT per byte broadcast; left to
// itself nvcc re-associates the masks of the 7-bit tricks into ~50 % more LOP3s, and the ALU pipe is the decoder's limiter).
// Host build (NRB200_HOST_EMUL, used by tests/host/packed_simd_check.cc only): the same functions over plain C emulations of the
// four PTX instructions, so every identity below is checked exhaustively per byte on the CPU before it ever reaches a GPU.
#pragma once
#include <cstdint>

#ifdef NRB200_HOST_EMUL
#define NRB200_SIMD inline
namespace nrb200 {
inline uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel)
{
  const uint64_t ab = ((uint64_t)b << 32) | a;
  uint32_t r = 0;
  for (int i = 0; i < 4; i++) {
    const uint32_t s = (sel >> (4 * i)) & 0xFu;
    uint32_t byte = (uint32_t)(ab >> (8 * (s & 7))) & 0xFFu;
    if (s & 8) byte = (byte & 0x80u) ? 0xFFu : 0u;   // replicate the sign bit
    r |= byte << (8 * i);
  }
  return r;
}
template <int LUT>
inline uint32_t lop3(uint32_t a, uint32_t b, uint32_t c)
{
  uint32_t r = 0;
  for (int i = 0; i < 32; i++) {
    const int idx = (((a >> i) & 1) << 2) | (((b >> i) & 1) << 1) | ((c >> i) & 1);
    r |= (uint32_t)((LUT >> idx) & 1) << i;
  }
  return r;
}
inline uint32_t add_fma(uint32_t a, uint32_t b, uint32_t one) { return a * one + b; }
inline uint32_t vabsdiffu4(uint32_t a, uint32_t b)
{
  uint32_t r = 0;
  for (int i = 0; i < 4; i++) {
    const int x = (a >> (8 * i)) & 0xFF, y = (b >> (8 * i)) & 0xFF;
    r |= (uint32_t)(x > y ? x - y : y - x) << (8 * i);
  }
  return r;
}
#else
#define NRB200_SIMD __device__ __forceinline__
namespace nrb200 {
NRB200_SIMD uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel)
{
  uint32_t r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
  return r;
}
template <int LUT>
NRB200_SIMD uint32_t lop3(uint32_t a, uint32_t b, uint32_t c)
{
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(r) : "r"(a), "r"(b), "r"(c), "n"(LUT));
  return r;
}
// a * one + b emitted as IMAD (one is a run-time 1 or 0xFFFFFFFF): same result as an add / subtract, but on the FMA pipe
// instead of the saturated ALU pipe
NRB200_SIMD uint32_t add_fma(uint32_t a, uint32_t b, uint32_t one)
{
  uint32_t r;
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(one), "r"(b));
  return r;
}
NRB200_SIMD uint32_t vabsdiffu4(uint32_t a, uint32_t b) { return __vabsdiffu4(a, b); }   // VABSDIFF4.U8, the one byte-SIMD ALU op in hardware
#endif

constexpr uint32_t kH = 0x80808080u, kL7 = 0x7f7f7f7fu;
constexpr uint32_t kNegC = 0x01010100u;   // kNegC - X == per-byte (256 - X_b) when every byte of X is in 128..255 (no borrow survives)
// LUT bytes: inputs a = 0xF0, b = 0xCC, c = 0xAA
constexpr int kLutSel = 0xCA;       // a ? b : c
constexpr int kLutOrAnd = 0xA8;     // (a | b) & c
constexpr int kLutOrBandC = 0xF8;   // a | (b & c)
constexpr int kLutXorAnd = 0x28;    // (a ^ b) & c
constexpr int kLutXor3 = 0x96;      // a ^ b ^ c
constexpr int kLutMajNot = 0x17;    // ~majority(a, b, c)
constexpr int kLutBorrow = 0x8E;    // (~a & b) | (~(a ^ b) & c): borrow out of each bit of a - b given the difference bits c

// 0xFF in every byte whose bit 7 is set
NRB200_SIMD uint32_t msb_mask(uint32_t x) { return prmt(x, 0u, 0xba98u); }
NRB200_SIMD uint32_t sel4(uint32_t m, uint32_t a, uint32_t b) { return lop3<kLutSel>(m, a, b); }

// One check-node input: the bn->cn message Q = subs_epi8(A, R_old) (nrLDPC_bnProc.h:325) of four lifts, from A' = A + 128 and
// R' = R + 128.  mag = min(|Q|, 127) (saturating the difference first never changes the clipped magnitude), qsm = mag | sign << 7.
// Sign: bit 7 of the borrow vector of the WORD subtraction A' - R'.  A borrow arriving from the byte below can only flip the
// outcome of a byte with A' == R', i.e. Q = 0, and the sign of a zero input never reaches an output: the other edges of the row
// see magnitude 0 (-0 is written as 0), and the edge itself is excluded from its own product.
NRB200_SIMD void cn_input(uint32_t aw, uint32_t ro, uint32_t mone, uint32_t &mag, uint32_t &qsm)
{
  const uint32_t dd = vabsdiffu4(aw, ro);                           // |A - R|
  mag = lop3<kLutOrAnd>(dd, msb_mask(dd), kL7);                     // min(|A - R|, 127)
  const uint32_t diff = add_fma(ro, aw, mone);                      // A' - R' (mod 2^32)
  qsm = lop3<kLutOrBandC>(mag, lop3<kLutBorrow>(aw, ro, diff), kH);
}

// Exclude-self two-minimum tracking on 7-bit magnitudes, state kept NEGATED so that every compare is one multiply-add on the FMA pipe
// (the ALU pipe is the decoder's limiter, the FMA pipe idles):
//   n1  = 0x7f - min1            n2p = 0x7f - min2 + 0x7f          (bytes 0..127 / 127..254: no byte ever carries or borrows)
//   mag > min1  <=>  bit 7 of mag + n1;       max(mag, min1) > min2  <=>  bit 7 of n2p - (0x7f - max(mag, min1))
// Strict compares are enough: on a tie both orders give the same minimum and the same runner-up.
struct TwoMin { uint32_t n1, n2p; };
NRB200_SIMD TwoMin twomin_init() { return TwoMin{0u, kL7}; }                  // min1 = min2 = 127
NRB200_SIMD void twomin(uint32_t mag, TwoMin &s, uint32_t one, uint32_t mone)
{
  const uint32_t nmag = add_fma(mag, kL7, mone);               // 0x7f - mag
  const uint32_t m1 = msb_mask(add_fma(mag, s.n1, one));       // mag > min1
  const uint32_t nt = sel4(m1, nmag, s.n1);                    // 0x7f - max(mag, min1)
  s.n1 = sel4(m1, s.n1, nmag);
  const uint32_t m2 = msb_mask(add_fma(nt, s.n2p, mone));      // max(mag, min1) > min2
  s.n2p = sel4(m2, s.n2p, add_fma(nt, kL7, one));
}
NRB200_SIMD uint32_t twomin_min1(const TwoMin &s, uint32_t mone) { return add_fma(s.n1, kL7, mone); }
NRB200_SIMD uint32_t twomin_min2(const TwoMin &s, uint32_t mone) { return add_fma(s.n2p, 0xFEFEFEFEu, mone); }
// two minima of the union of two edge sets: feed the other set's two minima through the tracker (the cluster decoder splits the degree-19
// rows between the two halves of a warp)
NRB200_SIMD void twomin_merge(TwoMin &s, const TwoMin &o, uint32_t one, uint32_t mone)
{
  twomin(twomin_min1(o, mone), s, one, mone);
  twomin(twomin_min2(o, mone), s, one, mone);
}

// offset-binary cn->bn message (R + 128) of one edge from the row's two minima and sign product (bit 7 = negative).
// p1 = min1 | 0x80, p2 = min2 | 0x80 are formed once per row; |Q| >= min1 always, so |Q| != min1 <=> bit 7 of |Q| + n1; the negative
// message 128 - mag is the per-byte two's complement of 128 + mag, one multiply-add on the FMA pipe (kNegC); -0 comes out as 0x80 like +0.
NRB200_SIMD uint32_t make_r(uint32_t qsm, uint32_t n1, uint32_t p1, uint32_t p2, uint32_t sgn, uint32_t one, uint32_t mone)
{
  const uint32_t ne = msb_mask(add_fma(qsm & kL7, n1, one));                            // 0xFF where |Q| != min1
  const uint32_t x = sel4(ne, p1, p2);                                                  // 128 + excluded minimum
  const uint32_t n = msb_mask(sgn ^ qsm);                                               // 0xFF where the other signs multiply to -1
  return sel4(n, add_fma(x, kNegC, mone), x);
}

}  // namespace nrb200
