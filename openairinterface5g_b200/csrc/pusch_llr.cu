// PUSCH/PDSCH single-layer max-log LLR computation (reference openair1/PHY/NR_TRANSPORT/nr_ulsch_llr_computation.c:45-312,
// UE mirror NR_UE_TRANSPORT/nr_dlsch_llr_computation.c:51-350): per resource element
//   QPSK   (y >> 3)
//   16QAM  y, subs(|h|^2 a, |y|)
//   64QAM  y, l1 = subs(mag_a, |y|), subs(mag_b, |l1|)
//   256QAM y, l1, l2 = subs(mag_b, |l1|), subs(mag_c, |l2|)
// with abs_epi16 (|-32768| stays -32768) and saturating subs_epi16, outputs interleaved per RE (Qm int16 each).
// Purely HBM bound: 4 B of y + 4 B per magnitude plane in, 2*Qm B out per RE; each thread handles 4 REs with 16-byte accesses.
#include "nrb200_ctx.h"

namespace nrb200 {

__device__ __forceinline__ int abs16w(int v) { return v == -32768 ? -32768 : abs(v); }
__device__ __forceinline__ int subs16(int a, int b) { return max(-32768, min(32767, a - b)); }
__device__ __forceinline__ unsigned pk(int lo, int hi) { return ((unsigned)lo & 0xFFFFu) | ((unsigned)hi << 16); }
__device__ __forceinline__ int lo16(unsigned w) { return (int)(short)(w & 0xFFFFu); }
__device__ __forceinline__ int hi16(unsigned w) { return (int)(short)(w >> 16); }

template <int QM>
__device__ __forceinline__ void llr_re(unsigned y, unsigned a, unsigned b, unsigned c, unsigned *o)
{
  if (QM == 2) { o[0] = pk(lo16(y) >> 3, hi16(y) >> 3); return; }
  o[0] = y;
  const int l1r = subs16(lo16(a), abs16w(lo16(y))), l1i = subs16(hi16(a), abs16w(hi16(y)));
  o[1] = pk(l1r, l1i);
  if (QM == 4) return;
  const int l2r = subs16(lo16(b), abs16w(l1r)), l2i = subs16(hi16(b), abs16w(l1i));
  o[2] = pk(l2r, l2i);
  if (QM == 6) return;
  o[3] = pk(subs16(lo16(c), abs16w(l2r)), subs16(hi16(c), abs16w(l2i)));
}

template <int QM>
__global__ void __launch_bounds__(256) pusch_llr_kernel(uint32_t nb_re, const unsigned *__restrict__ y, const unsigned *__restrict__ ma,
                                                        const unsigned *__restrict__ mb, const unsigned *__restrict__ mc, unsigned *__restrict__ out)
{
  constexpr int W = QM / 2;                     // output words per RE
  const uint32_t nq = nb_re >> 2;
  for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += gridDim.x * blockDim.x) {
    const uint4 yy = reinterpret_cast<const uint4 *>(y)[q];
    uint4 aa = make_uint4(0, 0, 0, 0), bb = aa, cc = aa;
    if (QM >= 4) aa = reinterpret_cast<const uint4 *>(ma)[q];
    if (QM >= 6) bb = reinterpret_cast<const uint4 *>(mb)[q];
    if (QM >= 8) cc = reinterpret_cast<const uint4 *>(mc)[q];
    unsigned o[4 * W];
    llr_re<QM>(yy.x, aa.x, bb.x, cc.x, o);
    llr_re<QM>(yy.y, aa.y, bb.y, cc.y, o + W);
    llr_re<QM>(yy.z, aa.z, bb.z, cc.z, o + 2 * W);
    llr_re<QM>(yy.w, aa.w, bb.w, cc.w, o + 3 * W);
    uint4 *dst = reinterpret_cast<uint4 *>(out + (size_t)q * 4 * W);
#pragma unroll
    for (int i = 0; i < W; i++) dst[i] = make_uint4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
  }
  // tail (nb_re not a multiple of 4)
  if (blockIdx.x == 0 && threadIdx.x < (nb_re & 3)) {
    const uint32_t i = (nq << 2) + threadIdx.x;
    unsigned o[W];
    llr_re<QM>(y[i], QM >= 4 ? ma[i] : 0u, QM >= 6 ? mb[i] : 0u, QM >= 8 ? mc[i] : 0u, o);
    for (int k = 0; k < W; k++) out[(size_t)i * W + k] = o[k];
  }
}

int launch_pusch_llr(int Qm, uint32_t nb_re, const int16_t *y, const int16_t *ma, const int16_t *mb, const int16_t *mc, int16_t *out, cudaStream_t st)
{
  if (nb_re == 0) return 0;
  const unsigned grid = std::min<unsigned>((nb_re / 4 + 255) / 256 + 1, 148 * 16);
  const unsigned *Y = (const unsigned *)y, *A = (const unsigned *)ma, *B = (const unsigned *)mb, *Cc = (const unsigned *)mc;
  switch (Qm) {
    case 2: pusch_llr_kernel<2><<<grid, 256, 0, st>>>(nb_re, Y, A, B, Cc, (unsigned *)out); break;
    case 4: pusch_llr_kernel<4><<<grid, 256, 0, st>>>(nb_re, Y, A, B, Cc, (unsigned *)out); break;
    case 6: pusch_llr_kernel<6><<<grid, 256, 0, st>>>(nb_re, Y, A, B, Cc, (unsigned *)out); break;
    case 8: pusch_llr_kernel<8><<<grid, 256, 0, st>>>(nb_re, Y, A, B, Cc, (unsigned *)out); break;
    default: return -4;
  }
  ctx().launches++;
  NRB200_CUDA_OK(cudaGetLastError(), "pusch_llr launch");
  return 0;
}

}  // namespace nrb200
