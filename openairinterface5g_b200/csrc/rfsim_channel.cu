// rfsimulator channel application: rxAddInput (radio/rfsimulator/apply_channelmod.c:55-111) for every receive antenna in one launch.
// One thread owns one output sample of one receive antenna and walks tx antennas and taps in the reference's order, so the double-precision sums are the
// reference's bit for bit (products and sums as separate IEEE operations: __dmul_rn / __dadd_rn / __dsub_rn keep nvcc from contracting them).  A CTA stages the
// taps of its receive antenna (nb_tx x L complex doubles) and the window of the circular tx buffer its 256 samples reach back into ((256 + L - 1) x nb_tx c16)
// in shared memory: every tx sample is read from global memory once per CTA instead of L times.  FP64-pipe bound: 8 double operations per tap and tx antenna.
// (Staging the samples already converted to double -- one I2F per staged sample instead of one per tap -- was measured SLOWER: 77.9 against 69.5 us for the
// 2 x 2 frame; the 16-byte shared loads per tap cost more than the conversions they save.)
#include "nrb200_ctx.h"
#include "../../include/nrb200_rfsim.h"
#include <cmath>
#include <cstring>

namespace nrb200 {

constexpr int kRfTpb = 256;

__global__ void __launch_bounds__(kRfTpb) rfsim_channel_kernel(int nb_tx, int nb_rx, int L, int dd, double pathLossLinear, double noise_per_sample,
                                                               const double *__restrict__ ch, const unsigned *__restrict__ sig, unsigned *__restrict__ out,
                                                               unsigned out_stride, int n, unsigned long long TS, unsigned CirSize, const double *__restrict__ noise)
{
  extern __shared__ __align__(16) unsigned char rf_smem[];
  double2 *s_ch = reinterpret_cast<double2 *>(rf_smem);                               // [nb_tx][L]
  unsigned *s_sig = reinterpret_cast<unsigned *>(s_ch + (size_t)nb_tx * L);           // [nb_tx][kRfTpb + L - 1], oldest sample first
  const int rx = blockIdx.y, i0 = blockIdx.x * kRfTpb, i = i0 + threadIdx.x;
  for (int t = threadIdx.x; t < nb_tx * L; t += kRfTpb) {
    const int tx = t / L, l = t - tx * L;
    s_ch[t] = reinterpret_cast<const double2 *>(ch)[(size_t)(rx + tx * nb_rx) * L + l];
  }
  // sample position p = TS + i - l - dd (64-bit unsigned arithmetic like the reference's uint64_t TS); the window starts at l = L - 1 of the CTA's first sample
  const unsigned long long p0 = TS + (unsigned long long)(long long)i0 - (unsigned long long)(L - 1) - (unsigned long long)dd;
  const int win = kRfTpb + L - 1;
  for (int t = threadIdx.x; t < win * nb_tx; t += kRfTpb) {
    const int w = t / nb_tx, tx = t - w * nb_tx;
    const unsigned idx = (unsigned)(((p0 + (unsigned long long)w) * (unsigned long long)nb_tx + (unsigned long long)tx + CirSize) % CirSize);
    s_sig[tx * win + w] = __ldg(sig + idx);                 // [tx][window]: neighbouring threads read neighbouring words (no bank conflicts for any nb_tx)
  }
  __syncthreads();
  if (i >= n) return;
  double rr = 0.0, ri = 0.0;
  for (int tx = 0; tx < nb_tx; tx++) {
    const double2 *c = s_ch + (size_t)tx * L;
    // tap l of sample i sits at window position (threadIdx.x + L - 1 - l)
    const unsigned *x = s_sig + (size_t)tx * win + (threadIdx.x + L - 1);
    for (int l = 0; l < L; l++) {
      const unsigned v = x[-l];
      const double xr = (double)(short)(v & 0xFFFFu), xi = (double)(short)(v >> 16);
      const double2 h = c[l];
      rr = __dadd_rn(rr, __dsub_rn(__dmul_rn(xr, h.x), __dmul_rn(xi, h.y)));
      ri = __dadd_rn(ri, __dadd_rn(__dmul_rn(xi, h.x), __dmul_rn(xr, h.y)));
    }
  }
  double ar = __dmul_rn(rr, pathLossLinear), ai = __dmul_rn(ri, pathLossLinear);
  double nr = 0.0, ni = 0.0;
  if (noise) {
    const double2 z = reinterpret_cast<const double2 *>(noise)[(size_t)rx * n + i];
    nr = __dmul_rn(noise_per_sample, z.x); ni = __dmul_rn(noise_per_sample, z.y);
  }
  // without noise the reference still adds noise_per_sample * 0.0 = +0.0, which changes nothing (x + 0.0 == x, and -0.0 + 0.0 rounds to the same integer 0)
  const long long qr = llround(__dadd_rn(ar, nr)), qi = llround(__dadd_rn(ai, ni));
  const unsigned o = out[(size_t)rx * out_stride + i];
  const unsigned r16 = ((unsigned)(int)(short)(o & 0xFFFFu) + (unsigned)qr) & 0xFFFFu, i16 = ((unsigned)(int)(short)(o >> 16) + (unsigned)qi) & 0xFFFFu;
  out[(size_t)rx * out_stride + i] = r16 | (i16 << 16);
}

static int rfsim_check(const nrb200_rfsim_chan_t &c, uint32_t n, uint32_t CirSize)
{
  if (c.nb_tx < 1 || c.nb_tx > 8 || c.nb_rx < 1 || c.nb_rx > 8 || c.channel_length < 1 || c.channel_length > 255 || n == 0 || CirSize == 0) return -4;
  return 0;
}

int launch_rfsim(const nrb200_rfsim_chan_t &c, const double *ch, const int16_t *sig, int16_t *out, uint32_t out_stride, uint32_t n, uint64_t TS, uint32_t CirSize,
                 const double *noise, cudaStream_t st)
{
  int rc = rfsim_check(c, n, CirSize);
  if (rc) return rc;
  const int L = (int)c.channel_length, dd = std::abs(c.channel_offset);
  const double pathLossLinear = std::pow(10, c.path_loss_dB / 20.0), noise_per_sample = std::pow(10, c.noise_power_dB / 10.0) * 256;   // :66-69, host libm like the reference
  const size_t smem = (size_t)c.nb_tx * L * sizeof(double2) + (size_t)(kRfTpb + L - 1) * c.nb_tx * 4;
  static bool attr_done[kMaxDevices] = {false};
  if (!attr_done[ctx().dev]) {
    NRB200_CUDA_OK(cudaFuncSetAttribute(rfsim_channel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 255 * 16 + (kRfTpb + 254) * 8 * 4), "rfsim attribute");
    attr_done[ctx().dev] = true;
  }
  rfsim_channel_kernel<<<dim3((n + kRfTpb - 1) / kRfTpb, c.nb_rx), kRfTpb, smem, st>>>((int)c.nb_tx, (int)c.nb_rx, L, dd, pathLossLinear, noise_per_sample, ch,
                                                                                      (const unsigned *)sig, (unsigned *)out, out_stride, (int)n,
                                                                                      (unsigned long long)TS, CirSize, noise);
  ctx().launches++;
  NRB200_CUDA_OK(cudaGetLastError(), "rfsim_channel launch");
  return 0;
}

}  // namespace nrb200
