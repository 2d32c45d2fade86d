// NR LDPC rate matching / bit interleaving (TX) and de-interleaving / rate recovery with HARQ soft combining / decoder-input
// packing (RX) as gather kernels: one CTA per code block segment, one thread per output element.
//
// Reference (all in openair1/PHY/CODING/nr_rate_matching.c unless noted):
//   bit selection        nr_rate_matching_ldpc      :424-505   circular buffer of Ncb = min(N, 3*Tbslbrm/(2C)) positions, start
//                                                               k0 = floor(index_k0[BG][rv]*Ncb/N)*Z, filler span [Foffset, Foffset+F) skipped,
//                                                               wrap-around repetition until E bits
//   interleaver          nr_interleaving_ldpc       :36-305    f[j*Qm + i] = e[i*E/Qm + j]
//   de-interleaver       nr_deinterleaving_ldpc     :310-388   int16 soft values, inverse mapping
//   rate recovery        nr_rate_matching_ldpc_rx   :507-603   w[ind] += soft[k] (int16, wraps), optional clear of the first Ncb entries
//   decoder input        nr_ulsch_decoding.c:195-210 / nr_dlsch_decoding.c:235-250
//                                                               z = [0 x 2Z | d[0..K-F-2Z) | 127 x F | d[K-2Z..)], packs_epi16 -> int8
// The reference walks the circular buffer sequentially; here every element computes its own source index:
// the non-filler positions form a ring of L = Foffset + max(0, Ncb - Foffset - F) slots, bit k sits on slot (r0 + k) mod L.
#include "nrb200_ctx.h"
#include "ldpc_common.cuh"
#include "../../include/nrb200_ldpc.h"

namespace nrb200 {

struct RmGeom {
  uint32_t N, Ncb, Foffset, F, L, r0;   // ring of non-filler positions and the slot of k0
};

__host__ __device__ inline RmGeom rm_geom(int BG, int Z, uint32_t Tbslbrm, uint32_t C, uint32_t F, uint32_t K, int rv)
{
  RmGeom g;
  g.N = (uint32_t)(BG == 1 ? 66 : 50) * Z;
  g.Ncb = g.N;
  if (Tbslbrm) { const uint32_t Nref = 3 * Tbslbrm / (2 * C); g.Ncb = g.N < Nref ? g.N : Nref; }
  g.F = F;
  g.Foffset = K - F - 2 * Z;
  const uint32_t k0tab[2][4] = {{0, 17, 33, 56}, {0, 13, 25, 43}};
  uint32_t ind = (k0tab[BG - 1][rv] * g.Ncb / g.N) * Z;
  if (ind >= g.Foffset && ind < g.Foffset + F) ind = g.Foffset + F;
  const uint32_t tail = g.Ncb > g.Foffset + F ? g.Ncb - g.Foffset - F : 0;
  g.L = g.Foffset + tail;
  g.r0 = ind < g.Foffset ? ind : (ind >= g.Ncb ? g.L : ind - F);   // rank of the start position on the ring
  if (g.r0 >= g.L) g.r0 = 0;                                       // k0 beyond the last usable position: the walk restarts at 0
  return g;
}
__device__ __forceinline__ uint32_t rm_slot_to_pos(const RmGeom &g, uint32_t slot) { return slot < g.Foffset ? slot : slot + g.F; }

// ---- TX: d (encoder output, one bit per byte, N per segment) -> f (E_r interleaved bits per segment, contiguous)
__global__ void rm_tx_kernel(nrb200_rm_desc_t p, const uint8_t *__restrict__ d, uint32_t d_stride, const uint32_t *__restrict__ E_seg,
                             const uint32_t *__restrict__ f_off, uint8_t *__restrict__ f)
{
  const uint32_t r = blockIdx.x;
  const uint32_t E = E_seg[r], EQm = E / p.Qm;
  const RmGeom g = rm_geom(p.BG, p.Z, p.Tbslbrm, p.C, p.F, p.K, p.rv);
  const uint8_t *w = d + (size_t)r * d_stride;
  uint8_t *out = f + f_off[r];
  for (uint32_t o = threadIdx.x; o < EQm * p.Qm; o += blockDim.x) {
    const uint32_t j = o / p.Qm, i = o - j * p.Qm;          // f[j*Qm + i] = e[i*EQm + j]
    const uint32_t k = i * EQm + j;
    out[o] = w[rm_slot_to_pos(g, (g.r0 + k) % g.L)];
  }
  // nr_interleaving_ldpc memset()s f and leaves the E % Qm tail zero
  for (uint32_t o = EQm * p.Qm + threadIdx.x; o < E; o += blockDim.x) out[o] = 0;
}

// ---- RX: soft (E_r interleaved int16 per segment) -> HARQ buffer d (int16, += with wrap) -> decoder input llr (int8)
// SoftT = int16_t (the CPU-compatible convention: demodulator LLRs) or int8_t (the offload convention: the caller already packed to int8)
// One segment is spread over gridDim.y CTAs: a ring slot's thread sums its repetitions, updates the soft buffer AND writes the slot's saturated decoder input
// itself (decoder position = buffer position + 2 Z), so nothing waits for anything; the positions no ring slot owns (the punctured 2 Z, the fillers, what lies
// beyond a limited buffer's Ncb) are filled by a second independent sweep.
template <typename SoftT>
__global__ void __launch_bounds__(256) rm_rx_kernel(nrb200_rm_desc_t p, const SoftT *__restrict__ soft, const uint32_t *__restrict__ E_seg,
                                                    const uint32_t *__restrict__ s_off, int16_t *__restrict__ harq, uint32_t harq_stride, int8_t *__restrict__ llr,
                                                    uint32_t llr_stride)
{
  const uint32_t r = blockIdx.x, tid = blockIdx.y * blockDim.x + threadIdx.x, nthr = gridDim.y * blockDim.x;
  const uint32_t E = E_seg[r], EQm = E / p.Qm;
  const RmGeom g = rm_geom(p.BG, p.Z, p.Tbslbrm, p.C, p.F, p.K, p.rv);
  const SoftT *in = soft + s_off[r];
  int16_t *w = harq + (size_t)r * harq_stride;
  const uint32_t kcZ = (uint32_t)(p.BG == 1 ? 68 : 52) * p.Z, twoZ = 2u * p.Z;
  int8_t *l = llr + (size_t)r * llr_stride;
  // rate recovery: every ring slot sums the soft values of all its repetitions (int16 arithmetic wraps like the reference's +=)
  for (uint32_t slot = tid; slot < g.L; slot += nthr) {
    const uint32_t pos = rm_slot_to_pos(g, slot);
    uint32_t acc = p.clear ? 0u : (uint32_t)(uint16_t)w[pos];
    for (uint32_t k = (slot + g.L - g.r0) % g.L; k < E; k += g.L) {
      const uint32_t i = k / EQm, j = k - i * EQm;           // e[i*EQm + j] = f[j*Qm + i]
      acc += (uint32_t)(uint16_t)(int16_t)in[j * p.Qm + i];
    }
    const int v = (int)(int16_t)(uint16_t)acc;
    w[pos] = (int16_t)v;
    l[pos + twoZ] = (int8_t)(v > 127 ? 127 : v < -128 ? -128 : v);     // decoder input: saturate(d) (nr_ulsch_decoding.c:195-210)
  }
  // everything else of the decoder input (kc*Z int8): punctured 2Z = 0, fillers = 127 (their soft-buffer span is cleared with the rest of Ncb on new data),
  // positions beyond Ncb = whatever the soft buffer holds there
  for (uint32_t i = tid; i < kcZ; i += nthr) {
    if (i < twoZ) { l[i] = 0; continue; }
    const uint32_t pos = i - twoZ;
    if (i >= p.K - p.F && i < p.K) { l[i] = 127; if (p.clear && pos < g.Ncb) w[pos] = 0; continue; }
    if (pos >= g.Ncb) { const int v = w[pos]; l[i] = (int8_t)(v > 127 ? 127 : v < -128 ? -128 : v); }
  }
}

// CTAs per segment: enough to put a transport block's few segments on the whole GPU, one for the hundreds of segments of a multi-user batch
static inline unsigned rm_rx_parts(uint32_t n_seg) { return n_seg >= 256 ? 1u : n_seg >= 64 ? 4u : n_seg >= 16 ? 8u : 16u; }

int launch_rm_tx(const nrb200_rm_desc_t &p, const uint8_t *d, uint32_t d_stride, const uint32_t *E, const uint32_t *off, uint8_t *f, cudaStream_t st)
{
  if (p.n_seg == 0) return 0;
  rm_tx_kernel<<<p.n_seg, 512, 0, st>>>(p, d, d_stride, E, off, f);
  ctx().launches++;
  NRB200_CUDA_OK(cudaGetLastError(), "rm_tx launch");
  return 0;
}

int launch_rm_rx(const nrb200_rm_desc_t &p, const int16_t *soft, const uint32_t *E, const uint32_t *off, int16_t *harq, uint32_t harq_stride,
                 int8_t *llr, uint32_t llr_stride, cudaStream_t st)
{
  if (p.n_seg == 0) return 0;
  rm_rx_kernel<int16_t><<<dim3(p.n_seg, rm_rx_parts(p.n_seg)), 256, 0, st>>>(p, soft, E, off, harq, harq_stride, llr, llr_stride);
  ctx().launches++;
  NRB200_CUDA_OK(cudaGetLastError(), "rm_rx launch");
  return 0;
}

int launch_rm_rx8(const nrb200_rm_desc_t &p, const int8_t *soft, const uint32_t *E, const uint32_t *off, int16_t *harq, uint32_t harq_stride,
                  int8_t *llr, uint32_t llr_stride, cudaStream_t st)
{
  if (p.n_seg == 0) return 0;
  rm_rx_kernel<int8_t><<<dim3(p.n_seg, rm_rx_parts(p.n_seg)), 256, 0, st>>>(p, soft, E, off, harq, harq_stride, llr, llr_stride);
  ctx().launches++;
  NRB200_CUDA_OK(cudaGetLastError(), "rm_rx launch");
  return 0;
}

}  // namespace nrb200
