// gNB PRACH detector: rx_nr_prach (openair1/PHY/NR_TRANSPORT/nr_prach.c:414-714), unrestricted set.
// The reference walks the 64 preambles; whenever a preamble starts a new root sequence it correlates every antenna with that root (conj(X_u) * rxsigF >> 15), runs one
// idft() per antenna, accumulates the powers and normalises; then it scans the preamble's window of NCS2 delay bins.  Here:
//   prach_corr_kernel    all (root, antenna) products, zero padded to the transform size, in one launch
//   dft_batch_internal   ONE batched IDFT-1024 / IDFT-256 over roots x antennas (the library's bit-exact Q15 transform)
//   prach_window_kernel  one CTA per root: antenna powers summed with 32-bit wrap, >> log2(size), / nb_rx into shared memory, then every preamble of the root reduces
//                        its window to (largest dB_fixed_times10, first bin that has it)
//   prach_pick_kernel    the reference's running maximum over (preamble, bin) in order == the lexicographically first pair holding the global maximum; timing advance
#include "nrb200_ctx.h"
#include "../../include/nrb200_prach.h"
#include "nr_db_table.h"

namespace nrb200 {

int dft_batch_internal(int N, int inverse, uint32_t n, const int16_t *d_in, int16_t *d_out, int scale, cudaStream_t st);   // dfts_internal.cu

__constant__ short c_db_table[256] = NRB200_DB_TABLE_TIMES10;

struct PrachGeom { int nb_rx, N_ZC, size, lg, NCS, NCS2, per_root, roots, fmt, mu, is_short; unsigned rx_stride; };

// dB_fixed_times10 (TOOLS/dB_routines.c:132-155)
__device__ __forceinline__ int prach_db(unsigned x)
{
  int v;
  if (x == 0) return 0;
  if (x & 0xff000000u) v = c_db_table[((x >> 24) & 255) - 1] + 3 * c_db_table[255];
  else if (x & 0x00ff0000u) v = c_db_table[((x >> 16) & 255) - 1] + 2 * c_db_table[255];
  else if (x & 0x0000ff00u) v = c_db_table[((x >> 8) & 255) - 1] + c_db_table[255];
  else v = c_db_table[(x & 255) - 1];
  return min(v, 900);
}

__global__ void __launch_bounds__(256) prach_corr_kernel(PrachGeom G, const unsigned *__restrict__ xu, const unsigned *__restrict__ rx, unsigned *__restrict__ prachF)
{
  const int ra = blockIdx.y, root = ra / G.nb_rx, a = ra - root * G.nb_rx, k = blockIdx.x * 256 + threadIdx.x;
  if (k >= G.size) return;
  unsigned v = 0;
  if (k < G.N_ZC) {
    const unsigned X = __ldg(xu + (size_t)root * 839 + k), r = __ldg(rx + (size_t)a * G.rx_stride + k);
    const int xr = (short)(X & 0xFFFFu), xi = (short)(X >> 16), rr = (short)(r & 0xFFFFu), ri = (short)(r >> 16);
    const int pr = (int)(short)(((int)((unsigned)(xr * rr) + (unsigned)(xi * ri))) >> 15), pi = (int)(short)(((int)((unsigned)(xr * ri) - (unsigned)(xi * rr))) >> 15);
    v = ((unsigned)pr & 0xFFFFu) | ((unsigned)pi << 16);
  }
  prachF[(size_t)ra * G.size + k] = v;
}

// per preamble: res[2 * p] = largest dB over the window, res[2 * p + 1] = first bin holding it
__global__ void __launch_bounds__(256) prach_window_kernel(PrachGeom G, const unsigned *__restrict__ t, int *__restrict__ res)
{
  __shared__ int s_pow[1024];
  const int root = blockIdx.x;
  for (int i = threadIdx.x; i < G.size; i += 256) {
    unsigned acc = 0;
    for (int a = 0; a < G.nb_rx; a++) {
      const unsigned v = __ldg(t + (size_t)(root * G.nb_rx + a) * G.size + i);
      const int r = (short)(v & 0xFFFFu), im = (short)(v >> 16);
      acc += (unsigned)(r * r) + (unsigned)(im * im);
    }
    s_pow[i] = ((int)acc >> G.lg) / G.nb_rx;
  }
  __syncthreads();
  // one warp per preamble of this root: lanes stride the window, keep (largest dB, first bin) and combine with shuffles -- no block-wide barrier per preamble
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int v = warp; v < G.per_root; v += 8) {
    const int p = root * G.per_root + v;
    if (p >= 64) break;
    int shift = (v * G.NCS) % G.N_ZC;                      // preamble_shift = -v NCS mod N_ZC (:470-474)
    shift = shift == 0 ? 0 : G.N_ZC - shift;
    const unsigned shift2 = shift == 0 ? 0u : (unsigned)((shift << G.lg) / G.N_ZC);
    int best = -1, bin = 0x7fffffff;
    for (int i = lane; i < G.NCS2; i += 32) {
      const unsigned b = shift2 + (unsigned)i;
      const int db = prach_db((unsigned)(b < (unsigned)G.size ? s_pow[b] : 0));
      if (db > best) { best = db; bin = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const int ob = __shfl_xor_sync(0xffffffffu, best, o), oi = __shfl_xor_sync(0xffffffffu, bin, o);
      if (ob > best || (ob == best && oi < bin)) { best = ob; bin = oi; }
    }
    if (lane == 0) { res[2 * p] = best; res[2 * p + 1] = bin; }
  }
}

__global__ void prach_pick_kernel(PrachGeom G, const int *__restrict__ res, int *__restrict__ out)
{
  if (threadIdx.x != 0) return;
  int e = 0, d = 0, p = 0;                                  // *max_preamble_energy = *max_preamble_delay = *max_preamble = 0 (:478-480)
  for (int q = 0; q < 64; q++)
    if (res[2 * q] > e) { e = res[2 * q]; d = res[2 * q + 1]; p = q; }
  unsigned ta = (unsigned)d;
  if (!G.is_short) {                                        // :693-698
    if (G.fmt == 0 || G.fmt == 1 || G.fmt == 2) ta = (unsigned)(d * 3 * (1 << G.mu) / 2);
    else if (G.fmt == 3) ta = (unsigned)(d * 3 * (1 << G.mu) / 8);
  } else ta = (unsigned)(d / 2);
  out[0] = p; out[1] = e; out[2] = (int)(ta & 0xFFFFu);
}

static int prach_geom(const nrb200_prach_t &d, PrachGeom *G)
{
  if (d.nb_rx < 1 || d.nb_rx > 8 || d.short_sequence > 1 || d.restricted_set != 0 || d.numerology > 4) return -4;
  G->is_short = (int)d.short_sequence; G->N_ZC = d.short_sequence ? 139 : 839; G->size = d.short_sequence ? 256 : 1024; G->lg = d.short_sequence ? 8 : 10;
  if (d.num_cs >= (uint32_t)G->N_ZC) return -4;
  G->nb_rx = (int)d.nb_rx; G->NCS = (int)d.num_cs; G->fmt = (int)d.prach_format; G->mu = (int)d.numerology;
  G->NCS2 = d.short_sequence ? (int)((d.num_cs << 8) / 139) : (int)((d.num_cs << 10) / 839);
  if (G->NCS2 == 0) G->NCS2 = G->N_ZC;
  G->per_root = G->NCS == 0 ? 1 : G->N_ZC / G->NCS;
  G->roots = (64 + G->per_root - 1) / G->per_root;
  G->rx_stride = d.rx_stride;
  return 0;
}

uint32_t prach_num_roots(const nrb200_prach_t &d) { PrachGeom G; return prach_geom(d, &G) == 0 ? (uint32_t)G.roots : 0u; }
// two planes of roots x antennas transforms (input and output) + the per-preamble results
size_t prach_scratch_bytes(const nrb200_prach_t &d) { PrachGeom G; return prach_geom(d, &G) == 0 ? (size_t)2 * G.roots * G.nb_rx * G.size * 4 + 64 * 8 : 0; }

int launch_prach(const nrb200_prach_t &d, const int16_t *xu, const int16_t *rxsigF, int32_t *out3, void *scratch, cudaStream_t st)
{
  PrachGeom G;
  int rc = prach_geom(d, &G);
  if (rc) return rc;
  if (G.rx_stride < (unsigned)G.N_ZC) return -4;
  const int n = G.roots * G.nb_rx;
  unsigned *pin = (unsigned *)scratch, *pout = pin + (size_t)n * G.size;
  int *res = (int *)(pout + (size_t)n * G.size);
  prach_corr_kernel<<<dim3((G.size + 255) / 256, n), 256, 0, st>>>(G, (const unsigned *)xu, (const unsigned *)rxsigF, pin);
  NRB200_CUDA_OK(cudaGetLastError(), "prach_corr launch");
  if ((rc = dft_batch_internal(G.size, 1, (uint32_t)n, (const int16_t *)pin, (int16_t *)pout, 1, st)) != 0) return rc;
  prach_window_kernel<<<G.roots, 256, 0, st>>>(G, pout, res);
  prach_pick_kernel<<<1, 32, 0, st>>>(G, res, out3);
  ctx().launches += 4;
  NRB200_CUDA_OK(cudaGetLastError(), "prach launch");
  return 0;
}

}  // namespace nrb200
