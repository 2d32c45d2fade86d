// Decoder kernel selection and launch.
#include "nrb200_ctx.h"
#include "ldpc_common.cuh"
#include "ldpc_decoder_generic.cuh"
#include "ldpc_decoder_packed.cuh"
#include <map>
#include <mutex>
#include <cstdlib>

namespace nrb200 {

static size_t generic_smem_bytes(const GraphDev &g)
{
  const size_t numLLR = (size_t)g.ncols * g.Z;
  auto a16 = [](size_t v) { return (v + 15) & ~(size_t)15; };
  return a16(sizeof(GraphDev)) + a16(numLLR) + a16((size_t)g.nreal * g.Z) + a16(numLLR) + a16((size_t)g.nrows * g.Z);
}

// device copies of the packed tables, keyed like Ctx::graphs
constexpr int kPackedDefaultThreadsZ384 = 768;
static std::mutex g_pk_mu;
static std::map<uint32_t, std::pair<PackedGraph *, PackedGraph>> g_pk;

static const PackedGraph *packed_graph(const GraphDev &h_g, const PackedGraph **host)
{
  const uint32_t key = ((uint32_t)h_g.BG << 24) | ((uint32_t)h_g.Z << 8) | (uint32_t)h_g.R;
  std::lock_guard<std::mutex> lk(g_pk_mu);
  auto it = g_pk.find(key);
  if (it == g_pk.end()) {
    PackedGraph pg;
    PackedGraph *d = nullptr;
    const int cap = h_g.Z == 384 ? kPackedMaxThreadsZ384 : kPackedMaxThreads;
    int maxt = h_g.Z == 384 ? kPackedDefaultThreadsZ384 : kPackedMaxThreads;
    if (const char *e = getenv("NRB200_PACKED_THREADS")) { int v = atoi(e); if (v >= 32 && v <= cap) maxt = v; }
    if (build_packed_graph(h_g, &pg, maxt) && (size_t)pg.total_bytes + sizeof(PackedGraph) + 64 <= (size_t)ctx().max_smem_optin) {
      if (cudaMalloc(&d, sizeof(PackedGraph)) != cudaSuccess) d = nullptr;
      else cudaMemcpy(d, &pg, sizeof(PackedGraph), cudaMemcpyHostToDevice);
    }
    it = g_pk.emplace(key, std::make_pair(d, pg)).first;
  }
  *host = &it->second.second;
  return it->second.first;
}

void packed_graph_cache_clear()
{
  std::lock_guard<std::mutex> lk(g_pk_mu);
  for (auto &kv : g_pk) if (kv.second.first) cudaFree(kv.second.first);
  g_pk.clear();
}

// Returns 0 or a negative error.  Asynchronous on `stream`.
int launch_decode(const GraphDev *d_g, const GraphDev &h_g, const DecodeArgs &a, cudaStream_t stream)
{
  Ctx &c = ctx();
  if (a.n_cb == 0) return 0;
  static const bool force_generic = getenv("NRB200_FORCE_GENERIC") != nullptr;
  const PackedGraph *h_pg = nullptr;
  const PackedGraph *d_pg = (h_g.Z % 4 == 0 && !force_generic) ? packed_graph(h_g, &h_pg) : nullptr;
  if (d_pg) {
    const size_t smem = (size_t)h_pg->total_bytes;
    // Z = 384 (K = 8448, both headline workloads) runs an instantiation with the row geometry as immediates
    void (*kern)(const PackedGraph *, DecodeArgs) = ldpc_decode_packed_kernel<0, kPackedMaxThreads>;
    int variant = 0;
    if (h_pg->Zw == 96 && !(a.quirks & 1)) {
      if (h_pg->nthreads > 864) { kern = ldpc_decode_packed_kernel<96, 960>; variant = 3; }
      else if (h_pg->nthreads > 768) { kern = ldpc_decode_packed_kernel<96, 864>; variant = 2; }
      else { kern = ldpc_decode_packed_kernel<96, 768>; variant = 1; }
    }
    static std::atomic<size_t> configured_pk[4];
    if (smem > configured_pk[variant].load()) {
      NRB200_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "packed smem attr");
      configured_pk[variant].store(smem);
    }
    kern<<<a.n_cb, h_pg->nthreads, smem, stream>>>(d_pg, a);
    c.launches++;
    NRB200_CUDA_OK(cudaGetLastError(), "packed decode launch");
    return 0;
  }
  {
    const size_t smem = generic_smem_bytes(h_g);
    if ((int)smem > c.max_smem_optin) return -3;
    static std::atomic<size_t> configured{0};
    if (smem > configured.load()) {
      NRB200_CUDA_OK(cudaFuncSetAttribute(ldpc_decode_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "smem attr");
      configured.store(smem);
    }
    int threads = h_g.Z * 2;
    threads = ((threads + 31) / 32) * 32;
    if (threads < 128) threads = 128;
    if (threads > 512) threads = 512;
    const unsigned grid = a.n_cb;
    ldpc_decode_generic_kernel<<<grid, threads, smem, stream>>>(d_g, a);
    c.launches++;
    NRB200_CUDA_OK(cudaGetLastError(), "decode launch");
  }
  return 0;
}

}  // namespace nrb200
