// Decoder kernel selection and launch.
#include "nrb200_ctx.h"
#include "ldpc_common.cuh"
#include "ldpc_decoder_generic.cuh"
#include "ldpc_decoder_packed.cuh"
#include "ldpc_decoder_cluster.cuh"
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
static std::map<uint64_t, std::pair<PackedGraph *, PackedGraph>> g_pk;   // key includes the device: tables live in that device's memory

static const PackedGraph *packed_graph(const GraphDev &h_g, const PackedGraph **host)
{
  const uint64_t key = ((uint64_t)ctx().dev << 40) | ((uint32_t)h_g.BG << 24) | ((uint32_t)h_g.Z << 8) | (uint32_t)h_g.R;
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

// cluster work partitions, keyed by graph and cluster size
static std::map<uint64_t, std::pair<ClusterSched *, ClusterSched>> g_cl;

static const ClusterSched *cluster_sched(const GraphDev &h_g, const PackedGraph &h_pg, int C, const ClusterSched **host)
{
  const uint64_t key = ((uint64_t)ctx().dev << 40) | ((uint64_t)C << 32) | ((uint32_t)h_g.BG << 24) | ((uint32_t)h_g.Z << 8) | (uint32_t)h_g.R;
  std::lock_guard<std::mutex> lk(g_pk_mu);
  auto it = g_cl.find(key);
  if (it == g_cl.end()) {
    ClusterSched cs;
    ClusterSched *d = nullptr;
    int T = C >= 8 ? 12 : C >= 4 ? 16 : 24;                 // warps per CTA: ~1-2 work items per warp and phase
    if (const char *e = getenv("NRB200_CLUSTER_WARPS")) { const int v = atoi(e); if (v >= 1 && v <= kClMaxWarps && C * v <= kClMaxLists) T = v; }
    if (build_cluster_sched(h_g, h_pg, C, T, &cs)) {
      if (cudaMalloc(&d, sizeof(ClusterSched)) != cudaSuccess) d = nullptr;
      else cudaMemcpy(d, &cs, sizeof(ClusterSched), cudaMemcpyHostToDevice);
    }
    it = g_cl.emplace(key, std::make_pair(d, cs)).first;
  }
  *host = &it->second.second;
  return it->second.first;
}

// debug: the cluster kernel's phase marks (64 per CTA of code block 0), see NRB200_CL_MARK
int debug_cluster_marks(long long *out)
{
  return cudaMemcpyFromSymbol(out, g_cl_marks, sizeof(long long) * kClMaxCtas * 64) == cudaSuccess ? 0 : -2;
}

void packed_graph_cache_clear()
{
  std::lock_guard<std::mutex> lk(g_pk_mu);
  for (auto &kv : g_pk) if (kv.second.first) cudaFree(kv.second.first);
  g_pk.clear();
  for (auto &kv : g_cl) if (kv.second.first) cudaFree(kv.second.first);
  g_cl.clear();
}

// CTAs per code block for a launch of n_cb blocks: a cluster per block while every cluster of the launch can be resident at once (the GPCs
// of a B200 hold 16 clusters of 8, 33 of 4, 74 of 2 with one 205 KB CTA per SM), one CTA per block beyond that -- there throughput, not the
// time of one block, is what matters and the single-CTA kernel spends the fewest SM-cycles per block.  NRB200_CLUSTER = 0 / 2 / 4 / 8 forces.
static int pick_cluster(uint32_t n_cb)
{
  static const int forced = []() { const char *e = getenv("NRB200_CLUSTER"); return e ? atoi(e) : -1; }();
  if (forced == 0 || forced == 2 || forced == 4 || forced == 8) return forced;
  return n_cb <= 15 ? 8 : n_cb <= 33 ? 4 : n_cb <= 74 ? 2 : 0;   // launch__cluster_max_active of the 8-CTA kernel is 15 on a B200 (ncu, profiles/)
}

int launch_decode_impl(const GraphDev *d_g, const GraphDev &h_g, const DecodeArgs &a, cudaStream_t stream, uint32_t load, int *cluster_out);

// Returns 0 or a negative error.  Asynchronous on `stream`.
int launch_decode(const GraphDev *d_g, const GraphDev &h_g, const DecodeArgs &a, cudaStream_t stream)
{
  return launch_decode_impl(d_g, h_g, a, stream, a.latency ? a.n_cb : 0xFFFFFFFFu, nullptr);   // throughput mode: never a cluster
}

// The low-latency path's launch: `load` = code blocks in flight on the device including this launch's (several small launches share the
// SMs, the cluster size is chosen for all of them together); *cluster_out = CTAs per block of the kernel that was launched (1 = no cluster).
int launch_decode_ll(const GraphDev *d_g, const GraphDev &h_g, const DecodeArgs &a, cudaStream_t stream, uint32_t load, int *cluster_out)
{
  return launch_decode_impl(d_g, h_g, a, stream, load, cluster_out);
}

int launch_decode_impl(const GraphDev *d_g, const GraphDev &h_g, const DecodeArgs &a, cudaStream_t stream, uint32_t load, int *cluster_out)
{
  Ctx &c = ctx();
  if (cluster_out) *cluster_out = 1;
  if (a.n_cb == 0) return 0;
  static const bool force_generic = getenv("NRB200_FORCE_GENERIC") != nullptr;
  const PackedGraph *h_pg = nullptr;
  const PackedGraph *d_pg = (h_g.Z % 4 == 0 && !force_generic) ? packed_graph(h_g, &h_pg) : nullptr;
  const int C = d_pg && h_pg->warp_items && !(a.quirks & 1) ? pick_cluster(load) : 0;
  if (C >= 2) {
    const ClusterSched *h_cs = nullptr;
    const ClusterSched *d_cs = cluster_sched(h_g, *h_pg, C, &h_cs);
    if (d_cs) {
      const size_t smem = (size_t)h_pg->total_bytes;
      const bool wide = h_cs->nthreads > 512;
      void (*kern)(const PackedGraph *, const ClusterSched *, DecodeArgs) =
          h_pg->Zw == 96 ? (wide ? ldpc_decode_cluster_kernel<96, 768> : ldpc_decode_cluster_kernel<96, 512>)
                         : (wide ? ldpc_decode_cluster_kernel<0, 768> : ldpc_decode_cluster_kernel<0, 512>);
      static std::atomic<size_t> configured_cl[kMaxDevices][4];
      const int v = (h_pg->Zw == 96 ? 2 : 0) + (wide ? 1 : 0);
      if (smem > configured_cl[c.dev][v].load()) {
        NRB200_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cluster smem attr");
        configured_cl[c.dev][v].store(smem);
      }
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(a.n_cb * (unsigned)C, 1, 1);
      cfg.blockDim = dim3((unsigned)h_cs->nthreads, 1, 1);
      cfg.dynamicSmemBytes = smem;
      cfg.stream = stream;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = (unsigned)C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      NRB200_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, d_pg, d_cs, a), "cluster decode launch");
      c.launches++;
      if (cluster_out) *cluster_out = C;
      return 0;
    }
  }
  if (d_pg) {
    const size_t smem = (size_t)h_pg->total_bytes;
    // Z = 384 (K = 8448, both headline workloads) runs an instantiation with the row geometry as immediates
    void (*kern)(const PackedGraph *, DecodeArgs) = ldpc_decode_packed_kernel<0, kPackedMaxThreads>;
    int variant = 0;
    if (h_pg->Zw == 96 && !(a.quirks & 1)) {
      if (h_pg->nthreads > 864) { kern = ldpc_decode_packed_kernel<96, 960>; variant = 3; }
      else if (h_pg->nthreads > 768) { kern = ldpc_decode_packed_kernel<96, 864>; variant = 2; }
      else if (a.ll_ctrl == nullptr && a.abort_flags == nullptr) { kern = ldpc_decode_packed_kernel<96, 768, true>; variant = 4; }   // the batch launches
      else { kern = ldpc_decode_packed_kernel<96, 768>; variant = 1; }
    }
    static std::atomic<size_t> configured_pk[kMaxDevices][5];
    if (smem > configured_pk[c.dev][variant].load()) {
      NRB200_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "packed smem attr");
      configured_pk[c.dev][variant].store(smem);
    }
    kern<<<a.n_cb, h_pg->nthreads, smem, stream>>>(d_pg, a);
    c.launches++;
    NRB200_CUDA_OK(cudaGetLastError(), "packed decode launch");
    return 0;
  }
  {
    const size_t smem = generic_smem_bytes(h_g);
    if ((int)smem > c.max_smem_optin) return -3;
    static std::atomic<size_t> configured[kMaxDevices];
    if (smem > configured[c.dev].load()) {
      NRB200_CUDA_OK(cudaFuncSetAttribute(ldpc_decode_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "smem attr");
      configured[c.dev].store(smem);
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
