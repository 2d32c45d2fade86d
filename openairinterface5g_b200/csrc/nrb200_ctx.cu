// Context management for libldpc_b200.so (see nrb200_ctx.h).  There is deliberately no CPU fallback anywhere:
// without a usable CUDA device init() fails and every entry point returns an error.
#include "nrb200_ctx.h"
#include "ldpc_common.cuh"
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace nrb200 {

void packed_graph_cache_clear();

static Ctx g_ctx[kMaxDevices];
static thread_local int tls_dev = -1;

int device_count()
{
  static const int n = []() { int k = 0; if (cudaGetDeviceCount(&k) != cudaSuccess) { cudaGetLastError(); k = 0; } return k < kMaxDevices ? k : kMaxDevices; }();
  return n;
}

int current_device()
{
  if (tls_dev < 0) {
    const int n = device_count();
    int want = 0;
    if (const char *s = getenv("NRB200_DEVICE")) want = atoi(s);
    else if (const char *s2 = getenv("LOCAL_RANK")) want = n > 0 ? atoi(s2) % n : 0;   // one process per GPU under torchrun
    if (want < 0 || want >= (n > 0 ? n : 1)) want = 0;
    tls_dev = want;
  }
  return tls_dev;
}

int set_current_device(int dev)
{
  if (dev < 0 || dev >= device_count()) return -1;
  tls_dev = dev;
  return 0;
}

Ctx &ctx() { return g_ctx[current_device()]; }

void Ctx::set_error(const char *where, cudaError_t e)
{
  std::lock_guard<std::mutex> lk(mu);
  last_error = std::string(where) + ": " + cudaGetErrorString(e);
  if (getenv("NRB200_VERBOSE")) fprintf(stderr, "[nrb200] %s\n", last_error.c_str());
}

bool Workspace::reserve(size_t in, size_t out, size_t aux)
{
  auto grow = [](void **d, void **h, size_t *cap, size_t need) -> bool {
    if (need <= *cap) return true;
    size_t n = need + need / 4 + 4096;
    if (*d) cudaFree(*d);
    if (*h) cudaFreeHost(*h);
    *d = nullptr; *h = nullptr; *cap = 0;
    if (cudaMalloc(d, n) != cudaSuccess) return false;
    if (cudaHostAlloc(h, n, cudaHostAllocDefault) != cudaSuccess) return false;
    *cap = n;
    return true;
  };
  return grow(&d_in, &h_in, &cap_in, in) && grow(&d_out, &h_out, &cap_out, out) && grow(&d_aux, &h_aux, &cap_aux, aux);
}

static const uint32_t kPoly[8] = {0x864cfb00u, 0x80006300u, 0xb2b11700u, 0x10210000u, 0x80F00000u, 0xc4200000u, 0x9B000000u, 0x84000000u};

int Ctx::init()
{
  std::lock_guard<std::mutex> lk(mu);
  if (inited) return 0;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    last_error = std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    return -1;
  }
  const int want = (int)(this - g_ctx);                 // Ctx i drives device i
  if (want < 0 || want >= n) { last_error = "no such CUDA device"; return -1; }
  if ((e = cudaSetDevice(want)) != cudaSuccess) { last_error = std::string("cudaSetDevice: ") + cudaGetErrorString(e); return -1; }
  dev = want;
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, dev);
  sm_count = p.multiProcessorCount;
  max_smem_optin = (int)p.sharedMemPerBlockOptin;
  // CRC tables: tab[j] = x^j mod g, left-aligned like the reference's crc values (crc_byte.c:46-58)
  std::vector<uint32_t> t(kCrcTableLen);
  for (int pi = 0; pi < 8; pi++) {
    uint32_t v = 0x80000000u;  // x^(deg-1) ... we want x^0 left-aligned at the polynomial's degree
    const int deg = pi <= 2 ? 24 : pi == 3 ? 16 : pi == 4 ? 12 : pi == 5 ? 11 : pi == 6 ? 8 : 6;
    v = 1u << (32 - deg);      // x^0 as a left-aligned deg-bit remainder
    for (int j = 0; j < kCrcTableLen; j++) {
      t[j] = v;
      const uint32_t top = v & 0x80000000u;
      v <<= 1;
      if (top) v ^= kPoly[pi];
    }
    if (cudaMalloc(&crc_tab[pi], kCrcTableLen * sizeof(uint32_t)) != cudaSuccess) { last_error = "cudaMalloc crc table"; return -1; }
    cudaMemcpy(crc_tab[pi], t.data(), kCrcTableLen * sizeof(uint32_t), cudaMemcpyHostToDevice);
    // long messages (transport-block CRC): shift[k][b] = x^(b + k * kCrcChunk) mod g, so a chunk's remainder can be moved to its place
    std::vector<uint32_t> sh((size_t)kCrcMaxChunks * 32, 0u);
    v = 1u << (32 - deg);
    for (int k = 0; k < kCrcMaxChunks; k++) {
      uint32_t u = v;
      for (int b = 0; b < deg; b++) {
        sh[(size_t)k * 32 + b] = u;
        const uint32_t top = u & 0x80000000u;
        u <<= 1;
        if (top) u ^= kPoly[pi];
      }
      for (int j = 0; j < kCrcChunk; j++) { const uint32_t top = v & 0x80000000u; v <<= 1; if (top) v ^= kPoly[pi]; }
    }
    if (cudaMalloc(&crc_shift[pi], sh.size() * sizeof(uint32_t)) != cudaSuccess) { last_error = "cudaMalloc crc shift table"; return -1; }
    cudaMemcpy(crc_shift[pi], sh.data(), sh.size() * sizeof(uint32_t), cudaMemcpyHostToDevice);
  }
  inited = true;
  return 0;
}

void Ctx::shutdown()
{
  std::lock_guard<std::mutex> lk(mu);
  if (!inited) return;
  cudaDeviceSynchronize();
  packed_graph_cache_clear();
  for (auto &kv : graphs) cudaFree(kv.second);
  for (auto &kv : enc_graphs) cudaFree(kv.second);
  graphs.clear(); graphs_host.clear(); enc_graphs.clear(); enc_graphs_host.clear();
  for (auto *w : pool) {
    if (w->stream) cudaStreamDestroy(w->stream);
    cudaFree(w->d_in); cudaFree(w->d_out); cudaFree(w->d_aux);
    cudaFreeHost(w->h_in); cudaFreeHost(w->h_out); cudaFreeHost(w->h_aux);
    delete w;
  }
  pool.clear();
  for (auto &p : crc_tab) { cudaFree(p); p = nullptr; }
  for (auto &p : crc_shift) { cudaFree(p); p = nullptr; }
  inited = false;
}

const GraphDev *Ctx::graph(int BG, int Z, int R, const GraphDev **host)
{
  if (BG < 1 || BG > 2 || Z < 2 || Z > 384 || R < 0 || R > 255) return nullptr;
  const uint32_t key = ((uint32_t)BG << 24) | ((uint32_t)Z << 8) | (uint32_t)R;
  std::lock_guard<std::mutex> lk(mu);
  auto it = graphs.find(key);
  if (it == graphs.end()) {
    GraphDev g;
    if (!build_graph(BG, Z, R, &g)) return nullptr;
    GraphDev *d = nullptr;
    if (cudaMalloc(&d, sizeof(GraphDev)) != cudaSuccess) return nullptr;
    cudaMemcpy(d, &g, sizeof(GraphDev), cudaMemcpyHostToDevice);
    graphs_host[key] = g;
    it = graphs.emplace(key, d).first;
  }
  if (host) *host = &graphs_host[key];
  return it->second;
}

const EncGraphDev *Ctx::enc_graph(int BG, int Z, const EncGraphDev **host)
{
  if (BG < 1 || BG > 2 || Z < 2 || Z > 384) return nullptr;
  const uint32_t key = ((uint32_t)BG << 16) | (uint32_t)Z;
  std::lock_guard<std::mutex> lk(mu);
  auto it = enc_graphs.find(key);
  if (it == enc_graphs.end()) {
    EncGraphDev g;
    if (!build_enc_graph(BG, Z, &g)) return nullptr;
    EncGraphDev *d = nullptr;
    if (cudaMalloc(&d, sizeof(EncGraphDev)) != cudaSuccess) return nullptr;
    cudaMemcpy(d, &g, sizeof(EncGraphDev), cudaMemcpyHostToDevice);
    enc_graphs_host[key] = g;
    it = enc_graphs.emplace(key, d).first;
  }
  if (host) *host = &enc_graphs_host[key];
  return it->second;
}

// want_buffers: hand out the pooled workspace with the largest staging buffers (a caller that will reserve()); otherwise the one with the
// smallest (a caller that only needs the stream).  Keeps the roles stable when several batches are in flight, so no call re-allocates
// pinned memory in steady state.
Workspace *Ctx::acquire(bool want_buffers)
{
  {
    std::lock_guard<std::mutex> lk(mu);
    if (!pool.empty()) {
      size_t best = 0;
      for (size_t i = 1; i < pool.size(); i++)
        if (want_buffers ? pool[i]->cap_in > pool[best]->cap_in : pool[i]->cap_in < pool[best]->cap_in) best = i;
      Workspace *w = pool[best];
      pool.erase(pool.begin() + (long)best);
      return w;
    }
  }
  Workspace *w = new Workspace();
  if (cudaStreamCreateWithFlags(&w->stream, cudaStreamNonBlocking) != cudaSuccess) { delete w; return nullptr; }
  return w;
}

void Ctx::release(Workspace *w)
{
  std::lock_guard<std::mutex> lk(mu);
  pool.push_back(w);
}

}  // namespace nrb200
