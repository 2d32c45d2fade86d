// Per-device CUDA context of libldpc_b200.so: lifted-graph cache, CRC tables, workspace pool.  One process can drive every visible GPU:
// there is one Ctx per device and ctx() returns the one of the calling thread's CURRENT device (nrb200_set_device(); default NRB200_DEVICE /
// LOCAL_RANK / 0), so all per-device state behind the entry points follows the thread's selection.
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <cstdint>
#include <map>
#include <mutex>
#include <string>
#include <vector>
#include "nrb200_graph.h"

namespace nrb200 {

// A stream plus grow-only device and pinned-host staging buffers; one is checked out per in-flight host call so
// concurrent callers (OAI's tpool workers, SURVEY.md section 8b "Threading") never share a stream.
struct Workspace {
  cudaStream_t stream = nullptr;
  void *d_in = nullptr, *d_out = nullptr, *d_aux = nullptr;
  void *h_in = nullptr, *h_out = nullptr, *h_aux = nullptr;
  size_t cap_in = 0, cap_out = 0, cap_aux = 0;
  bool reserve(size_t in, size_t out, size_t aux);
};

struct Ctx {
  std::mutex mu;
  std::atomic<bool> inited{false};             // read without the mutex by every entry point
  int dev = -1;
  int sm_count = 0;
  int max_smem_optin = 0;
  std::map<uint32_t, GraphDev *> graphs;        // key BG<<24 | Z<<8 | R -> device pointer
  std::map<uint32_t, GraphDev> graphs_host;
  std::map<uint32_t, EncGraphDev *> enc_graphs;  // key BG<<16 | Z
  std::map<uint32_t, EncGraphDev> enc_graphs_host;
  uint32_t *crc_tab[8] = {nullptr};              // device: x^j mod g for the 8 polynomials
  uint32_t *crc_shift[8] = {nullptr};            // device: x^(b + k * kCrcChunk) mod g, [k][32] (long-message CRC)
  std::vector<Workspace *> pool;
  std::atomic<uint64_t> launches{0};
  std::string last_error;

  int init();                                    // 0 ok, -1 no usable device
  void shutdown();
  const GraphDev *graph(int BG, int Z, int R, const GraphDev **host = nullptr);   // nullptr if invalid
  const EncGraphDev *enc_graph(int BG, int Z, const EncGraphDev **host = nullptr);
  Workspace *acquire(bool want_buffers = true);
  void release(Workspace *w);
  void set_error(const char *where, cudaError_t e);
};

constexpr int kMaxDevices = 16;
Ctx &ctx();                 // the calling thread's current device
int current_device();       // its index
int set_current_device(int dev);   // 0, or -1 if there is no such device
int device_count();         // visible CUDA devices (0 without a driver)

#define NRB200_CUDA_OK(call, where)                                  \
  do {                                                               \
    cudaError_t e__ = (call);                                        \
    if (e__ != cudaSuccess) { ctx().set_error(where, e__); return -2; } \
  } while (0)

}  // namespace nrb200
