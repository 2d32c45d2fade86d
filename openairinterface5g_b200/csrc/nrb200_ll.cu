// Low-latency per-call path behind LDPCdecoder (include/nrb200_ldpc.h part 1): what makes the four-symbol drop-in itself fast when
// unmodified OAI host code calls it once per segment from its tpool workers (nr_ulsch_decoding.c:435-468) or in ldpctest's serial loop
// (ldpctest.c:329-340).
//
//  * staging rows in MAPPED pinned host memory: the caller's LLRs are copied into a row (26 KB memcpy), the kernel reads the row over PCIe
//    itself and stores hard bits, iteration count and a completion byte back into the row; the calling thread spins on that byte.  No copy
//    engine, no stream synchronisation, no event: one kernel launch is the only driver call on the path.
//  * combining: callers queue their row and then take the launch lock; whoever holds it launches EVERYTHING queued at that moment as one
//    kernel (one cluster per code block), so callers that arrive while a launch is being issued ride together on the next one.  No timer,
//    no helper thread: a lone caller never waits for company.
//  * the kernel is the cluster decoder (ldpc_decoder_cluster.cuh) whenever the lifting size allows, so one block takes a few tens of
//    microseconds instead of ~100 us on a single SM.
//  * decode_abort_t is re-read by the waiting thread while it spins and mirrored into the row, where the kernel polls it once per iteration
//    (the reference's check_abort at the top of every iteration, nrLDPC_decoder.c:557-560).
#include "../../include/nrb200_ldpc.h"
#include "nrb200_ctx.h"
#include "ldpc_common.cuh"
#include <atomic>
#include <chrono>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace nrb200 {

int launch_decode(const GraphDev *d_g, const GraphDev &h_g, const DecodeArgs &a, cudaStream_t stream);

namespace {

constexpr int kRows = 128;                       // staging rows = calls in flight
constexpr size_t kInStride = 26624;              // >= 68 * 384 LLRs, 512-byte multiple
constexpr size_t kOutStride = 26624;             // one-bit-per-byte output modes need 68 * 384 bytes
constexpr int kStreams = 16;

struct Req {
  int row = -1;
  uint64_t sig = 0;                              // everything a launch must share: (BG, Z, R, numMaxIter, outMode, use_crc, crc_type, crc_len_bits)
  const GraphDev *dg = nullptr, *hg = nullptr;
  DecodeArgs a0;
  std::atomic<int> launched{0};                  // 0 queued, 1 launched (C valid), -1 launch failed
  int C = 1;
};

struct Pool {
  std::mutex mu;                                 // rows + queue
  std::mutex launch_mu;                          // one launcher at a time: this is where callers pile up and get combined
  bool ok = false, tried = false;
  int8_t *h_in = nullptr, *d_in = nullptr;
  uint8_t *h_out = nullptr, *d_out = nullptr;
  LlCtrl *h_ctrl = nullptr, *d_ctrl = nullptr;
  std::vector<int> free_rows;
  std::vector<Req *> queue;
  cudaStream_t streams[kStreams] = {nullptr};
  unsigned next_stream = 0;
  std::atomic<uint64_t> launches{0}, blocks{0};

  bool init()
  {
    std::lock_guard<std::mutex> lk(mu);
    if (tried) return ok;
    tried = true;
    void *p = nullptr;
    const size_t bytes = kRows * (kInStride + kOutStride) + kRows * sizeof(LlCtrl);
    if (cudaHostAlloc(&p, bytes, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return false; }
    void *d = nullptr;
    if (cudaHostGetDevicePointer(&d, p, 0) != cudaSuccess) { cudaFreeHost(p); cudaGetLastError(); return false; }
    std::memset(p, 0, bytes);
    h_in = (int8_t *)p; d_in = (int8_t *)d;
    h_out = (uint8_t *)p + kRows * kInStride; d_out = (uint8_t *)d + kRows * kInStride;
    h_ctrl = (LlCtrl *)((uint8_t *)p + kRows * (kInStride + kOutStride)); d_ctrl = (LlCtrl *)((uint8_t *)d + kRows * (kInStride + kOutStride));
    for (int i = 0; i < kStreams; i++)
      if (cudaStreamCreateWithFlags(&streams[i], cudaStreamNonBlocking) != cudaSuccess) return false;
    for (int i = kRows - 1; i >= 0; i--) free_rows.push_back(i);
    ok = true;
    return true;
  }
};

Pool &pool()
{
  static Pool p;
  return p;
}

inline void cpu_relax()
{
#if defined(__x86_64__)
  _mm_pause();
#else
  std::this_thread::yield();
#endif
}

// Launch everything that is queued right now: one kernel per group of requests that share a signature (at most kLlMaxBatch rows each).
void launch_queued(Pool &P)
{
  std::vector<Req *> q;
  {
    std::lock_guard<std::mutex> lk(P.mu);
    q.swap(P.queue);
  }
  while (!q.empty()) {
    Req *lead = q.front();
    std::vector<Req *> grp, rest;
    for (Req *r : q) ((r->sig == lead->sig && (int)grp.size() < kLlMaxBatch) ? grp : rest).push_back(r);
    DecodeArgs a = lead->a0;
    a.n_cb = (uint32_t)grp.size();
    a.llr = P.d_in; a.llr_stride = (uint32_t)kInStride;
    a.out = P.d_out; a.out_stride = (uint32_t)kOutStride;
    a.iters = nullptr; a.abort_flags = nullptr;
    a.ll_ctrl = P.d_ctrl; a.ll_seq = 1;
    for (size_t i = 0; i < grp.size(); i++) a.ll_rows[i] = (uint16_t)grp[i]->row;
    cudaStream_t st = P.streams[P.next_stream++ % kStreams];
    int C = 1;
    const int rc = launch_decode_ll(lead->dg, *lead->hg, a, st, &C);
    for (Req *r : grp) { r->C = C; r->launched.store(rc == 0 ? 1 : -1, std::memory_order_release); }
    P.launches++; P.blocks += grp.size();
    q.swap(rest);
  }
}

}  // namespace

// One blocking decode of one code block through the low-latency path.  `a0` carries the per-call decode parameters (fill_args), llr / out are
// the caller's buffers (any memory).  Returns 0 and *iters, or a negative error (-1 no device / pool, -2 CUDA error, -6 timeout).
int ll_decode_one(const GraphDev *dg, const GraphDev *hg, const DecodeArgs &a0, uint64_t sig, size_t in_bytes, size_t out_bytes, const int8_t *llr,
                  uint8_t *out, int32_t *iters, nrb200_decode_abort_t *ab)
{
  Pool &P = pool();
  if (!P.ok && !P.init()) return -1;
  if (in_bytes > kInStride || out_bytes > kOutStride) return -4;
  Req req;
  req.sig = sig; req.dg = dg; req.hg = hg; req.a0 = a0;
  // ---- a staging row
  for (unsigned spins = 0;; spins++) {
    {
      std::lock_guard<std::mutex> lk(P.mu);
      if (!P.free_rows.empty()) { req.row = P.free_rows.back(); P.free_rows.pop_back(); break; }
    }
    if (spins > 64) std::this_thread::yield(); else cpu_relax();
  }
  int8_t *h_in = P.h_in + (size_t)req.row * kInStride;
  uint8_t *h_out = P.h_out + (size_t)req.row * kOutStride;
  LlCtrl *ctrl = P.h_ctrl + req.row;
  std::memcpy(h_in, llr, in_bytes);
  if (a0.use_crc) std::memcpy(h_out, out, out_bytes);          // the reference leaves p_out untouched until a CRC check has run
  std::memset(ctrl->done, 0, sizeof(ctrl->done));
  ctrl->iters = 0;
  ctrl->abort = (ab && __atomic_load_n(&ab->failed, __ATOMIC_RELAXED)) ? 1 : 0;
  std::atomic_thread_fence(std::memory_order_release);
  // ---- queue, then combine at the launch lock
  {
    std::lock_guard<std::mutex> lk(P.mu);
    P.queue.push_back(&req);
  }
  if (req.launched.load(std::memory_order_acquire) == 0) {
    std::lock_guard<std::mutex> lk(P.launch_mu);
    if (req.launched.load(std::memory_order_acquire) == 0) launch_queued(P);   // mine and everybody's who queued while I waited for the lock
  }
  int rc = 0;
  int st;
  while ((st = req.launched.load(std::memory_order_acquire)) == 0) cpu_relax();
  if (st < 0) rc = -2;
  else {
    // ---- wait for the cluster's completion bytes (one per CTA); mirror the abort flag while waiting
    const uint64_t mask = req.C >= 8 ? ~0ull : ((1ull << (8 * req.C)) - 1ull);
    const uint64_t want = 0x0101010101010101ull & mask;
    const volatile uint64_t *done = reinterpret_cast<const volatile uint64_t *>(ctrl->done);
    const auto t0 = std::chrono::steady_clock::now();
    for (unsigned spins = 1;; spins++) {
      if ((*done & mask) == want) break;
      cpu_relax();
      if ((spins & 63u) == 0) {
        if (ab && !ctrl->abort && __atomic_load_n(&ab->failed, __ATOMIC_RELAXED)) *reinterpret_cast<volatile uint8_t *>(&ctrl->abort) = 1;
        if ((spins & 0xFFFFu) == 0 && std::chrono::steady_clock::now() - t0 > std::chrono::seconds(10)) { rc = -6; break; }
      }
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    if (rc == 0) {
      *iters = ctrl->iters;
      std::memcpy(out, h_out, out_bytes);
    } else {
      ctx().set_error("low-latency decode: no completion within 10 s", cudaPeekAtLastError());
    }
  }
  if (rc != -6) {                                                // a row whose kernel may still be running is never handed out again
    std::lock_guard<std::mutex> lk(P.mu);
    P.free_rows.push_back(req.row);
  }
  return rc;
}

void ll_stats(uint64_t *launches, uint64_t *blocks)
{
  *launches = pool().launches.load();
  *blocks = pool().blocks.load();
}

}  // namespace nrb200
