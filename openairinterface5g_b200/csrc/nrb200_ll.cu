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

int launch_encode(const EncGraphDev *d_g, const EncGraphDev &h_g, int K, uint32_t n_cb, const uint8_t *d_in, uint32_t in_stride,
                  uint8_t *d_out, uint32_t out_stride, cudaStream_t stream, uint8_t *done);
int launch_decode_ll(const GraphDev *d_g, const GraphDev &h_g, const DecodeArgs &a, cudaStream_t stream, uint32_t load, int *cluster_out);

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
  std::atomic<unsigned> next_stream{0};
  std::atomic<uint64_t> launches{0}, blocks{0};
  std::atomic<int> in_flight{0};               // code blocks launched and not yet seen complete
  std::atomic<uint64_t> ns_stage{0}, ns_launch{0}, ns_wait{0}, ns_out{0}, ns_device{0};   // where a call's time goes (sums over all calls, nanoseconds)

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

Pool &pool()                                     // one per device: the streams belong to it
{
  static Pool p[kMaxDevices];
  return p[current_device()];
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
    const int load = P.in_flight.fetch_add((int)grp.size()) + (int)grp.size();
    // callers keep arriving: size the clusters as if half as many blocks again were in flight, so the next launch still finds its SMs free
    const int rc = launch_decode_ll(lead->dg, *lead->hg, a, st, (uint32_t)(load + load / 2), &C);
    if (rc != 0) P.in_flight.fetch_sub((int)grp.size());
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
  auto now_ns = []() { return (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const uint64_t t_start = now_ns();
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
  const uint64_t t_staged = now_ns();
  // ---- queue, then combine at the launch lock
  {
    std::lock_guard<std::mutex> lk(P.mu);
    P.queue.push_back(&req);
  }
  if (req.launched.load(std::memory_order_acquire) == 0) {
    std::lock_guard<std::mutex> lk(P.launch_mu);                               // blocks (sleeps) while another caller launches: no CPU burnt
    if (req.launched.load(std::memory_order_acquire) == 0) launch_queued(P);   // mine and everybody's who queued while I waited for the lock
  }
  int rc = 0;
  int st;
  for (unsigned spins = 0; (st = req.launched.load(std::memory_order_acquire)) == 0; spins++) { if (spins > 2000) std::this_thread::yield(); else cpu_relax(); }
  const uint64_t t_launched = now_ns();
  uint64_t t_done = t_launched;
  if (st < 0) rc = -2;
  else {
    // ---- wait for the cluster's completion bytes (one per CTA); mirror the abort flag while waiting
    const uint64_t mask = req.C >= 8 ? ~0ull : ((1ull << (8 * req.C)) - 1ull);
    const uint64_t want = 0x0101010101010101ull & mask;
    const volatile uint64_t *done = reinterpret_cast<const volatile uint64_t *>(ctrl->done);
    const auto t0 = std::chrono::steady_clock::now();
    for (unsigned spins = 1;; spins++) {
      if ((*done & mask) == want) break;
      if (spins > 4000) std::this_thread::yield(); else cpu_relax();   // more callers than cores: let a runnable thread (a launcher) have this one
      if ((spins & 63u) == 0) {
        if (ab && !ctrl->abort && __atomic_load_n(&ab->failed, __ATOMIC_RELAXED)) *reinterpret_cast<volatile uint8_t *>(&ctrl->abort) = 1;
        if ((spins & 0xFFFFu) == 0 && std::chrono::steady_clock::now() - t0 > std::chrono::seconds(10)) { rc = -6; break; }
      }
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    t_done = now_ns();
    P.in_flight.fetch_sub(1);
    if (rc == 0) {
      P.ns_device += ctrl->t_end - ctrl->t_begin;
      *iters = ctrl->iters;
      std::memcpy(out, h_out, out_bytes);
    } else {
      ctx().set_error("low-latency decode: no completion within 10 s", cudaPeekAtLastError());
    }
  }
  P.ns_stage += t_staged - t_start; P.ns_launch += t_launched - t_staged; P.ns_wait += t_done - t_launched; P.ns_out += now_ns() - t_done;
  if (rc != -6) {                                                // a row whose kernel may still be running is never handed out again
    std::lock_guard<std::mutex> lk(P.mu);
    P.free_rows.push_back(req.row);
  }
  return rc;
}

// ------------------------------------------------------------------------------------------ encoder
// LDPCencoder's up to 8 segments per call (ldpc_encoder_optim8segmulti.c:62-63): payloads in, one-bit-per-byte code words out, through a slot of
// mapped pinned memory; one completion byte per segment.  No combining: a call already carries a group of segments.
namespace {
constexpr int kEncSlots = 24, kEncSegs = 8;
constexpr size_t kEncIn = 1088, kEncOut = 25344;                  // K / 8 <= 1056 bytes, 66 * 384 code bits per segment
struct EncPool {
  std::mutex mu;
  bool ok = false, tried = false;
  uint8_t *h = nullptr, *d = nullptr;
  std::vector<int> free_slots;
  static constexpr size_t kSlot = kEncSegs * (kEncIn + kEncOut) + 64;
  bool init()
  {
    std::lock_guard<std::mutex> lk(mu);
    if (tried) return ok;
    tried = true;
    void *p = nullptr, *dp = nullptr;
    if (cudaHostAlloc(&p, kEncSlots * kSlot, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return false; }
    if (cudaHostGetDevicePointer(&dp, p, 0) != cudaSuccess) { cudaFreeHost(p); cudaGetLastError(); return false; }
    std::memset(p, 0, kEncSlots * kSlot);
    h = (uint8_t *)p; d = (uint8_t *)dp;
    for (int i = kEncSlots - 1; i >= 0; i--) free_slots.push_back(i);
    return ok = true;
  }
};
EncPool &enc_pool() { static EncPool p[kMaxDevices]; return p[current_device()]; }
}  // namespace

int ll_encode(const EncGraphDev *dg, const EncGraphDev *hg, int K, uint32_t n, uint8_t **input, uint8_t **output, uint32_t kin, uint32_t nout)
{
  Pool &P = pool();
  EncPool &E = enc_pool();
  if ((!P.ok && !P.init()) || (!E.ok && !E.init())) return -1;
  if (n == 0) return 0;
  if (n > (uint32_t)kEncSegs || kin > kEncIn - 8 || nout > kEncOut) return -4;
  int slot = -1;
  for (unsigned spins = 0;; spins++) {
    {
      std::lock_guard<std::mutex> lk(E.mu);
      if (!E.free_slots.empty()) { slot = E.free_slots.back(); E.free_slots.pop_back(); break; }
    }
    if (spins > 64) std::this_thread::yield(); else cpu_relax();
  }
  const size_t base = (size_t)slot * EncPool::kSlot;
  uint8_t *h_in = E.h + base, *h_out = h_in + kEncSegs * kEncIn, *h_done = h_out + kEncSegs * kEncOut;
  for (uint32_t j = 0; j < n; j++) { std::memcpy(h_in + j * kEncIn, input[j], kin); std::memset(h_in + j * kEncIn + kin, 0, 8); }
  std::memset(h_done, 0, 8);
  std::atomic_thread_fence(std::memory_order_release);
  cudaStream_t st = P.streams[P.next_stream++ % kStreams];
  int rc = launch_encode(dg, *hg, K, n, E.d + base, (uint32_t)kEncIn, E.d + base + kEncSegs * kEncIn, (uint32_t)kEncOut, st, E.d + base + kEncSegs * (kEncIn + kEncOut));
  if (rc == 0) {
    const uint64_t mask = n >= 8 ? ~0ull : ((1ull << (8 * n)) - 1ull), want = 0x0101010101010101ull & mask;
    const volatile uint64_t *done = reinterpret_cast<const volatile uint64_t *>(h_done);
    const auto t0 = std::chrono::steady_clock::now();
    for (unsigned spins = 1;; spins++) {
      if ((*done & mask) == want) break;
      if (spins > 4000) std::this_thread::yield(); else cpu_relax();
      if ((spins & 0xFFFFu) == 0 && std::chrono::steady_clock::now() - t0 > std::chrono::seconds(10)) { rc = -6; break; }
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    if (rc == 0) for (uint32_t j = 0; j < n; j++) std::memcpy(output[j], h_out + j * kEncOut, nout);
  }
  if (rc != -6) {
    std::lock_guard<std::mutex> lk(E.mu);
    E.free_slots.push_back(slot);
  }
  return rc;
}

// called from LDPCinit: the pinned rows and the streams exist before the first decode call
int ll_warm() { return (pool().ok || pool().init()) && (enc_pool().ok || enc_pool().init()) ? 0 : -1; }

void ll_stats(uint64_t *launches, uint64_t *blocks)
{
  *launches = pool().launches.load();
  *blocks = pool().blocks.load();
}

// sums over all calls so far, nanoseconds: staging (row + memcpy in), queue + launch, wait for the kernel's completion bytes, copy out,
// and the device-side time of the block (%globaltimer, first instruction of CTA 0 to its completion byte)
void ll_timing(uint64_t out[5])
{
  Pool &P = pool();
  out[0] = P.ns_stage.load(); out[1] = P.ns_launch.load(); out[2] = P.ns_wait.load(); out[3] = P.ns_out.load(); out[4] = P.ns_device.load();
}

}  // namespace nrb200
