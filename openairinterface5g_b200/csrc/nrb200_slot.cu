// Slot-level entry points (include/nrb200_slot.h): the per-stage launches of one PUSCH / PDSCH slot issued back to back on the caller's stream.
// The sequencing is the reference's (nr_ulsch_demodulation.c:1447-1700 -> nr_ulsch_decoding.c:320-470 -> phy_procedures_nr_gNB.c:271-300 on the gNB,
// phy_procedures_nr_ue.c:520-760 on the UE, nr_dlsch.c:56-583 + nr_ru_procedures.c:55-140 for the transmitter); every stage is the entry point its
// stand-alone tests pin against the oracle, so the chain adds no arithmetic of its own.
#include "../../include/nrb200_slot.h"
#include "nrb200_ctx.h"
#include <cstring>
#include <algorithm>
#include <atomic>
#include <thread>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <mutex>

#define NRB200_EXPORT extern "C" __attribute__((visibility("default")))
using namespace nrb200;

// the slot-level OFDM front end is compiled into this library too (dfts_internal.cu), with hidden visibility
extern "C" int32_t nrb200_ofdm_mod_slot_dev(const nrb200_ofdm_slot_t *d, const int16_t *d_txdataF, int16_t *d_txdata, void *stream);
extern "C" int32_t nrb200_ofdm_demod_slot_dev(const nrb200_ofdm_slot_t *d, const int16_t *d_rxdata, const int16_t *d_timeshift, int16_t *d_rxdataF, void *stream);

NRB200_EXPORT int32_t nrb200_sch_slot_rx_dev(const nrb200_sch_rx_slot_t *d, const nrb200_sch_rx_bufs_t *b, void *stream)
{
  if (!d || !b) return -4;
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  // ---- nr_fep_full / nr_slot_fep: FFT windows -> rxdataF (rotation + time shift fused into the transform)
  if ((rc = nrb200_ofdm_demod_slot_dev(&d->ofdm, b->d_rxdata, b->d_timeshift, b->d_rxdataF, stream)) != 0) return rc;
  ctx().launches++;
  // ---- channel estimation on every DMRS port of the PDU
  nrb200_pusch_rx_t rx = d->rx;
  if (!d->use_estimates) {
    if ((rc = nrb200_pusch_chest_dev(&d->chest, b->d_rxdataF, b->d_est, b->d_chest_scratch, b->d_chest_state, stream)) != 0) return rc;
    if (rx.pdsch_ue && rx.nrOfLayers > 2) {
      // layers 3 and 4: DMRS ports port + 2 (and + 3) of the second CDM group, one more estimator call (the UE calls nr_pdsch_channel_estimation once per port,
      // phy_procedures_nr_ue.c); its planes and states go behind those of the first two ports, the scratch is reused in stream order
      nrb200_pusch_chest_t c2 = d->chest;
      c2.port = d->chest.port + 2; c2.n_ports = rx.nrOfLayers - 2;
      if ((rc = nrb200_pusch_chest_dev(&c2, b->d_rxdataF, b->d_est + (size_t)2 * 2 * rx.nb_rx * c2.ch_stride, b->d_chest_scratch, b->d_chest_state + 2 * 18, stream)) != 0)
        return rc;
    }
    if (rx.nrOfLayers == 2 && !rx.pdsch_ue) {   // the MMSE receiver takes max_ch / nvar from the estimator's state on the device (:1470-1524)
      rx.d_est_state = (uint64_t)(uintptr_t)b->d_chest_state;
      rx.est_state_ports = 2;
    }
  }
  // ---- level measurement + inner receiver (compensation / MMSE / zero forcing, LLRs, layer de-mapping, unscrambling)
  if ((rc = nrb200_pusch_log2_maxh_dev(&rx, b->d_est, b->d_level, stream)) != 0) return rc;
  if ((rc = nrb200_pusch_inner_rx_dev(&rx, b->d_rxdataF, b->d_est, b->d_level + 8, b->d_llr16, stream)) != 0) return rc;
  // ---- de-interleaving + rate recovery + HARQ combining + decoder-input packing
  if ((rc = nrb200_ldpc_rm_rx_batch_dev(&d->rm, b->d_llr16, b->d_E, b->d_Eoff, b->d_harq, b->harq_stride, b->d_llr8, b->llr8_stride, stream)) != 0) return rc;
  // ---- LDPC decode with the per-segment CRC stop
  nrb200_ldpc_batch_desc_t dd;
  std::memset(&dd, 0, sizeof(dd));
  dd.BG = d->rm.BG; dd.Z = d->rm.Z; dd.R = d->R; dd.numMaxIter = d->numMaxIter; dd.outMode = NRB200_OUTMODE_BIT;
  dd.use_crc = 1; dd.crc_type = (uint8_t)d->seg_crc_type; dd.crc_len_bits = d->crc_len_bits; dd.latency_mode = d->latency_mode;
  dd.n_cb = d->rm.n_seg; dd.llr_stride = b->llr8_stride; dd.out_stride = b->hard_stride;
  if ((rc = nrb200_ldpc_decode_batch_dev(&dd, b->d_llr8, b->d_hard, b->d_iters, stream)) != 0) return rc;
  // ---- nr_postDecode: the segments' payload bytes back to back, then the transport block's own CRC
  if (cudaMemcpy2DAsync(b->d_tb, d->seg_payload_bytes, b->d_hard, b->hard_stride, d->seg_payload_bytes, d->rm.n_seg, cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
    ctx().set_error("slot rx: transport block assembly", cudaGetLastError());
    return -2;
  }
  const uint32_t tb_bits = d->A + d->tb_crc_bits;
  return nrb200_crc_batch_dev(d->tb_crc_bits == 24 ? 0 : 3, 1, b->d_tb, (tb_bits + 7) / 8, tb_bits, b->d_tbcrc, stream);
}

NRB200_EXPORT int32_t nrb200_pdsch_slot_tx_dev(const nrb200_pdsch_tx_slot_t *d, const nrb200_pdsch_tx_bufs_t *b, void *stream)
{
  if (!d || !b) return -4;
  int rc;
  // ---- nr_dlsch_encoding: TB CRC + segmentation + per-segment CRC, LDPC encode, rate matching + interleaving
  if ((rc = nrb200_tb_segment_dev(d->rm.BG, d->A, b->d_payload, b->d_segs, b->seg_stride, b->d_seg_scratch, stream)) != 0) return rc;
  if ((rc = nrb200_ldpc_encode_batch_dev(d->rm.BG, d->rm.Z, (int)d->K, d->rm.n_seg, b->d_segs, b->seg_stride, b->d_cw, b->cw_stride, stream)) != 0) return rc;
  if ((rc = nrb200_ldpc_rm_tx_batch_dev(&d->rm, b->d_cw, b->cw_stride, b->d_E, b->d_Eoff, b->d_f, stream)) != 0) return rc;
  // ---- nr_generate_pdsch after the encoder, in one launch: scrambling, modulation, layer mapping, DMRS, resource mapping, precoding
  if ((rc = nrb200_pdsch_tx_slot_dev(&d->tx, b->d_f, b->d_txdataF, stream)) != 0) return rc;
  // ---- nr_feptx0: rotation + IDFT + cyclic prefix
  if ((rc = nrb200_ofdm_mod_slot_dev(&d->ofdm, b->d_txdataF, b->d_txdata, stream)) != 0) return rc;
  ctx().launches++;
  return 0;
}

// ------------------------------------------------------------------------------------------ transport-block level, host buffers (nr_ulsch_decoding)
namespace {
// segments of all transport blocks currently inside nrb200_ulsch_decode_tb_host: a lone block gets a cluster of SMs per segment (shortest time to its result);
// once more segments are in flight than the GPU has room for clusters, every segment takes one SM and the fewest SM-cycles (most blocks per second)
std::atomic<int> g_tb_segments_in_flight{0};
struct InFlight {
  int n, total;
  explicit InFlight(int n_) : n(n_), total(g_tb_segments_in_flight.fetch_add(n_) + n_) {}
  ~InFlight() { g_tb_segments_in_flight.fetch_sub(n); }
};
struct HarqBuf { int16_t *d = nullptr; size_t elems = 0; int dev = 0; };
std::mutex g_harq_mu;
std::map<uint64_t, HarqBuf> g_harq;

// the TB's soft buffers: n_seg x stride int16 on the calling thread's device, created zeroed; a key that changes shape (a reconfigured HARQ process) starts over
int16_t *harq_buffers(uint64_t key, size_t elems, bool *fresh)
{
  std::lock_guard<std::mutex> lk(g_harq_mu);
  HarqBuf &h = g_harq[key];
  *fresh = false;
  if (h.d != nullptr && (h.elems != elems || h.dev != ctx().dev)) {
    cudaSetDevice(h.dev); cudaFree(h.d); cudaSetDevice(ctx().dev);
    h = HarqBuf();
  }
  if (h.d == nullptr) {
    if (cudaMalloc(&h.d, elems * sizeof(int16_t)) != cudaSuccess) { g_harq.erase(key); return nullptr; }
    h.elems = elems; h.dev = ctx().dev; *fresh = true;
  }
  return h.d;
}
}  // namespace

// Wait for a stream without monopolising a core: poll, and after a while yield between polls.  With as many blocking callers as host cores a pure spin
// (cudaStreamSynchronize's default) starves whoever is preparing the next launch.
static cudaError_t polite_sync(cudaStream_t st)
{
  for (unsigned spins = 0;; spins++) {
    const cudaError_t e = cudaStreamQuery(st);
    if (e != cudaErrorNotReady) return e;
    if (spins > 24) std::this_thread::yield();
  }
}

NRB200_EXPORT int32_t nrb200_host_register(void *p, uint64_t bytes)
{
  { Ctx &cx = ctx(); if (!cx.inited && cx.init() != 0) return -1; cudaSetDevice(cx.dev); }
  cudaError_t e = cudaHostRegister(p, (size_t)bytes, cudaHostRegisterPortable);
  if (e == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); return 0; }
  if (e != cudaSuccess) { ctx().set_error("host_register", e); cudaGetLastError(); return -2; }
  return 0;
}
NRB200_EXPORT int32_t nrb200_host_unregister(void *p) { return cudaHostUnregister(p) == cudaSuccess ? 0 : (cudaGetLastError(), -2); }

NRB200_EXPORT int32_t nrb200_ulsch_harq_release(uint64_t harq_key)
{
  std::lock_guard<std::mutex> lk(g_harq_mu);
  for (auto it = g_harq.begin(); it != g_harq.end();) {
    if (harq_key == 0 || it->first == harq_key) { cudaSetDevice(it->second.dev); cudaFree(it->second.d); it = g_harq.erase(it); }
    else ++it;
  }
  return 0;
}

NRB200_EXPORT int32_t nrb200_ulsch_decode_tb_host(const nrb200_ulsch_tb_t *d, const int16_t *ulsch_llr, const uint32_t *E, const uint8_t *R, const uint8_t *clear,
                                                  uint8_t *const *c, int32_t *iters, int16_t *const *d_mirror)
{
  if (!d || !ulsch_llr || !E || !R || !clear || !c || !iters) return -4;
  { Ctx &cx = ctx(); if (!cx.inited && cx.init() != 0) return -1; cudaSetDevice(cx.dev); }
  const uint32_t n = d->rm.n_seg, Z = d->rm.Z;
  if (n == 0 || n != d->rm.C || n > 4 * 36 || (d->rm.BG != 1 && d->rm.BG != 2) || d->rm.K % 8) return -4;
  const uint32_t ncb = (d->rm.BG == 1 ? 66u : 50u) * Z, kc = (d->rm.BG == 1 ? 68u : 52u), llr_stride = (kc * Z + 63u) & ~63u, Kb = d->rm.K / 8;
  const uint32_t hard_stride = (kc * Z / 8 + 63u) & ~63u;
  size_t G = 0;
  for (uint32_t r = 0; r < n; r++) G += E[r];
  static const bool timing = [] { const char *e = getenv("NRB200_TB_TIMING"); return e && *e == '1'; }();
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto us = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::micro>(b - a).count(); };
  const auto t0 = now();
  const InFlight load((int)n);
  const uint8_t latency_mode = load.total <= 74 ? 1 : 0;
  Workspace *w = ctx().acquire();
  // d_in: LLRs | E / offset table;  d_out: hard bits | iteration counts;  d_aux: decoder inputs
  const size_t tab_off = (2 * G + 63) & ~(size_t)63, out_it = (size_t)n * hard_stride;
  // (the optional mirror of the soft buffers travels back through the input staging area once the kernels are done with it)
  const size_t in_bytes = std::max(tab_off + 8 * (size_t)n, d_mirror ? (size_t)n * ncb * 2 : (size_t)0);
  if (!w || !w->reserve(in_bytes, out_it + 4 * (size_t)n, (size_t)n * llr_stride)) { if (w) ctx().release(w); return -5; }
  bool fresh = false;
  int16_t *d_harq = harq_buffers(d->harq_key, (size_t)n * ncb, &fresh);
  if (!d_harq) { ctx().release(w); return -5; }
  int rc = 0;
  cudaStream_t st = w->stream;
  do {
    uint32_t *tab = (uint32_t *)((uint8_t *)w->h_in + tab_off);
    size_t off = 0;
    for (uint32_t r = 0; r < n; r++) { tab[r] = E[r]; tab[n + r] = (uint32_t)off; off += E[r]; }
    if (!d->llr_pinned) std::memcpy(w->h_in, ulsch_llr, 2 * G);
    const auto t1 = now();
    if (d->llr_pinned) {                                                    // the LLRs straight from the caller's page-locked buffer, the small table from the staging area
      if (cudaMemcpyAsync(w->d_in, ulsch_llr, 2 * G, cudaMemcpyHostToDevice, st) != cudaSuccess) { rc = -2; break; }
      if (cudaMemcpyAsync((uint8_t *)w->d_in + tab_off, tab, 8 * (size_t)n, cudaMemcpyHostToDevice, st) != cudaSuccess) { rc = -2; break; }
    } else if (cudaMemcpyAsync(w->d_in, w->h_in, tab_off + 8 * (size_t)n, cudaMemcpyHostToDevice, st) != cudaSuccess) { rc = -2; break; }
    if (fresh && cudaMemsetAsync(d_harq, 0, (size_t)n * ncb * 2, st) != cudaSuccess) { rc = -2; break; }
    const uint32_t *d_E = (const uint32_t *)((uint8_t *)w->d_in + tab_off), *d_off = d_E + n;
    // rate recovery: runs of segments with the same d_to_be_cleared flag (uniform in practice: one launch)
    for (uint32_t r0 = 0; r0 < n && rc == 0;) {
      uint32_t r1 = r0 + 1;
      while (r1 < n && (clear[r1] != 0) == (clear[r0] != 0)) r1++;
      nrb200_rm_desc_t rm = d->rm;
      rm.clear = clear[r0] ? 1 : 0; rm.n_seg = r1 - r0;
      rc = nrb200_ldpc_rm_rx_batch_dev(&rm, (const int16_t *)w->d_in, d_E + r0, d_off + r0, d_harq + (size_t)r0 * ncb, ncb, (int8_t *)w->d_aux + (size_t)r0 * llr_stride,
                                       llr_stride, st);
      r0 = r1;
    }
    if (rc) break;
    // decode: runs of segments with the same rate selector (E differs by one modulation symbol between the first and the last segments at most)
    for (uint32_t r0 = 0; r0 < n && rc == 0;) {
      uint32_t r1 = r0 + 1;
      while (r1 < n && R[r1] == R[r0]) r1++;
      nrb200_ldpc_batch_desc_t dd;
      std::memset(&dd, 0, sizeof(dd));
      dd.BG = d->rm.BG; dd.Z = (uint16_t)Z; dd.R = R[r0]; dd.numMaxIter = (uint8_t)d->numMaxIter; dd.outMode = NRB200_OUTMODE_BIT;
      dd.use_crc = 1; dd.crc_type = (uint8_t)d->crc_type; dd.crc_len_bits = d->crc_len_bits; dd.latency_mode = latency_mode;
      dd.n_cb = r1 - r0; dd.llr_stride = llr_stride; dd.out_stride = hard_stride;
      rc = nrb200_ldpc_decode_batch_dev(&dd, (const int8_t *)w->d_aux + (size_t)r0 * llr_stride, (uint8_t *)w->d_out + (size_t)r0 * hard_stride,
                                        (int32_t *)((uint8_t *)w->d_out + out_it) + r0, st);
      r0 = r1;
    }
    if (rc) break;
    if (cudaMemcpyAsync(w->h_out, w->d_out, out_it + 4 * (size_t)n, cudaMemcpyDeviceToHost, st) != cudaSuccess) { rc = -2; break; }
    const auto t2 = now();
    if (polite_sync(st) != cudaSuccess) { rc = -2; break; }
    const auto t3 = now();
    if (timing) fprintf(stderr, "ulsch_decode_tb_host: %u segments: stage %.1f us, enqueue %.1f us, wait %.1f us\n", n, us(t0, t1), us(t1, t2), us(t2, t3));
    const int32_t *it = (const int32_t *)((const uint8_t *)w->h_out + out_it);
    for (uint32_t r = 0; r < n; r++) {
      iters[r] = it[r];
      if (it[r] <= (int32_t)d->numMaxIter && c[r]) std::memcpy(c[r], (const uint8_t *)w->h_out + (size_t)r * hard_stride, Kb);
    }
    if (d_mirror) {
      if (cudaMemcpyAsync(w->h_in, d_harq, (size_t)n * ncb * 2, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) { rc = -2; break; }
      for (uint32_t r = 0; r < n; r++) if (d_mirror[r]) std::memcpy(d_mirror[r], (const int16_t *)w->h_in + (size_t)r * ncb, (size_t)ncb * 2);
    }
  } while (0);
  if (rc == -2) ctx().set_error("ulsch_decode_tb_host", cudaGetLastError());
  ctx().release(w);
  return rc;
}
