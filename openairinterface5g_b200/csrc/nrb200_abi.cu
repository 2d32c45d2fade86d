// C ABI of libldpc_b200.so (include/nrb200_ldpc.h): the four OAI loader symbols plus the batched extension.
#include "../../include/nrb200_ldpc.h"
#include "../../include/nrb200_rfsim.h"
#include "../../include/nrb200_prach.h"
#include "nrb200_ctx.h"
#include "ldpc_packed_graph.h"
#include "ldpc_common.cuh"
#include <algorithm>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

#define NRB200_EXPORT extern "C" __attribute__((visibility("default")))

namespace nrb200 {
int launch_decode(const GraphDev *d_g, const GraphDev &h_g, const DecodeArgs &a, cudaStream_t stream);
int ll_decode_one(const GraphDev *dg, const GraphDev *hg, const DecodeArgs &a0, uint64_t sig, size_t in_bytes, size_t out_bytes, const int8_t *llr,
                  uint8_t *out, int32_t *iters, nrb200_decode_abort_t *ab);
void ll_stats(uint64_t *launches, uint64_t *blocks);
int ll_warm();
void ll_timing(uint64_t out[5]);
int debug_cluster_marks(long long *out);
int launch_encode(const EncGraphDev *d_g, const EncGraphDev &h_g, int K, uint32_t n_cb, const uint8_t *d_in, uint32_t in_stride,
                  uint8_t *d_out, uint32_t out_stride, cudaStream_t stream, uint8_t *done = nullptr);
int ll_encode(const EncGraphDev *dg, const EncGraphDev *hg, int K, uint32_t n, uint8_t **input, uint8_t **output, uint32_t kin, uint32_t nout);
int launch_crc(int poly_id, uint32_t n_blk, const uint8_t *d_in, uint32_t stride, uint32_t bitlen, uint32_t *d_out, cudaStream_t stream);
int tb_segment_parms(int BG, uint32_t A, uint32_t out[6]);
int launch_tb_segment(int BG, uint32_t A, const uint8_t *d_payload, uint8_t *d_segs, uint32_t seg_stride, uint32_t *d_scratch, cudaStream_t stream);
int quirks_from_env();
int launch_gold(int mode, uint32_t c_init, uint32_t size, const uint8_t *in, uint32_t *out, int16_t *llr, cudaStream_t st);
uint32_t pusch_num_llr(const nrb200_pusch_rx_t &d);
int launch_pusch_level(const nrb200_pusch_rx_t &d, const int16_t *ch, int32_t *d_out9, uint32_t *d_count, cudaStream_t st);
int launch_pusch_rx(const nrb200_pusch_rx_t &d, const int16_t *rxF, const int16_t *ch, const int32_t *d_shift, int16_t *llr, cudaStream_t st);
size_t pusch_chest_scratch_bytes(const nrb200_pusch_chest_t &d);
size_t pusch_tp_scratch_bytes(const nrb200_pusch_rx_t &d);
int pusch_ptrs_layout(const nrb200_pusch_rx_t &d, uint32_t *mask, uint32_t *n_re);
uint32_t prach_num_roots(const nrb200_prach_t &d);
size_t prach_scratch_bytes(const nrb200_prach_t &d);
int launch_prach(const nrb200_prach_t &d, const int16_t *xu, const int16_t *rxsigF, int32_t *out3, void *scratch, cudaStream_t st);
int launch_rfsim(const nrb200_rfsim_chan_t &c, const double *ch, const int16_t *sig, int16_t *out, uint32_t out_stride, uint32_t n, uint64_t TS, uint32_t CirSize,
                 const double *noise, cudaStream_t st);
int lowpapr_sequence_host(uint32_t u, uint32_t v, uint32_t n_re, uint32_t scaling, int16_t *seq);
int launch_chest_time_avg(uint32_t N, uint32_t nb_rx, uint32_t ch_stride, uint32_t start_symbol, uint32_t nr_of_symbols, uint32_t dmrs_symb_pos, uint32_t rb_size,
                          int16_t *d_est, cudaStream_t st);
int pusch_dmrs_pilots_host(const nrb200_pusch_chest_t &d, int16_t *pil);
int launch_pusch_chest(const nrb200_pusch_chest_t &d, const int16_t *rxF, int16_t *est, void *d_scratch, int32_t *d_state, cudaStream_t st, int buf_symbol,
                       int tail_override = 0);
int launch_rm_rx8(const nrb200_rm_desc_t &p, const int8_t *soft, const uint32_t *E, const uint32_t *off, int16_t *harq, uint32_t harq_stride,
                  int8_t *llr, uint32_t llr_stride, cudaStream_t st);
int launch_modulate(int Qm, uint32_t length_bits, const uint8_t *bits, int16_t *out, cudaStream_t st);
uint32_t pdsch_tx_num_bits(const nrb200_pdsch_tx_t &d);
int launch_pdsch_tx(const nrb200_pdsch_tx_t &d, const uint8_t *f, int16_t *txF, cudaStream_t st);
int launch_pusch_llr(int Qm, uint32_t nb_re, const int16_t *y, const int16_t *ma, const int16_t *mb, const int16_t *mc, int16_t *out, cudaStream_t st);
int launch_rm_tx(const nrb200_rm_desc_t &p, const uint8_t *d, uint32_t d_stride, const uint32_t *E, const uint32_t *off, uint8_t *f, cudaStream_t st);
int launch_rm_rx(const nrb200_rm_desc_t &p, const int16_t *soft, const uint32_t *E, const uint32_t *off, int16_t *harq, uint32_t harq_stride,
                 int8_t *llr, uint32_t llr_stride, cudaStream_t st);
}
using namespace nrb200;

static inline unsigned long long rdtsc_now()
{
#if defined(__x86_64__)
  unsigned long long a, d;
  __asm__ volatile("rdtsc" : "=a"(a), "=d"(d));
  return (d << 32) | a;
#else
  return 0;
#endif
}

static int ensure_init()
{
  Ctx &c = ctx();
  if (!c.inited) { if (c.init() != 0) return -1; }
  cudaSetDevice(c.dev);   // the runtime's current device is per host thread
  return 0;
}

static int crc_poly_of_type(int crc_type) { return crc_type == 0 ? 0 : crc_type == 1 ? 1 : crc_type == 2 ? 3 : 6; }

static int fill_args(const nrb200_ldpc_batch_desc_t *d, const GraphDev &hg, DecodeArgs *a)
{
  const uint32_t numLLR = (uint32_t)hg.ncols * hg.Z;
  const uint32_t out_bytes = d->outMode == NRB200_OUTMODE_BIT ? (numLLR + 7) / 8 : numLLR;
  if (d->llr_stride < numLLR || d->out_stride < out_bytes) return -4;
  if (d->outMode > 2) return -4;
  std::memset(a, 0, sizeof(*a));
  a->n_cb = d->n_cb; a->llr_stride = d->llr_stride; a->out_stride = d->out_stride;
  a->numMaxIter = d->numMaxIter; a->outMode = d->outMode; a->use_crc = d->use_crc ? 1 : 0; a->latency = d->latency_mode ? 1 : 0;
  a->quirks = (uint8_t)quirks_from_env();
  if (a->use_crc) {
    if (d->crc_type > 3 || d->crc_len_bits % 8 || d->crc_len_bits < 32 || d->crc_len_bits > numLLR || d->crc_len_bits >= (uint32_t)kCrcTableLen) return -4;
    a->crc_len_bits = d->crc_len_bits;
    a->crc_tab = ctx().crc_tab[crc_poly_of_type(d->crc_type)];
  }
  return 0;
}

// ------------------------------------------------------------------------------------------ part 2: batch API
NRB200_EXPORT int32_t nrb200_ldpc_num_llr(int BG, int Z, int R)
{
  const int nc = ncols_for_rate(BG, R);
  if (nc < 0 || ils_of_z(Z) < 0) return -1;
  return nc * Z;
}

NRB200_EXPORT int32_t nrb200_ldpc_decode_batch_dev(const nrb200_ldpc_batch_desc_t *desc, const int8_t *d_llr, uint8_t *d_out,
                                                   int32_t *d_iters, void *stream)
{
  if (ensure_init()) return -1;
  const GraphDev *hg = nullptr;
  const GraphDev *dg = ctx().graph(desc->BG, desc->Z, desc->R, &hg);
  if (!dg) return -4;
  DecodeArgs a;
  if (int rc = fill_args(desc, *hg, &a)) return rc;
  a.llr = d_llr; a.out = d_out; a.iters = d_iters;
  return launch_decode(dg, *hg, a, (cudaStream_t)stream);
}

static bool is_pinned_host(const void *p)
{
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return at.type == cudaMemoryTypeHost;
}

// Host-buffer decode.  Pageable caller memory (OAI's stack arrays) is staged through the workspace's pinned buffers;
// page-locked caller memory is copied from / to directly.  Large batches are cut into chunks that alternate between two
// streams so the H2D copy of chunk i+1 overlaps the kernel of chunk i.
// The work is split in two halves so that a caller can keep several batches in flight (enqueue / dequeue, the shape of the
// bbdev calls of the reference's T2 offload, nrLDPC_decoder_offload.c:1048-1100): decode_submit() stages, enqueues every copy and
// kernel and returns; decode_finish() waits for the batch's two streams and hands the results over.
struct DecodeTicket {
  Workspace *w = nullptr, *w2 = nullptr;
  uint8_t *out = nullptr;
  int32_t *iters = nullptr;
  size_t n = 0, out_bytes = 0;
  bool pin_out = false;
  int rc = 0;
};

static DecodeTicket *decode_submit(const nrb200_ldpc_batch_desc_t *desc, const int8_t *llr, uint8_t *out, int32_t *iters, const uint8_t *abort_flags,
                                   int *rc_out)
{
  *rc_out = 0;
  if (ensure_init()) { *rc_out = -1; return nullptr; }
  const GraphDev *hg = nullptr;
  const GraphDev *dg = ctx().graph(desc->BG, desc->Z, desc->R, &hg);
  if (!dg) { *rc_out = -4; return nullptr; }
  DecodeArgs a0;
  if (int rc = fill_args(desc, *hg, &a0)) { *rc_out = rc; return nullptr; }
  const size_t n = desc->n_cb;
  DecodeTicket *t = new DecodeTicket();
  t->out = out; t->iters = iters; t->n = n;
  if (n == 0) return t;
  const size_t in_bytes = n * desc->llr_stride, out_bytes = n * desc->out_stride, aux_bytes = n * (sizeof(int32_t) + 1);
  t->out_bytes = out_bytes;
  Workspace *w = t->w = ctx().acquire();
  Workspace *w2 = t->w2 = n >= 64 ? ctx().acquire(false) : nullptr;
  if (!w || !w->reserve(in_bytes, out_bytes, aux_bytes)) { if (w) ctx().release(w); if (w2) ctx().release(w2); delete t; *rc_out = -5; return nullptr; }
  const bool pin_in = is_pinned_host(llr), pin_out = t->pin_out = is_pinned_host(out);
  int rc = 0;
  int32_t *d_it = (int32_t *)w->d_aux;
  uint8_t *d_ab = (uint8_t *)w->d_aux + n * sizeof(int32_t);
  uint8_t *h_ab = (uint8_t *)w->h_aux + n * sizeof(int32_t);
  if (abort_flags) std::memcpy(h_ab, abort_flags, n);
  const uint8_t *h_src = (const uint8_t *)llr;
  if (!pin_in) { std::memcpy(w->h_in, llr, in_bytes); h_src = (const uint8_t *)w->h_in; }
  uint8_t *h_dst = pin_out ? out : (uint8_t *)w->h_out;
  if (desc->use_crc && !pin_out) std::memcpy(w->h_out, out, out_bytes);   // the reference leaves p_out untouched until a CRC check runs
  // Chunks of one wave (one CTA per SM) alternate between the two streams: the first kernel starts after 1/7 of the input has crossed
  // PCIe, later chunks' CTAs fill the SMs the previous chunk's tail frees, and only the last wave's output is copied after the kernels.
  // NRB200_HOST_CHUNK_WAVES = waves per chunk (default 1).
  static const size_t waves = []() { const char *e = getenv("NRB200_HOST_CHUNK_WAVES"); const int v = e ? atoi(e) : 1; return (size_t)(v > 0 ? v : 1); }();
  const size_t per = w2 ? std::max<size_t>(1, (size_t)ctx().sm_count * waves) : n;
  cudaStream_t st[2] = {w->stream, w2 ? w2->stream : w->stream};
  for (size_t ci = 0, c0 = 0; c0 < n; ci++, c0 += per) {
    const size_t cn = std::min(per, n - c0);
    cudaStream_t s = st[ci & 1];
    const size_t io = c0 * desc->llr_stride, oo = c0 * desc->out_stride;
    if (cudaError_t ce = cudaMemcpyAsync((uint8_t *)w->d_in + io, h_src + io, cn * desc->llr_stride, cudaMemcpyHostToDevice, s)) { ctx().set_error("decode submit: H2D", ce); rc = -2; break; }
    if (abort_flags) cudaMemcpyAsync(d_ab + c0, h_ab + c0, cn, cudaMemcpyHostToDevice, s);
    if (desc->use_crc) cudaMemcpyAsync((uint8_t *)w->d_out + oo, h_dst + oo, cn * desc->out_stride, cudaMemcpyHostToDevice, s);
    DecodeArgs a = a0;
    a.n_cb = (uint32_t)cn;
    a.llr = (const int8_t *)w->d_in + io; a.out = (uint8_t *)w->d_out + oo; a.iters = d_it + c0;
    a.abort_flags = abort_flags ? d_ab + c0 : nullptr;
    if ((rc = launch_decode(dg, *hg, a, s)) != 0) break;
    if (cudaError_t ce = cudaMemcpyAsync(h_dst + oo, (uint8_t *)w->d_out + oo, cn * desc->out_stride, cudaMemcpyDeviceToHost, s)) { ctx().set_error("decode submit: D2H out", ce); rc = -2; break; }
    if (cudaError_t ce = cudaMemcpyAsync((int32_t *)w->h_aux + c0, d_it + c0, cn * sizeof(int32_t), cudaMemcpyDeviceToHost, s)) { ctx().set_error("decode submit: D2H iters", ce); rc = -2; break; }
  }
  t->rc = rc;
  return t;
}

static int decode_finish(DecodeTicket *t)
{
  int rc = t->rc;
  if (t->w) {
    cudaError_t e = cudaStreamSynchronize(t->w->stream);
    cudaError_t e2 = t->w2 ? cudaStreamSynchronize(t->w2->stream) : cudaSuccess;
    if (rc == 0 && (e != cudaSuccess || e2 != cudaSuccess)) { ctx().set_error("decode sync", e != cudaSuccess ? e : e2); rc = -2; }
    if (rc == 0) {
      if (!t->pin_out) std::memcpy(t->out, t->w->h_out, t->out_bytes);
      std::memcpy(t->iters, t->w->h_aux, t->n * sizeof(int32_t));
    }
    ctx().release(t->w);
    if (t->w2) ctx().release(t->w2);
  }
  delete t;
  return rc;
}

static int decode_host_impl(const nrb200_ldpc_batch_desc_t *desc, const int8_t *llr, uint8_t *out, int32_t *iters, const uint8_t *abort_flags)
{
  int rc = 0;
  DecodeTicket *t = decode_submit(desc, llr, out, iters, abort_flags, &rc);
  return t ? decode_finish(t) : rc;
}

NRB200_EXPORT int32_t nrb200_ldpc_decode_batch_host(const nrb200_ldpc_batch_desc_t *desc, const int8_t *llr, uint8_t *out, int32_t *iters)
{
  return decode_host_impl(desc, llr, out, iters, nullptr);
}

NRB200_EXPORT int32_t nrb200_ldpc_decode_batch_host_submit(const nrb200_ldpc_batch_desc_t *desc, const int8_t *llr, uint8_t *out, int32_t *iters,
                                                           void **ticket)
{
  if (!ticket) return -4;
  int rc = 0;
  *ticket = decode_submit(desc, llr, out, iters, nullptr, &rc);
  return *ticket ? 0 : rc;
}

NRB200_EXPORT int32_t nrb200_ldpc_decode_batch_host_wait(void *ticket)
{
  if (!ticket) return -4;
  return decode_finish((DecodeTicket *)ticket);
}

// Host arithmetic only (no GPU): the packed decoder's work schedule for (BG, Z, R) with at most max_threads threads, for inspection and the
// CPU tests.  info[0] = threads per CTA, [1] = work lists, [2] = 1 when a list belongs to a warp (items of 32 words) and 0 when to a bin of
// Z / 4 threads (whole rows), [3] / [4] = heaviest / mean check-node list (modelled warp instructions), [5] / [6] = same for the bit-node
// lists, [7] = 1 when every (row, chunk) and every (column of degree >= 2, chunk) appears in exactly one list.  Returns 0, -4 if the packed
// kernel does not serve this configuration.
NRB200_EXPORT int32_t nrb200_ldpc_packed_schedule_info(int BG, int Z, int R, int max_threads, int32_t *info)
{
  GraphDev *g = new GraphDev();
  PackedGraph *p = new PackedGraph();
  int rc = -4;
  if (info && build_graph(BG, Z, R, g) && build_packed_graph(*g, p, max_threads)) {
    const int chunks = p->warp_items ? p->Zw / 32 : 1;
    auto audit = [&](const int16_t *start, const int16_t *items, bool rows, int32_t *mx, int32_t *mean) {
      std::vector<int> seen((size_t)256 * 4, 0);
      long total = 0, worst = 0;
      for (int l = 0; l < p->nbins; l++) {
        long load = 0;
        for (int i = start[l]; i < start[l + 1]; i++) {
          const int id = items[i] & 0xFF, k = items[i] >> 8;
          seen[(size_t)id * 4 + k]++;
          load += rows ? 18 + 32 * (g->row_start[id + 1] - g->row_start[id]) + (g->row_p_col[id] >= 0 ? 51 : 11) : 65 + (27 * g->col_deg[id]) / 2;
        }
        total += load; worst = std::max(worst, load);
      }
      *mx = (int32_t)worst; *mean = (int32_t)(total / p->nbins);
      bool ok = true;
      const int n = rows ? g->nrows : g->ncols;
      for (int id = 0; id < n; id++)
        for (int k = 0; k < 4; k++) {
          const bool want = k < chunks && (rows || g->col_deg[id] >= 2);
          if (seen[(size_t)id * 4 + k] != (want ? 1 : 0)) ok = false;
        }
      return ok;
    };
    info[0] = p->nthreads; info[1] = p->nbins; info[2] = p->warp_items;
    const bool ok_cn = audit(p->cn_bin_start, p->cn_bin_rows, true, info + 3, info + 4);
    const bool ok_bn = audit(p->bn_bin_start, p->bn_bin_cols, false, info + 5, info + 6);
    info[7] = ok_cn && ok_bn ? 1 : 0;
    rc = 0;
  }
  delete g; delete p;
  return rc;
}

NRB200_EXPORT int32_t nrb200_device_index(void) { return ctx().inited ? ctx().dev : -1; }

// ------------------------------------------------------------------------------------------ one process, several GPUs
NRB200_EXPORT int32_t nrb200_device_count(void) { return device_count(); }
NRB200_EXPORT int32_t nrb200_set_device(int dev) { return set_current_device(dev) == 0 ? 0 : -4; }

// How many devices the OAI-facing entry points (LDPCdecoder / LDPCencoder / the offload convention) spread their calls over:
// NRB200_DEVICES = "all" or a count; default 1 (the thread's current device only).
static int abi_devices()
{
  static const int n = []() {
    const char *e = getenv("NRB200_DEVICES");
    if (!e) return 1;
    const int have = device_count();
    const int v = strcmp(e, "all") == 0 ? have : atoi(e);
    return v < 1 ? 1 : (v > have ? (have > 0 ? have : 1) : v);
  }();
  return n;
}

// The constant tables the decode kernels read for (BG, Z, R) -- lifted-graph descriptor + the packed decoder's shared-memory image description -- as
// one byte blob (host copy of what sits in device memory).  bench.py broadcasts rank 0's blob over NCCL at init and every rank compares its own
// against it ("NCCL broadcast of the base-graph matrices only at init").  Returns the blob size, or a negative error; copies min(size, cap) bytes.
NRB200_EXPORT int32_t nrb200_ldpc_graph_blob(int BG, int Z, int R, void *out, uint32_t cap)
{
  GraphDev *g = new GraphDev();
  PackedGraph *p = new PackedGraph();
  int32_t rc = -4;
  if (build_graph(BG, Z, R, g)) {
    const bool packed = (Z % 4 == 0) && build_packed_graph(*g, p, Z == 384 ? 768 : kPackedMaxThreads);
    const uint32_t total = (uint32_t)sizeof(GraphDev) + (packed ? (uint32_t)sizeof(PackedGraph) : 0u);
    if (out) {
      std::vector<uint8_t> b(total);
      std::memcpy(b.data(), g, sizeof(GraphDev));
      if (packed) std::memcpy(b.data() + sizeof(GraphDev), p, sizeof(PackedGraph));
      std::memcpy(out, b.data(), std::min(total, cap));
    }
    rc = (int32_t)total;
  }
  delete g; delete p;
  return rc;
}

// a code block (ulsch_id, segment r) goes to GPU hash(ulsch_id, r) mod nGPU, STICKY across HARQ rounds so that the library-owned int16
// soft buffer of the offload convention stays on one device.
NRB200_EXPORT int32_t nrb200_sticky_device(uint32_t ulsch_id, uint32_t r, uint32_t n_dev)
{
  if (n_dev == 0) return 0;
  uint32_t h = ulsch_id * 0x9E3779B1u + r * 0x85EBCA77u;
  h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12;
  return (int32_t)(h % n_dev);
}

// Blocking decode of one batch spread over `n_dev` GPUs (device 0 .. n_dev-1) of this process: contiguous shards of the batch, every device's
// copies and kernels enqueued before the first is waited for.  No collective: code blocks are independent.
NRB200_EXPORT int32_t nrb200_ldpc_decode_batch_host_multi(const nrb200_ldpc_batch_desc_t *desc, const int8_t *llr, uint8_t *out, int32_t *iters, int32_t n_dev)
{
  if (!desc || n_dev < 1 || n_dev > device_count()) return -4;
  const int saved = current_device();
  DecodeTicket *tk[kMaxDevices] = {nullptr};
  int rc = 0;
  const uint32_t n = desc->n_cb, per = (n + (uint32_t)n_dev - 1) / (uint32_t)n_dev;
  for (int d = 0; d < n_dev && rc == 0; d++) {
    const uint32_t c0 = std::min(n, (uint32_t)d * per), c1 = std::min(n, c0 + per);
    if (c1 == c0) continue;
    set_current_device(d);
    nrb200_ldpc_batch_desc_t sd = *desc;
    sd.n_cb = c1 - c0;
    tk[d] = decode_submit(&sd, llr + (size_t)c0 * desc->llr_stride, out + (size_t)c0 * desc->out_stride, iters + c0, nullptr, &rc);
  }
  for (int d = 0; d < n_dev; d++) {
    if (!tk[d]) continue;
    set_current_device(d);
    cudaSetDevice(ctx().dev);
    const int r2 = decode_finish(tk[d]);
    if (rc == 0) rc = r2;
  }
  set_current_device(saved);
  if (ctx().inited) cudaSetDevice(ctx().dev);
  return rc;
}
NRB200_EXPORT const char *nrb200_last_error(void)
{
  static thread_local std::string copy;       // other threads may replace Ctx::last_error at any time
  std::lock_guard<std::mutex> lk(ctx().mu);
  copy = ctx().last_error;
  return copy.c_str();
}
NRB200_EXPORT uint64_t nrb200_launch_count(void) { return ctx().launches.load(); }

// ------------------------------------------------------------------------------------------ part 1: OAI loader ABI
NRB200_EXPORT int32_t LDPCinit(void)
{
  if (ctx().init() != 0) return -1;
  cudaSetDevice(ctx().dev);
  return ll_warm() == 0 ? 0 : -1;
}
NRB200_EXPORT int32_t LDPCshutdown(void) { ctx().shutdown(); return 0; }

// Failure inside the library: OAI's callers decide success with `decodeIterations <= numMaxIter` (nr_ulsch_decoding.c:221,
// nr_dlsch_decoding.c:90) and know no negative return in this convention, so an internal error is reported the way the reference reports
// a block it could not decode: numMaxIter + 1 with the abort flag set; the cause is kept for nrb200_last_error().
static int32_t decoder_failure(const nrb200_ldpc_dec_params_t *p, nrb200_decode_abort_t *ab, const char *why)
{
  ctx().set_error(why, cudaPeekAtLastError());
  if (ab) {
    pthread_mutex_lock(&ab->mutex_failure);
    ab->failed = true;
    pthread_mutex_unlock(&ab->mutex_failure);
  }
  return (int32_t)p->numMaxIter + 1;
}

NRB200_EXPORT int32_t LDPCdecoder(nrb200_ldpc_dec_params_t *p, uint8_t harq_pid, uint8_t ulsch_id, uint8_t C, int8_t *p_llr,
                                  int8_t *p_out, nrb200_ldpc_time_stats_t *prof, nrb200_decode_abort_t *ab)
{
  (void)harq_pid; (void)ulsch_id; (void)C;
  const unsigned long long t0 = prof ? rdtsc_now() : 0;
  const int32_t numLLR = nrb200_ldpc_num_llr(p->BG, p->Z, p->R);
  if (numLLR < 0) return decoder_failure(p, ab, "LDPCdecoder: (BG, Z, R) is not an NR decoder configuration");
  if (abi_devices() > 1) {   // CPU-compatible convention: the host keeps the HARQ state, any device will do -- calls go round the GPUs
    static std::atomic<unsigned> rr{0};
    set_current_device((int)(rr++ % (unsigned)abi_devices()));
  }
  if (ensure_init()) return decoder_failure(p, ab, "LDPCdecoder: no CUDA device");
  nrb200_ldpc_batch_desc_t d;
  std::memset(&d, 0, sizeof(d));
  d.BG = p->BG; d.Z = p->Z; d.R = p->R; d.numMaxIter = p->numMaxIter; d.outMode = (uint8_t)p->outMode;
  d.n_cb = 1; d.llr_stride = (uint32_t)numLLR; d.out_stride = p->outMode == NRB200_OUTMODE_BIT ? (uint32_t)(numLLR + 7) / 8 : (uint32_t)numLLR;
  if (p->check_crc) {
    // The reference calls back into the host's check_crc (crc_byte.c:314) from inside the loop; the kernel evaluates the
    // same CRC on device (types CRC24_A/B, CRC16, CRC8).
    d.use_crc = 1; d.crc_type = (uint8_t)p->crc_type; d.crc_len_bits = (uint32_t)p->E;
  }
  int32_t it = 0;
  int rc;
  static const bool use_ll = []() { const char *e = getenv("NRB200_LL"); return !(e && atoi(e) == 0); }();
  if (use_ll) {
    // low-latency path (nrb200_ll.cu): mapped staging row, concurrent callers combined into one launch, cluster kernel, abort flag polled per iteration
    const GraphDev *hg = nullptr;
    const GraphDev *dg = ctx().graph(d.BG, d.Z, d.R, &hg);
    DecodeArgs a0;
    if (!dg || fill_args(&d, *hg, &a0) != 0) return decoder_failure(p, ab, "LDPCdecoder: invalid decode parameters");
    const uint64_t sig = ((uint64_t)d.BG << 56) | ((uint64_t)d.Z << 40) | ((uint64_t)d.R << 32) | ((uint64_t)d.numMaxIter << 24) | ((uint64_t)d.outMode << 22) |
                         ((uint64_t)d.use_crc << 21) | ((uint64_t)d.crc_type << 18) | (uint64_t)(d.crc_len_bits & 0x3FFFFu);
    rc = ll_decode_one(dg, hg, a0, sig, (size_t)numLLR, d.out_stride, p_llr, (uint8_t *)p_out, &it, ab);
  } else {
    uint8_t abort_now = 0;
    if (ab) {   // check_abort (defs_common.h:1008-1016)
      pthread_mutex_lock(&ab->mutex_failure);
      abort_now = ab->failed ? 1 : 0;
      pthread_mutex_unlock(&ab->mutex_failure);
    }
    rc = decode_host_impl(&d, p_llr, (uint8_t *)p_out, &it, &abort_now);
  }
  if (rc != 0) return decoder_failure(p, ab, "LDPCdecoder: device error");
  if (it > p->numMaxIter && ab) {   // set_abort (nrLDPC_decoder.c:190-193)
    pthread_mutex_lock(&ab->mutex_failure);
    ab->failed = true;
    pthread_mutex_unlock(&ab->mutex_failure);
  }
  if (prof) {   // stop_meas(&p_profiler->total) equivalent (time_meas.h:161-177); fine-grained fields are compiled out upstream
    const long long dt = (long long)(rdtsc_now() - t0);
    prof->total.trials++; prof->total.diff += dt; prof->total.p_time = dt; prof->total.diff_square += (double)dt * (double)dt;
    if (dt > prof->total.max) prof->total.max = dt;
  }
  return it;
}

NRB200_EXPORT void nrb200_ll_timing(uint64_t *out5) { if (out5) ll_timing(out5); }
NRB200_EXPORT int32_t nrb200_debug_cluster_marks(long long *out512) { return ensure_init() ? -1 : debug_cluster_marks(out512); }

// launches / code blocks of the low-latency path so far: blocks / launches = how many concurrent callers rode on one launch on average
NRB200_EXPORT void nrb200_ll_stats(uint64_t *launches, uint64_t *blocks) { if (launches && blocks) ll_stats(launches, blocks); }

// ------------------------------------------------------------------------------------------ encoder + CRC
NRB200_EXPORT int32_t nrb200_ldpc_encode_batch_dev(int BG, int Z, int K, uint32_t n_cb, const uint8_t *d_in, uint32_t in_stride,
                                                   uint8_t *d_out, uint32_t out_stride, void *stream)
{
  if (ensure_init()) return -1;
  const EncGraphDev *hg = nullptr;
  const EncGraphDev *dg = ctx().enc_graph(BG, Z, &hg);
  if (!dg || K != hg->nsys * Z) return -4;
  if (in_stride < (uint32_t)(K + 7) / 8 || out_stride < (uint32_t)(hg->ncols - 2) * Z) return -4;
  return launch_encode(dg, *hg, K, n_cb, d_in, in_stride, d_out, out_stride, (cudaStream_t)stream);
}

NRB200_EXPORT int32_t nrb200_ldpc_encode_batch_host(int BG, int Z, int K, uint32_t n_cb, const uint8_t *in, uint32_t in_stride, uint8_t *out,
                                                    uint32_t out_stride)
{
  if (ensure_init()) return -1;
  if (n_cb == 0) return 0;
  const size_t in_bytes = (size_t)n_cb * in_stride, out_bytes = (size_t)n_cb * out_stride;
  Workspace *w = ctx().acquire();
  if (!w || !w->reserve(in_bytes, out_bytes, 16)) { if (w) ctx().release(w); return -5; }
  int rc = 0;
  do {
    std::memcpy(w->h_in, in, in_bytes);
    if (cudaMemcpyAsync(w->d_in, w->h_in, in_bytes, cudaMemcpyHostToDevice, w->stream) != cudaSuccess) { rc = -2; break; }
    if ((rc = nrb200_ldpc_encode_batch_dev(BG, Z, K, n_cb, (const uint8_t *)w->d_in, in_stride, (uint8_t *)w->d_out, out_stride, w->stream)) != 0) break;
    if (cudaMemcpyAsync(w->h_out, w->d_out, out_bytes, cudaMemcpyDeviceToHost, w->stream) != cudaSuccess) { rc = -2; break; }
    cudaError_t e = cudaStreamSynchronize(w->stream);
    if (e != cudaSuccess) { ctx().set_error("encode sync", e); rc = -2; break; }
    std::memcpy(out, w->h_out, out_bytes);
  } while (0);
  ctx().release(w);
  return rc;
}

NRB200_EXPORT int32_t LDPCencoder(uint8_t **input, uint8_t **output, nrb200_ldpc_enc_params_t *impp)
{
  // segments 8*macro_num .. min(8*macro_num+8, n_segments) (ldpc_encoder_optim8segmulti.c:62-63)
  if (abi_devices() > 1) set_current_device((int)(impp->macro_num % (unsigned)abi_devices()));   // groups of 8 segments go round the GPUs
  if (ensure_init()) return -1;
  const unsigned s0 = 8 * impp->macro_num;
  const unsigned s1 = impp->n_segments > 8 * (impp->macro_num + 1) ? 8 * (impp->macro_num + 1) : impp->n_segments;
  if (s1 <= s0) return 0;
  const int BG = impp->BG, Z = (int)impp->Zc, K = (int)impp->K;
  const unsigned long long t0 = impp->tparity ? rdtsc_now() : 0;
  const uint32_t n = s1 - s0, kin = (uint32_t)(K + 7) / 8, nout = (uint32_t)((BG == 1 ? 66 : 50) * Z);
  const uint32_t in_stride = (kin + 15) & ~15u, out_stride = (nout + 15) & ~15u;
  static const bool use_ll = []() { const char *e = getenv("NRB200_LL"); return !(e && atoi(e) == 0); }();
  if (use_ll) {
    // low-latency path (nrb200_ll.cu): payloads and code words live in mapped pinned memory, the kernel reads and writes them over PCIe itself
    const EncGraphDev *hg = nullptr;
    const EncGraphDev *dg = ctx().enc_graph(BG, Z, &hg);
    if (!dg || K != hg->nsys * Z) return -1;
    const int rc = ll_encode(dg, hg, K, n, input + s0, output + s0, kin, nout);
    if (impp->tparity && rc == 0) {
      const long long dt = (long long)(rdtsc_now() - t0);
      impp->tparity->trials++; impp->tparity->diff += dt; impp->tparity->p_time = dt;
    }
    return rc == 0 ? 0 : -1;
  }
  Workspace *w = ctx().acquire();
  if (!w || !w->reserve((size_t)n * in_stride, (size_t)n * out_stride, 16)) { if (w) ctx().release(w); return -1; }
  int rc = 0;
  do {
    for (uint32_t j = 0; j < n; j++) std::memcpy((uint8_t *)w->h_in + (size_t)j * in_stride, input[s0 + j], kin);
    if (cudaMemcpyAsync(w->d_in, w->h_in, (size_t)n * in_stride, cudaMemcpyHostToDevice, w->stream) != cudaSuccess) { rc = -1; break; }
    if (nrb200_ldpc_encode_batch_dev(BG, Z, K, n, (const uint8_t *)w->d_in, in_stride, (uint8_t *)w->d_out, out_stride, w->stream) != 0) { rc = -1; break; }
    if (cudaMemcpyAsync(w->h_out, w->d_out, (size_t)n * out_stride, cudaMemcpyDeviceToHost, w->stream) != cudaSuccess) { rc = -1; break; }
    if (cudaStreamSynchronize(w->stream) != cudaSuccess) { rc = -1; break; }
    for (uint32_t j = 0; j < n; j++) std::memcpy(output[s0 + j], (uint8_t *)w->h_out + (size_t)j * out_stride, nout);
  } while (0);
  ctx().release(w);
  if (impp->tparity && rc == 0) {
    const long long dt = (long long)(rdtsc_now() - t0);
    impp->tparity->trials++; impp->tparity->diff += dt; impp->tparity->p_time = dt;
  }
  return rc;
}

NRB200_EXPORT int32_t nrb200_crc_batch_dev(int poly_id, uint32_t n_blk, const uint8_t *d_in, uint32_t stride, uint32_t bitlen, uint32_t *d_out,
                                           void *stream)
{
  if (ensure_init()) return -1;
  if (stride < (bitlen + 7) / 8) return -4;
  return launch_crc(poly_id, n_blk, d_in, stride, bitlen, d_out, (cudaStream_t)stream);
}

NRB200_EXPORT int32_t nrb200_crc_batch_host(int poly_id, uint32_t n_blk, const uint8_t *in, uint32_t stride, uint32_t bitlen, uint32_t *out)
{
  if (ensure_init()) return -1;
  if (n_blk == 0) return 0;
  const size_t in_bytes = (size_t)n_blk * stride, out_bytes = (size_t)n_blk * 4;
  Workspace *w = ctx().acquire();
  if (!w || !w->reserve(in_bytes, out_bytes, 16)) { if (w) ctx().release(w); return -5; }
  int rc = 0;
  do {
    std::memcpy(w->h_in, in, in_bytes);
    if (cudaMemcpyAsync(w->d_in, w->h_in, in_bytes, cudaMemcpyHostToDevice, w->stream) != cudaSuccess) { rc = -2; break; }
    if ((rc = nrb200_crc_batch_dev(poly_id, n_blk, (const uint8_t *)w->d_in, stride, bitlen, (uint32_t *)w->d_out, w->stream)) != 0) break;
    if (cudaMemcpyAsync(w->h_out, w->d_out, out_bytes, cudaMemcpyDeviceToHost, w->stream) != cudaSuccess) { rc = -2; break; }
    if (cudaStreamSynchronize(w->stream) != cudaSuccess) { rc = -2; break; }
    std::memcpy(out, w->h_out, out_bytes);
  } while (0);
  ctx().release(w);
  return rc;
}

NRB200_EXPORT int32_t nrb200_tb_segment_parms(int BG, uint32_t A, uint32_t out[6]) { return (BG == 1 || BG == 2) && out ? tb_segment_parms(BG, A, out) : -1; }

NRB200_EXPORT int32_t nrb200_tb_segment_dev(int BG, uint32_t A, const uint8_t *d_payload, uint8_t *d_segs, uint32_t seg_stride, uint32_t *d_scratch, void *stream)
{
  if (ensure_init()) return -1;
  return launch_tb_segment(BG, A, d_payload, d_segs, seg_stride, d_scratch, (cudaStream_t)stream);
}

NRB200_EXPORT int32_t nrb200_tb_segment_host(int BG, uint32_t A, const uint8_t *payload, uint8_t *segs, uint32_t seg_stride)
{
  uint32_t q[6];
  if ((BG != 1 && BG != 2) || A == 0 || (A & 7) || tb_segment_parms(BG, A, q) < 0) return -4;
  if (ensure_init()) return -1;
  const size_t in_bytes = A / 8, out_bytes = (size_t)q[0] * seg_stride;
  Workspace *w = ctx().acquire();
  if (!w || !w->reserve(in_bytes + 64, out_bytes, 16)) { if (w) ctx().release(w); return -5; }
  int rc = 0;
  do {
    std::memcpy(w->h_in, payload, in_bytes);
    if (cudaMemcpyAsync(w->d_in, w->h_in, in_bytes, cudaMemcpyHostToDevice, w->stream) != cudaSuccess) { rc = -2; break; }
    if (cudaMemsetAsync(w->d_out, 0, out_bytes, w->stream) != cudaSuccess) { rc = -2; break; }
    if ((rc = launch_tb_segment(BG, A, (const uint8_t *)w->d_in, (uint8_t *)w->d_out, seg_stride, (uint32_t *)w->d_aux, w->stream)) != 0) break;
    if (cudaMemcpyAsync(w->h_out, w->d_out, out_bytes, cudaMemcpyDeviceToHost, w->stream) != cudaSuccess) { rc = -2; break; }
    if (cudaStreamSynchronize(w->stream) != cudaSuccess) { rc = -2; break; }
    std::memcpy(segs, w->h_out, out_bytes);
  } while (0);
  ctx().release(w);
  return rc;
}

// ------------------------------------------------------------------------------------------ part 3: rate matching
static int rm_check(const nrb200_rm_desc_t *d)
{
  if ((d->BG != 1 && d->BG != 2) || ils_of_z(d->Z) < 0 || d->rv > 3) return -4;
  if (d->Qm != 1 && d->Qm != 2 && d->Qm != 4 && d->Qm != 6 && d->Qm != 8) return -4;
  if (d->K != (uint32_t)(d->BG == 1 ? 22 : 10) * d->Z || d->F + 2u * d->Z > d->K || d->C == 0) return -4;
  return 0;
}

NRB200_EXPORT int32_t nrb200_ldpc_rm_tx_batch_dev(const nrb200_rm_desc_t *desc, const uint8_t *d_d, uint32_t d_stride, const uint32_t *d_E,
                                                  const uint32_t *d_foff, uint8_t *d_f, void *stream)
{
  if (ensure_init()) return -1;
  if (int rc = rm_check(desc)) return rc;
  if (d_stride < (uint32_t)(desc->BG == 1 ? 66 : 50) * desc->Z) return -4;
  return launch_rm_tx(*desc, d_d, d_stride, d_E, d_foff, d_f, (cudaStream_t)stream);
}

NRB200_EXPORT int32_t nrb200_ldpc_rm_tx_batch_host(const nrb200_rm_desc_t *desc, const uint8_t *d, uint32_t d_stride, const uint32_t *E, uint8_t *f)
{
  if (ensure_init()) return -1;
  if (int rc = rm_check(desc)) return rc;
  const uint32_t n = desc->n_seg;
  if (n == 0) return 0;
  std::vector<uint32_t> tab(2 * (size_t)n);
  size_t tot = 0;
  const uint32_t Foffset = desc->K - desc->F - 2u * desc->Z;
  for (uint32_t r = 0; r < n; r++) { if (E[r] < Foffset) return -4; tab[r] = E[r]; tab[n + r] = (uint32_t)tot; tot += E[r]; }   // "Foffset > E" is an error upstream (:454)
  const size_t in_bytes = (size_t)n * d_stride;
  Workspace *w = ctx().acquire();
  if (!w || !w->reserve(in_bytes, tot, 8 * (size_t)n)) { if (w) ctx().release(w); return -5; }
  int rc = 0;
  do {
    std::memcpy(w->h_in, d, in_bytes);
    std::memcpy(w->h_aux, tab.data(), 8 * (size_t)n);
    if (cudaMemcpyAsync(w->d_in, w->h_in, in_bytes, cudaMemcpyHostToDevice, w->stream) != cudaSuccess) { rc = -2; break; }
    if (cudaMemcpyAsync(w->d_aux, w->h_aux, 8 * (size_t)n, cudaMemcpyHostToDevice, w->stream) != cudaSuccess) { rc = -2; break; }
    if ((rc = launch_rm_tx(*desc, (const uint8_t *)w->d_in, d_stride, (const uint32_t *)w->d_aux, (const uint32_t *)w->d_aux + n, (uint8_t *)w->d_out, w->stream)) != 0) break;
    if (cudaMemcpyAsync(w->h_out, w->d_out, tot, cudaMemcpyDeviceToHost, w->stream) != cudaSuccess) { rc = -2; break; }
    if (cudaStreamSynchronize(w->stream) != cudaSuccess) { rc = -2; break; }
    std::memcpy(f, w->h_out, tot);
  } while (0);
  ctx().release(w);
  return rc;
}

NRB200_EXPORT int32_t nrb200_ldpc_rm_rx_batch_dev(const nrb200_rm_desc_t *desc, const int16_t *d_soft, const uint32_t *d_E, const uint32_t *d_soff,
                                                  int16_t *d_harq, uint32_t harq_stride, int8_t *d_llr, uint32_t llr_stride, void *stream)
{
  if (ensure_init()) return -1;
  if (int rc = rm_check(desc)) return rc;
  if (harq_stride < (uint32_t)(desc->BG == 1 ? 66 : 50) * desc->Z || llr_stride < (uint32_t)(desc->BG == 1 ? 68 : 52) * desc->Z) return -4;
  return launch_rm_rx(*desc, d_soft, d_E, d_soff, d_harq, harq_stride, d_llr, llr_stride, (cudaStream_t)stream);
}

NRB200_EXPORT int32_t nrb200_ldpc_rm_rx_batch_host(const nrb200_rm_desc_t *desc, const int16_t *soft, const uint32_t *E, int16_t *harq,
                                                   uint32_t harq_stride, int8_t *llr, uint32_t llr_stride)
{
  if (ensure_init()) return -1;
  if (int rc = rm_check(desc)) return rc;
  const uint32_t n = desc->n_seg;
  if (n == 0) return 0;
  if (harq_stride < (uint32_t)(desc->BG == 1 ? 66 : 50) * desc->Z || llr_stride < (uint32_t)(desc->BG == 1 ? 68 : 52) * desc->Z) return -4;
  std::vector<uint32_t> tab(2 * (size_t)n);
  size_t tot = 0;
  const uint32_t Foffset = desc->K - desc->F - 2u * desc->Z;
  for (uint32_t r = 0; r < n; r++) { if (E[r] < Foffset || E[r] % desc->Qm) return -4; tab[r] = E[r]; tab[n + r] = (uint32_t)tot; tot += E[r]; }
  const size_t soft_bytes = 2 * tot, harq_bytes = 2 * (size_t)n * harq_stride, llr_bytes = (size_t)n * llr_stride;
  Workspace *w = ctx().acquire();
  if (!w || !w->reserve(soft_bytes + harq_bytes + 64, llr_bytes, 8 * (size_t)n)) { if (w) ctx().release(w); return -5; }
  int rc = 0;
  do {
    const size_t hoff = (soft_bytes + 15) & ~(size_t)15;
    std::memcpy(w->h_in, soft, soft_bytes);
    std::memcpy((uint8_t *)w->h_in + hoff, harq, harq_bytes);
    std::memcpy(w->h_aux, tab.data(), 8 * (size_t)n);
    if (cudaMemcpyAsync(w->d_in, w->h_in, hoff + harq_bytes, cudaMemcpyHostToDevice, w->stream) != cudaSuccess) { rc = -2; break; }
    if (cudaMemcpyAsync(w->d_aux, w->h_aux, 8 * (size_t)n, cudaMemcpyHostToDevice, w->stream) != cudaSuccess) { rc = -2; break; }
    int16_t *d_harq = (int16_t *)((uint8_t *)w->d_in + hoff);
    if ((rc = launch_rm_rx(*desc, (const int16_t *)w->d_in, (const uint32_t *)w->d_aux, (const uint32_t *)w->d_aux + n, d_harq, harq_stride,
                           (int8_t *)w->d_out, llr_stride, w->stream)) != 0) break;
    if (cudaMemcpyAsync(w->h_out, w->d_out, llr_bytes, cudaMemcpyDeviceToHost, w->stream) != cudaSuccess) { rc = -2; break; }
    if (cudaMemcpyAsync((uint8_t *)w->h_in + hoff, d_harq, harq_bytes, cudaMemcpyDeviceToHost, w->stream) != cudaSuccess) { rc = -2; break; }
    if (cudaStreamSynchronize(w->stream) != cudaSuccess) { rc = -2; break; }
    std::memcpy(llr, w->h_out, llr_bytes);
    std::memcpy(harq, (uint8_t *)w->h_in + hoff, harq_bytes);
  } while (0);
  ctx().release(w);
  return rc;
}

// ------------------------------------------------------------------------------------------ part 4: demodulation LLRs
NRB200_EXPORT int32_t nrb200_pusch_llr_dev(int Qm, uint32_t nb_re, const int16_t *d_rxF, const int16_t *d_mag_a, const int16_t *d_mag_b,
                                           const int16_t *d_mag_c, int16_t *d_llr, void *stream)
{
  if (ensure_init()) return -1;
  if ((Qm >= 4 && !d_mag_a) || (Qm >= 6 && !d_mag_b) || (Qm >= 8 && !d_mag_c)) return -4;
  return launch_pusch_llr(Qm, nb_re, d_rxF, d_mag_a, d_mag_b, d_mag_c, d_llr, (cudaStream_t)stream);
}

NRB200_EXPORT int32_t nrb200_pusch_llr_host(int Qm, uint32_t nb_re, const int16_t *rxF, const int16_t *mag_a, const int16_t *mag_b, const int16_t *mag_c,
                                            int16_t *llr)
{
  if (ensure_init()) return -1;
  if (Qm != 2 && Qm != 4 && Qm != 6 && Qm != 8) return -4;
  if ((Qm >= 4 && !mag_a) || (Qm >= 6 && !mag_b) || (Qm >= 8 && !mag_c)) return -4;
  if (nb_re == 0) return 0;
  const size_t plane = ((size_t)nb_re * 4 + 15) & ~(size_t)15, out_bytes = (size_t)nb_re * Qm * 2;
  Workspace *w = ctx().acquire();
  if (!w || !w->reserve(4 * plane, out_bytes + 16, 16)) { if (w) ctx().release(w); return -5; }
  int rc = 0;
  do {
    uint8_t *h = (uint8_t *)w->h_in;
    std::memcpy(h, rxF, (size_t)nb_re * 4);
    if (Qm >= 4) std::memcpy(h + plane, mag_a, (size_t)nb_re * 4);
    if (Qm >= 6) std::memcpy(h + 2 * plane, mag_b, (size_t)nb_re * 4);
    if (Qm >= 8) std::memcpy(h + 3 * plane, mag_c, (size_t)nb_re * 4);
    if (cudaMemcpyAsync(w->d_in, w->h_in, 4 * plane, cudaMemcpyHostToDevice, w->stream) != cudaSuccess) { rc = -2; break; }
    const uint8_t *d = (const uint8_t *)w->d_in;
    if ((rc = launch_pusch_llr(Qm, nb_re, (const int16_t *)d, (const int16_t *)(d + plane), (const int16_t *)(d + 2 * plane), (const int16_t *)(d + 3 * plane),
                               (int16_t *)w->d_out, w->stream)) != 0) break;
    if (cudaMemcpyAsync(w->h_out, w->d_out, out_bytes, cudaMemcpyDeviceToHost, w->stream) != cudaSuccess) { rc = -2; break; }
    if (cudaStreamSynchronize(w->stream) != cudaSuccess) { rc = -2; break; }
    std::memcpy(llr, w->h_out, out_bytes);
  } while (0);
  ctx().release(w);
  return rc;
}

// ------------------------------------------------------------------------------------------ scrambling + QAM mapper
static inline uint32_t gold_cinit(uint32_t q, uint32_t Nid, uint32_t n_RNTI) { return (n_RNTI << 15) + (q << 14) + Nid; }

NRB200_EXPORT int32_t nrb200_scramble_dev(const uint8_t *d_in, uint32_t size, uint32_t q, uint32_t Nid, uint32_t n_RNTI, uint32_t *d_out, void *stream)
{
  if (ensure_init()) return -1;
  return launch_gold(0, gold_cinit(q, Nid, n_RNTI), size, d_in, d_out, nullptr, (cudaStream_t)stream);
}
NRB200_EXPORT int32_t nrb200_unscramble_llr_dev(int16_t *d_llr, uint32_t size, uint32_t q, uint32_t Nid, uint32_t n_RNTI, void *stream)
{
  if (ensure_init()) return -1;
  return launch_gold(1, gold_cinit(q, Nid, n_RNTI), size, nullptr, nullptr, d_llr, (cudaStream_t)stream);
}
NRB200_EXPORT int32_t nrb200_modulate_dev(const uint32_t *d_bits, uint32_t length_bits, int Qm, int16_t *d_out, void *stream)
{
  if (ensure_init()) return -1;
  return launch_modulate(Qm, length_bits, (const uint8_t *)d_bits, d_out, (cudaStream_t)stream);
}

// "copy in, run, copy out" for the small host-buffer variants (64 zero bytes of slack after the input)
template <typename F>
static int host_roundtrip(const void *in, size_t in_bytes, void *out, size_t out_bytes, bool inplace, F &&run)
{
  if (ensure_init()) return -1;
  Workspace *w = ctx().acquire();
  if (!w || !w->reserve(in_bytes + 64, out_bytes + 64, 16)) { if (w) ctx().release(w); return -5; }
  int rc = 0;
  do {
    std::memcpy(w->h_in, in, in_bytes);
    std::memset((uint8_t *)w->h_in + in_bytes, 0, 64);
    if (cudaMemcpyAsync(w->d_in, w->h_in, in_bytes + 64, cudaMemcpyHostToDevice, w->stream) != cudaSuccess) { rc = -2; break; }
    if ((rc = run(w)) != 0) break;
    void *src = inplace ? w->d_in : w->d_out;
    if (cudaMemcpyAsync(w->h_out, src, out_bytes, cudaMemcpyDeviceToHost, w->stream) != cudaSuccess) { rc = -2; break; }
    if (cudaStreamSynchronize(w->stream) != cudaSuccess) { rc = -2; break; }
    std::memcpy(out, w->h_out, out_bytes);
  } while (0);
  ctx().release(w);
  return rc;
}

NRB200_EXPORT int32_t nrb200_scramble_host(const uint8_t *in, uint32_t size, uint32_t q, uint32_t Nid, uint32_t n_RNTI, uint32_t *out)
{
  if (size == 0) return 0;
  return host_roundtrip(in, size, out, 4 * (size_t)((size + 31) >> 5), false, [&](Workspace *w) {
    return launch_gold(0, gold_cinit(q, Nid, n_RNTI), size, (const uint8_t *)w->d_in, (uint32_t *)w->d_out, nullptr, w->stream); });
}
NRB200_EXPORT int32_t nrb200_unscramble_llr_host(int16_t *llr, uint32_t size, uint32_t q, uint32_t Nid, uint32_t n_RNTI)
{
  if (size == 0) return 0;
  return host_roundtrip(llr, 2 * (size_t)size, llr, 2 * (size_t)size, true, [&](Workspace *w) {
    return launch_gold(1, gold_cinit(q, Nid, n_RNTI), size, nullptr, nullptr, (int16_t *)w->d_in, w->stream); });
}
NRB200_EXPORT int32_t nrb200_modulate_host(const uint32_t *bits, uint32_t length_bits, int Qm, int16_t *out)
{
  if (Qm != 2 && Qm != 4 && Qm != 6 && Qm != 8) return -4;
  if (length_bits / Qm == 0) return 0;
  return host_roundtrip(bits, (length_bits + 7) / 8, out, 4 * (size_t)(length_bits / Qm), false, [&](Workspace *w) {
    return launch_modulate(Qm, length_bits, (const uint8_t *)w->d_in, (int16_t *)w->d_out, w->stream); });
}

// ------------------------------------------------------------------------------------------ PDSCH transmitter after the encoder
NRB200_EXPORT uint32_t nrb200_pdsch_tx_num_bits(const nrb200_pdsch_tx_t *d) { return d ? pdsch_tx_num_bits(*d) : 0; }

NRB200_EXPORT int32_t nrb200_pdsch_tx_slot_dev(const nrb200_pdsch_tx_t *d, const uint8_t *d_f, int16_t *d_txF, void *stream)
{
  if (ensure_init() || !d) return -1;
  return launch_pdsch_tx(*d, d_f, d_txF, (cudaStream_t)stream);
}

NRB200_EXPORT int32_t nrb200_pdsch_tx_slot_host(const nrb200_pdsch_tx_t *d, const uint8_t *f, int16_t *txdataF)
{
  if (ensure_init() || !d) return -1;
  const uint32_t G = pdsch_tx_num_bits(*d);
  if (G == 0) return -4;
  const size_t tx_bytes = (size_t)d->nb_tx * 14 * d->fft_size * 4;
  Workspace *w = ctx().acquire();
  if (!w || !w->reserve(G + 64, tx_bytes, 16)) { if (w) ctx().release(w); return -5; }
  nrb200_pdsch_tx_t dd = *d;
  dd.tx_stride = 14 * d->fft_size;
  int rc = 0;
  do {
    std::memcpy(w->h_in, f, G);
    std::memcpy(w->h_out, txdataF, tx_bytes);
    if (cudaMemcpyAsync(w->d_in, w->h_in, G, cudaMemcpyHostToDevice, w->stream) != cudaSuccess) { rc = -2; break; }
    if (cudaMemcpyAsync(w->d_out, w->h_out, tx_bytes, cudaMemcpyHostToDevice, w->stream) != cudaSuccess) { rc = -2; break; }
    if ((rc = launch_pdsch_tx(dd, (const uint8_t *)w->d_in, (int16_t *)w->d_out, w->stream)) != 0) break;
    if (cudaMemcpyAsync(w->h_out, w->d_out, tx_bytes, cudaMemcpyDeviceToHost, w->stream) != cudaSuccess) { rc = -2; break; }
    if (cudaStreamSynchronize(w->stream) != cudaSuccess) { rc = -2; break; }
    std::memcpy(txdataF, w->h_out, tx_bytes);
  } while (0);
  ctx().release(w);
  return rc;
}

// ------------------------------------------------------------------------------------------ PUSCH inner receiver (one layer)
// completion counters of the level kernels ("last CTA combines"): one per launch in flight, handed out round robin so that slots processed concurrently on
// different streams never share one (the kernel leaves its counter at zero)
static uint32_t *pusch_counter()
{
  constexpr uint32_t kCounters = 4096;
  static uint32_t *dd[kMaxDevices] = {nullptr};
  static std::mutex mu;
  static uint32_t ticket = 0;
  std::lock_guard<std::mutex> lk(mu);
  uint32_t *&d = dd[ctx().dev];
  // 32 words per launch: [0] the counter, [1..16] the per-plane levels of the UE receiver with up to 4 layers x 4 antennas
  if (!d && cudaMalloc(&d, kCounters * 128) == cudaSuccess) cudaMemset(d, 0, kCounters * 128);
  return d ? d + 32 * (ticket++ % kCounters) : nullptr;
}

NRB200_EXPORT uint32_t nrb200_pusch_num_llr(const nrb200_pusch_rx_t *d) { return d ? pusch_num_llr(*d) : 0; }

NRB200_EXPORT int32_t nrb200_pusch_log2_maxh_dev(const nrb200_pusch_rx_t *d, const int16_t *d_ch, int32_t *d_out, void *stream)
{
  if (ensure_init() || !d) return -1;
  uint32_t *cnt = pusch_counter();
  if (!cnt) return -5;
  return launch_pusch_level(*d, d_ch, d_out, cnt, (cudaStream_t)stream);
}

NRB200_EXPORT int32_t nrb200_pusch_inner_rx_dev(const nrb200_pusch_rx_t *d, const int16_t *d_rxF, const int16_t *d_ch, const int32_t *d_log2_maxh,
                                                int16_t *d_llr, void *stream)
{
  if (ensure_init() || !d) return -1;
  return launch_pusch_rx(*d, d_rxF, d_ch, d_log2_maxh, d_llr, (cudaStream_t)stream);
}

NRB200_EXPORT int32_t nrb200_pusch_inner_rx_host(const nrb200_pusch_rx_t *d, const int16_t *rxdataF, const int16_t *ul_ch_estimates, int16_t *llr,
                                                 int32_t *log2_maxh_out)
{
  if (ensure_init() || !d) return -1;
  if (d->d_est_state != 0) return -4;                                     // a device address: the _dev entry points only
  const uint32_t n_llr = pusch_num_llr(*d);
  if (n_llr == 0) return -4;
  const size_t plane = (size_t)d->nb_rx * 14 * d->fft_size * 4, est_plane = plane * (d->nrOfLayers >= 2 ? d->nrOfLayers : 1);
  const size_t tp_bytes = d->transform_precoding ? pusch_tp_scratch_bytes(*d) : 0;        // the transforms' input / output planes, behind the level slots
  Workspace *w = ctx().acquire();
  if (!w || !w->reserve(plane + est_plane, (size_t)n_llr * 2 + 64, 64 + tp_bytes + 128 + 128)) { if (w) ctx().release(w); return -5; }
  int rc = 0;
  do {
    std::memcpy(w->h_in, rxdataF, plane);
    std::memcpy((uint8_t *)w->h_in + plane, ul_ch_estimates, est_plane);
    if (cudaMemcpyAsync(w->d_in, w->h_in, plane + est_plane, cudaMemcpyHostToDevice, w->stream) != cudaSuccess) { rc = -2; break; }
    nrb200_pusch_rx_t e = *d;
    e.rx_stride = e.ch_stride = 14 * d->fft_size;
    e.d_tp_scratch = tp_bytes ? (uint64_t)(uintptr_t)((uint8_t *)w->d_aux + 64) : 0;
    e.d_ptrs_state = (uint64_t)(uintptr_t)((uint8_t *)w->d_aux + 64 + tp_bytes);           // PT-RS: 14 phases + status behind the transforms' planes
    const int16_t *d_rx = (const int16_t *)w->d_in, *d_ch = (const int16_t *)((uint8_t *)w->d_in + plane);
    const bool measure = d->log2_maxh == 0xFFFFFFFFu;
    int32_t *d_lvl = (int32_t *)w->d_aux;
    if (measure) {
      e.log2_maxh = 0;
      // one workspace, one stream: give the level kernel its own completion counter slot in the workspace
      uint32_t *cnt = (uint32_t *)((uint8_t *)w->d_aux + 64 + tp_bytes + 128);             // counter + per-plane levels (32 words)
      if (cudaMemsetAsync(cnt, 0, 4, w->stream) != cudaSuccess) { rc = -2; break; }
      if ((rc = launch_pusch_level(e, d_ch, d_lvl, cnt, w->stream)) != 0) break;
    }
    if ((rc = launch_pusch_rx(e, d_rx, d_ch, measure ? d_lvl + 8 : nullptr, (int16_t *)w->d_out, w->stream)) != 0) break;
    if (cudaMemcpyAsync(w->h_out, w->d_out, (size_t)n_llr * 2, cudaMemcpyDeviceToHost, w->stream) != cudaSuccess) { rc = -2; break; }
    if (measure && cudaMemcpyAsync(w->h_aux, d_lvl, 36, cudaMemcpyDeviceToHost, w->stream) != cudaSuccess) { rc = -2; break; }
    if (cudaStreamSynchronize(w->stream) != cudaSuccess) { rc = -2; break; }
    std::memcpy(llr, w->h_out, (size_t)n_llr * 2);
    if (log2_maxh_out) *log2_maxh_out = measure ? ((int32_t *)w->h_aux)[8] : (int32_t)d->log2_maxh;
  } while (0);
  ctx().release(w);
  return rc;
}

NRB200_EXPORT int32_t nrb200_pdsch_ptrs_layout(const nrb200_pusch_rx_t *d, uint32_t *ptrs_symbols, uint32_t *ptrs_re_per_symbol)
{
  return d ? pusch_ptrs_layout(*d, ptrs_symbols, ptrs_re_per_symbol) : -1;
}

// ------------------------------------------------------------------------------------------ gNB PRACH detector (rx_nr_prach)
NRB200_EXPORT uint32_t nrb200_prach_num_roots(const nrb200_prach_t *d) { return d ? prach_num_roots(*d) : 0; }
NRB200_EXPORT uint64_t nrb200_prach_scratch_bytes(const nrb200_prach_t *d) { return d ? prach_scratch_bytes(*d) : 0; }

NRB200_EXPORT int32_t nrb200_rx_nr_prach_dev(const nrb200_prach_t *d, const int16_t *d_X_u, const int16_t *d_rxsigF, int32_t *d_out, void *d_scratch, void *stream)
{
  if (ensure_init() || !d || !d_X_u || !d_rxsigF || !d_out || !d_scratch) return -1;
  return launch_prach(*d, d_X_u, d_rxsigF, d_out, d_scratch, (cudaStream_t)stream);
}

NRB200_EXPORT int32_t nrb200_rx_nr_prach_host(const nrb200_prach_t *d, const int16_t *X_u, const int16_t *rxsigF, uint16_t *max_preamble, uint16_t *max_preamble_energy,
                                              uint16_t *max_preamble_delay)
{
  if (ensure_init() || !d || !X_u || !rxsigF) return -1;
  const uint32_t roots = prach_num_roots(*d);
  if (roots == 0) return -4;
  const uint32_t N_ZC = d->short_sequence ? 139 : 839;
  const size_t xu_b = (size_t)roots * 839 * 4, rx_b = (size_t)d->nb_rx * N_ZC * 4, sc_b = prach_scratch_bytes(*d);
  Workspace *w = ctx().acquire();
  if (!w || !w->reserve(xu_b + rx_b, 64, sc_b + 64)) { if (w) ctx().release(w); return -5; }
  int rc = 0;
  do {
    std::memcpy(w->h_in, X_u, xu_b);
    std::memcpy((uint8_t *)w->h_in + xu_b, rxsigF, rx_b);
    if (cudaMemcpyAsync(w->d_in, w->h_in, xu_b + rx_b, cudaMemcpyHostToDevice, w->stream) != cudaSuccess) { rc = -2; break; }
    nrb200_prach_t e = *d;
    e.rx_stride = N_ZC;
    if ((rc = launch_prach(e, (const int16_t *)w->d_in, (const int16_t *)((uint8_t *)w->d_in + xu_b), (int32_t *)w->d_out, w->d_aux, w->stream)) != 0) break;
    if (cudaMemcpyAsync(w->h_out, w->d_out, 12, cudaMemcpyDeviceToHost, w->stream) != cudaSuccess || cudaStreamSynchronize(w->stream) != cudaSuccess) { rc = -2; break; }
    const int32_t *o = (const int32_t *)w->h_out;
    if (max_preamble) *max_preamble = (uint16_t)o[0];
    if (max_preamble_energy) *max_preamble_energy = (uint16_t)o[1];
    if (max_preamble_delay) *max_preamble_delay = (uint16_t)o[2];
  } while (0);
  ctx().release(w);
  return rc;
}

// ------------------------------------------------------------------------------------------ rfsimulator channel application (rxAddInput)
NRB200_EXPORT int32_t nrb200_rfsim_rx_add_input_dev(const nrb200_rfsim_chan_t *c, const double *d_ch, const int16_t *d_input_sig, int16_t *d_out, uint32_t out_stride,
                                                    uint32_t nbSamples, uint64_t TS, uint32_t CirSize, const double *d_noise, void *stream)
{
  if (ensure_init() || !c || !d_ch || !d_input_sig || !d_out) return -1;
  if (out_stride < nbSamples) return -4;
  return launch_rfsim(*c, d_ch, d_input_sig, d_out, out_stride, nbSamples, TS, CirSize, d_noise, (cudaStream_t)stream);
}

NRB200_EXPORT int32_t nrb200_rfsim_rx_add_input_host(const nrb200_rfsim_chan_t *c, const double *ch, const int16_t *input_sig, int16_t *out, uint32_t out_stride,
                                                     uint32_t nbSamples, uint64_t TS, uint32_t CirSize, const double *noise)
{
  if (ensure_init() || !c || !ch || !input_sig || !out) return -1;
  if (out_stride < nbSamples || c->nb_tx < 1 || c->nb_tx > 8 || c->nb_rx < 1 || c->nb_rx > 8 || c->channel_length < 1 || c->channel_length > 255) return -4;
  const size_t ch_b = (size_t)c->nb_tx * c->nb_rx * c->channel_length * 16, sig_b = ((size_t)CirSize * 4 + 15) & ~(size_t)15;
  const size_t nz_b = noise ? (size_t)c->nb_rx * nbSamples * 16 : 0, out_b = (size_t)c->nb_rx * out_stride * 4;
  Workspace *w = ctx().acquire();
  if (!w || !w->reserve(ch_b + sig_b + nz_b, out_b + 64, 64)) { if (w) ctx().release(w); return -5; }
  int rc = 0;
  do {
    uint8_t *h = (uint8_t *)w->h_in, *dv = (uint8_t *)w->d_in;
    std::memcpy(h, ch, ch_b);
    std::memcpy(h + ch_b, input_sig, (size_t)CirSize * 4);
    if (noise) std::memcpy(h + ch_b + sig_b, noise, nz_b);
    std::memcpy(w->h_out, out, out_b);
    if (cudaMemcpyAsync(dv, h, ch_b + sig_b + nz_b, cudaMemcpyHostToDevice, w->stream) != cudaSuccess ||
        cudaMemcpyAsync(w->d_out, w->h_out, out_b, cudaMemcpyHostToDevice, w->stream) != cudaSuccess) { rc = -2; break; }
    if ((rc = launch_rfsim(*c, (const double *)dv, (const int16_t *)(dv + ch_b), (int16_t *)w->d_out, out_stride, nbSamples, TS, CirSize,
                           noise ? (const double *)(dv + ch_b + sig_b) : nullptr, w->stream)) != 0) break;
    if (cudaMemcpyAsync(w->h_out, w->d_out, out_b, cudaMemcpyDeviceToHost, w->stream) != cudaSuccess || cudaStreamSynchronize(w->stream) != cudaSuccess) { rc = -2; break; }
    std::memcpy(out, w->h_out, out_b);
  } while (0);
  ctx().release(w);
  return rc;
}

NRB200_EXPORT uint64_t nrb200_pusch_tp_scratch_bytes(const nrb200_pusch_rx_t *d) { return d ? pusch_tp_scratch_bytes(*d) : 0; }

// ------------------------------------------------------------------------------------------ PUSCH channel estimation
NRB200_EXPORT int32_t nrb200_lowpapr_sequence_host(uint32_t u, uint32_t v, uint32_t n_re, uint32_t scaling, int16_t *seq) { return lowpapr_sequence_host(u, v, n_re, scaling, seq); }
NRB200_EXPORT int32_t nrb200_pusch_dmrs_pilots_host(const nrb200_pusch_chest_t *d, int16_t *pilots) { return d && pilots ? pusch_dmrs_pilots_host(*d, pilots) : -4; }

NRB200_EXPORT uint64_t nrb200_pusch_chest_scratch_bytes(const nrb200_pusch_chest_t *d) { return d ? pusch_chest_scratch_bytes(*d) : 0; }

NRB200_EXPORT int32_t nrb200_pusch_chest_dev(const nrb200_pusch_chest_t *d, const int16_t *d_rxF, int16_t *d_est, void *d_scratch, int32_t *d_state, void *stream)
{
  if (ensure_init() || !d) return -1;
  return launch_pusch_chest(*d, d_rxF, d_est, d_scratch, d_state, (cudaStream_t)stream, -1);
}

NRB200_EXPORT int32_t nrb200_chest_time_avg_dev(uint32_t fft_size, uint32_t nb_rx, uint32_t ch_stride, uint32_t start_symbol, uint32_t nr_of_symbols,
                                                uint32_t dmrs_symb_pos, uint32_t rb_size, int16_t *d_est, void *stream)
{
  if (ensure_init()) return -1;
  return launch_chest_time_avg(fft_size, nb_rx, ch_stride, start_symbol, nr_of_symbols, dmrs_symb_pos, rb_size, d_est, (cudaStream_t)stream);
}

NRB200_EXPORT int32_t nrb200_chest_time_avg_host(uint32_t fft_size, uint32_t nb_rx, uint32_t start_symbol, uint32_t nr_of_symbols, uint32_t dmrs_symb_pos,
                                                 uint32_t rb_size, int16_t *est)
{
  if (ensure_init() || !est) return -1;
  if (nb_rx < 1 || nb_rx > 64 || fft_size < 12) return -4;
  const size_t sym = (size_t)fft_size * 4, plane = 14 * sym, all = plane * nb_rx;
  Workspace *w = ctx().acquire();
  if (!w || !w->reserve(all, 16, 16)) { if (w) ctx().release(w); return -5; }
  int rc = 0;
  do {
    std::memcpy(w->h_in, est, all);
    if (cudaMemcpyAsync(w->d_in, w->h_in, all, cudaMemcpyHostToDevice, w->stream) != cudaSuccess) { rc = -2; break; }
    rc = launch_chest_time_avg(fft_size, nb_rx, 14 * fft_size, start_symbol, nr_of_symbols, dmrs_symb_pos, rb_size, (int16_t *)w->d_in, w->stream);
    if (rc < 0) break;
    const int first = rc;
    for (uint32_t a = 0; a < nb_rx; a++)      // only the first DMRS symbol changes
      if (cudaMemcpyAsync((uint8_t *)w->h_in + a * plane + first * sym, (uint8_t *)w->d_in + a * plane + first * sym, sym, cudaMemcpyDeviceToHost, w->stream) != cudaSuccess) { rc = -2; break; }
    if (rc < 0) break;
    if (cudaStreamSynchronize(w->stream) != cudaSuccess) { rc = -2; break; }
    for (uint32_t a = 0; a < nb_rx; a++) std::memcpy((uint8_t *)est + a * plane + first * sym, (uint8_t *)w->h_in + a * plane + first * sym, sym);
  } while (0);
  ctx().release(w);
  return rc;
}

NRB200_EXPORT int32_t nrb200_pusch_chest_host(const nrb200_pusch_chest_t *d, const int16_t *rxdataF, int16_t *ul_ch_estimates, int32_t *state5)
{
  if (ensure_init() || !d) return -1;
  if (d->nb_rx < 1 || d->nb_rx > 8 || d->symbol > 13 || d->n_ports > 2) return -4;
  const uint32_t np = d->n_ports == 0 ? 1 : d->n_ports;
  // only the DMRS symbol travels: [nb_rx][N + 4] in (4 c16 of the next symbol follow: the variants' pointer shift can read one of them),
  // [n_ports * nb_rx][N] out
  const size_t sym = (size_t)d->fft_size * 4, row = sym + 16, plane = sym * d->nb_rx, scratch = pusch_chest_scratch_bytes(*d);
  const int tail = d->symbol < 13 ? 4 : 0;
  const size_t seq_bytes = d->transform_precoding ? (size_t)24 * d->rb_size : 0;     // the caller's low-PAPR sequence travels behind the state
  if (d->transform_precoding && d->lowpapr_seq == 0) return -4;
  Workspace *w = ctx().acquire();
  if (!w || !w->reserve(row * d->nb_rx, plane * np, scratch + 256 + seq_bytes)) { if (w) ctx().release(w); return -5; }
  int rc = 0;
  do {
    for (uint32_t a = 0; a < d->nb_rx; a++) {
      std::memset((uint8_t *)w->h_in + row * a + sym, 0, 16);
      std::memcpy((uint8_t *)w->h_in + row * a, (const uint8_t *)rxdataF + ((size_t)a * 14 + d->symbol) * sym, sym + 4 * (size_t)tail);
    }
    if (cudaMemcpyAsync(w->d_in, w->h_in, row * d->nb_rx, cudaMemcpyHostToDevice, w->stream) != cudaSuccess) { rc = -2; break; }
    nrb200_pusch_chest_t e = *d;
    e.rx_stride = d->fft_size + 4; e.ch_stride = d->fft_size;                // staged as a one-symbol slot (buffer symbol 0)
    int32_t *d_state = (int32_t *)((uint8_t *)w->d_aux + scratch);
    if (seq_bytes) {
      std::memcpy((uint8_t *)w->h_aux + scratch + 256, reinterpret_cast<const void *>((uintptr_t)d->lowpapr_seq), seq_bytes);
      if (cudaMemcpyAsync((uint8_t *)w->d_aux + scratch + 256, (uint8_t *)w->h_aux + scratch + 256, seq_bytes, cudaMemcpyHostToDevice, w->stream) != cudaSuccess) { rc = -2; break; }
      e.lowpapr_seq = (uint64_t)(uintptr_t)((uint8_t *)w->d_aux + scratch + 256);
    }
    rc = launch_pusch_chest(e, (const int16_t *)w->d_in, (int16_t *)w->d_out, w->d_aux, d_state, w->stream, 0, tail);
    if (rc != 0) break;
    if (cudaMemcpyAsync(w->h_out, w->d_out, plane * np, cudaMemcpyDeviceToHost, w->stream) != cudaSuccess) { rc = -2; break; }
    if (cudaMemcpyAsync(w->h_aux, d_state, 18 * 4 * np, cudaMemcpyDeviceToHost, w->stream) != cudaSuccess) { rc = -2; break; }
    if (cudaStreamSynchronize(w->stream) != cudaSuccess) { rc = -2; break; }
    for (uint32_t q = 0; q < np; q++)
      for (uint32_t a = 0; a < d->nb_rx; a++)
        std::memcpy((uint8_t *)ul_ch_estimates + ((size_t)(q * d->nb_rx + a) * 14 + d->symbol) * sym, (uint8_t *)w->h_out + sym * (q * d->nb_rx + a), sym);
    if (state5) for (uint32_t q = 0; q < np; q++) std::memcpy(state5 + 5 * q, (int32_t *)w->h_aux + 18 * q, 20);
  } while (0);
  ctx().release(w);
  return rc;
}

// ------------------------------------------------------------------------------------------ offload calling convention
// Library-owned HARQ soft buffers, one per (ulsch_id, segment) like the accelerator's internal HARQ memory
// (harq_combined_input.offset = ulsch_id * 64 * LDPC_MAX_CB_SIZE + r * LDPC_MAX_CB_SIZE, nrLDPC_decoder_offload.c:545-546).
struct OffloadHarq { int16_t *d = nullptr; int dev = 0; std::mutex mu; };   // mu: the reference serialises calls on one buffer with decode_mutex
static std::mutex g_oh_mu;
static std::map<uint32_t, OffloadHarq> g_oh_store;                            // node based: entries never move
static OffloadHarq *offload_harq(uint8_t ulsch_id, uint8_t r)
{
  std::lock_guard<std::mutex> lk(g_oh_mu);
  const uint32_t key = ((uint32_t)ctx().dev << 16) | ((uint32_t)ulsch_id << 8) | r;   // a buffer lives on the device its (ulsch_id, r) is pinned to
  OffloadHarq &h = g_oh_store[key];
  if (h.d == nullptr) {
    if (cudaMalloc(&h.d, (size_t)66 * 384 * 2) != cudaSuccess) { h.d = nullptr; return nullptr; }
    cudaMemset(h.d, 0, (size_t)66 * 384 * 2);
    h.dev = ctx().dev;
  }
  return &h;
}
// frees every soft buffer of the offload convention (the _t2 library's LDPCshutdown; at most 65536 keys x 50 KB can accumulate otherwise)
NRB200_EXPORT int32_t nrb200_ldpc_offload_release(void)
{
  std::lock_guard<std::mutex> lk(g_oh_mu);
  for (auto &kv : g_oh_store) {
    std::lock_guard<std::mutex> lk2(kv.second.mu);
    if (kv.second.d) { cudaSetDevice(kv.second.dev); cudaFree(kv.second.d); kv.second.d = nullptr; }
  }
  cudaSetDevice(ctx().dev);
  return 0;
}

NRB200_EXPORT int32_t nrb200_ldpc_offload_init(void) { return LDPCinit(); }

NRB200_EXPORT int32_t nrb200_ldpc_offload_decode(const nrb200_ldpc_dec_params_t *p, uint8_t harq_pid, uint8_t ulsch_id, uint8_t r, const int8_t *llr,
                                                 uint8_t *out)
{
  (void)harq_pid;
  if (abi_devices() > 1) set_current_device(nrb200_sticky_device(ulsch_id, r, (uint32_t)abi_devices()));   // the segment's soft buffer lives there
  if (ensure_init() || !p) return -1;
  const int32_t numLLR = nrb200_ldpc_num_llr(p->BG, p->Z, p->R);
  if (numLLR < 0 || p->E <= 0 || p->E > 32768 * 4) return -4;
  nrb200_rm_desc_t rd;
  std::memset(&rd, 0, sizeof(rd));
  rd.BG = p->BG; rd.Z = p->Z; rd.Qm = p->Qm; rd.rv = p->rv; rd.clear = p->setCombIn ? 0 : 1; rd.C = 1; rd.Tbslbrm = 0; rd.F = p->F;
  rd.K = (uint32_t)(p->BG == 1 ? 22 : 10) * p->Z; rd.n_seg = 1;
  if (int rc = rm_check(&rd)) return rc;
  nrb200_ldpc_batch_desc_t d;
  std::memset(&d, 0, sizeof(d));
  const uint32_t kcZ = (uint32_t)(p->BG == 1 ? 68 : 52) * p->Z;
  d.BG = p->BG; d.Z = p->Z; d.R = p->R; d.numMaxIter = p->numMaxIter; d.outMode = NRB200_OUTMODE_BIT; d.n_cb = 1; d.llr_stride = kcZ;
  d.out_stride = ((uint32_t)numLLR + 7) / 8;
  const GraphDev *hg = nullptr;
  const GraphDev *dg = ctx().graph(d.BG, d.Z, d.R, &hg);
  if (!dg) return -4;
  DecodeArgs a;
  if (int rc = fill_args(&d, *hg, &a)) return rc;
  OffloadHarq *oh = offload_harq(ulsch_id, r);
  if (!oh) return -5;
  std::lock_guard<std::mutex> key_lock(oh->mu);
  int16_t *harq = oh->d;
  if (!harq) return -5;
  Workspace *w = ctx().acquire();
  if (!w || !w->reserve((size_t)p->E + 64, kcZ + d.out_stride + 64, 64)) { if (w) ctx().release(w); return -5; }
  int rc = 0;
  int32_t it = -1;
  do {
    std::memcpy(w->h_in, llr, (size_t)p->E);
    uint32_t *h_meta = (uint32_t *)w->h_aux;
    h_meta[0] = (uint32_t)p->E; h_meta[1] = 0;
    if (cudaMemcpyAsync(w->d_in, w->h_in, (size_t)p->E, cudaMemcpyHostToDevice, w->stream) != cudaSuccess) { rc = -2; break; }
    if (cudaMemcpyAsync(w->d_aux, w->h_aux, 8, cudaMemcpyHostToDevice, w->stream) != cudaSuccess) { rc = -2; break; }
    int8_t *d_llr8 = (int8_t *)w->d_out;
    uint8_t *d_hard = (uint8_t *)w->d_out + ((kcZ + 63) & ~63u);
    if ((rc = launch_rm_rx8(rd, (const int8_t *)w->d_in, (const uint32_t *)w->d_aux, (const uint32_t *)w->d_aux + 1, harq, 66 * 384, d_llr8, kcZ, w->stream)) != 0) break;
    a.llr = d_llr8; a.out = d_hard; a.iters = (int32_t *)((uint8_t *)w->d_aux + 16);
    if ((rc = launch_decode(dg, *hg, a, w->stream)) != 0) break;
    if (cudaMemcpyAsync(w->h_out, d_hard, rd.K / 8, cudaMemcpyDeviceToHost, w->stream) != cudaSuccess) { rc = -2; break; }
    if (cudaMemcpyAsync((uint8_t *)w->h_aux + 16, (uint8_t *)w->d_aux + 16, 4, cudaMemcpyDeviceToHost, w->stream) != cudaSuccess) { rc = -2; break; }
    if (cudaStreamSynchronize(w->stream) != cudaSuccess) { rc = -2; break; }
    std::memcpy(out, w->h_out, rd.K / 8);
    it = *(int32_t *)((uint8_t *)w->h_aux + 16);
  } while (0);
  ctx().release(w);
  return rc != 0 ? (rc < 0 ? rc : -1) : it;
}

NRB200_EXPORT int32_t nrb200_ldpc_offload_encode(const uint8_t *in, uint8_t *out, const nrb200_ldpc_enc_params_t *impp)
{
  if (ensure_init() || !impp || !in || !out) return -1;
  const int BG = impp->BG, Z = (int)impp->Zc;
  const EncGraphDev *hg = nullptr;
  const EncGraphDev *dg = ctx().enc_graph(BG, Z, &hg);
  if (!dg || impp->K != (uint32_t)hg->nsys * Z || impp->E == 0) return -4;
  nrb200_rm_desc_t rd;
  std::memset(&rd, 0, sizeof(rd));
  rd.BG = (uint8_t)BG; rd.Z = (uint16_t)Z; rd.Qm = impp->Qm; rd.rv = impp->rv; rd.C = 1; rd.Tbslbrm = 0; rd.F = impp->F; rd.K = impp->K; rd.n_seg = 1;
  if (int rc = rm_check(&rd)) return rc;
  const uint32_t nout = (uint32_t)(hg->ncols - 2) * Z, kin = impp->K / 8;
  Workspace *w = ctx().acquire();
  if (!w || !w->reserve(kin + 64, (size_t)nout + impp->E + 128, 64)) { if (w) ctx().release(w); return -5; }
  int rc = 0;
  do {
    std::memcpy(w->h_in, in, kin);
    uint32_t *h_meta = (uint32_t *)w->h_aux;
    h_meta[0] = impp->E; h_meta[1] = 0;
    if (cudaMemcpyAsync(w->d_in, w->h_in, kin, cudaMemcpyHostToDevice, w->stream) != cudaSuccess) { rc = -2; break; }
    if (cudaMemcpyAsync(w->d_aux, w->h_aux, 8, cudaMemcpyHostToDevice, w->stream) != cudaSuccess) { rc = -2; break; }
    uint8_t *d_d = (uint8_t *)w->d_out, *d_f = (uint8_t *)w->d_out + ((nout + 63) & ~63u);
    if ((rc = launch_encode(dg, *hg, (int)impp->K, 1, (const uint8_t *)w->d_in, kin, d_d, nout, w->stream)) != 0) break;
    if ((rc = launch_rm_tx(rd, d_d, nout, (const uint32_t *)w->d_aux, (const uint32_t *)w->d_aux + 1, d_f, w->stream)) != 0) break;
    if (cudaMemcpyAsync(w->h_out, d_f, impp->E, cudaMemcpyDeviceToHost, w->stream) != cudaSuccess) { rc = -2; break; }
    if (cudaStreamSynchronize(w->stream) != cudaSuccess) { rc = -2; break; }
    std::memcpy(out, w->h_out, impp->E);
  } while (0);
  ctx().release(w);
  return rc;
}
