// 4096-point Q15 DFT / IDFT (the 100 MHz, mu = 1 OFDM size: every symbol of BASELINE configs 3-5) as a persistent, TMA-fed kernel.
// Included by dfts.cu after the butterfly primitives; same arithmetic, same order of operations as dft_kernel (bit exact with
// oai_dfts.c:2553-2663 and the sub-transforms it calls), different data movement:
//
//   * persistent CTAs (one transform per iteration), the NEXT transform's 16 KB of IQ samples arrive by one bulk copy
//     (cp.async.bulk global -> shared, completion on an mbarrier) while the current one is computed, and the result leaves by a bulk store
//     (cp.async.bulk shared -> global) that the next iteration overlaps: no thread issues a global load or store for the samples
//   * three passes over shared memory, each with a layout that makes its loads AND its stores bank-conflict free (the generic kernel pays
//     8-way conflicts on the digit-reversed leaf loads and 4-way on the 64-byte leaf stores):
//       leaves  thread t reads x[t + 256 m] (consecutive threads, consecutive words), writes element j of its 16-point result to
//               plane j, slot 16 a + 80 b + g (a, b = the two radix-4 digits the next pass combines, g = the sub-transform) -- plane stride 304
//       pass A  (levels 64 and 256, fused) thread (g, k): reads plane k slots 16 a + 80 b + g, writes sub-transform g element k + 16 a + 64 b
//               to row g of a 16 x 258 array
//       pass B  (levels 1024 and 4096, fused) thread k': reads element k' of the 16 rows, writes y[k' + 256 a + 1024 b] in natural order
//   * the slot-level OFDM work stays fused: TX rotation on the leaf loads, RX rotation + time-shift on the last pass's stores, cyclic prefix
//     as a second bulk store of the tail of the same shared-memory buffer.
#pragma once

namespace dft4096 {

constexpr int kN = 4096, kXW = 16 * 258, kYW = 16 * 304, kStageWords = kXW + kYW;
constexpr size_t kSmemBytes = 2 * (size_t)kStageWords * 4;

__device__ __forceinline__ unsigned saddr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *b, unsigned count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(saddr(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *b, unsigned bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(saddr(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *b, unsigned parity)
{
  asm volatile(
      "{\n\t.reg .pred p;\n"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n"
      "DONE_%=:\n\t}" ::"r"(saddr(b)), "r"(parity) : "memory");
}
// global -> shared bulk copy (TMA, 1-D): 16-byte aligned addresses, size a multiple of 16
__device__ __forceinline__ void bulk_load(void *dst, const void *src, unsigned bytes, unsigned long long *b)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(saddr(dst)), "l"(src), "r"(bytes), "r"(saddr(b))
               : "memory");
}
__device__ __forceinline__ void bulk_store(void *dst, const void *src, unsigned bytes)
{
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(saddr(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ bool al16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace dft4096

// MODE as in dft_kernel: 0 = n plain transforms, contiguous; 1 = OFDM TX (rotation, IDFT, cyclic prefix); 2 = OFDM RX (window gather, DFT, rotation).
template <int MODE, bool INV>
__global__ void __launch_bounds__(256, 3) dft4096_kernel(DftPlan P, TwOffsets O, const short *__restrict__ tw, const unsigned *__restrict__ in,
                                                         unsigned *__restrict__ out, unsigned n, SlotIO S)
{
  using namespace dft4096;
  extern __shared__ __align__(128) unsigned sm[];
  __shared__ __align__(8) unsigned long long mbar[2];
  constexpr bool inv = INV;
  constexpr int N = kN;
  const int t = threadIdx.x;
  if (t == 0) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // where transform tr's samples come from / go to, and whether a bulk copy can carry them (16-byte alignment, no ring wrap)
  auto src_of = [&](unsigned tr) -> const unsigned * {
    if (MODE == 0) return in + (size_t)tr * N;
    const unsigned ant = tr / S.n_symb, l = tr - ant * S.n_symb;
    if (MODE == 1) return in + (size_t)ant * S.f_stride + (size_t)l * N;
    unsigned k = S.t_off[l];
    if (S.t_ring) k %= S.t_ring;
    return in + (size_t)ant * S.t_stride + k;
  };
  auto in_bulk = [&](unsigned tr) -> bool {
    if (MODE != 2) return true;                                   // alignment of the base and the strides is checked on the host
    const unsigned l = tr % S.n_symb;
    unsigned k = S.t_off[l];
    if (S.t_ring) { k %= S.t_ring; if (k + N > S.t_ring) return false; }
    return al16(src_of(tr));
  };
  unsigned ph0 = 0, ph1 = 0;
  if (t == 0 && blockIdx.x < n && in_bulk(blockIdx.x)) { mbar_expect_tx(&mbar[0], N * 4); bulk_load(sm, src_of(blockIdx.x), N * 4, &mbar[0]); }

  const uint2 *d16 = reinterpret_cast<const uint2 *>(tw + O.dps16[inv ? 1 : 0]), *d64 = reinterpret_cast<const uint2 *>(tw + O.dps64[inv ? 1 : 0]);
  unsigned it = 0;
  for (unsigned tr = blockIdx.x; tr < n; tr += gridDim.x, it++) {
    const int s = (int)(it & 1u);
    unsigned *X = sm + s * kStageWords, *Y = X + kXW;
    const unsigned ant = MODE == 0 ? 0u : tr / S.n_symb, l = MODE == 0 ? 0u : tr - ant * S.n_symb;
    // ---- next transform's samples: one bulk copy into the other stage (its last readers finished before the previous iteration's store barrier)
    const unsigned nxt = tr + gridDim.x;
    if (t == 0) {
      if (nxt < n && in_bulk(nxt)) { mbar_expect_tx(&mbar[s ^ 1], N * 4); bulk_load(sm + (s ^ 1) * kStageWords, src_of(nxt), N * 4, &mbar[s ^ 1]); }
      bulk_wait_read<1>();                                        // the store that read Y of this stage two iterations ago is done with it
    }
    // ---- this transform's samples
    if (in_bulk(tr)) {
      mbar_wait(&mbar[s], s ? ph1 : ph0);
      if (s) ph1 ^= 1u; else ph0 ^= 1u;
    } else {                                                      // unaligned window or ring wrap: gather by hand
      const unsigned ring = S.t_ring, k0 = ring ? S.t_off[l] % ring : S.t_off[l];
      const unsigned *base = in + (size_t)ant * S.t_stride;
      for (int i = t; i < N; i += 256) { unsigned k = k0 + i; if (ring && k >= ring) k -= ring; X[i] = base[k]; }
    }
    __syncthreads();
    // ---- leaves: thread t owns the 16-point transform of x[t + 256 m]
    {
      cx v[16];
#pragma unroll
      for (int m = 0; m < 16; m++) {
        unsigned w = X[t + 256 * m];
        if (MODE == 1 && S.rotate) {
          const int j = range_pos(S, (unsigned)(t + 256 * m));
          if (j >= 0) w = rotate_c16(w, S.rot[l][0], S.rot[l][1], (unsigned)j < (S.r_len & ~7u));
        }
        v[m] = unpack(w);
      }
      cx A[4][4];
#pragma unroll
      for (int c = 0; c < 4; c++) bfly4_sat(v[c], v[4 + c], v[8 + c], v[12 + c], inv, A[0][c], A[1][c], A[2][c], A[3][c]);
      // digits of t: t = d1 + 4 d2 + 16 d3 + 64 d4; the leaf's natural slot is d4 + 4 d3 + 16 d2 + 64 d1 = a + 4 b + 16 g
      const int slot = 16 * (t >> 6) + 80 * ((t >> 4) & 3) + ((t >> 2) & 3) + 4 * (t & 3);
#pragma unroll
      for (int k1 = 0; k1 < 4; k1++) {
        cx b1 = cmult2dp(pack(A[k1][1]), d16 + k1), b2 = cmult2dp(pack(A[k1][2]), d16 + 4 + k1), b3 = cmult2dp(pack(A[k1][3]), d16 + 8 + k1);
        cx y0, y1, y2, y3;
        bfly4_sat(A[k1][0], b1, b2, b3, inv, y0, y1, y2, y3);
        Y[(k1)*304 + slot] = pack(y0); Y[(k1 + 4) * 304 + slot] = pack(y1); Y[(k1 + 8) * 304 + slot] = pack(y2); Y[(k1 + 12) * 304 + slot] = pack(y3);
      }
    }
    __syncthreads();
    // ---- pass A: levels 64 and 256.  Thread (g, k) holds elements k + 16 a + 64 b of sub-transform g
    {
      const int g = t & 15, k = t >> 4;
      const unsigned *p = Y + k * 304 + g;
      unsigned *q = X + g * 258 + k;
      cx v[4][4];
#pragma unroll
      for (int b = 0; b < 4; b++) {
        // level 64 (saturating butterfly, hand-rounded table pair, >> 3): x0 unpacked for the adds, x1..x3 stay packed for the dot products
        const cx x0 = unpack(p[80 * b]);
        const cx a1 = cmult2dp(p[16 + 80 * b], d64 + k), a2 = cmult2dp(p[32 + 80 * b], d64 + 16 + k), a3 = cmult2dp(p[48 + 80 * b], d64 + 32 + k);
        bfly4_sat(x0, a1, a2, a3, inv, v[0][b], v[1][b], v[2][b], v[3][b]);
#pragma unroll
        for (int a = 0; a < 4; a++) { v[a][b].r >>= 3; v[a][b].i >>= 3; }
      }
      if (inv) {                                            // inverse: level 256 is the 32-bit butterfly -> packed form, straight to shared memory
        const uint2 *t4 = reinterpret_cast<const uint2 *>(tw + O.dp256i);
#pragma unroll
        for (int a = 0; a < 4; a++) {
          const int kk = k + 16 * a;
          unsigned y0, y1, y2, y3;
          bfly4_32dp(pack(v[a][0]), pack(v[a][1]), pack(v[a][2]), pack(v[a][3]), t4 + kk, t4 + 64 + kk, t4 + 128 + kk, true, true, y0, y1, y2, y3);
          q[16 * a] = y0; q[16 * a + 64] = y1; q[16 * a + 128] = y2; q[16 * a + 192] = y3;
        }
      } else {
        const uint2 *d256 = reinterpret_cast<const uint2 *>(tw + O.dps256f);
#pragma unroll
        for (int a = 0; a < 4; a++) {                        // forward level 256: saturating butterfly, >> 1
          const int kk = k + 16 * a;
          cx y0, y1, y2, y3;
          bfly4_sat(v[a][0], cmult2dp(pack(v[a][1]), d256 + kk), cmult2dp(pack(v[a][2]), d256 + 64 + kk), cmult2dp(pack(v[a][3]), d256 + 128 + kk), false, y0, y1, y2, y3);
          y0.r >>= 1; y0.i >>= 1; y1.r >>= 1; y1.i >>= 1; y2.r >>= 1; y2.i >>= 1; y3.r >>= 1; y3.i >>= 1;
          q[16 * a] = pack(y0); q[16 * a + 64] = pack(y1); q[16 * a + 128] = pack(y2); q[16 * a + 192] = pack(y3);
        }
      }
    }
    __syncthreads();
    // ---- pass B: levels 1024 and 4096 (32-bit butterflies on packed words).  Thread k' holds element k' of the 16 sub-transforms; results in natural order
    {
      const int k = t;
      unsigned v[4][4];
#pragma unroll
      for (int b = 0; b < 4; b++)
#pragma unroll
        for (int a = 0; a < 4; a++) v[a][b] = X[(a + 4 * b) * 258 + k];
      const uint2 *t1 = reinterpret_cast<const uint2 *>(tw + O.dp1024[inv ? 1 : 0]), *t2 = reinterpret_cast<const uint2 *>(tw + O.dp4096[inv ? 1 : 0]);
#pragma unroll
      for (int b = 0; b < 4; b++)
        bfly4_32dp(v[0][b], v[1][b], v[2][b], v[3][b], t1 + k, t1 + 256 + k, t1 + 512 + k, inv, true, v[0][b], v[1][b], v[2][b], v[3][b]);
#pragma unroll
      for (int a = 0; a < 4; a++) {
        const int kk = k + 256 * a;
        bfly4_32dp(v[a][0], v[a][1], v[a][2], v[a][3], t2 + kk, t2 + 1024 + kk, t2 + 2048 + kk, inv, P.scale != 0, v[a][0], v[a][1], v[a][2], v[a][3]);
      }
#pragma unroll
      for (int b = 0; b < 4; b++)
#pragma unroll
        for (int a = 0; a < 4; a++) {
          const unsigned i = (unsigned)(k + 256 * a + 1024 * b);
          unsigned w = v[a][b];
          if (MODE == 2 && S.rotate) {
            const int j = range_pos(S, i);
            if (j >= 0) {
              w = rotate_c16(w, S.rot[l][0], wrap16(-S.rot[l][1]), (unsigned)j < (S.r_len & ~7u));
              if ((unsigned)j < (S.r_len & ~3u)) w = mult_c16(w, __ldg(S.timeshift + i));
            }
          }
          Y[i] = w;
        }
    }
    // ---- result out: bulk store(s) from Y when the destination is 16-byte aligned, by hand otherwise
    fence_async_smem();
    __syncthreads();
    if (MODE == 1) {
      unsigned *o = out + (size_t)ant * S.t_stride + S.t_off[l];
      const unsigned cp = S.prefix[l];
      if (al16(o) && (cp & 3u) == 0u) {
        if (t == 0) {
          bulk_store(o + cp, Y, N * 4);
          if (cp) bulk_store(o, Y + (N - cp), cp * 4);            // cyclic prefix = the last cp samples
          bulk_commit();
        }
      } else {
        for (int i = t; i < N; i += 256) {
          const unsigned w = Y[i];
          o[cp + i] = w;
          if ((unsigned)i >= (unsigned)N - cp) o[i - ((unsigned)N - cp)] = w;
        }
        if (t == 0) bulk_commit();                                // keeps the group count in step with the iteration count
      }
    } else {
      unsigned *o = MODE == 0 ? out + (size_t)tr * N : out + (size_t)ant * S.f_stride + (size_t)l * N;
      if (t == 0) { bulk_store(o, Y, N * 4); bulk_commit(); }
    }
  }
  if (t == 0) bulk_wait_read<0>();                                // shared memory must outlive the last store's reads
}
