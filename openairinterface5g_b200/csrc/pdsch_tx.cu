// gNB PDSCH transmitter after the encoder, one launch per PDSCH: scrambling, QAM mapping, layer mapping, DMRS generation, resource mapping with the
// reference's amplitude scaling, identity precoding -- straight from the rate-matched bits (one per byte) into txdataF.
// Reference: openair1/PHY/NR_TRANSPORT/nr_dlsch.c nr_generate_pdsch :56-583 (scrambling :160, modulation :175, nr_layer_mapping :192, resource mapping
// :236-478, precoding :490-530), DMRS sequence nr_init_pdsch_dmrs NR_REFSIG/nr_gold.c:78-96, port tables NR_TRANSPORT/nr_sch_dmrs.c:35-100,
// allowed_xlsch_re_in_dmrs_symbol NR_REFSIG/dmrs_nr.c:37-62.
// The reference makes five passes over five buffers (scrambled words, mod_symbs, tx_layers, txdataF_precoding, txdataF).  Here one thread owns one
// sub-carrier of one symbol for all layers: it computes which modulation symbol lands there (closed forms for the reference's running counters m, dmrs_idx,
// k', n), reads the Qm bits per layer, XORs the Gold bits (jump-ahead words staged per CTA), looks the symbol up, scales it exactly as the reference does
// (mulhrs in groups of four per contiguous piece, the doubled-amplitude leftovers at the end of a piece, truncation in DMRS symbols) and writes txdataF once.
#include <algorithm>
#include "nrb200_ctx.h"
#include "gold_seq.cuh"
#include "../../include/nrb200_ldpc.h"

namespace nrb200 {

struct PdschTxGeom {
  int N, nb_tx, nl, Qm, type, cdm, amp, start_sc, nb_re, upper, rem;
  unsigned tx_stride, c_init;
  int n_sym, sym[14], is_dmrs[14];
  unsigned m_base[14], dmrs_cinit[14];
  int delta[4], wf1[4];            // per layer: DMRS comb offset and Wf(1) (Wf(0) = Wt = 1 for the supported ports)
  int dmrs_idx0;
  unsigned ptrs_pos;               // PT-RS symbols (set_ptrs_symb_idx); 0: no PT-RS
  int ptrs_K12, ptrs_q0, ptrs_n;   // PT-RS REs of such a symbol: q0 + j * K12, j < n (is_ptrs_subcarrier relative to the allocation's first sub-carrier)
  int pm;                          // > 0: wideband non-identity precoding with pmw (one PRG over the allocation)
  short pmw[4][4][2];              // nfapi_nr_pm_pdu_t.weights[layer][antenna] {Re, Im}
};

__device__ __forceinline__ int t_wrap16(int v) { return (int)(short)v; }

template <int QM>
__global__ void __launch_bounds__(256) pdsch_tx_kernel(PdschTxGeom G, const GoldTables *__restrict__ T, const uint32_t *__restrict__ modtab, const uint8_t *__restrict__ f,
                                                       unsigned *__restrict__ txF)
{
  __shared__ uint32_t s_gold[(256 * 4 * QM) / 32 + 2];
  __shared__ uint32_t s_dmrs[12];
  const int k = blockIdx.y, symbol = G.sym[k], is_dmrs = G.is_dmrs[k];
  const bool is_ptrs = (G.ptrs_pos >> symbol) & 1u;                               // never a DMRS symbol
  const int i0 = blockIdx.x * 256, i = i0 + threadIdx.x;
  const uint32_t *tab = modtab + (QM == 2 ? 0 : QM == 4 ? 4 : QM == 6 ? 20 : 84);
  // PT-RS REs of this symbol below index j
  auto ptrs_below = [&](int j) -> int { return j <= G.ptrs_q0 ? 0 : min(G.ptrs_n, (j - G.ptrs_q0 - 1) / G.ptrs_K12 + 1); };
  // number of data REs of this symbol below index j (per layer)
  auto rank_below = [&](int j) -> int {
    if (is_ptrs) return j - ptrs_below(j);
    if (!is_dmrs) return j;
    if (G.type == 0) return G.cdm == 1 ? (j >> 1) : 0;
    const int g6 = j / 6, r = j - 6 * g6;
    return g6 * (6 - 2 * G.cdm) + max(0, r - 2 * G.cdm);
  };
  const unsigned bit0 = (G.m_base[k] + (unsigned)rank_below(i0)) * (unsigned)(G.nl * QM);
  {
    const unsigned w0 = bit0 >> 5, nw = ((bit0 + 256u * (unsigned)(G.nl * QM) + 31u) >> 5) - w0;
    for (unsigned w = threadIdx.x; w < nw; w += 256) s_gold[w] = gold_word(T, G.c_init, w0 + w);
  }
  unsigned dw0 = 0;
  if (is_dmrs) {
    const int j0 = G.type == 0 ? (i0 >> 1) : 2 * (i0 / 6);
    dw0 = (2u * (unsigned)(G.dmrs_idx0 + j0)) >> 5;
    if (threadIdx.x < 12) s_dmrs[threadIdx.x] = gold_word(T, G.dmrs_cinit[k], dw0 + threadIdx.x);
  } else if (is_ptrs) {                                                            // the pilots are the first 2 n_ptrs bits of the symbol's DMRS sequence (:296)
    dw0 = (2u * (unsigned)ptrs_below(i0)) >> 5;
    if (threadIdx.x < 12) s_dmrs[threadIdx.x] = gold_word(T, G.dmrs_cinit[k], dw0 + threadIdx.x);
  }
  __syncthreads();
  if (i >= G.nb_re) return;
  unsigned lay[4] = {0u, 0u, 0u, 0u};                                              // this RE's value on every layer (txdataF_precoding[layer][symbol][k])
  int kk = G.start_sc + i;
  if (kk >= G.N) kk -= G.N;
  const size_t o = (size_t)symbol * G.N + kk;
  // modulation symbol number c of the code word: Qm bits (one per byte) XOR Gold bits -> table
  auto modsym = [&](unsigned c) -> unsigned {
    const unsigned b = c * QM, rel = b - ((bit0 >> 5) << 5);
    unsigned idx = 0;
#pragma unroll
    for (int q = 0; q < QM; q++) idx |= (unsigned)(__ldg(f + b + q) & 1u) << q;
    const unsigned long long g = ((unsigned long long)s_gold[(rel >> 5) + 1] << 32) | s_gold[rel >> 5];
    idx ^= (unsigned)(g >> (rel & 31u)) & ((1u << QM) - 1u);
    return __ldg(tab + idx);
  };
  if (is_ptrs) {
    // PT-RS symbol (:300-376): the per-RE branch of the reference -- pilots on every layer, data with the truncating scaling
    const int d = i - G.ptrs_q0;
    const bool pilot = d >= 0 && d % G.ptrs_K12 == 0;
    unsigned v;
    if (pilot) {
      const unsigned b = 2u * (unsigned)(d / G.ptrs_K12), rel = b - (dw0 << 5);
      const unsigned x = __ldg(modtab + ((s_dmrs[rel >> 5] >> (rel & 31u)) & 3u));
      v = ((unsigned)t_wrap16(((int)(short)(x & 0xFFFFu) * G.amp) >> 15) & 0xFFFFu) | ((unsigned)t_wrap16(((int)(short)(x >> 16) * G.amp) >> 15) << 16);
    }
    const unsigned m = G.m_base[k] + (unsigned)rank_below(i);
#pragma unroll
    for (int l = 0; l < 4; l++) {
      if (l >= G.nl) continue;
      if (!pilot) {
        const unsigned x = modsym(m * G.nl + l);
        v = ((unsigned)t_wrap16(((int)(short)(x & 0xFFFFu) * G.amp) >> 15) & 0xFFFFu) | ((unsigned)t_wrap16(((int)(short)(x >> 16) * G.amp) >> 15) << 16);
      }
      lay[l] = v;
    }
  } else if (!is_dmrs) {
    const unsigned m = G.m_base[k] + (unsigned)i;
    const int j = i < G.upper ? i : i - G.upper, len = i < G.upper ? G.upper : G.rem;
    const bool body = j < (len & ~3);
#pragma unroll
    for (int l = 0; l < 4; l++) {
      if (l < G.nl) {
        const unsigned x = modsym(m * G.nl + l);
        const int xr = (int)(short)(x & 0xFFFFu), xi = (int)(short)(x >> 16);
        const int r = body ? t_wrap16((xr * G.amp + 0x4000) >> 15) : t_wrap16(((xr * G.amp) >> 14) + 1);
        const int im = body ? t_wrap16((xi * G.amp + 0x4000) >> 15) : t_wrap16(((xi * G.amp) >> 14) + 1);
        lay[l] = ((unsigned)r & 0xFFFFu) | ((unsigned)im << 16);
      }
    }
  } else {
    int g6 = 0, r;
    if (G.type == 0) r = i & 1; else { g6 = i / 6; r = i - 6 * g6; }
    const bool data_ok = G.type == 0 ? r >= G.cdm : r >= 2 * G.cdm;
    const unsigned m = G.m_base[k] + (unsigned)rank_below(i);
#pragma unroll
    for (int l = 0; l < 4; l++) {
      if (l >= G.nl) continue;
      const int d = r - G.delta[l];
      unsigned v = 0;
      if (G.type == 0 ? d == 0 : (d == 0 || d == 1)) {
        const int j = G.type == 0 ? (i >> 1) : 2 * g6 + d, kp = G.type == 0 ? (j & 1) : d;
        const unsigned b = 2u * (unsigned)(G.dmrs_idx0 + j), rel = b - (dw0 << 5);
        const unsigned x = __ldg(modtab + ((s_dmrs[rel >> 5] >> (rel & 31u)) & 3u));
        const int w = (kp ? G.wf1[l] : 1) * G.amp;
        v = ((unsigned)t_wrap16(((int)(short)(x & 0xFFFFu) * w) >> 15) & 0xFFFFu) | ((unsigned)t_wrap16(((int)(short)(x >> 16) * w) >> 15) << 16);
      } else if (data_ok) {
        const unsigned x = modsym(m * G.nl + l);
        v = ((unsigned)t_wrap16(((int)(short)(x & 0xFFFFu) * G.amp) >> 15) & 0xFFFFu) | ((unsigned)t_wrap16(((int)(short)(x >> 16) * G.amp) >> 15) << 16);
      }
      lay[l] = v;
    }
  }
  if (G.pm == 0) {                                                                 // identity precoding: layer l -> antenna l, antennas beyond the layers are zeroed
#pragma unroll
    for (int l = 0; l < 4; l++) if (l < G.nl) txF[(size_t)l * G.tx_stride + o] = lay[l];
    for (int a = G.nl; a < G.nb_tx; a++) txF[(size_t)a * G.tx_stride + o] = 0;
    return;
  }
  // Non-identity precoding (nr_dlsch.c:536-590): with one PRG the reference takes the RBs two at a time (the last one alone when rb_size is odd).  A group
  // that ends below the symbol's last sub-carrier runs nr_layer_precoder_simd -- per-layer products truncated to 16 bits, SATURATING sum over the layers --
  // any other group nr_layer_precoder_cm, whose c16maddShift sum WRAPS (MODULATION/nr_modulation.c:702-815).
  const int g = i / 24, cnt = (G.nb_re - 24 * g) >= 24 ? 24 : 12;
  int sc_g = G.start_sc + 24 * g;
  if (sc_g >= G.N) sc_g -= G.N;
  const bool wraps = sc_g + cnt >= G.N;
  for (int a = 0; a < G.nb_tx; a++) {
    int yr = 0, yi = 0;
#pragma unroll
    for (int l = 0; l < 4; l++) {
      if (l < G.nl) {
        const int xr = (int)(short)(lay[l] & 0xFFFFu), xi = (int)(short)(lay[l] >> 16), wr = G.pmw[l][a][0], wi = G.pmw[l][a][1];
        const int pr = t_wrap16((xr * wr - xi * wi) >> 15), pi = t_wrap16((xr * wi + xi * wr) >> 15);
        if (wraps) { yr = t_wrap16(yr + pr); yi = t_wrap16(yi + pi); }
        else { yr = max(-32768, min(32767, yr + pr)); yi = max(-32768, min(32767, yi + pi)); }
      }
    }
    txF[(size_t)a * G.tx_stride + o] = ((unsigned)yr & 0xFFFFu) | ((unsigned)yi << 16);
  }
}

static int make_tx_geom(const nrb200_pdsch_tx_t &d, PdschTxGeom *G, uint32_t *n_bits)
{
  static const int8_t grp1[4] = {0, 0, 1, 1}, dl1[4] = {0, 0, 1, 1}, wf1[4] = {1, -1, 1, -1};
  static const int8_t grp2[6] = {0, 0, 1, 1, 2, 2}, dl2[6] = {0, 0, 2, 2, 4, 4}, wf2[6] = {1, -1, 1, -1, 1, -1};
  const int Qm = d.qam_mod_order, nl = d.nrOfLayers, type = d.dmrs_config_type, cdm = d.num_dmrs_cdm_grps_no_data;
  if ((Qm != 2 && Qm != 4 && Qm != 6 && Qm != 8) || nl < 1 || nl > 4 || d.nb_tx < (uint32_t)nl || d.nb_tx > 8 || type > 1 || cdm < 1 || cdm > (type == 0 ? 2 : 3) ||
      d.rb_size < 1 || d.fft_size < 12 * d.rb_size || d.nr_of_symbols < 1 || d.start_symbol_index + d.nr_of_symbols > 14 || d.slot > 159 || d.scid > 1)
    return -4;
  G->N = d.fft_size; G->nb_tx = d.nb_tx; G->nl = nl; G->Qm = Qm; G->type = type; G->cdm = cdm; G->amp = (int16_t)d.amp; G->nb_re = 12 * d.rb_size;
  int sc = d.first_carrier_offset + (d.rb_start + d.bwp_start) * 12;
  if (sc >= (int)d.fft_size) sc -= d.fft_size;
  if (sc < 0 || sc >= (int)d.fft_size) return -4;
  G->start_sc = sc;
  G->upper = G->nb_re; G->rem = 0;
  if (sc + G->nb_re > G->N) { G->rem = G->nb_re + sc - G->N; G->upper = G->N - sc; }
  G->tx_stride = d.tx_stride; G->c_init = (d.rnti << 15) + d.data_scrambling_id;
  G->dmrs_idx0 = (d.rb_start + d.bwp_start) * (type == 0 ? 6 : 4);
  G->pm = (int)d.pm_idx;
  if (G->pm > 0) {
    if (d.nb_tx > 4 || d.nb_tx < 2) return -4;                              // weights[4][4]; "No precoding can be done with a single antenna port"
    for (int l = 0; l < 4; l++) for (int a = 0; a < 4; a++) { G->pmw[l][a][0] = d.pm_weights[l][a][0]; G->pmw[l][a][1] = d.pm_weights[l][a][1]; }
  }
  for (int l = 0; l < nl; l++) {
    int port = 0;
    if (d.dmrs_ports) { int found = -1; port = -1; for (int i = 0; i < 12; i++) if ((d.dmrs_ports >> i) & 1) { if (++found == l) { port = i; break; } } }   // get_dmrs_port
    if (port < 0 || port >= (type == 0 ? 4 : 6)) return -4;                        // double-symbol ports (Wt = -1) are not covered
    const int grp = type == 0 ? grp1[port] : grp2[port];
    G->delta[l] = type == 0 ? dl1[port] : dl2[port]; G->wf1[l] = type == 0 ? wf1[port] : wf2[port];
    // the port's own CDM group must be one without data; and the configuration in which the reference maps one RE too many and reads beyond its
    // modulation buffer (type 2, two groups, delta != 0, fft_size % 6 == 4: allowed_xlsch_re_in_dmrs_symbol admits k == start_sc) is refused
    if (grp >= cdm || (type == 1 && G->delta[l] != 0 && cdm == 2 && d.fft_size % 6 == 4)) return -4;
  }
  G->ptrs_pos = 0; G->ptrs_K12 = 24; G->ptrs_q0 = 0; G->ptrs_n = 0;
  if (d.ptrs) {                                                                   // :98-111; set_ptrs_symb_idx / is_ptrs_subcarrier of NR_REFSIG/ptrs_nr.c
    const int K = (int)d.ptrs_freq_density, nb = (int)d.rb_size;
    if ((K != 2 && K != 4) || d.ptrs_time_density > 2 || d.ptrs_re_offset >= 12) return -4;
    const int L = 1 << d.ptrs_time_density, last = (int)(d.start_symbol_index + d.nr_of_symbols) - 1;
    int i = 0, l_ref = (int)d.start_symbol_index;
    while (l_ref + i * L <= last) {
      int hit = -1;
      for (int l = l_ref + i * L; l >= std::max(l_ref + (i - 1) * L + 1, l_ref); l--) if ((d.dl_dmrs_symb_pos >> l) & 1u) { hit = l; break; }
      if (hit >= 0) { l_ref = hit; i = 1; continue; }
      G->ptrs_pos |= 1u << (l_ref + i * L);
      i++;
    }
    const int k_rb_ref = (nb % K == 0) ? (int)(d.rnti & 0xFFFFu) % K : (int)(d.rnti & 0xFFFFu) % (nb % K);
    G->ptrs_K12 = 12 * K; G->ptrs_q0 = (int)d.ptrs_re_offset + 12 * k_rb_ref; G->ptrs_n = (nb + K - 1) / K;
  }
  const int per_dmrs = G->nb_re - d.rb_size * cdm * (type == 0 ? 6 : 4);
  unsigned m = 0;
  G->n_sym = 0;
  for (uint32_t s = d.start_symbol_index; s < d.start_symbol_index + d.nr_of_symbols; s++) {
    const int dm = (d.dl_dmrs_symb_pos >> s) & 1, kx = G->n_sym++;
    G->sym[kx] = s; G->is_dmrs[kx] = dm; G->m_base[kx] = m;
    const unsigned long long x2 = (1ULL << 17) * (14ULL * d.slot + s + 1) * (((unsigned long long)d.dl_dmrs_scrambling_id << 1) + 1) + (((unsigned long long)d.dl_dmrs_scrambling_id << 1) + d.scid);
    G->dmrs_cinit[kx] = (unsigned)(x2 % (1ULL << 31));
    m += dm ? per_dmrs : G->nb_re - (((G->ptrs_pos >> s) & 1u) ? G->ptrs_n : 0);
  }
  // the reference derives the length from the whole dlDmrsSymbPos mask (get_num_dmrs); DMRS symbols outside the allocation would desynchronise it
  int n_mask = 0, n_in = 0;
  for (int s = 0; s < 14; s++) { n_mask += (d.dl_dmrs_symb_pos >> s) & 1; n_in += ((d.dl_dmrs_symb_pos >> s) & 1) && s >= (int)d.start_symbol_index && s < (int)(d.start_symbol_index + d.nr_of_symbols); }
  if (n_mask != n_in) return -4;
  if (n_bits) *n_bits = m * (unsigned)(nl * Qm);
  return 0;
}

uint32_t pdsch_tx_num_bits(const nrb200_pdsch_tx_t &d)
{
  PdschTxGeom G;
  uint32_t n = 0;
  return make_tx_geom(d, &G, &n) == 0 ? n : 0;
}

int launch_pdsch_tx(const nrb200_pdsch_tx_t &d, const uint8_t *f, int16_t *txF, cudaStream_t st)
{
  PdschTxGeom G;
  int rc = make_tx_geom(d, &G, nullptr);
  if (rc) return rc;
  if (scramble_mod_init() != 0) return -5;
  const dim3 grid((G.nb_re + 255) / 256, G.n_sym);
  switch (G.Qm) {
    case 2: pdsch_tx_kernel<2><<<grid, 256, 0, st>>>(G, gold_tables_dev(), mod_tables_dev(), f, (unsigned *)txF); break;
    case 4: pdsch_tx_kernel<4><<<grid, 256, 0, st>>>(G, gold_tables_dev(), mod_tables_dev(), f, (unsigned *)txF); break;
    case 6: pdsch_tx_kernel<6><<<grid, 256, 0, st>>>(G, gold_tables_dev(), mod_tables_dev(), f, (unsigned *)txF); break;
    default: pdsch_tx_kernel<8><<<grid, 256, 0, st>>>(G, gold_tables_dev(), mod_tables_dev(), f, (unsigned *)txF); break;
  }
  ctx().launches++;
  NRB200_CUDA_OK(cudaGetLastError(), "pdsch_tx launch");
  return 0;
}

}  // namespace nrb200
