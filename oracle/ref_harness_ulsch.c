/* TEST INFRASTRUCTURE ONLY -- never linked into the product.
 * The reference-side CALLER of nr_ulsch_decoding (openair1/PHY/NR_TRANSPORT/nr_ulsch_decoding.c:320): a PHY_VARS_gNB with the thread pool and the response
 * FIFO the way init_gNB_Tpool leaves them, one ULSCH from the reference's own new_gNB_ulsch, a PDU as the scheduler fills it; it calls the function by name
 * and then pulls the per-segment results off gNB->respDecode like phy_procedures_gNB_uespec_RX does (SCHED_NR/phy_procedures_nr_gNB.c:905-925), including
 * nr_postDecode's copy of the decoded segments into the transport block.
 * Built twice (oracle/build_ref.sh, integration/build_shims.sh):
 *   oracle/_ref/libref_ulsch.so        with the reference's nr_ulsch_decoding and the compiled reference decoder (oai_libs/libldpc.so) behind ldpc_interface
 *   oracle/_ref/libshimtest_ulsch.so   with integration/oai_shim_ulsch_decoding.c linked ahead of the same object: the call lands in libldpc_b200.so */
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "PHY/defs_gNB.h"
#include "PHY/CODING/nrLDPC_extern.h"
#include "PHY/CODING/coding_defs.h"
#include "PHY/NR_TRANSPORT/nr_transport_proto.h"
#include "PHY/NR_TRANSPORT/nr_ulsch.h"

ldpc_interface_t ldpc_interface, ldpc_interface_offload;
NR_gNB_PHY_STATS_t *get_phy_stats(PHY_VARS_gNB *gNB, uint16_t rnti) { (void)gNB; (void)rnti; return NULL; }

/* binds ldpc_interface to a codec library by path, the lookups of load_LDPClib (nrLDPC_load.c:46-71); returns LDPCinit's value */
int refh_ulsch_bind_ldpc(const char *so)
{
  void *h = dlopen(so, RTLD_NOW | RTLD_LOCAL);
  if (!h) { fprintf(stderr, "refh_ulsch_bind_ldpc: %s\n", dlerror()); return -1; }
  ldpc_interface.LDPCinit = (LDPC_initfunc_t *)dlsym(h, "LDPCinit");
  ldpc_interface.LDPCshutdown = (LDPC_shutdownfunc_t *)dlsym(h, "LDPCshutdown");
  ldpc_interface.LDPCdecoder = (LDPC_decoderfunc_t *)dlsym(h, "LDPCdecoder");
  ldpc_interface.LDPCencoder = (LDPC_encoderfunc_t *)dlsym(h, "LDPCencoder");
  if (!ldpc_interface.LDPCdecoder) return -2;
  return ldpc_interface.LDPCinit ? ldpc_interface.LDPCinit() : 0;
}

enum { U_N_RB_UL, U_RB_SIZE, U_QM, U_NL, U_TBS_BYTES, U_RV, U_BG, U_TBSLBRM, U_MAX_ITER, U_NEW_DATA, U_ROUND, U_NB_RX, U_COUNT };

#define REFH_MAX_CTX 64
static PHY_VARS_gNB *g_gNBs[REFH_MAX_CTX];          /* one gNB instance per calling thread of the multi-UE benchmark */
static int g_n_rb_ul[REFH_MAX_CTX];

/* One PUSCH through nr_ulsch_decoding + the caller's collection loop.  llr: G int16.  Outputs: info[0] = C, [1] = K, [2] = Z, [3] = F, [4] = llrLen;
 * iters[r]; c_out: C x (K / 8) bytes (harq_process->c[r]); tb_out: the transport block as nr_postDecode assembles it (harq_process->b, TBS + 3 bytes);
 * d_out (optional): C x 66 Z | 50 Z int16 (harq_process->d[r]).  p[U_NEW_DATA] = 1 starts a new transport block (harq_to_be_cleared), 0 combines.
 * Returns nr_ulsch_decoding's return value. */
int refh_ulsch_decode_ctx(int ctx, const int32_t *p, const int16_t *llr, int G, int32_t *info, int32_t *iters, uint8_t *c_out, uint8_t *tb_out, int16_t *d_out)
{
  if (ctx < 0 || ctx >= REFH_MAX_CTX) return -101;
  if (!g_gNBs[ctx]) {
    PHY_VARS_gNB *g_gNB = g_gNBs[ctx] = calloc(1, sizeof(*g_gNB));
    crcTableInit();                                                        /* phy_init_nr_gNB does (PHY/INIT/nr_init.c) */
    initNamedTpool("n", &g_gNB->threadPool, false, "gNB-tpool");          /* no worker threads: jobs run in pushTpool, like nr_ulsim without -C */
    initNotifiedFIFO(&g_gNB->respDecode);
    g_gNB->ulsch = calloc(1, sizeof(NR_gNB_ULSCH_t));
    g_gNB->pusch_vars = calloc(1, sizeof(NR_gNB_PUSCH));
  }
  PHY_VARS_gNB *gNB = g_gNBs[ctx];
  int n_rb_ul = g_n_rb_ul[ctx];
  if (n_rb_ul != p[U_N_RB_UL]) {                                          /* the ULSCH's segment buffers are sized by the carrier (init_nr_transport) */
    if (n_rb_ul) free_gNB_ulsch(&gNB->ulsch[0], (uint16_t)n_rb_ul);
    gNB->ulsch[0] = new_gNB_ulsch((uint8_t)p[U_MAX_ITER], (uint16_t)p[U_N_RB_UL]);
    gNB->ulsch[0].rnti = 0x1234;
    g_n_rb_ul[ctx] = p[U_N_RB_UL];
  }
  NR_DL_FRAME_PARMS *fp = &gNB->frame_parms;
  fp->nb_antennas_rx = p[U_NB_RX]; fp->N_RB_UL = p[U_N_RB_UL];
  NR_gNB_ULSCH_t *ulsch = &gNB->ulsch[0];
  NR_UL_gNB_HARQ_t *hp = ulsch->harq_process;
  ulsch->max_ldpc_iterations = (uint8_t)p[U_MAX_ITER];
  hp->round = (uint8_t)p[U_ROUND];
  if (p[U_NEW_DATA]) hp->harq_to_be_cleared = true;
  nfapi_nr_pusch_pdu_t pdu;
  memset(&pdu, 0, sizeof(pdu));
  pdu.rb_size = p[U_RB_SIZE]; pdu.qam_mod_order = p[U_QM]; pdu.nrOfLayers = p[U_NL]; pdu.mcs_index = 9; pdu.target_code_rate = 6160;
  pdu.pusch_data.tb_size = p[U_TBS_BYTES]; pdu.pusch_data.rv_index = p[U_RV];
  pdu.maintenance_parms_v3.ldpcBaseGraph = p[U_BG]; pdu.maintenance_parms_v3.tbSizeLbrmBytes = p[U_TBSLBRM];
  /* pusch_vars->llr is allocated ONCE per ULSCH at start-up in OAI (init_nr_transport) and lives as long as the gNB: the caller's LLRs are copied into such a
   * buffer (grow-only, never freed) instead of being handed over in place.  The interposer page-locks the buffer it is given; a numpy array that is freed (and
   * possibly unmapped) afterwards would leave a stale registration behind that later host buffers at the same address inherit. */
  static short *g_llr[REFH_MAX_CTX];
  static size_t g_llr_cap[REFH_MAX_CTX];
  if ((size_t)G > g_llr_cap[ctx]) {
    const size_t cap = (size_t)G > ((size_t)3 << 20) ? (size_t)G : ((size_t)3 << 20);        /* 273 PRB x 14 symbols x 256QAM x 4 layers = 2.9 M LLRs */
    if (posix_memalign((void **)&g_llr[ctx], 4096, cap * sizeof(short)) != 0) return -102;   /* an outgrown buffer stays allocated (and registered) */
    g_llr_cap[ctx] = cap;
  }
  memcpy(g_llr[ctx], llr, (size_t)G * sizeof(short));
  const int rc = nr_ulsch_decoding(gNB, 0, g_llr[ctx], fp, &pdu, 100, 4, 0, (uint32_t)G);
  if (rc < 0) return rc;
  const int Kb = hp->K >> 3;
  for (int n = 0; n < rc; n++) {
    notifiedFIFO_elt_t *req = pullTpool(&gNB->respDecode, &gNB->threadPool);
    if (!req) return -100;
    ldpcDecode_t *rd = (ldpcDecode_t *)NotifiedFifoData(req);
    const int r = rd->segment_r;
    iters[r] = rd->decodeIterations;
    if (rd->decodeIterations <= rd->decoderParms.numMaxIter)            /* nr_postDecode: the segment's payload joins the transport block */
      memcpy(hp->b + rd->offset, hp->c[r], rd->Kr_bytes - (hp->F >> 3) - (hp->C > 1 ? 3 : 0));
    delNotifiedFIFO_elt(req);
  }
  info[0] = hp->C; info[1] = hp->K; info[2] = hp->Z; info[3] = hp->F; info[4] = hp->llrLen;
  const int ncb = (p[U_BG] == 1 ? 66 : 50) * hp->Z;
  for (int r = 0; r < (int)hp->C; r++) {
    memcpy(c_out + (size_t)r * Kb, hp->c[r], Kb);
    if (d_out) memcpy(d_out + (size_t)r * ncb, hp->d[r], 2 * (size_t)ncb);
  }
  memcpy(tb_out, hp->b, (size_t)p[U_TBS_BYTES] + 3);
  return rc;
}

int refh_ulsch_decode(const int32_t *p, const int16_t *llr, int G, int32_t *info, int32_t *iters, uint8_t *c_out, uint8_t *tb_out, int16_t *d_out)
{
  return refh_ulsch_decode_ctx(0, p, llr, G, info, iters, c_out, tb_out, d_out);
}

/* n back-to-back calls on context ctx (new data every time), timed here so that no interpreter sits between the calls; returns seconds, < 0 on error */
#include <time.h>
double refh_ulsch_decode_loop(int ctx, int n, const int32_t *p, const int16_t *llr, int G, int32_t *info, int32_t *iters, uint8_t *c_out, uint8_t *tb_out)
{
  struct timespec a, b;
  clock_gettime(CLOCK_MONOTONIC, &a);
  for (int i = 0; i < n; i++)
    if (refh_ulsch_decode_ctx(ctx, p, llr, G, info, iters, c_out, tb_out, NULL) < 0) return -1.0;
  clock_gettime(CLOCK_MONOTONIC, &b);
  return (double)(b.tv_sec - a.tv_sec) + 1e-9 * (double)(b.tv_nsec - a.tv_nsec);
}
