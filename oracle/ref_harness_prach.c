/* TEST INFRASTRUCTURE ONLY -- never linked into the product.
 * Flat C entry points around the UNMODIFIED gNB PRACH detector rx_nr_prach (openair1/PHY/NR_TRANSPORT/nr_prach.c:414-714) and the root-sequence generator
 * compute_nr_prach_seq (NR_TRANSPORT/nr_prach_common.c:100-152).  The harness fills the fields of PHY_VARS_gNB / nfapi_nr_prach_config_t / nfapi_nr_prach_pdu_t the
 * detector reads (buffers sized like nr_init.c:281-283) and hands it rxsigF, the per-antenna PRACH sub-carriers rx_nr_prach_ru leaves behind.  idft stays OAI's
 * function pointer, bound at run time to oracle/_ref/libref_dfts.so. */
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "PHY/defs_gNB.h"
#include "PHY/NR_TRANSPORT/nr_transport_proto.h"
#include "PHY/NR_TRANSPORT/nr_transport_common_proto.h"
#include "PHY/TOOLS/tools_defs.h"

dftfunc_t dft;
idftfunc_t idft;
int refh_prach_init(const char *dfts_so)
{
  void *h = dlopen(dfts_so, RTLD_NOW | RTLD_LOCAL);
  if (!h) { fprintf(stderr, "%s\n", dlerror()); return -1; }
  void (*autoinit)(void) = (void (*)(void))dlsym(h, "dfts_autoinit");
  dft = (dftfunc_t)dlsym(h, "dft"); idft = (idftfunc_t)dlsym(h, "idft");
  if (!autoinit || !dft || !idft) return -2;
  autoinit();
  return 0;
}

/* X_u out: [64][839] c16 as compute_nr_prach_seq leaves gNB->X_u */
void refh_prach_seq(int short_sequence, int num_sequences, int rootSequenceIndex, int16_t *xu_out)
{
  c16_t (*X)[839] = calloc(64, sizeof(*X));
  compute_nr_prach_seq((uint8_t)short_sequence, (uint8_t)num_sequences, (uint8_t)rootSequenceIndex, X);
  memcpy(xu_out, X, sizeof(c16_t) * 64 * 839);
  free(X);
}
int refh_db_fixed_times10(uint32_t x) { return dB_fixed_times10(x); }

enum { R_NB_RX, R_SHORT, R_ROOT, R_NUM_ROOTS, R_NCS, R_FORMAT, R_MU, R_COUNT };
/* xu: [64][839] c16; rxsigF: [nb_rx][N_ZC] c16.  out3: max_preamble, max_preamble_energy, max_preamble_delay */
int refh_rx_nr_prach(const int32_t *p, const int16_t *xu, const int16_t *rxsigF, int32_t *out3)
{
  const int nrx = p[R_NB_RX], N_ZC = p[R_SHORT] ? 139 : 839;
  PHY_VARS_gNB *gNB = calloc(1, sizeof(*gNB));
  gNB->frame_parms.numerology_index = p[R_MU]; gNB->frame_parms.N_RB_UL = 273; gNB->frame_parms.ofdm_symbol_size = 4096;
  gNB->gNB_config.carrier_config.num_rx_ant.value = nrx;
  nfapi_nr_prach_config_t *cfg = &gNB->gNB_config.prach_config;
  cfg->prach_sequence_length.value = p[R_SHORT]; cfg->restricted_set_config.value = 0;
  cfg->num_prach_fd_occasions_list = calloc(1, sizeof(*cfg->num_prach_fd_occasions_list));
  cfg->num_prach_fd_occasions_list[0].prach_root_sequence_index.value = p[R_ROOT];
  cfg->num_prach_fd_occasions_list[0].num_root_sequences.value = p[R_NUM_ROOTS];
  cfg->num_prach_fd_occasions_list[0].k1.value = 0;
  memcpy(gNB->X_u, xu, sizeof(gNB->X_u));
  gNB->prach_vars.prachF = calloc(1024 * 2, sizeof(int16_t));
  gNB->prach_vars.prach_ifft = calloc(1024 * 2, sizeof(int32_t));
  gNB->prach_vars.rxsigF = calloc(nrx, sizeof(int16_t *));
  for (int a = 0; a < nrx; a++) {
    posix_memalign((void **)&gNB->prach_vars.rxsigF[a], 32, 4 * 1024);
    memset(gNB->prach_vars.rxsigF[a], 0, 4 * 1024);
    memcpy(gNB->prach_vars.rxsigF[a], rxsigF + 2 * (size_t)a * N_ZC, 4 * (size_t)N_ZC);
  }
  nfapi_nr_prach_pdu_t pdu;
  memset(&pdu, 0, sizeof(pdu));
  pdu.num_ra = 0; pdu.num_cs = p[R_NCS]; pdu.prach_format = p[R_FORMAT];
  uint16_t mp = 0, me = 0, md = 0;
  rx_nr_prach(gNB, &pdu, 0, 0, 0, &mp, &me, &md);
  out3[0] = mp; out3[1] = me; out3[2] = md;
  for (int a = 0; a < nrx; a++) free(gNB->prach_vars.rxsigF[a]);
  free(gNB->prach_vars.rxsigF); free(gNB->prach_vars.prachF); free(gNB->prach_vars.prach_ifft); free(cfg->num_prach_fd_occasions_list); free(gNB);
  return 0;
}
