/* TEST INFRASTRUCTURE ONLY -- never linked into the product.
 * The reference's side of the LDPC plug-in boundary, compiled against the reference's OWN types (openair1/PHY/CODING/nrLDPC_defs.h, nrLDPC_extern.h,
 * nrLDPC_decoder/nrLDPC_types.h): what load_LDPClib does after load_module_version_shlib has resolved the four names (nrLDPC_load.c:46-71; dlopen flags of
 * common/utils/load_module_shlib.c:160), followed by the calls ldpctest makes (TESTBENCH/ldpctest.c:269-340).  The library under test is handed OAI's
 * structures, not this repo's mirror of them, so any layout or calling-convention difference shows up as a parity failure.
 * (load_module_shlib.c itself needs the config module and libconfig, which this image does not have; its symbol lookup is restated here.) */
#include <dlfcn.h>
#include <pthread.h>
#include <stdio.h>
#include <string.h>
#include "PHY/CODING/nrLDPC_defs.h"
#include "PHY/CODING/nrLDPC_extern.h"

static ldpc_interface_t itf;
static void *handle;

int refh_loader_open(const char *so_path)
{
  handle = dlopen(so_path, RTLD_LAZY | RTLD_NODELETE | RTLD_GLOBAL);
  if (!handle) { fprintf(stderr, "refh_loader_open: %s\n", dlerror()); return -1; }
  itf.LDPCinit = (LDPC_initfunc_t *)dlsym(handle, "LDPCinit");
  itf.LDPCshutdown = (LDPC_shutdownfunc_t *)dlsym(handle, "LDPCshutdown");
  itf.LDPCdecoder = (LDPC_decoderfunc_t *)dlsym(handle, "LDPCdecoder");
  itf.LDPCencoder = (LDPC_encoderfunc_t *)dlsym(handle, "LDPCencoder");
  if (!itf.LDPCinit || !itf.LDPCshutdown || !itf.LDPCdecoder || !itf.LDPCencoder) return -2;
  dlclose(handle);                      /* the loader closes the handle right after the lookup; RTLD_NODELETE keeps the library resident */
  return itf.LDPCinit();                /* AssertFatal(itf->LDPCinit() == 0) */
}

/* n_segments payloads of K/8 bytes -> n_segments code words, one bit per byte, (BG1: 66, BG2: 50) * Zc each (ldpctest.c:269-284) */
int refh_loader_encode(int BG, int Zc, int Kb, int K, int n_segments, const uint8_t *in, uint8_t *out)
{
  uint8_t *ip[64], *op[64];
  const int nout = (BG == 1 ? 66 : 50) * Zc;
  if (n_segments > 64) return -1;
  for (int j = 0; j < n_segments; j++) { ip[j] = (uint8_t *)in + (size_t)j * (K / 8); op[j] = out + (size_t)j * nout; }
  encoder_implemparams_t impp = {.n_segments = n_segments, .macro_num = 0, .gen_code = 0, .tinput = NULL, .tprep = NULL, .tparity = NULL, .toutput = NULL,
                                 .Kb = Kb, .Zc = Zc, .BG = BG, .K = K};
  int rc = 0;
  for (int m = 0; m < (n_segments + 7) / 8 && rc == 0; m++) { impp.macro_num = m; rc = itf.LDPCencoder(ip, op, &impp); }
  return rc;
}

/* one blocking call per segment, like ldpctest.c:329-340; returns the iteration count of the last segment, iters[] receives all */
int refh_loader_decode(int BG, int Zc, int R, int max_iter, int block_length, int n_segments, int llr_stride, const int8_t *llr, int out_stride, uint8_t *out,
                       int32_t *iters)
{
  t_nrLDPC_time_stats prof;
  decode_abort_t ab;
  memset(&prof, 0, sizeof(prof));
  init_abort(&ab);
  int n = 0;
  for (int j = 0; j < n_segments; j++) {
    t_nrLDPC_dec_params dp;
    memset(&dp, 0, sizeof(dp));
    dp.BG = BG; dp.Z = Zc; dp.R = R; dp.numMaxIter = max_iter; dp.outMode = nrLDPC_outMode_BIT; dp.E = block_length;
    itf.LDPCinit();                     /* ldpctest calls it again per segment (ldpctest.c:326) */
    set_abort(&ab, false);
    n = itf.LDPCdecoder(&dp, 0, 0, 0, (int8_t *)llr + (size_t)j * llr_stride, (int8_t *)out + (size_t)j * out_stride, &prof, &ab);
    iters[j] = n;
  }
  return n;
}

int refh_loader_close(void) { return itf.LDPCshutdown(); }
