/* TEST INFRASTRUCTURE ONLY -- never linked into the product.
 * Flat C entry points around the UNMODIFIED reference OFDM front end (ofdm_mod.c, slot_fep_nr.c, cmult_sv.c, cmult_vv.c,
 * nr_modulation.c compiled from /root/reference by build_ref.sh into libref_ofdm.so).  The harness fills the few NR_DL_FRAME_PARMS
 * fields those functions read, owns the dft/idft function-pointer globals the softmodem normally gets from dfts_load.c and binds
 * them to the compiled reference libref_dfts.so. */
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "PHY/defs_nr_common.h"
#include "PHY/TOOLS/tools_defs.h"
#include <time.h>
/* wall time of the last call into the reference function(s), excluding the harness's own allocation and copying (cpu_baseline of the DL slot chain) */
static double g_last_s;
static inline double now_s(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }


void nr_normal_prefix_mod(c16_t *txdataF, c16_t *txdata, uint8_t nsymb, const NR_DL_FRAME_PARMS *frame_parms, uint32_t slot);
void apply_nr_rotation_TX(const NR_DL_FRAME_PARMS *fp, c16_t *txdataF, const c16_t *symbol_rotation, int slot, int nb_rb, int first_symbol, int nsymb);
void apply_nr_rotation_RX(NR_DL_FRAME_PARMS *frame_parms, c16_t *rxdataF, c16_t *rot, int slot, int nb_rb, int soffset, int first_symbol, int nsymb);
int nr_slot_fep_ul(NR_DL_FRAME_PARMS *frame_parms, int32_t *rxdata, int32_t *rxdataF, unsigned char symbol, unsigned char Ns, int sample_offset);
void init_symbol_rotation(NR_DL_FRAME_PARMS *fp);
void init_timeshift_rotation(NR_DL_FRAME_PARMS *fp);

dftfunc_t dft;
idftfunc_t idft;
void *get_softmodem_params(void) { static char z[4096]; return z; }
signed char dB_fixed(unsigned int x) { (void)x; return 0; }
int32_t signal_energy(int32_t *a, uint32_t n) { (void)a; (void)n; return 0; }
int is_pmch_subframe(uint32_t f, int s, void *p) { (void)f; (void)s; (void)p; return 0; }

int refh_ofdm_init(const char *dfts_so)
{
  void *h = dlopen(dfts_so, RTLD_NOW | RTLD_LOCAL);
  if (!h) { fprintf(stderr, "refh_ofdm_init: %s\n", dlerror()); return -1; }
  int (*autoinit)(void) = (int (*)(void))dlsym(h, "dfts_autoinit");
  dft = (dftfunc_t)dlsym(h, "dft");
  idft = (idftfunc_t)dlsym(h, "idft");
  if (!autoinit || !dft || !idft) return -2;
  autoinit();
  return 0;
}

static uint32_t h_samples_per_slot(int slot, const NR_DL_FRAME_PARMS *fp)
{
  if (fp->numerology_index == 0) return fp->samples_per_subframe;
  return (slot % (fp->slots_per_subframe / 2)) ? fp->samples_per_slotN0 : fp->samples_per_slot0;
}
static uint32_t h_slot_timestamp(int slot, const NR_DL_FRAME_PARMS *fp, uint8_t ahead)
{
  uint32_t s = 0;
  for (int i = ahead ? slot : 0; i < (ahead ? slot + ahead : slot); i++) s += h_samples_per_slot(i, fp);
  return s;
}

static NR_DL_FRAME_PARMS *fill(int N, int mu, int nb_rb, int divisor)
{
  static NR_DL_FRAME_PARMS fp;
  memset(&fp, 0, sizeof(fp));
  fp.ofdm_symbol_size = N;
  fp.numerology_index = mu;
  fp.slots_per_subframe = 1 << mu;
  fp.slots_per_frame = 10 << mu;
  fp.symbols_per_slot = 14;
  fp.N_RB_DL = fp.N_RB_UL = nb_rb;
  fp.first_carrier_offset = N - nb_rb * 6;
  fp.nb_prefix_samples = N / 128 * 9;
  fp.nb_prefix_samples0 = N / 128 * (9 + (1 << mu));
  fp.samples_per_slotN0 = (fp.nb_prefix_samples + N) * 14;
  fp.samples_per_slot0 = fp.nb_prefix_samples0 + 13 * fp.nb_prefix_samples + 14 * N;
  fp.samples_per_subframe = (fp.nb_prefix_samples0 + N) * 2 + (fp.nb_prefix_samples + N) * (14 * fp.slots_per_subframe - 2);
  fp.samples_per_frame = 10 * fp.samples_per_subframe;
  fp.get_samples_per_slot = h_samples_per_slot;
  fp.get_samples_slot_timestamp = h_slot_timestamp;
  fp.ofdm_offset_divisor = divisor;
  return &fp;
}

/* symbol_rotation[0] (DL carrier), [1] (UL carrier): 224 {re,im} pairs each; timeshift: N pairs */
void refh_rotation_tables(int N, int mu, int nb_rb, int divisor, double dl_freq, double ul_freq, int16_t *rot_dl, int16_t *rot_ul, int16_t *timeshift)
{
  NR_DL_FRAME_PARMS *fp = fill(N, mu, nb_rb, divisor);
  fp->dl_CarrierFreq = (uint64_t)dl_freq;
  fp->ul_CarrierFreq = (uint64_t)ul_freq;
  init_symbol_rotation(fp);
  init_timeshift_rotation(fp);
  memcpy(rot_dl, fp->symbol_rotation[0], 224 * 4);
  memcpy(rot_ul, fp->symbol_rotation[1], 224 * 4);
  memcpy(timeshift, fp->timeshift_symbol_rotation, (size_t)N * 4);
}

/* nr_feptx_ofdm order: apply_nr_rotation_TX (in place on txdataF) then the slot-wise CP-OFDM modulator.  txdata = the slot's samples. */
void refh_ofdm_tx_slot(int N, int mu, int nb_rb, int slot, int nsymb, const int16_t *rot /* 224 pairs or NULL */, int16_t *txdataF, int16_t *txdata)
{
  NR_DL_FRAME_PARMS *fp = fill(N, mu, nb_rb, 8);
  const double t0 = now_s();
  if (rot) apply_nr_rotation_TX(fp, (c16_t *)txdataF, (const c16_t *)rot, slot, nb_rb, 0, nsymb);
  nr_normal_prefix_mod((c16_t *)txdataF, (c16_t *)txdata, (uint8_t)nsymb, fp, (uint32_t)slot);
  g_last_s = now_s() - t0;
}
double refh_ofdm_last_seconds(void) { return g_last_s; }

/* nr_fep_full order: nr_slot_fep_ul for the 14 symbols of `slot`, then apply_nr_rotation_RX (phase + timeshift compensation).
 * rxdata = one frame of samples (samples_per_frame c16); returns samples_per_frame. */
int refh_ofdm_rx_slot(int N, int mu, int nb_rb, int slot, int divisor, int sample_offset, const int16_t *rot_ul, int16_t *rxdata, int16_t *rxdataF)
{
  NR_DL_FRAME_PARMS *fp = fill(N, mu, nb_rb, divisor);
  if (!rxdata) return (int)fp->samples_per_frame;
  init_timeshift_rotation(fp);
  for (int l = 0; l < 14; l++) nr_slot_fep_ul(fp, (int32_t *)rxdata, (int32_t *)rxdataF, (unsigned char)l, (unsigned char)slot, sample_offset);
  if (rot_ul) apply_nr_rotation_RX(fp, (c16_t *)rxdataF, (c16_t *)rot_ul, slot, nb_rb, 0, 0, 14);
  return (int)fp->samples_per_frame;
}
