/* Link-time interposers for the OAI RU front end: nr_feptx0 and nr_fep_full on the GPU with host C unchanged (hook
 * the OFDM front end one level above the per-symbol dft()/idft() plug-in, where a whole slot's symbols are visible).  Same prototypes as
 * openair1/SCHED_NR/nr_ru_procedures.c:53 and :228, compiled against OAI's headers.
 *   nr_feptx0(ru, slot, first_symbol, num_symbols, aa)  the IDFT + cyclic prefix of num_symbols symbols of tx antenna aa: txdataF_BF[aa] -> txdata[aa]
 *                                                      (the phase pre-compensation apply_nr_rotation_TX has been applied by the caller, nr_feptx_ofdm / nr_feptx)
 *   nr_fep_full(ru, slot)                               the 14 DFTs of every rx antenna of slot proc->tti_rx: rxdata[aa] (frame ring, N_TA_offset) -> rxdataF[aa]
 * Each becomes ONE call of the library's slot-level OFDM entry point (one launch) instead of 14 x antennas dft()/idft() calls.  Symbol positions and prefix
 * lengths are derived from NR_DL_FRAME_PARMS the way the reference derives them (:63-72 and MODULATION/slot_fep_nr.c:236-244).
 * Test: tests/test_gpu_interpose.py through oracle/ref_harness_ru.c (reference-side caller with OAI's RU_t) against the pinned oracle. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "PHY/defs_RU.h"
#include "PHY/defs_nr_common.h"
#define NRB200_NO_OAI_LOADER_PROTOTYPES
#include "nrb200_dfts.h"

static unsigned prefix_of(const NR_DL_FRAME_PARMS *fp, int abs_symbol)
{
  if (fp->Ncp == 1) return fp->nb_prefix_samples;
  return (abs_symbol % (0x7 << fp->numerology_index)) ? fp->nb_prefix_samples : fp->nb_prefix_samples0;
}

void nr_feptx0(RU_t *ru, int tti_tx, int first_symbol, int num_symbols, int aa)
{
  NR_DL_FRAME_PARMS *fp = ru->nr_frame_parms;
  const int N = fp->ofdm_symbol_size, slot = tti_tx;
  nrb200_ofdm_slot_t d;
  memset(&d, 0, sizeof(d));
  d.fft_size = N; d.n_symb = num_symbols; d.n_ant = 1; d.rotate = 0; d.nb_rb = fp->N_RB_DL; d.first_carrier_offset = fp->first_carrier_offset;
  unsigned off = fp->get_samples_slot_timestamp(slot, fp, 0);
  for (int l = 0; l < first_symbol; l++) off += prefix_of(fp, slot * fp->symbols_per_slot + l) + N;
  for (int l = 0; l < num_symbols; l++) {
    d.prefix[l] = prefix_of(fp, slot * fp->symbols_per_slot + first_symbol + l);
    d.t_off[l] = off;
    off += d.prefix[l] + N;
  }
  const int16_t *fin[1] = {(const int16_t *)&ru->common.txdataF_BF[aa][first_symbol * N]};
  int16_t *tout[1] = {(int16_t *)ru->common.txdata[aa]};
  const int rc = nrb200_ofdm_mod_slot_host(&d, fin, tout);
  if (rc != 0) { fprintf(stderr, "nrb200 shim: nr_feptx0: nrb200_ofdm_mod_slot_host failed (rc = %d): %s\n", rc, nrb200_dfts_last_error()); abort(); }
}

void nr_fep_full(RU_t *ru, int slot)
{
  (void)slot;                                            /* the reference works on proc->tti_rx too (:228-250) */
  NR_DL_FRAME_PARMS *fp = ru->nr_frame_parms;
  const int N = fp->ofdm_symbol_size, Ns = ru->proc.tti_rx, nrx = fp->nb_antennas_rx;
  const int offset = (Ns % RU_RX_SLOT_DEPTH) * (fp->symbols_per_slot * N);
  nrb200_ofdm_slot_t d;
  memset(&d, 0, sizeof(d));
  d.fft_size = N; d.n_symb = fp->symbols_per_slot; d.n_ant = nrx; d.rotate = 0; d.nb_rb = fp->N_RB_UL; d.first_carrier_offset = fp->first_carrier_offset;
  d.t_ring = fp->samples_per_frame;
  unsigned off = fp->get_samples_slot_timestamp(Ns, fp, 0);
  for (int l = 0; l < fp->symbols_per_slot; l++) {
    off += prefix_of(fp, Ns * fp->symbols_per_slot + l);                     /* first sample after the prefix of symbol l */
    const unsigned start = off + N * l - fp->nb_prefix_samples / fp->ofdm_offset_divisor;   /* 1/divisor of the CP early, against ISI */
    d.t_off[l] = (start + fp->samples_per_frame - (unsigned)ru->N_TA_offset) % fp->samples_per_frame;
  }
  const int16_t *tin[8];
  int16_t *fout[8];
  if (nrx > 8) { fprintf(stderr, "nrb200 shim: nr_fep_full: %d rx antennas\n", nrx); abort(); }
  for (int a = 0; a < nrx; a++) { tin[a] = (const int16_t *)ru->common.rxdata[a]; fout[a] = (int16_t *)&ru->common.rxdataF[a][offset]; }
  const int rc = nrb200_ofdm_demod_slot_host(&d, tin, NULL, fout);
  if (rc != 0) { fprintf(stderr, "nrb200 shim: nr_fep_full: nrb200_ofdm_demod_slot_host failed (rc = %d): %s\n", rc, nrb200_dfts_last_error()); abort(); }
}
