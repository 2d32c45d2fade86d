/* TEST INFRASTRUCTURE ONLY -- never linked into the product.
 * The reference-side CALLER of the RU front-end functions nr_feptx0 / nr_fep_full (openair1/SCHED_NR/nr_ru_procedures.c:53, :228): an RU_t with the frame
 * parameters, the txdataF_BF / txdata / rxdata / rxdataF buffers and the RX proc the way init_nr_ru / nr_phy_init_RU leave them; calls the functions by name
 * like nr_feptx_ofdm / ru_thread do.  Linked against integration/oai_shim_ru_ofdm.c (integration/build_shims.sh -> oracle/_ref/libshimtest_ru.so). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "PHY/defs_RU.h"
#include "PHY/defs_nr_common.h"

void nr_feptx0(RU_t *ru, int tti_tx, int first_symbol, int num_symbols, int aa);
void nr_fep_full(RU_t *ru, int slot);

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
static NR_DL_FRAME_PARMS *fill(int N, int mu, int nb_rb, int divisor, int nrx)
{
  NR_DL_FRAME_PARMS *fp = calloc(1, sizeof(*fp));
  fp->ofdm_symbol_size = N; fp->numerology_index = mu; fp->slots_per_subframe = 1 << mu; fp->slots_per_frame = 10 << mu; fp->symbols_per_slot = 14;
  fp->N_RB_DL = fp->N_RB_UL = nb_rb; fp->first_carrier_offset = N - nb_rb * 6; fp->nb_antennas_rx = nrx; fp->nb_antennas_tx = nrx;
  fp->nb_prefix_samples = N / 128 * 9; fp->nb_prefix_samples0 = N / 128 * (9 + (1 << mu));
  fp->samples_per_slotN0 = (fp->nb_prefix_samples + N) * 14;
  fp->samples_per_slot0 = fp->nb_prefix_samples0 + 13 * fp->nb_prefix_samples + 14 * N;
  fp->samples_per_subframe = (fp->nb_prefix_samples0 + N) * 2 + (fp->nb_prefix_samples + N) * (14 * fp->slots_per_subframe - 2);
  fp->samples_per_frame = 10 * fp->samples_per_subframe;
  fp->get_samples_per_slot = h_samples_per_slot; fp->get_samples_slot_timestamp = h_slot_timestamp; fp->ofdm_offset_divisor = divisor;
  return fp;
}

/* txdataF: [nb_tx][14 N] c16 (already phase pre-compensated); txdata out: [nb_tx][samples_per_frame] c16, written by nr_feptx0 in `chunks` calls per antenna */
int refh_ru_feptx(int N, int mu, int nb_rb, int slot, int nb_tx, int chunks, const int16_t *txdataF, int16_t *txdata)
{
  RU_t *ru = calloc(1, sizeof(*ru));
  NR_DL_FRAME_PARMS *fp = ru->nr_frame_parms = fill(N, mu, nb_rb, 8, nb_tx);
  ru->common.txdataF_BF = calloc(nb_tx, sizeof(int32_t *));
  ru->common.txdata = calloc(nb_tx, sizeof(int32_t *));
  for (int a = 0; a < nb_tx; a++) {
    posix_memalign((void **)&ru->common.txdataF_BF[a], 32, 4 * (size_t)14 * N);
    memcpy(ru->common.txdataF_BF[a], txdataF + 2 * (size_t)a * 14 * N, 4 * (size_t)14 * N);
    posix_memalign((void **)&ru->common.txdata[a], 32, 4 * (size_t)fp->samples_per_frame);
    memset(ru->common.txdata[a], 0, 4 * (size_t)fp->samples_per_frame);
  }
  const int per = 14 / chunks;                      /* nr_feptx_ofdm: one call for the slot; nr_feptx (thread pool): half slots */
  for (int a = 0; a < nb_tx; a++)
    for (int c = 0; c < chunks; c++) nr_feptx0(ru, slot, c * per, c == chunks - 1 ? 14 - c * per : per, a);
  for (int a = 0; a < nb_tx; a++) memcpy(txdata + 2 * (size_t)a * fp->samples_per_frame, ru->common.txdata[a], 4 * (size_t)fp->samples_per_frame);
  return (int)fp->samples_per_frame;
}

/* rxdata: [nb_rx][samples_per_frame] c16; rxdataF out: [nb_rx][14 N] c16 of slot `slot` (taken from its place in the 4-slot ring) */
int refh_ru_fep_full(int N, int mu, int nb_rb, int slot, int nb_rx, int divisor, int n_ta_offset, const int16_t *rxdata, int16_t *rxdataF)
{
  RU_t *ru = calloc(1, sizeof(*ru));
  NR_DL_FRAME_PARMS *fp = ru->nr_frame_parms = fill(N, mu, nb_rb, divisor, nb_rx);
  if (!rxdata) return (int)fp->samples_per_frame;
  ru->N_TA_offset = n_ta_offset; ru->proc.tti_rx = slot;
  ru->common.rxdata = calloc(nb_rx, sizeof(int32_t *));
  ru->common.rxdataF = calloc(nb_rx, sizeof(int32_t *));
  for (int a = 0; a < nb_rx; a++) {
    posix_memalign((void **)&ru->common.rxdata[a], 32, 4 * (size_t)fp->samples_per_frame);
    memcpy(ru->common.rxdata[a], rxdata + 2 * (size_t)a * fp->samples_per_frame, 4 * (size_t)fp->samples_per_frame);
    posix_memalign((void **)&ru->common.rxdataF[a], 32, 4 * (size_t)4 * 14 * N);
    memset(ru->common.rxdataF[a], 0, 4 * (size_t)4 * 14 * N);
  }
  nr_fep_full(ru, slot);
  const int offset = (slot % RU_RX_SLOT_DEPTH) * 14 * N;
  for (int a = 0; a < nb_rx; a++) memcpy(rxdataF + 2 * (size_t)a * 14 * N, &ru->common.rxdataF[a][offset], 4 * (size_t)14 * N);
  return (int)fp->samples_per_frame;
}
