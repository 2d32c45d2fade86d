/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the gNB PRACH detector rx_nr_prach (openair1/PHY/NR_TRANSPORT/nr_prach.c:414-714), unrestricted set:
 * per root sequence the received PRACH sub-carriers of every antenna are multiplied by the conjugated root (>> 15, truncating), zero-padded to 1024 (long sequences)
 * or 256 (short ones) points and taken through idft(IDFT_1024 / IDFT_256, scale 1); the powers are summed over the antennas, >> log2(size) and divided by the
 * number of antennas; each of the 64 preambles then searches its window of NCS2 delay bins (cyclic shift preamble_shift = -v NCS mod N_ZC, bin = shift << log2 / N_ZC)
 * for the largest dB_fixed_times10 (TOOLS/dB_routines.c:132-155), first maximum wins.  Pinned bit-exactly against the compiled reference through
 * oracle/ref_harness_prach.c (tests/test_oracle_vs_reference.py).  Only tests/, smoke() and bench.py's cpu_baseline leg may link this.
 * Bins beyond the transform size read as zero: the reference's buffer is cleared up to 2048 entries for long sequences and its window never gets there; for short
 * sequences only 256 entries are cleared and the last window can touch entry 256, which holds whatever an earlier long-sequence occasion left (not restated).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "nrb200_oracle.h"
#include "nr_db_table.h"

static const int16_t db_tab[256] = NRB200_DB_TABLE_TIMES10;
int orc_db_fixed_times10(uint32_t x)
{
  int v;
  if (x == 0) return 0;
  if (x & 0xff000000u) v = db_tab[((x >> 24) & 255) - 1] + 3 * db_tab[255];
  else if (x & 0x00ff0000u) v = db_tab[((x >> 16) & 255) - 1] + 2 * db_tab[255];
  else if (x & 0x0000ff00u) v = db_tab[((x >> 8) & 255) - 1] + db_tab[255];
  else v = db_tab[(x & 255) - 1];
  return v > 900 ? 900 : v;
}

/* xu: [64][839] c16 (gNB->X_u); rxsigF: [nb_rx][N_ZC] c16.  out3: max_preamble, max_preamble_energy, max_preamble_delay (after the timing-advance scaling) */
int orc_rx_nr_prach(int nb_rx, int short_sequence, int NCS, int prach_fmt, int mu, const int16_t *xu, const int16_t *rxsigF, int32_t *out3)
{
  const int N_ZC = short_sequence ? 139 : 839, size = short_sequence ? 256 : 1024, lg = short_sequence ? 8 : 10;
  int NCS2 = short_sequence ? ((NCS << 8) / 139) : ((NCS << 10) / 839);
  if (NCS2 == 0) NCS2 = N_ZC;
  int16_t *prachF = calloc(2 * 1024, 2), *tmp = calloc(2 * 2048, 2);
  int32_t *ifft = calloc(2048, 4);
  int old = 99, shift = 0;
  uint16_t best_e = 0, best_d = 0, best_p = 0;
  for (int pi = 0; pi < 64; pi++) {
    const int off = NCS == 0 ? pi : pi / (N_ZC / NCS);
    if (off != old) {
      old = off; shift = 0;
      const int16_t *X = xu + 2 * (size_t)off * 839;
      memset(ifft, 0, 4 * (size_t)size);
      memset(prachF, 0, 2 * 2 * 1024);
      for (int a = 0; a < nb_rx; a++) {
        const int16_t *r = rxsigF + 2 * (size_t)a * N_ZC;
        for (int k = 0; k < N_ZC; k++) {
          prachF[2 * k] = (int16_t)(((int32_t)X[2 * k] * r[2 * k] + (int32_t)X[2 * k + 1] * r[2 * k + 1]) >> 15);
          prachF[2 * k + 1] = (int16_t)(((int32_t)X[2 * k] * r[2 * k + 1] - (int32_t)X[2 * k + 1] * r[2 * k]) >> 15);
        }
        orc_dft(size, 1, prachF, tmp, 1);
        for (int i = 0; i < size; i++)
          ifft[i] = (int32_t)((uint32_t)ifft[i] + (uint32_t)((int32_t)tmp[2 * i] * tmp[2 * i]) + (uint32_t)((int32_t)tmp[2 * i + 1] * tmp[2 * i + 1]));
      }
      for (int i = 0; i < size; i++) ifft[i] = (ifft[i] >> lg) / nb_rx;
    } else {
      shift -= NCS;
      if (shift < 0) shift += N_ZC;
    }
    const uint32_t shift2 = shift == 0 ? 0 : (uint32_t)((shift << lg) / N_ZC);
    for (int i = 0; i < NCS2; i++) {
      const uint32_t b = shift2 + (uint32_t)i;
      const int32_t lev = b < (uint32_t)size ? ifft[b] : 0;
      const int16_t levdB = (int16_t)orc_db_fixed_times10((uint32_t)lev);
      if (levdB > best_e) { best_e = (uint16_t)levdB; best_d = (uint16_t)i; best_p = (uint16_t)pi; }
    }
  }
  if (!short_sequence) {
    if (prach_fmt == 0 || prach_fmt == 1 || prach_fmt == 2) best_d = (uint16_t)(best_d * 3 * (1 << mu) / 2);
    else if (prach_fmt == 3) best_d = (uint16_t)(best_d * 3 * (1 << mu) / 8);
  } else best_d = (uint16_t)(best_d / 2);
  out3[0] = best_p; out3[1] = best_e; out3[2] = best_d;
  free(prachF); free(tmp); free(ifft);
  return 0;
}
