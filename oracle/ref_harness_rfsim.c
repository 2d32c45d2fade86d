/* TEST INFRASTRUCTURE ONLY -- never linked into the product.
 * Flat C entry point around the UNMODIFIED rfsimulator channel application rxAddInput (radio/rfsimulator/apply_channelmod.c:55-111), compiled from the reference
 * tree next to this file (oracle/build_ref.sh -> oracle/_ref/libref_rfsim.so).  The two things the function takes from the rest of the simulator are supplied here:
 * gaussZiggurat (openair1/SIMULATION/TOOLS/rangen_double.c: a sequential generator) returns the caller's pre-drawn noise samples in call order, so the noise term is
 * a plain input; signal_energy is only read by a debug log line. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <openair1/SIMULATION/TOOLS/sim.h>
#include "radio/rfsimulator/rfsimulator.h"

static const double *g_noise;
static size_t g_noise_pos;
double gaussZiggurat(double mean, double variance) { (void)mean; (void)variance; return g_noise ? g_noise[g_noise_pos++] : 0.0; }
int32_t signal_energy(int32_t *input, uint32_t length) { (void)input; (void)length; return 1; }

/* ch: [nb_tx * nb_rx][channel_length] {r, i} doubles, plane index rxAnt + txAnt * nb_rx (channel_desc_t.ch); input_sig: the circular buffer (CirSize c16, tx antennas
 * interleaved); out: nbSamples c16 of receive antenna rxAnt, ACCUMULATED into; noise: 2 * nbSamples draws (r then i per sample) or NULL */
void refh_rfsim_rx_add_input(int nb_tx, int nb_rx, int channel_length, int channel_offset, double path_loss_dB, float noise_power_dB, const double *ch,
                             const int16_t *input_sig, int16_t *out, int rxAnt, int nbSamples, uint64_t TS, uint32_t CirSize, const double *noise)
{
  channel_desc_t d;
  memset(&d, 0, sizeof(d));
  d.nb_tx = (uint8_t)nb_tx; d.nb_rx = (uint8_t)nb_rx; d.channel_length = (uint8_t)channel_length; d.channel_offset = channel_offset;
  d.path_loss_dB = path_loss_dB; d.noise_power_dB = noise_power_dB;
  struct complexd **planes = calloc((size_t)nb_tx * nb_rx, sizeof(*planes));
  for (int p = 0; p < nb_tx * nb_rx; p++) planes[p] = (struct complexd *)(ch + 2 * (size_t)p * channel_length);
  d.ch = planes;
  g_noise = noise; g_noise_pos = 0;
  rxAddInput((const c16_t *)input_sig, (c16_t *)out, rxAnt, &d, nbSamples, TS, CirSize);
  g_noise = NULL;
  free(planes);
}
