/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the rfsimulator's channel application rxAddInput (radio/rfsimulator/apply_channelmod.c:55-111):
 * for every output sample of one receive antenna, the double-precision sum over tx antennas and channel taps of tap x past tx sample (tx antennas interleaved in
 * a circular buffer), scaled by the linear path loss, plus noise_per_sample x a standard normal draw, rounded with lround and ACCUMULATED into the int16 output.
 * Pinned bit-exactly against the compiled reference through oracle/ref_harness_rfsim.c (tests/test_oracle_vs_reference.py), where gaussZiggurat hands back the
 * caller's draws in call order (real part first).  Plain IEEE double arithmetic in the reference's order, no fused multiply-add (oracle/_ref is built -mavx2 only).
 * Only tests/, smoke() and bench.py's cpu_baseline leg may link this.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include "nrb200_oracle.h"

void orc_rfsim_rx_add_input(int nb_tx, int nb_rx, int channel_length, int channel_offset, double path_loss_dB, float noise_power_dB, const double *ch,
                            const int16_t *input_sig, int16_t *out, int rxAnt, int nbSamples, uint64_t TS, uint32_t CirSize, const double *noise)
{
  const double pathLossLinear = pow(10, path_loss_dB / 20.0);
  const double noise_per_sample = pow(10, noise_power_dB / 10.0) * 256;
  const int dd = abs(channel_offset);
  for (int i = 0; i < nbSamples; i++) {
    volatile double rr = 0.0, ri = 0.0;                     /* volatile products below: no contraction whatever the flags */
    for (int txAnt = 0; txAnt < nb_tx; txAnt++) {
      const double *c = ch + 2 * (size_t)(rxAnt + txAnt * nb_rx) * channel_length;
      for (int l = 0; l < channel_length; l++) {
        const int idx = (int)(((TS + i - l - dd) * nb_tx + txAnt + CirSize) % CirSize);
        const int16_t xr = input_sig[2 * idx], xi = input_sig[2 * idx + 1];
        volatile double a = xr * c[2 * l], b = xi * c[2 * l + 1], e = xi * c[2 * l], f = xr * c[2 * l + 1];
        rr += a - b;
        ri += e + f;
      }
    }
    volatile double pr = rr * pathLossLinear, pi = ri * pathLossLinear;
    volatile double nr = noise_per_sample * (noise ? noise[2 * i] : 0.0), ni = noise_per_sample * (noise ? noise[2 * i + 1] : 0.0);
    out[2 * i] = (int16_t)(uint16_t)(uint32_t)((int32_t)out[2 * i] + (int32_t)lround(pr + nr));
    out[2 * i + 1] = (int16_t)(uint16_t)(uint32_t)((int32_t)out[2 * i + 1] + (int32_t)lround(pi + ni));
  }
}
