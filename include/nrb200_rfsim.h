/* b200-nrphy: the rfsimulator's channel application on the GPU ("the rfsimulator channel convolution for large multi-UE rfsim runs").
 *
 * Replaces rxAddInput (radio/rfsimulator/apply_channelmod.c:55-111) as simulator.c:975-981 calls it -- once per receive antenna of a connected peer, every read of
 * the simulated radio: for each output sample the double-precision sum over tx antennas and channel taps of tap x past tx sample (tx antennas interleaved in the
 * peer's circular buffer), times the linear path loss, plus noise_per_sample x a standard-normal draw, rounded (lround) and ADDED to the int16 output.
 * One library call does every receive antenna.  Arithmetic: IEEE double, the reference's order of operations, no fused multiply-add -- bit-exact against the compiled
 * reference built without -mfma (what oracle/_ref uses; a -march=native build of OAI may contract the products and differ in the last place).
 *
 * The reference draws its noise from gaussZiggurat (openair1/SIMULATION/TOOLS/rangen_double.c), a sequential generator with process-wide state: the draws are an INPUT
 * here, in the order the reference consumes them -- antenna by antenna, per sample real part first: noise[rx][i][2].  NULL = no noise term.
 * Test: tests/test_gpu_rfsim.py against the oracle that tests/test_oracle_vs_reference.py pins to the real function. */
#ifndef NRB200_RFSIM_H
#define NRB200_RFSIM_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct nrb200_rfsim_chan_s {        /* the fields of channel_desc_t (openair1/SIMULATION/TOOLS/sim.h:60-129) rxAddInput reads */
  uint32_t nb_tx, nb_rx;                    /* 1..8 each */
  uint32_t channel_length;                  /* taps per antenna pair (uint8_t in the reference: <= 255) */
  int32_t channel_offset;                   /* extra delay in samples (its absolute value is used, like the reference) */
  double path_loss_dB;                      /* total path gain: pathLossLinear = pow(10, path_loss_dB / 20) */
  float noise_power_dB;                     /* noise_per_sample = pow(10, noise_power_dB / 10) * 256 */
  uint32_t reserved;
} nrb200_rfsim_chan_t;

/* ch: channel_desc_t.ch flattened, [nb_tx * nb_rx][channel_length] {re, im} doubles, plane index rxAnt + txAnt * nb_rx.
 * input_sig: the peer's circular buffer, CirSize c16 {re, im} int16, tx antennas interleaved (sample t of antenna a at ((t * nb_tx + a) % CirSize)).
 * out: [nb_rx][out_stride] c16, nbSamples per antenna are accumulated into (the caller clears them before the first peer, simulator.c:956-957).
 * TS: time stamp of the first output sample (t->nextRxTstamp).  noise: [nb_rx][nbSamples][2] doubles or NULL. */
int32_t nrb200_rfsim_rx_add_input_dev(const nrb200_rfsim_chan_t *c, const double *d_ch, const int16_t *d_input_sig, int16_t *d_out, uint32_t out_stride,
                                      uint32_t nbSamples, uint64_t TS, uint32_t CirSize, const double *d_noise, void *stream);
int32_t nrb200_rfsim_rx_add_input_host(const nrb200_rfsim_chan_t *c, const double *ch, const int16_t *input_sig, int16_t *out, uint32_t out_stride,
                                       uint32_t nbSamples, uint64_t TS, uint32_t CirSize, const double *noise);

#ifdef __cplusplus
}
#endif
#endif
