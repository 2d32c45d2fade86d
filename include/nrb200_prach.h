/* b200-nrphy: the gNB's PRACH detector on the GPU ("PUCCH/PRACH detectors").
 *
 * Replaces rx_nr_prach (openair1/PHY/NR_TRANSPORT/nr_prach.c:414-714) for the unrestricted set (restricted_set_config == 0): correlation of the received PRACH
 * sub-carriers of every antenna with every root sequence in use, idft(IDFT_1024 | IDFT_256), power summed over the antennas, and the search of the 64 preambles'
 * delay windows for the largest dB_fixed_times10 -- all roots and antennas in one batched transform instead of one idft() call per (root, antenna).
 * Integer arithmetic, bit-exact: detected preamble, energy (0.1 dB units) and timing advance are the reference's.
 * Test: tests/test_gpu_prach.py against the oracle that tests/test_oracle_vs_reference.py pins to the real function. */
#ifndef NRB200_PRACH_H
#define NRB200_PRACH_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct nrb200_prach_s {
  uint32_t nb_rx;                           /* gNB_config.carrier_config.num_rx_ant (1..8) */
  uint32_t short_sequence;                  /* prach_config.prach_sequence_length: 0 = 839 (formats 0-3), 1 = 139 (A1 ... C2) */
  uint32_t num_cs;                          /* prach_pdu->num_cs (N_CS; 0 = one preamble per root) */
  uint32_t prach_format;                    /* prach_pdu->prach_format: scales the timing advance of long sequences (:693-698) */
  uint32_t numerology;                      /* frame_parms.numerology_index */
  uint32_t restricted_set;                  /* prach_config.restricted_set_config: must be 0 (the high-speed sets return -4) */
  uint32_t rx_stride;                       /* _dev: c16 between antennas of rxsigF (>= 839 | 139) */
  uint32_t reserved;
} nrb200_prach_t;

/* number of root sequences the 64 preambles use = rows of X_u read: 64 for num_cs == 0, else ceil(64 / (N_ZC / num_cs)); 0 if the descriptor is invalid */
uint32_t nrb200_prach_num_roots(const nrb200_prach_t *d);
uint64_t nrb200_prach_scratch_bytes(const nrb200_prach_t *d);
/* X_u: gNB->X_u as compute_nr_prach_seq leaves it (NR_TRANSPORT/nr_prach_common.c:100-152), [roots][839] c16, root i at row i.
 * rxsigF: gNB->prach_vars.rxsigF, [nb_rx][rx_stride] c16 (the PRACH occasion's sub-carriers rx_nr_prach_ru extracted).
 * out: 3 int32 = *max_preamble, *max_preamble_energy, *max_preamble_delay (timing advance, after the format scaling). */
int32_t nrb200_rx_nr_prach_dev(const nrb200_prach_t *d, const int16_t *d_X_u, const int16_t *d_rxsigF, int32_t *d_out, void *d_scratch, void *stream);
int32_t nrb200_rx_nr_prach_host(const nrb200_prach_t *d, const int16_t *X_u, const int16_t *rxsigF, uint16_t *max_preamble, uint16_t *max_preamble_energy,
                                uint16_t *max_preamble_delay);

#ifdef __cplusplus
}
#endif
#endif
