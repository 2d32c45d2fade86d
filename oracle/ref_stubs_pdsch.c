/* TEST INFRASTRUCTURE ONLY.  Symbols nr_dlsch_demodulation.c references on paths ref_harness_pdsch.c never takes (PT-RS), plus the time-measurement
 * globals of the softmodem. */
#include <stdio.h>
#include <stdlib.h>
double cpuf = 1.0;
#define REFH_DEAD(name) void name(void) { fprintf(stderr, "ref_harness_pdsch: unexpected call of " #name "\n"); abort(); }
REFH_DEAD(nr_pdsch_ptrs_processing) REFH_DEAD(signal_energy)
