/* TEST INFRASTRUCTURE ONLY.  Symbols nr_dl_channel_estimation.c references on paths ref_harness_uechest.c never takes (PBCH/PDCCH/PT-RS/SRS
 * estimators, radio config): they abort if reached. */
#include <stdio.h>
#include <stdlib.h>
char openair0_cfg[65536];
#define REFH_DEAD(name) void name(void) { fprintf(stderr, "ref_harness_uechest: unexpected call of " #name "\n"); abort(); }
REFH_DEAD(dB_fixed) REFH_DEAD(nr_ptrs_cpe_estimation) REFH_DEAD(nr_ptrs_process_slot) REFH_DEAD(set_ptrs_symb_idx)
REFH_DEAD(signal_energy)
