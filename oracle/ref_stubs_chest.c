/* TEST INFRASTRUCTURE ONLY.  Symbols nr_ul_channel_estimation.c references on paths ref_harness_chest.c never takes (PT-RS, SRS): they abort if reached. */
#include <stdio.h>
#include <stdlib.h>
#define REFH_DEAD(name) void name(void) { fprintf(stderr, "ref_harness_chest: unexpected call of " #name "\n"); abort(); }
REFH_DEAD(dB_fixed) REFH_DEAD(nr_ptrs_cpe_estimation) REFH_DEAD(nr_ptrs_process_slot) REFH_DEAD(set_ptrs_symb_idx)
