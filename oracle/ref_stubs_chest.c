/* TEST INFRASTRUCTURE ONLY.  Symbols nr_ul_channel_estimation.c references on paths ref_harness_chest.c never takes (transform precoding,
 * PT-RS, SRS): they abort if reached. */
#include <stdio.h>
#include <stdlib.h>
void *gNB_dmrs_lowpaprtype1_sequence[30 * 2 * 128];
#define REFH_DEAD(name) void name(void) { fprintf(stderr, "ref_harness_chest: unexpected call of " #name "\n"); abort(); }
REFH_DEAD(dB_fixed) REFH_DEAD(get_index_for_dmrs_lowpapr_seq) REFH_DEAD(nr_ptrs_cpe_estimation) REFH_DEAD(nr_ptrs_process_slot) REFH_DEAD(set_ptrs_symb_idx)
