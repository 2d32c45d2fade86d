/* TEST INFRASTRUCTURE ONLY.  Symbols nr_ulsch_demodulation.c references on paths ref_harness_pusch.c never takes (channel estimation,
 * PT-RS, measurements): they abort if reached.  dft/idft are the softmodem's function-pointer globals. */
#include <stdio.h>
#include <stdlib.h>
void *dft, *idft;
#define REFH_DEAD(name) void name(void) { fprintf(stderr, "ref_harness_pusch: unexpected call of " #name "\n"); abort(); }
REFH_DEAD(get_dmrs_port) REFH_DEAD(get_next_dmrs_symbol_in_slot) REFH_DEAD(get_ptrs_symbols_in_slot) REFH_DEAD(nr_chest_time_domain_avg)
REFH_DEAD(nr_codeword_unscrambling_init) REFH_DEAD(nr_get_G) REFH_DEAD(nr_gnb_measurements)
REFH_DEAD(nr_pusch_channel_estimation) REFH_DEAD(nr_pusch_ptrs_processing) REFH_DEAD(set_ptrs_symb_idx) REFH_DEAD(signal_energy_nodc)
