/* TEST INFRASTRUCTURE ONLY.  Symbols the sources of libref_pdsch_ptrs.so (ref_harness_pdsch.c with -DREFH_PTRS: the real nr_rx_pdsch together with the real
 * nr_pdsch_ptrs_processing / ptrs_nr.c) reference on paths the harness never takes, plus the softmodem's globals. */
#include <stdio.h>
#include <stdlib.h>
double cpuf = 1.0;
char openair0_cfg[65536];
#define REFH_DEAD(name) void name(void) { fprintf(stderr, "ref_harness_pdsch (PT-RS): unexpected call of " #name "\n"); abort(); }
REFH_DEAD(dB_fixed) REFH_DEAD(signal_energy)
/* dft / idft are the loader's function pointers (tools_defs.h); nothing on the receiver's path calls them */
void *dft, *idft;
