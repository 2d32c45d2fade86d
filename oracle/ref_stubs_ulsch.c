/* TEST INFRASTRUCTURE ONLY.  Symbols nr_ulsch_decoding.c and its helpers reference on paths ref_harness_ulsch.c never takes (the T2 offload branch, NACK
 * indications, tracing): they abort if reached. */
#include <stdio.h>
#include <stdlib.h>
#define REFH_DEAD(name) void name(void) { fprintf(stderr, "ref_harness_ulsch: unexpected call of " #name "\n"); abort(); }
REFH_DEAD(nr_fill_indication) REFH_DEAD(threadCreate)
