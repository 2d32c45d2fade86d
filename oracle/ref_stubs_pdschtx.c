/* TEST INFRASTRUCTURE ONLY. Process-level symbols nr_dlsch.c and its callees expect from the softmodem executable; none does anything here. */
#include <stddef.h>
#include <stdint.h>
void *dft = NULL, *idft = NULL;
void *get_softmodem_params(void) { static char z[4096]; return z; }
void vcd_signal_dumper_dump_function_by_name(int n, int v) { (void)n; (void)v; }
void vcd_signal_dumper_dump_variable_by_name(int n, unsigned long v) { (void)n; (void)v; }
double cpuf = 1.0;
uint64_t get_softmodem_optmask(void) { return 0; }
void nr_gen_ref_conj_symbols(void) {}
void mult_cpx_vector(void) {}
