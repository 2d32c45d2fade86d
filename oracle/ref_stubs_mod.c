/* TEST INFRASTRUCTURE ONLY. Extra process-level symbols nr_modulation.c expects from the softmodem executable (the dft/idft
 * function-pointer globals of dfts_load.c and get_softmodem_params); never called by the functions the tests use. */
#include <stddef.h>
void *dft = NULL, *idft = NULL;
void *get_softmodem_params(void) { static char z[4096]; return z; }
