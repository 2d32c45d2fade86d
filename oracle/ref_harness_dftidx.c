/* TEST INFRASTRUCTURE ONLY.  OAI's own dft_size_idx_t / idft_size_idx_t enumerator order, expanded from the FOREACH_DFTSZ / FOREACH_IDFTSZ lists of OAI's header
 * (openair1/PHY/TOOLS/tools_defs.h:404-499, 530-545), so that the size index the library's dft() / idft() entry points receive is pinned to the reference. */
#include "PHY/TOOLS/tools_defs.h"
#define NRB200_SZ_NUM(Sz) Sz,
static const int dft_sizes[] = {FOREACH_DFTSZ(NRB200_SZ_NUM)};
static const int idft_sizes[] = {FOREACH_IDFTSZ(NRB200_SZ_NUM)};
int refh_dft_count(void) { return (int)DFT_SIZE_IDXTABLESIZE; }
int refh_idft_count(void) { return (int)IDFT_SIZE_IDXTABLESIZE; }
int refh_dft_size_at(int idx) { return idx >= 0 && idx < (int)DFT_SIZE_IDXTABLESIZE ? dft_sizes[idx] : -1; }
int refh_idft_size_at(int idx) { return idx >= 0 && idx < (int)IDFT_SIZE_IDXTABLESIZE ? idft_sizes[idx] : -1; }
/* the enumerators OAI's OFDM code asks for by size (get_dft / get_idft abort on other sizes) */
int refh_get_dft4096(void) { return (int)get_dft(4096); }
int refh_get_idft4096(void) { return (int)get_idft(4096); }
