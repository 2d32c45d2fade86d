/*
 * nrb200_dfts.h -- C ABI of libdfts_b200.so, the B200-native drop-in for OpenAirInterface's loadable DFT library
 * ("libdfts.so": openair1/PHY/TOOLS/dfts_load.c:47-61 dlsym()s `dft` and `idft`, the loader calls `dfts_autoinit`).
 *
 * Arithmetic: the reference's Q15 fixed point, bit exact (oai_dfts.c), for the OFDM sizes
 *   64 128 256 512 768 1024 1536 2048 3072 4096 6144 8192   (both directions).
 * The DFT-s-OFDM / PRACH sizes of FOREACH_DFTSZ (12..3240 and > 8192) are not implemented yet: calling them aborts loudly
 * (there is no CPU fallback in this library).
 */
#ifndef NRB200_DFTS_H
#define NRB200_DFTS_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Part 1: the OAI plug-in ABI (tools_defs.h:514-521).  sizeidx = position in FOREACH_DFTSZ / FOREACH_IDFTSZ (tools_defs.h:404-499);
 * sigF/sig = interleaved {re, im} int16, one transform; scale_flag as in the reference (0 = none, 1 = 1/sqrt(N) overall). */
void dft(uint8_t sizeidx, int16_t *sigF, int16_t *sig, unsigned char scale_flag);
void idft(uint8_t sizeidx, int16_t *sigF, int16_t *sig, unsigned char scale_flag);
int dfts_autoinit(void);   /* called by load_module_shlib when present (load_module_shlib.c:174-191); returns 0, -1 without a GPU */

/* Part 2: batched extension -- n transforms of the same size, contiguous (2*N int16 each). */
int32_t nrb200_dft_batch_dev(int N, int inverse, uint32_t n, const int16_t *d_in, int16_t *d_out, int scale, void *stream);
int32_t nrb200_dft_batch_host(int N, int inverse, uint32_t n, const int16_t *in, int16_t *out, int scale);
/* N for a reference size index (dft_size_idx_t / idft_size_idx_t), -1 if out of range */
int32_t nrb200_dft_size_of_index(int inverse, int sizeidx);
int32_t nrb200_dft_supported(int N);
const char *nrb200_dfts_last_error(void);
uint64_t nrb200_dfts_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif
