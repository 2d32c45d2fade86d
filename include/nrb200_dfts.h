/*
 * nrb200_dfts.h -- C ABI of libdfts_b200.so, the B200-native drop-in for OpenAirInterface's loadable DFT library
 * ("libdfts.so": openair1/PHY/TOOLS/dfts_load.c:47-61 dlsym()s `dft` and `idft`, the loader calls `dfts_autoinit`).
 *
 * Arithmetic: the reference's Q15 fixed point, bit exact (oai_dfts.c), for
 *   the OFDM sizes 64 128 256 512 768 1024 1536 2048 3072 4096 6144 8192 (both directions, one transform per call), and
 *   the DFT-s-OFDM family 12 24 36 48 60 72 96 108 120 144 180 192 216 240 288 300 324 360 384 432 480 540 576 600 648 720 864 900 960 972 1080 1152 1200
 *   1296 1440 1500 1620 1728 1800 1920 1944 2160 2304 2400 2592 2700 2880 2916 3000 3240 (forward only, as in the reference; like the reference's entry
 *   points oai_dfts.c:4352-7706 every call transforms FOUR interleaved sequences: c16 number 4 n + l is element n of transform l, 4 N c16 in and out).
 *   DFT_2304: the reference's function combines uninitialised stack (it calls the single-transform dft768 on four-way data, oai_dfts.c:7288-7310) and is not
 *   reproducible; this library returns the 768 x 3 transform the function documents.
 *   the sizes above 8192 (12288 16384 18432 24576 36864 49152 both directions, 32768 / 98304, idft 65536): radix-3 / 4 / 2 levels over global memory on top of the
 *   shared-memory transforms; 32768, 65536 and 98304 crash or read beyond their tables in the reference (DESIGN.md defect 12) and are checked against a float DFT only.
 * 9216 and 73728 are AssertFatal("Need to do this") in the reference and abort here too (there is no CPU fallback in this library).
 */
#ifndef NRB200_DFTS_H
#define NRB200_DFTS_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Part 1: the OAI plug-in ABI (tools_defs.h:514-521).  sizeidx = position in FOREACH_DFTSZ / FOREACH_IDFTSZ (tools_defs.h:404-499);
 * sigF/sig = interleaved {re, im} int16, one transform; scale_flag as in the reference (0 = none, 1 = 1/sqrt(N) overall). */
/* (a translation unit that also includes OAI's own tools_defs.h -- an interposer under integration/ -- defines NRB200_NO_OAI_LOADER_PROTOTYPES: there
 * `dft` / `idft` are OAI's function-pointer globals) */
#ifndef NRB200_NO_OAI_LOADER_PROTOTYPES
void dft(uint8_t sizeidx, int16_t *sigF, int16_t *sig, unsigned char scale_flag);
void idft(uint8_t sizeidx, int16_t *sigF, int16_t *sig, unsigned char scale_flag);
int dfts_autoinit(void);   /* called by load_module_shlib when present (load_module_shlib.c:174-191); returns 0, -1 without a GPU */
#endif

/* Part 2: batched extension -- n calls of the same size, contiguous (2*N int16 each; 8*N int16 each for the four-way DFT-s-OFDM sizes). */
int32_t nrb200_dft_batch_dev(int N, int inverse, uint32_t n, const int16_t *d_in, int16_t *d_out, int scale, void *stream);
int32_t nrb200_dft_batch_host(int N, int inverse, uint32_t n, const int16_t *in, int16_t *out, int scale);
/* N for a reference size index (dft_size_idx_t / idft_size_idx_t), -1 if out of range */
int32_t nrb200_dft_size_of_index(int inverse, int sizeidx);
int32_t nrb200_dft_supported(int N);
const char *nrb200_dfts_last_error(void);
uint64_t nrb200_dfts_launch_count(void);
/* one process, several GPUs: the library keeps one context per device; selects the device the CALLING THREAD's following calls run on
 * (default NRB200_DEVICE, else LOCAL_RANK, else 0).  Device pointers handed to the *_dev entry points must belong to it. */
int32_t nrb200_dfts_set_device(int dev);

/* Part 3: slot-level OFDM front end -- one launch per slot instead of one dft()/idft() call per symbol and antenna, with the work the
 * reference does around each transform fused into the kernel's load and store phases.  All buffers are interleaved {re, im} int16
 * ("c16"); offsets and strides count c16 elements.
 *   nrb200_ofdm_mod_slot_*   replaces apply_nr_rotation_TX + nr_feptx0/PHY_ofdm_mod(CYCLIC_PREFIX)   (ofdm_mod.c:337-376, :130-281;
 *                            nr_ru_procedures.c:52-140): rotate -> IDFT -> cyclic prefix.  The rotated txdataF is not written back.
 *   nrb200_ofdm_demod_slot_* replaces the nr_slot_fep_ul loop of nr_fep_full + apply_nr_rotation_RX  (slot_fep_nr.c:223-332;
 *                            nr_ru_procedures.c:228-262): FFT-window gather (1/ofdm_offset_divisor of the CP early, frame ring wrap)
 *                            -> DFT -> phase and timeshift compensation. */
typedef struct nrb200_ofdm_slot_s {
  uint32_t fft_size;             /* fp->ofdm_symbol_size: one of the sizes nrb200_dft_supported() accepts */
  uint32_t n_symb;               /* symbols handled by this call (1..14), symbol l uses entry l of the arrays below */
  uint32_t n_ant;
  uint32_t f_stride;             /* _dev: c16 between antennas in the frequency-domain buffer (symbol l at l * fft_size) */
  uint32_t t_stride;             /* _dev: c16 between antennas in the time-domain buffer */
  uint32_t t_off[14];            /* TX: first CP sample of symbol l; RX: first sample of symbol l's FFT window */
  uint32_t prefix[14];           /* TX: CP length of symbol l (fp->nb_prefix_samples0 or nb_prefix_samples) */
  uint32_t t_ring;               /* RX: when non-zero, time-domain indices are taken modulo t_ring (fp->samples_per_frame) */
  uint32_t rotate;               /* 0: transform (+CP) only; 1: also apply_nr_rotation_TX / apply_nr_rotation_RX */
  uint32_t nb_rb;                /* N_RB_DL / N_RB_UL: the rotation covers the two half-band ranges the reference covers */
  uint32_t first_carrier_offset; /* fp->first_carrier_offset */
  int16_t rot[14][2];            /* fp->symbol_rotation[link][(slot % slots_per_subframe) * 14 + l] as stored (RX conjugates it itself) */
} nrb200_ofdm_slot_t;
int32_t nrb200_ofdm_mod_slot_dev(const nrb200_ofdm_slot_t *d, const int16_t *d_txdataF, int16_t *d_txdata, void *stream);
int32_t nrb200_ofdm_demod_slot_dev(const nrb200_ofdm_slot_t *d, const int16_t *d_rxdata, const int16_t *d_timeshift, int16_t *d_rxdataF, void *stream);
/* host buffers, one pointer per antenna like ru->common.txdataF_BF[aa] / txdata[aa] / rxdata[aa] / rxdataF[aa]; f_stride / t_stride are
 * ignored; timeshift = fp->timeshift_symbol_rotation (fft_size c16; may be NULL when rotate == 0) */
int32_t nrb200_ofdm_mod_slot_host(const nrb200_ofdm_slot_t *d, const int16_t *const *txdataF, int16_t *const *txdata);
int32_t nrb200_ofdm_demod_slot_host(const nrb200_ofdm_slot_t *d, const int16_t *const *rxdata, const int16_t *timeshift, int16_t *const *rxdataF);

#ifdef __cplusplus
}
#endif
#endif
