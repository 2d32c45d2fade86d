// The Q15 transform kernels of dfts.cu, compiled into libldpc_b200.so with a hidden C ABI: nrb200::dft_batch_internal serves the PUSCH delay
// estimator (nr_est_delay takes an IDFT of the least-squares estimate, common/utils/nr/nr_common.c:968-990).
#define NRB200_DFTS_INTERNAL 1
#include "dfts.cu"
