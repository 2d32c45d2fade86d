// Decoder kernel selection and launch.
#include "nrb200_ctx.h"
#include "ldpc_common.cuh"
#include "ldpc_decoder_generic.cuh"
#include <cstdlib>

namespace nrb200 {

static size_t generic_smem_bytes(const GraphDev &g)
{
  const size_t numLLR = (size_t)g.ncols * g.Z;
  auto a16 = [](size_t v) { return (v + 15) & ~(size_t)15; };
  return a16(sizeof(GraphDev)) + a16(numLLR) + a16((size_t)g.nreal * g.Z) + a16(numLLR) + a16((size_t)g.nrows * g.Z);
}

// Returns 0 or a negative error.  Asynchronous on `stream`.
int launch_decode(const GraphDev *d_g, const GraphDev &h_g, const DecodeArgs &a, cudaStream_t stream)
{
  Ctx &c = ctx();
  if (a.n_cb == 0) return 0;
  {
    const size_t smem = generic_smem_bytes(h_g);
    if ((int)smem > c.max_smem_optin) return -3;
    static std::atomic<size_t> configured{0};
    if (smem > configured.load()) {
      NRB200_CUDA_OK(cudaFuncSetAttribute(ldpc_decode_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "smem attr");
      configured.store(smem);
    }
    int threads = h_g.Z * 2;
    threads = ((threads + 31) / 32) * 32;
    if (threads < 128) threads = 128;
    if (threads > 512) threads = 512;
    const unsigned grid = a.n_cb;
    ldpc_decode_generic_kernel<<<grid, threads, smem, stream>>>(d_g, a);
    c.launches++;
    NRB200_CUDA_OK(cudaGetLastError(), "decode launch");
  }
  return 0;
}

}  // namespace nrb200
