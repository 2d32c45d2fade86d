// libldpc_b200_t2.so -- the "offload" flavour of OAI's loadable codec ABI (ldpc_interface_offload, loaded as libldpc_t2.so when the
// softmodem runs with --ldpc-offload-enable; reference openair1/PHY/CODING/nrLDPC_decoder/nrLDPC_decoder_offload.c:1034-1145).
// Same four symbols as libldpc_b200.so, offload semantics: one segment per call, E raw LLRs in, rate recovery and HARQ combining inside.
// A forwarder: the work is nrb200_ldpc_offload_{decode,encode} in libldpc_b200.so (found next to this file through $ORIGIN).
#include "../../include/nrb200_ldpc.h"

#define SHIM_EXPORT extern "C" __attribute__((visibility("default")))

SHIM_EXPORT int32_t LDPCinit(void) { return nrb200_ldpc_offload_init(); }
SHIM_EXPORT int32_t LDPCshutdown(void) { return nrb200_ldpc_offload_release(); }   // free_LDPClib: the soft buffers go with the module

SHIM_EXPORT int32_t LDPCdecoder(nrb200_ldpc_dec_params_t *p, uint8_t harq_pid, uint8_t ulsch_id, uint8_t C, int8_t *p_llr, int8_t *p_out,
                                nrb200_ldpc_time_stats_t *prof, nrb200_decode_abort_t *ab)
{
  (void)prof; (void)ab;   // both are NULL in offload mode (nr_ulsch_decoding.c:264-266)
  return nrb200_ldpc_offload_decode(p, harq_pid, ulsch_id, C, p_llr, (uint8_t *)p_out);
}

SHIM_EXPORT int32_t LDPCencoder(uint8_t **input, uint8_t **output, nrb200_ldpc_enc_params_t *impp)
{
  return nrb200_ldpc_offload_encode(input[0], output[0], impp);
}
