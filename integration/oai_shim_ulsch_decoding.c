/* Link-time interposer for OpenAirInterface: nr_ulsch_decoding on the GPU with host C unchanged -- a whole transport block per library call.
 *
 * The reference (openair1/PHY/NR_TRANSPORT/nr_ulsch_decoding.c:320-470) queues one job per code-block segment on the gNB's thread pool; each job
 * (nr_processULSegment, :121-230) de-interleaves, rate-recovers into harq_process->d[r], packs the decoder input and makes one blocking LDPCdecoder call.
 * The caller (phy_procedures_gNB_uespec_RX -> nr_postDecode, SCHED_NR/phy_procedures_nr_gNB.c:240-330) then pulls C results from gNB->respDecode, copies
 * harq_process->c[r] into the transport block and checks its CRC.  This file DEFINES nr_ulsch_decoding with the reference's prototype
 * (PHY/NR_TRANSPORT/nr_transport_proto.h), keeps every side effect the caller and the statistics code read -- harq_process C / K / Z / F / llrLen / TBS /
 * processedSegments, the d_to_be_cleared flags, the per-UE statistics, harq_process->c[r], one ldpcDecode_t per segment on gNB->respDecode with
 * decodeIterations filled in -- and replaces the C jobs by ONE call of nrb200_ulsch_decode_tb_host (include/nrb200_slot.h): one rate-recovery launch and
 * one decode launch for all segments, soft buffers resident on the GPU and keyed by the HARQ process.  The unchanged caller's pull / nr_postDecode loop
 * runs as before.  Link it ahead of libPHY_NR with -Wl,--allow-multiple-definition (the reference's own definition shares its object file with
 * new_gNB_ulsch / free_gNB_ulsch, which stay in use); integration/build_shims.sh shows the command.
 * Not reproduced: the abort flag (a failed segment stops its siblings early; the block is lost either way) and the ldpc_offload_flag branch (this IS the
 * offload).  NRB200_SHIM_MIRROR_HARQ=1 also copies the combined soft buffers back into harq_process->d[r] after every call.
 * Test: tests/test_gpu_interpose.py drives it through oracle/ref_harness_ulsch.c next to the reference's own nr_ulsch_decoding + CPU decoder. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include "PHY/defs_gNB.h"
#include "PHY/CODING/coding_extern.h"
#include "PHY/CODING/coding_defs.h"
#include "PHY/NR_TRANSPORT/nr_transport_proto.h"
#include "PHY/NR_TRANSPORT/nr_ulsch.h"
#include "PHY/NR_TRANSPORT/nr_dlsch.h"
#define NRB200_NO_OAI_LOADER_PROTOTYPES
#include "nrb200_slot.h"

/* pusch_vars->llr is allocated once per ULSCH at start-up (init_nr_transport): page-lock each such buffer the first time it is seen (and again if a longer
 * stretch of it is used) so that the GPU reads it in place.  A small table under a mutex: the function is called from several threads for different ULSCHs. */
static int llr_is_pinned(short *llr, size_t bytes)
{
  static pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
  static struct { short *p; size_t bytes; } tab[64];
  int ok = 0;
  pthread_mutex_lock(&mu);
  int slot = -1;
  for (int i = 0; i < 64; i++) {
    if (tab[i].p == llr) { slot = i; break; }
    if (slot < 0 && tab[i].p == NULL) slot = i;
  }
  if (slot >= 0) {
    if (tab[slot].p == llr && tab[slot].bytes >= bytes) ok = 1;
    else {
      if (tab[slot].p == llr) nrb200_host_unregister(llr);
      tab[slot].p = NULL;
      if (nrb200_host_register(llr, bytes) == 0) { tab[slot].p = llr; tab[slot].bytes = bytes; ok = 1; }
    }
  }
  pthread_mutex_unlock(&mu);
  return ok;
}

int nr_ulsch_decoding(PHY_VARS_gNB *gNB, uint8_t ULSCH_id, short *ulsch_llr, NR_DL_FRAME_PARMS *frame_parms, nfapi_nr_pusch_pdu_t *pusch_pdu, uint32_t frame,
                      uint8_t nr_tti_rx, uint8_t harq_pid, uint32_t G)
{
  if (!ulsch_llr) return -1;
  NR_gNB_ULSCH_t *ulsch = &gNB->ulsch[ULSCH_id];
  NR_gNB_PUSCH *pusch = &gNB->pusch_vars[ULSCH_id];
  NR_UL_gNB_HARQ_t *hp = ulsch->harq_process;
  if (!hp) return -1;
  const int Qm = pusch_pdu->qam_mod_order, nl = pusch_pdu->nrOfLayers, rv = pusch_pdu->pusch_data.rv_index, BG = pusch_pdu->maintenance_parms_v3.ldpcBaseGraph;
  hp->processedSegments = 0;
  hp->TBS = pusch_pdu->pusch_data.tb_size;
  const uint32_t A = hp->TBS << 3;
  /* per-UE statistics, as the reference updates them before decoding (:359-372) */
  NR_gNB_PHY_STATS_t *stats = get_phy_stats(gNB, ulsch->rnti);
  if (stats) {
    stats->frame = frame;
    stats->ulsch_stats.round_trials[hp->round]++;
    for (int a = 0; a < frame_parms->nb_antennas_rx; a++) {
      stats->ulsch_stats.power[a] = dB_fixed_x10(pusch->ulsch_power[a]);
      stats->ulsch_stats.noise_power[a] = dB_fixed_x10(pusch->ulsch_noise_power[a]);
    }
    if (!hp->harq_to_be_cleared) {
      stats->ulsch_stats.current_Qm = Qm;
      stats->ulsch_stats.current_RI = nl;
      stats->ulsch_stats.total_bytes_tx += hp->TBS;
    }
  }
  /* C, K, Zc, F of the transport block: OAI's own nr_segmentation in its parameters-only mode */
  nr_segmentation(NULL, NULL, lenWithCrc(1, A), &hp->C, &hp->K, &hp->Z, &hp->F, BG);
  int room = MAX_NUM_NR_ULSCH_SEGMENTS_PER_LAYER * nl;
  if ((int)hp->C > room) return -1;
  if (pusch_pdu->rb_size != 273) room = room * pusch_pdu->rb_size / 273 + 1;
  if ((int)hp->C > room) return -1;
  const int C = hp->C;
  if (hp->harq_to_be_cleared) {
    for (int r = 0; r < C; r++) hp->d_to_be_cleared[r] = true;
    hp->harq_to_be_cleared = false;
  }
  set_abort(&hp->abort_decode, false);

  nrb200_ulsch_tb_t d;
  memset(&d, 0, sizeof(d));
  d.rm.BG = BG; d.rm.Z = hp->Z; d.rm.Qm = Qm; d.rm.rv = rv; d.rm.C = C; d.rm.n_seg = C; d.rm.Tbslbrm = pusch_pdu->maintenance_parms_v3.tbSizeLbrmBytes;
  d.rm.F = hp->F; d.rm.K = hp->K;
  d.numMaxIter = ulsch->max_ldpc_iterations;
  d.crc_type = crcType(C, A); d.crc_len_bits = lenWithCrc(C, A);
  d.harq_key = (uint64_t)(uintptr_t)hp;
  uint32_t E[C];
  uint8_t R[C], clear[C];
  int32_t iters[C];
  for (int r = 0; r < C; r++) {
    E[r] = nr_get_E(G, C, Qm, nl, r);
    R[r] = nr_get_R_ldpc_decoder(rv, E[r], BG, hp->Z, &hp->llrLen, hp->round);
    clear[r] = hp->d_to_be_cleared[r];
    memset(hp->c[r], 0, hp->K >> 3);                                  /* nr_processULSegment :177 */
  }
  d.llr_pinned = llr_is_pinned(ulsch_llr, 2 * (size_t)G);
  static int mirror = -1;
  if (mirror < 0) { const char *e = getenv("NRB200_SHIM_MIRROR_HARQ"); mirror = e && *e == '1'; }
  const int rc = nrb200_ulsch_decode_tb_host(&d, ulsch_llr, E, R, clear, (uint8_t *const *)hp->c, iters, mirror ? (int16_t *const *)hp->d : NULL);
  if (rc != 0) { fprintf(stderr, "nrb200 shim: nrb200_ulsch_decode_tb_host failed (rc = %d: %s)\n", rc, nrb200_last_error()); abort(); }

  /* one result per segment on the response FIFO, filled like the reference fills its jobs (:437-463) */
  uint32_t offset = 0, r_offset = 0;
  for (int r = 0; r < C; r++) {
    hp->d_to_be_cleared[r] = false;
    union ldpcReqUnion id = {.s = {ulsch->rnti, frame, nr_tti_rx, 0, 0}};
    notifiedFIFO_elt_t *req = newNotifiedFIFO_elt(sizeof(ldpcDecode_t), id.p, &gNB->respDecode, NULL);
    ldpcDecode_t *rd = (ldpcDecode_t *)NotifiedFifoData(req);
    memset(rd, 0, sizeof(*rd));
    rd->gNB = gNB; rd->ulsch_harq = hp; rd->ulsch = ulsch; rd->ulsch_llr = ulsch_llr; rd->ulsch_id = ULSCH_id; rd->harq_pid = harq_pid;
    rd->decoderParms.BG = BG; rd->decoderParms.Z = hp->Z; rd->decoderParms.R = R[r]; rd->decoderParms.numMaxIter = ulsch->max_ldpc_iterations;
    rd->decoderParms.outMode = 0; rd->decoderParms.crc_type = d.crc_type; rd->decoderParms.E = d.crc_len_bits; rd->decoderParms.check_crc = check_crc;
    rd->Kc = BG == 2 ? 52 : 68; rd->segment_r = r; rd->nbSegments = C; rd->E = E[r]; rd->A = A; rd->Qm = Qm; rd->r_offset = r_offset;
    rd->Kr_bytes = hp->K >> 3; rd->rv_index = rv; rd->offset = offset; rd->tbslbrm = pusch_pdu->maintenance_parms_v3.tbSizeLbrmBytes;
    rd->decodeIterations = iters[r];
    pushNotifiedFIFO(&gNB->respDecode, req);
    r_offset += E[r];
    offset += (hp->K >> 3) - (hp->F >> 3) - (C > 1 ? 3 : 0);
  }
  return C;
}
