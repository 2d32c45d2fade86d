"""Host-side binding of libldpc_b200.so -- the B200-native drop-in for OAI's loadable LDPC codec.

This mirrors the reference's `ldpc_interface_t` (openair1/PHY/CODING/nrLDPC_defs.h:68-87): `LDPCinit`, `LDPCshutdown`,
`LDPCdecoder`, `LDPCencoder` with the same parameter structs, plus the batched extension declared in
include/nrb200_ldpc.h.  Everything computes on the GPU; if the shared object or a CUDA device is missing the calls raise --
there is no CPU fallback.  PyTorch is used only to own device memory / streams for the device-resident entry points.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.environ.get("NRB200_LIB") or os.path.join(_HERE, "libldpc_b200.so")   # NRB200_LIB: A/B builds of tools/ (never set in production)

_u8p = C.POINTER(C.c_uint8)
_i8p = C.POINTER(C.c_int8)

OUTMODE_BIT, OUTMODE_BITINT8, OUTMODE_LLRINT8 = 0, 1, 2
CRC24_A, CRC24_B, CRC16, CRC8 = 0, 1, 2, 3          # coding_defs.h:33-36
POLY_24A, POLY_24B, POLY_24C, POLY_16, POLY_12, POLY_11, POLY_8, POLY_6 = range(8)


class TimeStats(C.Structure):          # include/nrb200_ldpc.h nrb200_time_stats_t
    _fields_ = [("in_", C.c_longlong), ("diff", C.c_longlong), ("p_time", C.c_longlong), ("diff_square", C.c_double),
                ("max", C.c_longlong), ("trials", C.c_int), ("meas_flag", C.c_int), ("meas_name", C.c_char_p),
                ("meas_index", C.c_int), ("meas_enabled", C.c_int), ("tpoolmsg", C.c_void_p), ("tstatptr", C.c_void_p)]


class LdpcTimeStats(C.Structure):
    _fields_ = [(n, TimeStats) for n in ("llr2llrProcBuf", "llr2CnProcBuf", "cnProc", "cnProcPc", "bnProcPc", "bnProc",
                                         "cn2bnProcBuf", "bn2cnProcBuf", "llrRes2llrOut", "llr2bit", "total")]


class DecParams(C.Structure):          # t_nrLDPC_dec_params
    _fields_ = [("BG", C.c_uint8), ("Z", C.c_uint16), ("R", C.c_uint8), ("F", C.c_uint16), ("Qm", C.c_uint8), ("rv", C.c_uint8),
                ("numMaxIter", C.c_uint8), ("E", C.c_int), ("outMode", C.c_int), ("crc_type", C.c_int),
                ("check_crc", C.c_void_p), ("setCombIn", C.c_uint8)]


class DecodeAbort(C.Structure):        # decode_abort_t
    _fields_ = [("mutex", C.c_uint64 * 5), ("failed", C.c_bool)]     # pthread_mutex_t: 40 bytes, 8-byte aligned (x86-64 glibc) -> sizeof 48 like the C struct


class EncParams(C.Structure):          # encoder_implemparams_t
    _fields_ = [("n_segments", C.c_uint), ("macro_num", C.c_uint), ("gen_code", C.c_ubyte),
                ("tinput", C.c_void_p), ("tprep", C.c_void_p), ("tparity", C.c_void_p), ("toutput", C.c_void_p),
                ("Kr", C.c_int), ("Kb", C.c_uint32), ("Zc", C.c_uint32), ("harq", C.c_void_p), ("BG", C.c_uint8),
                ("output", C.c_void_p), ("K", C.c_uint32), ("F", C.c_uint32), ("Qm", C.c_uint8), ("E", C.c_uint32),
                ("G", C.c_uint), ("rv", C.c_uint8)]


class BatchDesc(C.Structure):          # nrb200_ldpc_batch_desc_t
    _fields_ = [("BG", C.c_uint8), ("Z", C.c_uint16), ("R", C.c_uint8), ("numMaxIter", C.c_uint8), ("outMode", C.c_uint8),
                ("crc_type", C.c_uint8), ("use_crc", C.c_uint8), ("latency_mode", C.c_uint8), ("crc_len_bits", C.c_uint32), ("n_cb", C.c_uint32),
                ("llr_stride", C.c_uint32), ("out_stride", C.c_uint32)]


class RmDesc(C.Structure):             # nrb200_rm_desc_t
    _fields_ = [("BG", C.c_uint8), ("Z", C.c_uint16), ("Qm", C.c_uint8), ("rv", C.c_uint8), ("clear", C.c_uint8), ("C", C.c_uint32),
                ("Tbslbrm", C.c_uint32), ("F", C.c_uint32), ("K", C.c_uint32), ("n_seg", C.c_uint32)]


class Nrb200Error(RuntimeError):
    pass


def ncols_for_rate(BG, R):
    """Columns the decoder LUT for rate selector R spans (nrLDPCdecoder_defs.h:53-57,80-84)."""
    return {(1, 13): 68, (1, 23): 35, (1, 89): 27, (2, 15): 52, (2, 13): 32, (2, 23): 17}[(BG, R)]


class PuschRxDesc(C.Structure):       # nrb200_pusch_rx_t (field names of nfapi_nr_pusch_pdu_t / NR_DL_FRAME_PARMS)
    _fields_ = [(n, C.c_uint32) for n in ("fft_size", "nb_rx", "rb_start", "bwp_start", "rb_size", "first_carrier_offset", "qam_mod_order",
                                          "start_symbol_index", "nr_of_symbols", "ul_dmrs_symb_pos", "dmrs_config_type", "num_dmrs_cdm_grps_no_data",
                                          "log2_maxh", "rx_stride", "ch_stride", "unscramble", "rnti", "data_scrambling_id", "nrOfLayers", "noise_var", "max_ch", "pdsch_ue")] + \
               [("d_est_state", C.c_uint64), ("est_state_ports", C.c_uint32), ("transform_precoding", C.c_uint32), ("d_tp_scratch", C.c_uint64)] + \
               [(n, C.c_uint32) for n in ("ptrs", "ptrs_time_density", "ptrs_freq_density", "ptrs_re_offset", "ptrs_slot", "ptrs_nscid", "ptrs_dmrs_scrambling_id",
                                          "ptrs_reserved")] + [("d_ptrs_state", C.c_uint64)]

    def set_ptrs(self, time_density, freq_density, re_offset, slot, nscid, dmrs_scrambling_id, d_state=0):
        """PT-RS at the UE (one layer): the fields of fapi_nr_dl_config_dlsch_pdu_rel15_t nr_pdsch_ptrs_processing reads; self.rnti is dlsch[0].rnti."""
        self.ptrs, self.ptrs_time_density, self.ptrs_freq_density, self.ptrs_re_offset = 1, time_density, freq_density, re_offset
        self.ptrs_slot, self.ptrs_nscid, self.ptrs_dmrs_scrambling_id, self.d_ptrs_state = slot, nscid, dmrs_scrambling_id, d_state
        return self


class PrachDesc(C.Structure):         # nrb200_prach_t
    _fields_ = [(n, C.c_uint32) for n in ("nb_rx", "short_sequence", "num_cs", "prach_format", "numerology", "restricted_set", "rx_stride", "reserved")]


class RfsimChan(C.Structure):         # nrb200_rfsim_chan_t
    _fields_ = [("nb_tx", C.c_uint32), ("nb_rx", C.c_uint32), ("channel_length", C.c_uint32), ("channel_offset", C.c_int32), ("path_loss_dB", C.c_double),
                ("noise_power_dB", C.c_float), ("reserved", C.c_uint32)]


class PuschChestDesc(C.Structure):    # nrb200_pusch_chest_t
    _fields_ = [(n, C.c_uint32) for n in ("fft_size", "nb_rx", "slot", "symbol", "port", "rb_start", "bwp_start", "rb_size", "first_carrier_offset", "scid",
                                          "ul_dmrs_scrambling_id", "rx_stride", "ch_stride", "n_ports", "pdsch_ue", "dmrs_config_type", "chest_freq",
                                          "transform_precoding")] + [("lowpapr_seq", C.c_uint64)]

    def set_lowpapr(self, seq):
        """Transform precoding: seq = the 6 * rb_size c16 low-PAPR type-1 sequence -- a numpy int16 array (host entry points) or a torch CUDA tensor
        (_dev entry points); None switches back to the Gold-sequence DMRS.  The descriptor keeps the buffer alive."""
        self._seq = seq
        if seq is None:
            self.transform_precoding, self.lowpapr_seq = 0, 0
        else:
            self.transform_precoding = 1
            self.lowpapr_seq = seq.data_ptr() if hasattr(seq, "data_ptr") else seq.ctypes.data
        return self


class PdschTxDesc(C.Structure):       # nrb200_pdsch_tx_t (field names of nfapi_nr_dl_tti_pdsch_pdu_rel15_t / NR_DL_FRAME_PARMS)
    _fields_ = [(n, C.c_uint32) for n in ("fft_size", "nb_tx", "slot", "rb_start", "bwp_start", "rb_size", "first_carrier_offset", "qam_mod_order", "nrOfLayers",
                                          "start_symbol_index", "nr_of_symbols", "dl_dmrs_symb_pos", "dmrs_config_type", "num_dmrs_cdm_grps_no_data", "dmrs_ports",
                                          "scid", "dl_dmrs_scrambling_id", "data_scrambling_id", "rnti", "amp", "tx_stride", "pm_idx")] + [("pm_weights", C.c_int16 * 32)] + \
               [(n, C.c_uint32) for n in ("ptrs", "ptrs_time_density", "ptrs_freq_density", "ptrs_re_offset")]

    def set_ptrs(self, time_density, freq_density, re_offset):
        """PT-RS insertion (pduBitmap & 1): PTRSTimeDensity (log2 of L), PTRSFreqDensity (K), PTRSReOffset."""
        self.ptrs, self.ptrs_time_density, self.ptrs_freq_density, self.ptrs_re_offset = 1, time_density, freq_density, re_offset
        return self

    def set_precoding(self, pm_idx, weights):
        """Wideband precoding matrix: weights [layers <= 4][antennas <= 4][2] int16 (nfapi_nr_pm_pdu_t.weights); pm_idx 0 = identity."""
        w = np.zeros((4, 4, 2), np.int16)
        if weights is not None:
            a = np.asarray(weights, dtype=np.int16)
            w[:a.shape[0], :a.shape[1]] = a
        self.pm_idx = pm_idx
        self.pm_weights = (C.c_int16 * 32)(*[int(x) for x in w.reshape(-1)])
        return self


def _ofdm_desc_cls():
    from .ofdm import OfdmSlotDesc
    return OfdmSlotDesc


class SchRxSlotDesc(C.Structure):     # nrb200_sch_rx_slot_t (include/nrb200_slot.h)
    pass


class SchRxBufs(C.Structure):         # nrb200_sch_rx_bufs_t
    _fields_ = [(n, C.c_void_p) for n in ("d_rxdata", "d_timeshift", "d_rxdataF", "d_est", "d_chest_scratch", "d_chest_state", "d_level", "d_llr16", "d_E", "d_Eoff",
                                          "d_harq", "d_llr8", "d_hard", "d_iters", "d_tb", "d_tbcrc")] + \
               [(n, C.c_uint32) for n in ("harq_stride", "llr8_stride", "hard_stride", "reserved")]


class PdschTxSlotDesc(C.Structure):   # nrb200_pdsch_tx_slot_t
    pass


class PdschTxBufs(C.Structure):       # nrb200_pdsch_tx_bufs_t
    _fields_ = [(n, C.c_void_p) for n in ("d_payload", "d_segs", "d_seg_scratch", "d_cw", "d_E", "d_Eoff", "d_f", "d_txdataF", "d_txdata")] + \
               [("seg_stride", C.c_uint32), ("cw_stride", C.c_uint32)]


def _late_fields():
    """The slot descriptors embed nrb200_ofdm_slot_t, whose mirror lives in ofdm.py (which imports nothing from here)."""
    if not hasattr(SchRxSlotDesc, "ofdm"):
        O = _ofdm_desc_cls()
        SchRxSlotDesc._fields_ = [("ofdm", O), ("chest", PuschChestDesc), ("rx", PuschRxDesc), ("rm", RmDesc), ("R", C.c_uint8), ("numMaxIter", C.c_uint8),
                                  ("use_estimates", C.c_uint8), ("latency_mode", C.c_uint8), ("crc_len_bits", C.c_uint32), ("seg_crc_type", C.c_uint32),
                                  ("A", C.c_uint32), ("tb_crc_bits", C.c_uint32), ("seg_payload_bytes", C.c_uint32)]
        PdschTxSlotDesc._fields_ = [("tx", PdschTxDesc), ("ofdm", O), ("rm", RmDesc), ("A", C.c_uint32), ("K", C.c_uint32)]


class LdpcLib:
    """ldpc_interface_t equivalent bound to libldpc_b200.so."""

    def __init__(self, path=_SO):
        if not os.path.exists(path):
            raise Nrb200Error(f"{path} not built: run `python -c 'import __graft_entry__ as g; g.build()'` (nvcc, sm_100a)")
        L = self.lib = C.CDLL(path)
        L.LDPCinit.restype = C.c_int32
        L.LDPCshutdown.restype = C.c_int32
        L.LDPCdecoder.restype = C.c_int32
        L.LDPCdecoder.argtypes = [C.POINTER(DecParams), C.c_uint8, C.c_uint8, C.c_uint8, _i8p, _i8p, C.c_void_p, C.c_void_p]
        L.LDPCencoder.restype = C.c_int32
        L.LDPCencoder.argtypes = [C.POINTER(_u8p), C.POINTER(_u8p), C.POINTER(EncParams)]
        L.nrb200_ldpc_num_llr.argtypes = [C.c_int] * 3
        L.nrb200_ldpc_decode_batch_dev.argtypes = [C.POINTER(BatchDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.nrb200_ldpc_decode_batch_host.argtypes = [C.POINTER(BatchDesc), C.c_void_p, C.c_void_p, C.c_void_p]
        L.nrb200_ldpc_decode_batch_host_submit.argtypes = [C.POINTER(BatchDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
        L.nrb200_ldpc_decode_batch_host_wait.argtypes = [C.c_void_p]
        L.nrb200_ldpc_encode_batch_dev.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p]
        L.nrb200_ldpc_encode_batch_host.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32]
        L.nrb200_crc_batch_dev.argtypes = [C.c_int, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
        L.nrb200_crc_batch_host.argtypes = [C.c_int, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
        L.nrb200_ldpc_rm_tx_batch_dev.argtypes = [C.POINTER(RmDesc), C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.nrb200_ldpc_rm_rx_batch_dev.argtypes = [C.POINTER(RmDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p]
        L.nrb200_ldpc_rm_tx_batch_host.argtypes = [C.POINTER(RmDesc), C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
        L.nrb200_ldpc_rm_rx_batch_host.argtypes = [C.POINTER(RmDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32]
        L.nrb200_pusch_llr_host.argtypes = [C.c_int, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.nrb200_pusch_llr_dev.argtypes = [C.c_int, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.nrb200_pusch_dmrs_pilots_host.argtypes = [C.c_void_p, C.c_void_p]
        L.nrb200_pusch_chest_scratch_bytes.argtypes = [C.c_void_p]
        L.nrb200_pusch_chest_scratch_bytes.restype = C.c_uint64
        L.nrb200_pusch_chest_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.nrb200_pusch_chest_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.nrb200_pusch_num_llr.argtypes = [C.c_void_p]
        L.nrb200_pusch_num_llr.restype = C.c_uint32
        L.nrb200_pusch_log2_maxh_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.nrb200_pusch_inner_rx_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.nrb200_pusch_inner_rx_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.nrb200_scramble_dev.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
        L.nrb200_unscramble_llr_dev.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
        L.nrb200_modulate_dev.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p]
        L.nrb200_scramble_host.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
        L.nrb200_unscramble_llr_host.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
        L.nrb200_modulate_host.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_void_p]
        L.nrb200_last_error.restype = C.c_char_p
        L.nrb200_launch_count.restype = C.c_uint64
        self._inited = False

    # ---- lifecycle (load_LDPClib / free_LDPClib, nrLDPC_load.c:46-76)
    def init(self):
        if self.lib.LDPCinit() != 0:
            raise Nrb200Error("LDPCinit failed: " + self.last_error())
        self._inited = True
        return 0

    def shutdown(self):
        self._inited = False
        return self.lib.LDPCshutdown()

    # ---- slot-level entry points (include/nrb200_slot.h): one call per slot, device resident, stream ordered
    def sch_slot_rx_torch(self, desc, bufs, device):
        import torch
        self._check(self.lib.nrb200_sch_slot_rx_dev(C.byref(desc), C.byref(bufs), C.c_void_p(torch.cuda.current_stream(device).cuda_stream)), "sch_slot_rx_dev")

    def pdsch_slot_tx_torch(self, desc, bufs, device):
        import torch
        self._check(self.lib.nrb200_pdsch_slot_tx_dev(C.byref(desc), C.byref(bufs), C.c_void_p(torch.cuda.current_stream(device).cuda_stream)), "pdsch_slot_tx_dev")

    # ---- one process, several GPUs
    def device_count(self):
        return int(self.lib.nrb200_device_count())

    def set_device(self, dev):
        """Selects the device the calling thread's following calls run on."""
        self._check(self.lib.nrb200_set_device(int(dev)), "set_device")

    def sticky_device(self, ulsch_id, r, n_dev):
        return int(self.lib.nrb200_sticky_device(int(ulsch_id), int(r), int(n_dev)))

    def decode_batch_host_multi(self, BG, Z, R, numMaxIter, llr, n_dev, outMode=OUTMODE_BIT, use_crc=0, crc_len_bits=0, crc_type=0, out=None, iters=None):
        """decode_batch_host spread over devices 0 .. n_dev-1 of this process."""
        llr = np.ascontiguousarray(llr, dtype=np.int8)
        n_cb, stride = llr.shape
        n = ncols_for_rate(BG, R) * Z
        ob = (n + 7) // 8 if outMode == OUTMODE_BIT else n
        if out is None:
            out = np.zeros((n_cb, ob), dtype=np.uint8)
        if iters is None:
            iters = np.zeros(n_cb, dtype=np.int32)
        d = self._desc(BG, Z, R, numMaxIter, outMode, n_cb, stride, out.shape[1], use_crc, crc_len_bits, crc_type)
        self._check(self.lib.nrb200_ldpc_decode_batch_host_multi(C.byref(d), llr.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p),
                                                                iters.ctypes.data_as(C.c_void_p), int(n_dev)), "decode_batch_host_multi")
        return iters, out

    def last_error(self):
        return (self.lib.nrb200_last_error() or b"").decode()

    def launch_count(self):
        return int(self.lib.nrb200_launch_count())

    def _check(self, rc, what):
        if rc != 0:
            raise Nrb200Error(f"{what} failed (rc={rc}): {self.last_error()}")

    # ---- the per-code-block ABI, exactly as OAI calls it
    def LDPCdecoder(self, BG, Z, R, numMaxIter, llr, outMode=OUTMODE_BIT, E=0, crc_type=0, check_crc=None, abort=None, profiler=None):
        """Returns (iterations, out).  `check_crc`: None => parity-check stop; any truthy value => CRC stop (evaluated on device
        with the reference's check_crc semantics).  `abort`: optional DecodeAbort shared between segments."""
        n = ncols_for_rate(BG, R) * Z
        llr = np.ascontiguousarray(llr, dtype=np.int8)
        assert llr.size >= n
        out = np.zeros(n if outMode != OUTMODE_BIT else (n + 7) // 8, dtype=np.int8)
        p = DecParams()
        p.BG, p.Z, p.R, p.numMaxIter, p.outMode, p.E, p.crc_type = BG, Z, R, numMaxIter, outMode, E, crc_type
        p.check_crc = 1 if check_crc else None
        it = self.lib.LDPCdecoder(C.byref(p), 0, 0, 0, llr.ctypes.data_as(_i8p), out.ctypes.data_as(_i8p),
                                  C.cast(C.byref(profiler), C.c_void_p) if profiler is not None else None,
                                  C.cast(C.byref(abort), C.c_void_p) if abort is not None else None)
        if it < 0:
            raise Nrb200Error("LDPCdecoder failed: " + self.last_error())
        return it, (out.view(np.uint8) if outMode == OUTMODE_BIT else out)

    def LDPCencoder(self, BG, Z, K, payloads):
        """payloads: (n_seg, ceil(K/8)) uint8.  Encodes in groups of 8 (macro_num) like nr_dlsch_coding.c:389-395.
        Returns (n_seg, 66Z|50Z) uint8, one bit per byte."""
        payloads = np.ascontiguousarray(payloads, dtype=np.uint8)
        nseg = payloads.shape[0]
        nout = (66 if BG == 1 else 50) * Z
        outs = np.zeros((nseg, nout), dtype=np.uint8)
        inp = (_u8p * nseg)(*[C.cast(payloads[i].ctypes.data, _u8p) for i in range(nseg)])
        oup = (_u8p * nseg)(*[C.cast(outs[i].ctypes.data, _u8p) for i in range(nseg)])
        ip = EncParams()
        ip.n_segments, ip.Kb, ip.Zc, ip.BG, ip.K = nseg, (22 if BG == 1 else 10), Z, BG, K
        for m in range((nseg + 7) // 8):
            ip.macro_num = m
            self._check(self.lib.LDPCencoder(inp, oup, C.byref(ip)), "LDPCencoder")
        return outs

    # ---- batched extension, host buffers (end-to-end: H2D + kernel + D2H inside the call)
    def _desc(self, BG, Z, R, numMaxIter, outMode, n_cb, llr_stride, out_stride, use_crc=0, crc_len_bits=0, crc_type=0, latency_mode=0):
        d = BatchDesc()
        d.BG, d.Z, d.R, d.numMaxIter, d.outMode, d.latency_mode = BG, Z, R, numMaxIter, outMode, latency_mode
        d.use_crc, d.crc_len_bits, d.crc_type = use_crc, crc_len_bits, crc_type
        d.n_cb, d.llr_stride, d.out_stride = n_cb, llr_stride, out_stride
        return d

    def decode_batch_host(self, BG, Z, R, numMaxIter, llr, outMode=OUTMODE_BIT, use_crc=0, crc_len_bits=0, crc_type=0, out=None, iters=None, latency_mode=0):
        llr = np.ascontiguousarray(llr, dtype=np.int8)
        n_cb, stride = llr.shape
        n = ncols_for_rate(BG, R) * Z
        ob = (n + 7) // 8 if outMode == OUTMODE_BIT else n
        if out is None:
            out = np.zeros((n_cb, ob), dtype=np.uint8)
        if iters is None:
            iters = np.zeros(n_cb, dtype=np.int32)
        d = self._desc(BG, Z, R, numMaxIter, outMode, n_cb, stride, out.shape[1], use_crc, crc_len_bits, crc_type, latency_mode)
        self._check(self.lib.nrb200_ldpc_decode_batch_host(C.byref(d), llr.ctypes.data, out.ctypes.data, iters.ctypes.data), "decode_batch_host")
        return iters, out

    def decode_batch_host_submit(self, BG, Z, R, numMaxIter, llr, out, iters, outMode=OUTMODE_BIT, use_crc=0, crc_len_bits=0, crc_type=0):
        """First half of decode_batch_host (enqueue): returns a ticket for decode_batch_host_wait.  llr (int8, C-contiguous, 2-D), out and
        iters are used in place and must stay alive and untouched until the wait; results are valid only after it."""
        assert llr.dtype == np.int8 and llr.flags.c_contiguous and out.flags.c_contiguous and iters.dtype == np.int32
        n_cb, stride = llr.shape
        d = self._desc(BG, Z, R, numMaxIter, outMode, n_cb, stride, out.shape[1], use_crc, crc_len_bits, crc_type)
        t = C.c_void_p()
        self._check(self.lib.nrb200_ldpc_decode_batch_host_submit(C.byref(d), llr.ctypes.data, out.ctypes.data, iters.ctypes.data, C.byref(t)),
                    "decode_batch_host_submit")
        return (t, llr, out, iters)   # the tuple keeps the buffers alive

    def decode_batch_host_wait(self, ticket):
        """Second half (dequeue): blocks until the batch of `ticket` is complete; returns (iters, out)."""
        t, _, out, iters = ticket
        self._check(self.lib.nrb200_ldpc_decode_batch_host_wait(t), "decode_batch_host_wait")
        return iters, out

    def encode_batch_host(self, BG, Z, K, payloads):
        payloads = np.ascontiguousarray(payloads, dtype=np.uint8)
        n_cb, stride = payloads.shape
        nout = (66 if BG == 1 else 50) * Z
        out = np.zeros((n_cb, nout), dtype=np.uint8)
        self._check(self.lib.nrb200_ldpc_encode_batch_host(BG, Z, K, n_cb, payloads.ctypes.data, stride, out.ctypes.data, nout), "encode_batch_host")
        return out

    def crc_batch_host(self, poly_id, data, bitlen):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        n, stride = data.shape
        out = np.zeros(n, dtype=np.uint32)
        self._check(self.lib.nrb200_crc_batch_host(poly_id, n, data.ctypes.data, stride, bitlen, out.ctypes.data), "crc_batch_host")
        return out

    def crc_batch_torch(self, poly_id, data, bitlen, out=None):
        """data: (n, stride) uint8 on the device; returns uint32 (n,) left-aligned CRCs like crc_byte.c."""
        import torch
        n, stride = data.shape
        if out is None:
            out = torch.empty(n, dtype=torch.int32, device=data.device)
        self._check(self.lib.nrb200_crc_batch_dev(poly_id, n, data.data_ptr(), stride, bitlen, out.data_ptr(),
                                                  torch.cuda.current_stream(data.device).cuda_stream), "crc_batch_dev")
        return out

    # ---- TB CRC attachment + nr_segmentation on the device
    def tb_segment_parms(self, BG, A):
        """Scalar part of nr_segmentation for a transport block of A payload bits (host arithmetic).  Returns dict(C, K, Z, F, Kprime, L, Kb)."""
        q = (C.c_uint32 * 6)()
        kb = self.lib.nrb200_tb_segment_parms(BG, A, q)
        if kb < 0:
            raise ValueError("transport block too large")
        return {"C": q[0], "K": q[1], "Z": q[2], "F": q[3], "Kprime": q[4], "L": q[5], "Kb": kb}

    def tb_segment_host(self, BG, payload):
        """payload: uint8[A / 8].  Returns (C, K / 8) uint8 segments: TB CRC attached, per-segment CRC24B, zero filler."""
        p = np.ascontiguousarray(payload, dtype=np.uint8)
        q = self.tb_segment_parms(BG, p.size * 8)
        out = np.zeros((q["C"], q["K"] // 8), dtype=np.uint8)
        self.lib.nrb200_tb_segment_host.argtypes = [C.c_int, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32]
        self._check(self.lib.nrb200_tb_segment_host(BG, p.size * 8, p.ctypes.data, out.ctypes.data, out.shape[1]), "tb_segment_host")
        return out

    def tb_segment_torch(self, BG, A, payload, segs, scratch):
        """Device-resident: payload uint8[A / 8], segs (C, >= K / 8) uint8 out, scratch int32[1]."""
        import torch
        self.lib.nrb200_tb_segment_dev.argtypes = [C.c_int, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
        self._check(self.lib.nrb200_tb_segment_dev(BG, A, payload.data_ptr(), segs.data_ptr(), segs.shape[1], scratch.data_ptr(),
                                                   torch.cuda.current_stream(payload.device).cuda_stream), "tb_segment_dev")
        return segs

    # ---- rate matching / interleaving around the codec (nr_rate_matching.c), one transport block per call
    def _rmdesc(self, BG, Z, Qm, rv, C_, Tbslbrm, F, n_seg, clear=0):
        d = RmDesc()
        d.BG, d.Z, d.Qm, d.rv, d.clear, d.C, d.Tbslbrm, d.F, d.K, d.n_seg = BG, Z, Qm, rv, clear, C_, Tbslbrm, F, (22 if BG == 1 else 10) * Z, n_seg
        return d

    def rm_tx_host(self, BG, Z, Qm, rv, C_, Tbslbrm, F, d, E):
        """nr_rate_matching_ldpc + nr_interleaving_ldpc for the n_seg segments of a TB.  d: (n_seg, 66Z|50Z) 0/1 bytes, E: per-segment
        lengths.  Returns the concatenated f (sum(E) bytes)."""
        d = np.ascontiguousarray(d, dtype=np.uint8)
        E = np.ascontiguousarray(E, dtype=np.uint32)
        f = np.zeros(int(E.sum()), dtype=np.uint8)
        desc = self._rmdesc(BG, Z, Qm, rv, C_, Tbslbrm, F, d.shape[0])
        self._check(self.lib.nrb200_ldpc_rm_tx_batch_host(C.byref(desc), d.ctypes.data, d.shape[1], E.ctypes.data, f.ctypes.data), "rm_tx_batch_host")
        return f

    def rm_rx_host(self, BG, Z, Qm, rv, C_, Tbslbrm, F, soft, E, harq, clear):
        """nr_deinterleaving_ldpc + nr_rate_matching_ldpc_rx + decoder-input packing.  soft: concatenated int16 LLRs, harq: (n_seg, >=N)
        int16 updated in place.  Returns (n_seg, 68Z|52Z) int8 decoder inputs."""
        soft = np.ascontiguousarray(soft, dtype=np.int16)
        E = np.ascontiguousarray(E, dtype=np.uint32)
        assert harq.dtype == np.int16 and harq.flags.c_contiguous
        n = harq.shape[0]
        kcz = (68 if BG == 1 else 52) * Z
        llr = np.zeros((n, kcz), dtype=np.int8)
        desc = self._rmdesc(BG, Z, Qm, rv, C_, Tbslbrm, F, n, clear)
        self._check(self.lib.nrb200_ldpc_rm_rx_batch_host(C.byref(desc), soft.ctypes.data, E.ctypes.data, harq.ctypes.data, harq.shape[1], llr.ctypes.data, kcz), "rm_rx_batch_host")
        return llr

    def rm_tx_torch(self, BG, Z, Qm, rv, C_, Tbslbrm, F, d, E, foff, f):
        """Device-resident rm_tx: d (n_seg, 66Z|50Z) uint8, E / foff uint32 device vectors (lengths, offsets into f), f uint8 out."""
        import torch
        desc = self._rmdesc(BG, Z, Qm, rv, C_, Tbslbrm, F, d.shape[0])
        self._check(self.lib.nrb200_ldpc_rm_tx_batch_dev(C.byref(desc), d.data_ptr(), d.shape[1], E.data_ptr(), foff.data_ptr(), f.data_ptr(),
                                                         torch.cuda.current_stream(d.device).cuda_stream), "rm_tx_batch_dev")
        return f

    def rm_rx_torch(self, BG, Z, Qm, rv, C_, Tbslbrm, F, soft, E, soff, harq, llr, clear=1):
        """Device-resident rm_rx: soft int16 (sum E), harq (n_seg, >= 66Z|50Z) int16 in place, llr (n_seg, >= 68Z|52Z) int8 out."""
        import torch
        desc = self._rmdesc(BG, Z, Qm, rv, C_, Tbslbrm, F, harq.shape[0], clear)
        self._check(self.lib.nrb200_ldpc_rm_rx_batch_dev(C.byref(desc), soft.data_ptr(), E.data_ptr(), soff.data_ptr(), harq.data_ptr(), harq.shape[1],
                                                         llr.data_ptr(), llr.shape[1], torch.cuda.current_stream(soft.device).cuda_stream), "rm_rx_batch_dev")
        return llr

    # ---- scrambling (nr_scrambling.c) and QAM mapper (nr_modulation.c)
    def scramble_host(self, in_bits, q, Nid, n_RNTI):
        x = np.ascontiguousarray(in_bits, dtype=np.uint8)
        out = np.zeros((x.size + 31) // 32, dtype=np.uint32)
        self._check(self.lib.nrb200_scramble_host(x.ctypes.data, x.size, q, Nid, n_RNTI, out.ctypes.data), "scramble_host")
        return out

    def unscramble_llr_host(self, llr, q, Nid, n_RNTI):
        y = np.ascontiguousarray(llr, dtype=np.int16).copy()
        self._check(self.lib.nrb200_unscramble_llr_host(y.ctypes.data, y.size, q, Nid, n_RNTI), "unscramble_llr_host")
        return y

    def modulate_host(self, packed_words, length_bits, Qm):
        x = np.ascontiguousarray(packed_words, dtype=np.uint32)
        out = np.zeros(2 * (length_bits // Qm), dtype=np.int16)
        self._check(self.lib.nrb200_modulate_host(x.ctypes.data, length_bits, Qm, out.ctypes.data), "modulate_host")
        return out

    def scramble_torch(self, in_bits, q, Nid, n_RNTI, out):
        import torch
        self._check(self.lib.nrb200_scramble_dev(in_bits.data_ptr(), in_bits.numel(), q, Nid, n_RNTI, out.data_ptr(),
                                                 torch.cuda.current_stream(in_bits.device).cuda_stream), "scramble_dev")
        return out

    def unscramble_llr_torch(self, llr, q, Nid, n_RNTI):
        import torch
        self._check(self.lib.nrb200_unscramble_llr_dev(llr.data_ptr(), llr.numel(), q, Nid, n_RNTI, torch.cuda.current_stream(llr.device).cuda_stream),
                    "unscramble_llr_dev")
        return llr

    def modulate_torch(self, packed_words, length_bits, Qm, out):
        import torch
        self._check(self.lib.nrb200_modulate_dev(packed_words.data_ptr(), length_bits, Qm, out.data_ptr(),
                                                 torch.cuda.current_stream(out.device).cuda_stream), "modulate_dev")
        return out

    # ---- PUSCH channel estimation, DMRS type 1 (nr_ul_channel_estimation.c)
    def pusch_dmrs_pilots(self, desc):
        pil = np.zeros(2 * 6 * desc.rb_size, dtype=np.int16)
        self._check(self.lib.nrb200_pusch_dmrs_pilots_host(C.addressof(desc), pil.ctypes.data), "pusch_dmrs_pilots_host")
        return pil

    def lowpapr_sequence(self, u, v, n_re, scaling=32767):
        """Low-PAPR base sequence r_{u,v} as the reference generates it (n_re = 30 or >= 36); None for the table-driven lengths (see the header)."""
        seq = np.zeros(2 * n_re, dtype=np.int16)
        rc = self.lib.nrb200_lowpapr_sequence_host(u, v, n_re, scaling, C.c_void_p(seq.ctypes.data))
        return seq if rc == 0 else None

    def pusch_tp_scratch_bytes(self, desc):
        self.lib.nrb200_pusch_tp_scratch_bytes.restype = C.c_uint64
        return int(self.lib.nrb200_pusch_tp_scratch_bytes(C.c_void_p(C.addressof(desc))))

    def pusch_chest_host(self, desc, rxdataF, ul_ch_estimates=None):
        """rxdataF [nb_rx][14][N][2] int16 -> (ul_ch_estimates with symbol desc.symbol rewritten, state int32[5] = max_ch, nvar, est_delay, pos, val)."""
        x = np.ascontiguousarray(rxdataF, dtype=np.int16)
        np_ = max(1, int(desc.n_ports))
        est = np.zeros((np_ * x.shape[0],) + x.shape[1:], np.int16) if ul_ch_estimates is None else np.ascontiguousarray(ul_ch_estimates, dtype=np.int16)
        st = np.zeros(5 * np_, dtype=np.int32)
        self._check(self.lib.nrb200_pusch_chest_host(C.addressof(desc), x.ctypes.data, est.ctypes.data, st.ctypes.data), "pusch_chest_host")
        return est, st

    def pusch_chest_torch(self, desc, rxdataF, ul_ch_estimates, scratch, state):
        import torch
        self._check(self.lib.nrb200_pusch_chest_dev(C.addressof(desc), rxdataF.data_ptr(), ul_ch_estimates.data_ptr(), scratch.data_ptr(), state.data_ptr(),
                                                    torch.cuda.current_stream(rxdataF.device).cuda_stream), "pusch_chest_dev")
        return ul_ch_estimates

    def chest_time_avg_host(self, est, num_symbols, start_symbol, dmrs_bitmap, num_rbs):
        """nr_chest_time_domain_avg: est [nb_rx][14][N][2] int16 -> (averaged copy, first DMRS symbol)."""
        e = np.ascontiguousarray(est, dtype=np.int16).copy()
        rc = self.lib.nrb200_chest_time_avg_host(e.shape[2], e.shape[0], start_symbol, num_symbols, dmrs_bitmap, num_rbs, C.c_void_p(e.ctypes.data))
        if rc < 0:
            self._check(rc, "chest_time_avg_host")
        return e, rc

    def chest_time_avg_torch(self, est, num_symbols, start_symbol, dmrs_bitmap, num_rbs):
        import torch
        rc = self.lib.nrb200_chest_time_avg_dev(est.shape[2], est.shape[0], 14 * est.shape[2], start_symbol, num_symbols, dmrs_bitmap, num_rbs,
                                                C.c_void_p(est.data_ptr()), C.c_void_p(torch.cuda.current_stream(est.device).cuda_stream))
        if rc < 0:
            self._check(rc, "chest_time_avg_dev")
        return rc

    def pusch_chest_scratch_bytes(self, desc):
        return int(self.lib.nrb200_pusch_chest_scratch_bytes(C.addressof(desc)))

    # ---- single-layer PUSCH inner receiver (nr_ulsch_demodulation.c inner_rx + log2_maxh measurement)
    def pusch_num_llr(self, desc):
        return int(self.lib.nrb200_pusch_num_llr(C.addressof(desc)))

    def prach_num_roots(self, desc):
        self.lib.nrb200_prach_num_roots.restype = C.c_uint32
        return int(self.lib.nrb200_prach_num_roots(C.c_void_p(C.addressof(desc))))

    def rx_nr_prach_host(self, desc, xu, rxsigF):
        """rx_nr_prach: xu [>= roots][839][2] int16 (gNB->X_u), rxsigF [nb_rx][N_ZC][2] int16.  Returns (max_preamble, max_preamble_energy, max_preamble_delay)."""
        x = np.ascontiguousarray(xu, dtype=np.int16); r = np.ascontiguousarray(rxsigF, dtype=np.int16)
        o = (C.c_uint16 * 3)()
        f = self.lib.nrb200_rx_nr_prach_host
        f.argtypes = [C.c_void_p] * 6
        self._check(f(C.addressof(desc), x.ctypes.data, r.ctypes.data, C.addressof(o), C.addressof(o) + 2, C.addressof(o) + 4), "rx_nr_prach_host")
        return int(o[0]), int(o[1]), int(o[2])

    def rx_nr_prach_torch(self, desc, xu, rxsigF, out3, scratch):
        import torch
        f = self.lib.nrb200_rx_nr_prach_dev
        f.argtypes = [C.c_void_p] * 6
        self._check(f(C.addressof(desc), xu.data_ptr(), rxsigF.data_ptr(), out3.data_ptr(), scratch.data_ptr(), torch.cuda.current_stream(xu.device).cuda_stream),
                    "rx_nr_prach_dev")
        return out3

    def prach_scratch_bytes(self, desc):
        self.lib.nrb200_prach_scratch_bytes.restype = C.c_uint64
        return int(self.lib.nrb200_prach_scratch_bytes(C.c_void_p(C.addressof(desc))))

    def rfsim_rx_add_input_host(self, nb_tx, nb_rx, offset, pl_dB, noise_dB, ch, sig, out, TS, noise=None):
        """rxAddInput for every receive antenna: ch [nb_tx * nb_rx][L][2] float64 (plane rx + tx * nb_rx), sig [CirSize][2] int16 (tx antennas interleaved),
        out [nb_rx][n][2] int16 (accumulated into; a copy is returned), noise [nb_rx][n][2] float64 or None."""
        c = np.ascontiguousarray(ch, dtype=np.float64); s = np.ascontiguousarray(sig, dtype=np.int16); o = np.ascontiguousarray(out, dtype=np.int16).copy()
        d = RfsimChan(nb_tx, nb_rx, c.shape[1], offset, pl_dB, noise_dB, 0)
        nz = None if noise is None else np.ascontiguousarray(noise, dtype=np.float64)
        f = self.lib.nrb200_rfsim_rx_add_input_host
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32, C.c_void_p]
        self._check(f(C.addressof(d), c.ctypes.data, s.ctypes.data, o.ctypes.data, o.shape[1], o.shape[1], TS, s.shape[0], None if nz is None else nz.ctypes.data),
                    "rfsim_rx_add_input_host")
        return o

    def rfsim_rx_add_input_torch(self, d, ch, sig, out, TS, noise=None):
        """Device-resident variant: torch CUDA tensors with the shapes above; out is updated in place."""
        import torch
        f = self.lib.nrb200_rfsim_rx_add_input_dev
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p]
        self._check(f(C.addressof(d), ch.data_ptr(), sig.data_ptr(), out.data_ptr(), out.shape[1], out.shape[1], TS, sig.shape[0],
                      None if noise is None else noise.data_ptr(), torch.cuda.current_stream(out.device).cuda_stream), "rfsim_rx_add_input_dev")
        return out

    def pdsch_ptrs_layout(self, desc):
        """(PT-RS symbol mask, PT-RS REs per PT-RS symbol) of a descriptor with ptrs = 1 (set_ptrs_symb_idx / nr_ptrs_cpe_estimation's bookkeeping)."""
        mask, n = C.c_uint32(0), C.c_uint32(0)
        self._check(self.lib.nrb200_pdsch_ptrs_layout(C.c_void_p(C.addressof(desc)), C.byref(mask), C.byref(n)), "pdsch_ptrs_layout")
        return mask.value, n.value

    def pusch_inner_rx_host(self, desc, rxdataF, ul_ch_estimates):
        """rxdataF: [nb_rx][14][N][2] int16; ul_ch_estimates: [nb_rx * layers][14][N][2] (index layer * nb_rx + rx).
        desc.log2_maxh == 0xFFFFFFFF: measure it like nr_rx_pusch_tp does.
        Returns (llr int16[G], log2_maxh)."""
        x = np.ascontiguousarray(rxdataF, dtype=np.int16)
        h = np.ascontiguousarray(ul_ch_estimates, dtype=np.int16)
        n = self.pusch_num_llr(desc)
        if n == 0:
            raise ValueError("invalid PUSCH descriptor")
        llr = np.zeros(n, dtype=np.int16)
        sh = C.c_int32(-1)
        self._check(self.lib.nrb200_pusch_inner_rx_host(C.addressof(desc), x.ctypes.data, h.ctypes.data, llr.ctypes.data, C.addressof(sh)), "pusch_inner_rx_host")
        return llr, sh.value

    def pusch_inner_rx_torch(self, desc, rxdataF, ul_ch_estimates, llr, level=None):
        """Device-resident variant; `level` = int32[9] scratch: when given, log2_maxh is measured on the device first (stream ordered)."""
        import torch
        st = torch.cuda.current_stream(rxdataF.device).cuda_stream
        if level is not None:
            self._check(self.lib.nrb200_pusch_log2_maxh_dev(C.addressof(desc), ul_ch_estimates.data_ptr(), level.data_ptr(), st), "pusch_log2_maxh_dev")
        self._check(self.lib.nrb200_pusch_inner_rx_dev(C.addressof(desc), rxdataF.data_ptr(), ul_ch_estimates.data_ptr(),
                                                       0 if level is None else level.data_ptr() + 32, llr.data_ptr(), st), "pusch_inner_rx_dev")
        return llr

    # ---- gNB PDSCH transmitter after the encoder (nr_generate_pdsch from scrambling to txdataF)
    def pdsch_tx_num_bits(self, desc):
        self.lib.nrb200_pdsch_tx_num_bits.restype = C.c_uint32
        return int(self.lib.nrb200_pdsch_tx_num_bits(C.c_void_p(C.addressof(desc))))

    def pdsch_tx_slot_host(self, desc, f, txdataF=None):
        """f: uint8 bits (one per element); txdataF [nb_tx][14][N][2] int16 in/out (zeros when None).  Returns txdataF."""
        G = self.pdsch_tx_num_bits(desc)
        if G == 0:
            raise ValueError("invalid PDSCH descriptor")
        b = np.ascontiguousarray(f, dtype=np.uint8)
        assert b.size == G, (b.size, G)
        out = np.zeros((desc.nb_tx, 14, desc.fft_size, 2), np.int16) if txdataF is None else np.ascontiguousarray(txdataF, dtype=np.int16).copy()
        self.lib.nrb200_pdsch_tx_slot_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        self._check(self.lib.nrb200_pdsch_tx_slot_host(C.addressof(desc), b.ctypes.data, out.ctypes.data), "pdsch_tx_slot_host")
        return out

    def pdsch_tx_slot_torch(self, desc, f, txdataF):
        """Device-resident variant: f uint8[G], txdataF int16 [nb_tx][14][N][2] (desc.tx_stride must be set)."""
        import torch
        st = torch.cuda.current_stream(f.device).cuda_stream
        self.lib.nrb200_pdsch_tx_slot_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        self._check(self.lib.nrb200_pdsch_tx_slot_dev(C.addressof(desc), f.data_ptr(), txdataF.data_ptr(), st), "pdsch_tx_slot_dev")
        return txdataF

    # ---- demodulation: nr_ulsch_compute_llr (single layer, max-log)
    def pusch_llr_host(self, Qm, rxF, mag_a=None, mag_b=None, mag_c=None):
        rxF = np.ascontiguousarray(rxF, dtype=np.int16)
        n = rxF.size // 2
        mk = lambda a: np.ascontiguousarray(a, dtype=np.int16) if a is not None else None
        a, b, c = mk(mag_a), mk(mag_b), mk(mag_c)
        out = np.zeros(n * Qm, dtype=np.int16)
        p = lambda v: v.ctypes.data if v is not None else None
        self._check(self.lib.nrb200_pusch_llr_host(Qm, n, rxF.ctypes.data, p(a), p(b), p(c), out.ctypes.data), "pusch_llr_host")
        return out

    def pusch_llr_torch(self, Qm, rxF, mag_a=None, mag_b=None, mag_c=None, out=None):
        import torch
        n = rxF.numel() // 2
        if out is None:
            out = torch.empty(n * Qm, dtype=torch.int16, device=rxF.device)
        p = lambda v: v.data_ptr() if v is not None else None
        st = torch.cuda.current_stream(rxF.device).cuda_stream
        self._check(self.lib.nrb200_pusch_llr_dev(Qm, n, rxF.data_ptr(), p(mag_a), p(mag_b), p(mag_c), out.data_ptr(), st), "pusch_llr_dev")
        return out

    # ---- batched extension, device-resident torch tensors (asynchronous on torch's current stream)
    def decode_batch_torch(self, BG, Z, R, numMaxIter, llr, outMode=OUTMODE_BIT, use_crc=0, crc_len_bits=0, crc_type=0, out=None, iters=None, latency_mode=0):
        import torch
        assert llr.is_cuda and llr.dtype == torch.int8 and llr.is_contiguous() and llr.dim() == 2
        n_cb, stride = llr.shape
        n = ncols_for_rate(BG, R) * Z
        ob = (n + 7) // 8 if outMode == OUTMODE_BIT else n
        if out is None:
            out = torch.empty((n_cb, ob), dtype=torch.uint8, device=llr.device)
        if iters is None:
            iters = torch.empty(n_cb, dtype=torch.int32, device=llr.device)
        d = self._desc(BG, Z, R, numMaxIter, outMode, n_cb, stride, out.shape[1], use_crc, crc_len_bits, crc_type, latency_mode)
        st = torch.cuda.current_stream(llr.device).cuda_stream
        self._check(self.lib.nrb200_ldpc_decode_batch_dev(C.byref(d), llr.data_ptr(), out.data_ptr(), iters.data_ptr(), st), "decode_batch_dev")
        return iters, out

    def encode_batch_torch(self, BG, Z, K, payloads, out=None):
        import torch
        assert payloads.is_cuda and payloads.dtype == torch.uint8 and payloads.is_contiguous()
        n_cb, stride = payloads.shape
        nout = (66 if BG == 1 else 50) * Z
        if out is None:
            out = torch.empty((n_cb, nout), dtype=torch.uint8, device=payloads.device)
        st = torch.cuda.current_stream(payloads.device).cuda_stream
        self._check(self.lib.nrb200_ldpc_encode_batch_dev(BG, Z, K, n_cb, payloads.data_ptr(), stride, out.data_ptr(), out.shape[1], st), "encode_batch_dev")
        return out


class OffloadLdpcLib:
    """libldpc_b200_t2.so: the same four symbols with OAI's "offload" semantics (ldpc_interface_offload, nr_ulsch_decoding.c:230-300,
    nr_dlsch_coding.c:362-384): one segment per call, the library de-interleaves, rate-recovers and HARQ-combines."""

    def __init__(self, path=os.path.join(_HERE, "libldpc_b200_t2.so"), init=True):
        if not os.path.exists(path):
            raise RuntimeError(f"{path} not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = self.lib = C.CDLL(path)
        L.LDPCdecoder.argtypes = [C.POINTER(DecParams), C.c_uint8, C.c_uint8, C.c_uint8, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.LDPCencoder.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(EncParams)]
        if init and L.LDPCinit() != 0:
            raise Nrb200Error("LDPCinit (offload) failed: no usable CUDA device")

    def LDPCdecoder(self, BG, Z, R, numMaxIter, E, Qm, rv, F, llr_E, ulsch_id=0, r=0, harq_pid=0, setCombIn=0):
        p = DecParams()
        p.BG, p.Z, p.R, p.F, p.Qm, p.rv, p.numMaxIter, p.E, p.outMode, p.setCombIn = BG, Z, R, F, Qm, rv, numMaxIter, E, OUTMODE_BIT, setCombIn
        x = np.ascontiguousarray(llr_E, dtype=np.int8)
        assert x.size >= E
        K = (22 if BG == 1 else 10) * Z
        out = np.zeros(K // 8, dtype=np.uint8)
        it = self.lib.LDPCdecoder(C.byref(p), harq_pid, ulsch_id, r, x.ctypes.data, out.ctypes.data, None, None)
        if it < 0:
            raise Nrb200Error(f"offload LDPCdecoder rc={it}")
        return it, out

    def LDPCencoder(self, BG, Z, K, F, Qm, rv, E, segment):
        ip = EncParams()
        ip.n_segments, ip.BG, ip.Zc, ip.K, ip.F, ip.Qm, ip.rv, ip.E = 1, BG, Z, K, F, Qm, rv, E
        seg = np.ascontiguousarray(segment, dtype=np.uint8)
        out = np.zeros(E, dtype=np.uint8)
        inp = (C.c_void_p * 1)(seg.ctypes.data)
        oup = (C.c_void_p * 1)(out.ctypes.data)
        rc = self.lib.LDPCencoder(inp, oup, C.byref(ip))
        if rc != 0:
            raise Nrb200Error(f"offload LDPCencoder rc={rc}")
        return out


_lib = None


def load_LDPClib():
    """Process-wide instance, initialised (mirrors load_LDPClib + LDPCinit, nrLDPC_load.c:46-71)."""
    global _lib
    if _lib is None:
        _lib = LdpcLib()
        _lib.init()
    return _lib
