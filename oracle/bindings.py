"""TEST INFRASTRUCTURE ONLY -- ctypes bindings for the CPU oracle (oracle/liboracle.so, the C restatement) and,
when present, the compiled unmodified reference (oracle/_ref/libref_*.so, built by oracle/build_ref.sh).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg import this module.
The product package (openairinterface5g_b200) never does.
"""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REFDIR = os.path.join(HERE, "_ref")

_u8p = C.POINTER(C.c_uint8)
_i8p = C.POINTER(C.c_int8)
_i16p = C.POINTER(C.c_int16)


class TimeStats(C.Structure):  # common/utils/time_meas.h:61-75
    _fields_ = [("in_", C.c_longlong), ("diff", C.c_longlong), ("p_time", C.c_longlong), ("diff_square", C.c_double),
                ("max", C.c_longlong), ("trials", C.c_int), ("meas_flag", C.c_int), ("meas_name", C.c_char_p),
                ("meas_index", C.c_int), ("meas_enabled", C.c_int), ("tpoolmsg", C.c_void_p), ("tstatptr", C.c_void_p)]


class LdpcTimeStats(C.Structure):  # nrLDPC_types.h:115-127
    _fields_ = [(n, TimeStats) for n in ("llr2llrProcBuf", "llr2CnProcBuf", "cnProc", "cnProcPc", "bnProcPc", "bnProc",
                                         "cn2bnProcBuf", "bn2cnProcBuf", "llrRes2llrOut", "llr2bit", "total")]


CHECK_CRC_T = C.CFUNCTYPE(C.c_int, _u8p, C.c_uint32, C.c_uint8)


class DecParams(C.Structure):  # nrLDPC_types.h:84-97
    _fields_ = [("BG", C.c_uint8), ("Z", C.c_uint16), ("R", C.c_uint8), ("F", C.c_uint16), ("Qm", C.c_uint8), ("rv", C.c_uint8),
                ("numMaxIter", C.c_uint8), ("E", C.c_int), ("outMode", C.c_int), ("crc_type", C.c_int),
                ("check_crc", C.c_void_p), ("setCombIn", C.c_uint8)]


class DecodeAbort(C.Structure):  # defs_common.h:996-1000 (pthread_mutex_t is 40 bytes on x86-64 glibc)
    _fields_ = [("mutex", C.c_uint8 * 40), ("failed", C.c_bool)]


class EncParams(C.Structure):  # nrLDPC_defs.h:40-66
    _fields_ = [("n_segments", C.c_uint), ("macro_num", C.c_uint), ("gen_code", C.c_ubyte),
                ("tinput", C.c_void_p), ("tprep", C.c_void_p), ("tparity", C.c_void_p), ("toutput", C.c_void_p),
                ("Kr", C.c_int), ("Kb", C.c_uint32), ("Zc", C.c_uint32), ("harq", C.c_void_p), ("BG", C.c_uint8),
                ("output", C.c_void_p), ("K", C.c_uint32), ("F", C.c_uint32), ("Qm", C.c_uint8), ("E", C.c_uint32),
                ("G", C.c_uint), ("rv", C.c_uint8)]


def ncols_for_rate(BG, R):
    return {(1, 13): 68, (1, 23): 35, (1, 89): 27, (2, 15): 52, (2, 13): 32, (2, 23): 17}[(BG, R)]


def _ptr(a, t):
    return a.ctypes.data_as(t)


# ------------------------------------------------------------------------------------------------ oracle (C restatement)
class ChestParms(C.Structure):       # orc_chest_t
    _fields_ = [(n, C.c_int32) for n in ("fft_size", "nb_rx", "slot", "symbol", "port", "rb_start", "bwp_start", "rb_size", "first_carrier_offset", "scid",
                                         "dmrs_scrambling_id", "dmrs_type", "chest_freq")]


class PuschParms(C.Structure):       # orc_pusch_t
    _fields_ = [(n, C.c_int32) for n in ("fft_size", "nb_rx", "rb_start", "bwp_start", "rb_size", "first_carrier_offset", "Qm", "ul_dmrs_symb_pos",
                                         "dmrs_config_type", "num_dmrs_cdm_grps_no_data")]


class PtrsParms(C.Structure):        # orc_ptrs_t
    _fields_ = [(n, C.c_int32) for n in ("on", "L", "K", "re_offset", "rnti", "slot", "nscid", "nid")]


class PdschTxParms(C.Structure):     # orc_pdsch_tx_t
    _fields_ = [(n, C.c_int32) for n in ("fft_size", "nb_tx", "slot", "rb_start", "bwp_start", "rb_size", "first_carrier_offset", "Qm", "nrOfLayers", "start_symbol",
                                         "nr_of_symbols", "dl_dmrs_symb_pos", "dmrs_config_type", "num_dmrs_cdm_grps_no_data", "dmrs_ports", "scid",
                                         "dl_dmrs_scrambling_id", "data_scrambling_id", "rnti", "amp", "pm_idx")] + [("pm_weights", C.c_int16 * 32)] + \
               [(n, C.c_int32) for n in ("ptrs_on", "ptrs_L", "ptrs_K", "ptrs_re_offset")]

    def set_ptrs(self, L_log2, K, re_offset):
        self.ptrs_on, self.ptrs_L, self.ptrs_K, self.ptrs_re_offset = 1, L_log2, K, re_offset
        return self

    def ptrs_res(self):
        """PT-RS REs of the slot per layer (harq->unav_res)."""
        if not self.ptrs_on:
            return 0
        i, l_ref, L, last, mask = 0, self.start_symbol, 1 << self.ptrs_L, self.start_symbol + self.nr_of_symbols - 1, 0
        while l_ref + i * L <= last:                       # set_ptrs_symb_idx
            hit = [l for l in range(l_ref + i * L, max(l_ref + (i - 1) * L + 1, l_ref) - 1, -1) if (self.dl_dmrs_symb_pos >> l) & 1]
            if hit:
                l_ref, i = hit[0], 1
                continue
            mask |= 1 << (l_ref + i * L)
            i += 1
        return bin(mask).count("1") * ((self.rb_size + self.ptrs_K - 1) // self.ptrs_K)

    def set_precoding(self, pm_idx, weights):
        """weights [4 layers][4 antennas][2] int16 (nfapi_nr_pm_pdu_t.weights); pm_idx 0 = identity."""
        self.pm_idx = pm_idx
        w = np.zeros((4, 4, 2), np.int16)
        if weights is not None:
            a = np.asarray(weights, dtype=np.int16)
            w[:a.shape[0], :a.shape[1]] = a
        self.pm_weights = (C.c_int16 * 32)(*[int(x) for x in w.reshape(-1)])
        return self

    def G(self):
        n_dmrs_sym = bin(self.dl_dmrs_symb_pos & (((1 << self.nr_of_symbols) - 1) << self.start_symbol)).count("1")
        per = self.num_dmrs_cdm_grps_no_data * (6 if self.dmrs_config_type == 0 else 4)
        return ((12 * self.nr_of_symbols - per * bin(self.dl_dmrs_symb_pos).count("1")) * self.rb_size - self.ptrs_res()) * self.nrOfLayers * self.Qm



def _rfsim_call(fn, nb_tx, nb_rx, L, offset, pl_dB, noise_dB, ch, sig, out, rx_ant, TS, cir, noise):
    fn.restype = None
    fn.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_uint64, C.c_uint32, C.c_void_p]
    c = np.ascontiguousarray(ch, dtype=np.float64); s = np.ascontiguousarray(sig, dtype=np.int16)
    assert c.shape == (nb_tx * nb_rx, L, 2) and s.shape == (cir, 2)
    o = np.ascontiguousarray(out, dtype=np.int16).copy()
    n = None if noise is None else np.ascontiguousarray(noise, dtype=np.float64)
    fn(nb_tx, nb_rx, L, offset, pl_dB, noise_dB, c.ctypes.data, s.ctypes.data, o.ctypes.data, rx_ant, o.shape[0], TS, cir, None if n is None else n.ctypes.data)
    return o


class Oracle:
    def __init__(self):
        so = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(so):
            subprocess.check_call(["make", "-C", HERE, "liboracle.so"])
        L = self.lib = C.CDLL(so)
        L.orc_crc.restype = C.c_uint32
        L.orc_crc.argtypes = [C.c_int, _u8p, C.c_uint32]
        L.orc_check_crc.argtypes = [_u8p, C.c_uint32, C.c_int]
        L.orc_ldpc_decode.argtypes = [C.c_int] * 5 + [_i8p, _i8p, C.c_int, C.c_uint32, C.c_int, C.c_int]
        L.orc_ldpc_encode.argtypes = [C.c_int, C.c_int, C.c_int, _u8p, _u8p]
        L.orc_rate_matching_tx.argtypes = [C.c_uint32, C.c_int, C.c_int, _u8p, _u8p, C.c_int, C.c_uint32, C.c_uint32, C.c_int, C.c_uint32]
        L.orc_rate_matching_rx.argtypes = [C.c_uint32, C.c_int, C.c_int, _i16p, _i16p, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32]
        L.orc_interleave.argtypes = [C.c_uint32, C.c_int, _u8p, _u8p]
        L.orc_interleave.restype = None
        L.orc_deinterleave.argtypes = [C.c_uint32, C.c_int, _i16p, _i16p]
        L.orc_deinterleave.restype = None
        L.orc_get_R_ldpc_decoder.argtypes = [C.c_int] * 4 + [C.POINTER(C.c_int), C.c_int]
        L.orc_segmentation.argtypes = [_u8p, C.POINTER(_u8p), C.c_uint, C.POINTER(C.c_uint), C.POINTER(C.c_uint),
                                       C.POINTER(C.c_uint), C.POINTER(C.c_uint), C.c_int]

    def ils_of_z(self, Z):
        return self.lib.orc_ils_of_z(int(Z))

    def crc(self, poly_id, data, bitlen):
        d = np.ascontiguousarray(data, dtype=np.uint8)
        return int(self.lib.orc_crc(poly_id, _ptr(d, _u8p), bitlen))

    def check_crc(self, data, n, crc_type):
        d = np.ascontiguousarray(data, dtype=np.uint8)
        return int(self.lib.orc_check_crc(_ptr(d, _u8p), n, crc_type))

    def decode(self, BG, Z, R, max_iter, llr, out_mode=0, use_crc=0, crc_len_bits=0, crc_type=0, abort_in=0):
        n = ncols_for_rate(BG, R) * Z
        llr = np.ascontiguousarray(llr, dtype=np.int8)
        assert llr.size >= n
        out = np.zeros(n + 64, dtype=np.int8)
        it = self.lib.orc_ldpc_decode(BG, Z, R, max_iter, out_mode, _ptr(llr, _i8p), _ptr(out, _i8p), use_crc, crc_len_bits, crc_type, abort_in)
        return it, (out[:(n + 7) // 8].view(np.uint8) if out_mode == 0 else out[:n])

    def encode(self, BG, Z, K, payload):
        p = np.ascontiguousarray(payload, dtype=np.uint8)
        out = np.zeros((68 if BG == 1 else 52) * Z, dtype=np.uint8)
        rc = self.lib.orc_ldpc_encode(BG, Z, K, _ptr(p, _u8p), _ptr(out, _u8p))
        assert rc == 0, rc
        return out[:(66 if BG == 1 else 50) * Z]

    def rate_matching_tx(self, Tbslbrm, BG, Z, w, C_, F, Foffset, rv, E):
        w = np.ascontiguousarray(w, dtype=np.uint8)
        e = np.zeros(E, dtype=np.uint8)
        rc = self.lib.orc_rate_matching_tx(Tbslbrm, BG, Z, _ptr(w, _u8p), _ptr(e, _u8p), C_, F, Foffset, rv, E)
        return rc, e

    def rate_matching_rx(self, Tbslbrm, BG, Z, w, soft, C_, rv, clear, E, F, Foffset):
        soft = np.ascontiguousarray(soft, dtype=np.int16)
        assert w.dtype == np.int16 and w.flags.c_contiguous
        return self.lib.orc_rate_matching_rx(Tbslbrm, BG, Z, _ptr(w, _i16p), _ptr(soft, _i16p), C_, rv, clear, E, F, Foffset)

    def interleave(self, E, Qm, e):
        e = np.ascontiguousarray(e, dtype=np.uint8)
        f = np.zeros(E, dtype=np.uint8)
        self.lib.orc_interleave(E, Qm, _ptr(e, _u8p), _ptr(f, _u8p))
        return f

    def deinterleave(self, E, Qm, f):
        f = np.ascontiguousarray(f, dtype=np.int16)
        e = np.zeros(E, dtype=np.int16)
        self.lib.orc_deinterleave(E, Qm, _ptr(e, _i16p), _ptr(f, _i16p))
        return e

    def get_R(self, rv, E, BG, Z, llrLen, rnd):
        ll = C.c_int(llrLen)
        r = self.lib.orc_get_R_ldpc_decoder(rv, E, BG, Z, C.byref(ll), rnd)
        return r, ll.value

    def scramble(self, in_bits, q, Nid, rnti):
        x = np.ascontiguousarray(in_bits, dtype=np.uint8)
        out = np.zeros((x.size + 31) // 32, dtype=np.uint32)
        self.lib.orc_scramble.restype = None
        self.lib.orc_scramble(x.ctypes.data_as(C.c_void_p), C.c_uint32(x.size), C.c_uint32(q), C.c_uint32(Nid), C.c_uint32(rnti), out.ctypes.data_as(C.c_void_p))
        return out

    def unscramble_llr(self, llr, q, Nid, rnti):
        y = np.ascontiguousarray(llr, dtype=np.int16).copy()
        self.lib.orc_unscramble_llr.restype = None
        self.lib.orc_unscramble_llr(y.ctypes.data_as(C.c_void_p), C.c_uint32(y.size), C.c_uint32(q), C.c_uint32(Nid), C.c_uint32(rnti))
        return y

    def modulate(self, packed_bits, length, Qm):
        x = np.ascontiguousarray(packed_bits).view(np.uint8)
        out = np.zeros(2 * (length // Qm), dtype=np.int16)
        self.lib.orc_modulate.restype = None
        self.lib.orc_modulate(x.ctypes.data_as(C.c_void_p), C.c_uint32(length), Qm, out.ctypes.data_as(C.c_void_p))
        return out

    def ulsch_llr(self, Qm, rxF, maga, magb, magc):
        rxF = np.ascontiguousarray(rxF, dtype=np.int16)
        n = rxF.size // 2
        out = np.zeros(n * Qm, dtype=np.int16)
        mk = lambda a: np.ascontiguousarray(a, dtype=np.int16).ctypes.data_as(C.c_void_p) if a is not None else None
        self.lib.orc_ulsch_llr.restype = None
        self.lib.orc_ulsch_llr(Qm, rxF.ctypes.data_as(C.c_void_p), mk(maga), mk(magb), mk(magc), out.ctypes.data_as(C.c_void_p), C.c_uint32(n))
        return out

    # ---- PUSCH channel estimation (DMRS type 1)
    def pusch_dmrs_pilots(self, P):
        pil = np.zeros(2 * 6 * P.rb_size, np.int16)
        self.lib.orc_pusch_dmrs_pilots(C.byref(P), pil.ctypes.data_as(C.c_void_p))
        return pil

    def pusch_channel_estimation(self, P, rxdataF):
        x = np.ascontiguousarray(rxdataF, dtype=np.int16)
        est = np.zeros(P.nb_rx * 14 * P.fft_size * 2, np.int16)
        out = np.zeros(5, np.int32)
        self.lib.orc_pusch_channel_estimation(C.byref(P), x.ctypes.data_as(C.c_void_p), est.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
        return est.reshape(P.nb_rx, 14, P.fft_size, 2), out

    def chest_time_domain_avg(self, est, num_symbols, start_symbol, dmrs_bitmap, num_rbs):
        """est [nb_rx][14][N][2] int16 -> averaged copy (nr_chest_time_domain_avg)."""
        e = np.ascontiguousarray(est, dtype=np.int16).copy()
        rc = self.lib.orc_chest_time_domain_avg(e.shape[2], e.shape[0], num_symbols, start_symbol, dmrs_bitmap, num_rbs, e.ctypes.data_as(C.c_void_p))
        assert rc == 0
        return e

    def pdsch_channel_estimation(self, P, rxdataF):
        x = np.ascontiguousarray(rxdataF, dtype=np.int16)
        est = np.zeros(P.nb_rx * 14 * P.fft_size * 2, np.int16)
        self.lib.orc_pdsch_channel_estimation(C.byref(P), x.ctypes.data_as(C.c_void_p), est.ctypes.data_as(C.c_void_p))
        return est.reshape(P.nb_rx, 14, P.fft_size, 2)

    # ---- transform precoding (DFT-s-OFDM): low-PAPR DMRS for the estimator, frequency equalisation + nr_idft in the one-layer inner receiver
    def lowpapr_seq(self, u, v, M_ZC, scaling=32767):
        out = np.zeros(2 * M_ZC, np.int16)
        rc = self.lib.orc_lowpapr_seq(u, v, M_ZC, scaling, out.ctypes.data_as(C.c_void_p))
        return out if rc == 0 else None

    def chest_set_lowpapr(self, seq):
        """seq: the 6 * rb_size c16 low-PAPR sequence the estimator correlates with from now on; None switches back to the Gold-sequence DMRS."""
        self._lowpapr = None if seq is None else np.ascontiguousarray(seq, dtype=np.int16)
        self.lib.orc_chest_set_lowpapr(None if seq is None else self._lowpapr.ctypes.data_as(C.c_void_p))

    def pusch_set_transform_precoding(self, on):
        self.lib.orc_pusch_set_transform_precoding(int(on))

    def nr_idft(self, z, M):
        y = np.ascontiguousarray(z, dtype=np.int16).copy()
        rc = self.lib.orc_nr_idft(y.ctypes.data_as(C.c_void_p), M)
        return y if rc == 0 else None

    # ---- single-layer PUSCH inner receiver
    def pusch_nb_re(self, P, symbol):
        return int(self.lib.orc_pusch_nb_re(C.byref(P), symbol))

    def pusch_log2_maxh(self, P, meas_symbol, ch_symbol, rxdataF, ch_est):
        x = np.ascontiguousarray(rxdataF, dtype=np.int16); h = np.ascontiguousarray(ch_est, dtype=np.int16)
        avg = np.zeros(8, np.int32)
        r = self.lib.orc_pusch_log2_maxh(C.byref(P), meas_symbol, ch_symbol, x.ctypes.data_as(C.c_void_p), h.ctypes.data_as(C.c_void_p), avg.ctypes.data_as(C.c_void_p))
        return int(r), avg[:P.nb_rx].copy()

    def pusch_inner_rx_symbol(self, P, symbol, ch_symbol, shift, rxdataF, ch_est):
        x = np.ascontiguousarray(rxdataF, dtype=np.int16); h = np.ascontiguousarray(ch_est, dtype=np.int16)
        blen = (P.rb_size * 12 + 15) & ~15
        valid = self.pusch_nb_re(P, symbol)
        llr = np.zeros(valid * P.Qm, np.int16); comp = np.zeros(2 * blen, np.int16)
        self.lib.orc_pusch_inner_rx_symbol(C.byref(P), symbol, ch_symbol, shift, x.ctypes.data_as(C.c_void_p), h.ctypes.data_as(C.c_void_p),
                                           llr.ctypes.data_as(C.c_void_p), comp.ctypes.data_as(C.c_void_p))
        return llr, comp

    def pusch_log2_maxh_2l(self, P, meas_symbol, ch_symbol, max_ch, rxdataF, ch_est):
        x = np.ascontiguousarray(rxdataF, dtype=np.int16); h = np.ascontiguousarray(ch_est, dtype=np.int16)
        avg = np.zeros(8, np.int32)
        r = self.lib.orc_pusch_log2_maxh_2l(C.byref(P), meas_symbol, ch_symbol, max_ch, x.ctypes.data_as(C.c_void_p), h.ctypes.data_as(C.c_void_p), avg.ctypes.data_as(C.c_void_p))
        return int(r), avg[:2 * P.nb_rx].copy()

    def pusch_inner_rx_symbol_2l(self, P, symbol, ch_symbol, shift, nvar, rxdataF, ch_est):
        x = np.ascontiguousarray(rxdataF, dtype=np.int16); h = np.ascontiguousarray(ch_est, dtype=np.int16)
        blen = (P.rb_size * 12 + 15) & ~15
        valid = self.pusch_nb_re(P, symbol)
        llr = np.zeros((2, valid * P.Qm), np.int16); comp = np.zeros((2, 2 * blen), np.int16)
        self.lib.orc_pusch_inner_rx_symbol_2l.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        rc = self.lib.orc_pusch_inner_rx_symbol_2l(C.addressof(P), symbol, ch_symbol, shift, nvar, x.ctypes.data, h.ctypes.data, llr.ctypes.data, comp.ctypes.data)
        assert rc == valid, rc
        return llr, comp

    def pdsch_rx_slot(self, P, start_symbol, nr_symbols, rxdataF, dl_ch_est, nl=1):
        """UE-side PDSCH receiver for a whole slot (nl = 1: MRC; nl = 2: zero forcing, dl_ch_est [2 * nb_rx] planes).  Returns (llr int16[G], log2_maxh)."""
        x = np.ascontiguousarray(rxdataF, dtype=np.int16); h = np.ascontiguousarray(dl_ch_est, dtype=np.int16)
        llr = np.zeros(nl * 14 * 12 * P.rb_size * P.Qm + 64, np.int16)
        sh = C.c_int32(0)
        args = (start_symbol, nr_symbols, x.ctypes.data_as(C.c_void_p), h.ctypes.data_as(C.c_void_p), llr.ctypes.data_as(C.c_void_p), C.byref(sh))
        n = self.lib.orc_pdsch_rx_slot(C.byref(P), *args) if nl == 1 else self.lib.orc_pdsch_rx_slot_nl(C.byref(P), nl, *args)
        return llr[:n].copy(), sh.value

    def gold_words(self, c_init, n_words):
        """n_words 32-bit words of the Gold sequence lte_gold_generic produces for c_init."""
        out = np.zeros(n_words, np.uint32)
        self.lib.orc_gold_words(C.c_uint32(c_init), C.c_uint32(n_words), out.ctypes.data_as(C.c_void_p))
        return out

    def rfsim_rx_add_input(self, nb_tx, nb_rx, L, offset, pl_dB, noise_dB, ch, sig, out, rx_ant, TS, cir, noise=None):
        """rxAddInput: ch [nb_tx * nb_rx][L][2] float64 (plane rx + tx * nb_rx), sig [CirSize][2] int16 (tx antennas interleaved), out [n][2] int16 accumulated into."""
        return _rfsim_call(self.lib.orc_rfsim_rx_add_input, nb_tx, nb_rx, L, offset, pl_dB, noise_dB, ch, sig, out, rx_ant, TS, cir, noise)

    def db_fixed_times10(self, x):
        return int(self.lib.orc_db_fixed_times10(C.c_uint32(x)))

    def rx_nr_prach(self, nb_rx, short_sequence, NCS, prach_fmt, mu, xu, rxsigF):
        """PRACH detector: xu [64][839][2] int16 (gNB->X_u), rxsigF [nb_rx][N_ZC][2] int16.  Returns (max_preamble, max_preamble_energy, max_preamble_delay)."""
        x = np.ascontiguousarray(xu, dtype=np.int16); r = np.ascontiguousarray(rxsigF, dtype=np.int16)
        assert x.shape == (64, 839, 2) and r.shape == (nb_rx, 139 if short_sequence else 839, 2)
        out = np.zeros(3, np.int32)
        self.lib.orc_rx_nr_prach(nb_rx, short_sequence, NCS, prach_fmt, mu, x.ctypes.data_as(C.c_void_p), r.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
        return tuple(int(v) for v in out)

    def ptrs_symbols(self, start_symbol, nr_symbols, L_log2, dmrs_pos):
        self.lib.orc_ptrs_symbols.restype = C.c_uint32
        return int(self.lib.orc_ptrs_symbols(start_symbol, nr_symbols, 1 << L_log2, C.c_uint32(dmrs_pos)))

    def pdsch_rx_slot_ptrs(self, P, T, start_symbol, nr_symbols, rxdataF, dl_ch_est):
        """UE-side one-layer PDSCH receiver with PT-RS (T = PtrsParms).  Returns (llr, log2_maxh, phase int16[14][2], ptrs_re int32[14])."""
        x = np.ascontiguousarray(rxdataF, dtype=np.int16); h = np.ascontiguousarray(dl_ch_est, dtype=np.int16)
        llr = np.zeros(14 * 12 * P.rb_size * P.Qm + 64, np.int16)
        sh = C.c_int32(0); ph = np.zeros((14, 2), np.int16); nre = np.zeros(14, np.int32)
        n = self.lib.orc_pdsch_rx_slot_ptrs(C.byref(P), C.byref(T), start_symbol, nr_symbols, x.ctypes.data_as(C.c_void_p), h.ctypes.data_as(C.c_void_p),
                                            llr.ctypes.data_as(C.c_void_p), C.byref(sh), ph.ctypes.data_as(C.c_void_p), nre.ctypes.data_as(C.c_void_p))
        return llr[:n].copy(), sh.value, ph, nre

    def pdsch_tx_slot(self, P, bits):
        """gNB PDSCH transmitter after the encoder: bits (uint8, one per element, P.G() of them) -> txdataF [nb_tx][14][N][2] (zeros where nothing is mapped)."""
        b = np.ascontiguousarray(bits, dtype=np.uint8)
        assert b.size == P.G(), (b.size, P.G())
        out = np.zeros((P.nb_tx, 14, P.fft_size, 2), np.int16)
        self.lib.orc_pdsch_tx_slot.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        rc = self.lib.orc_pdsch_tx_slot(C.addressof(P), b.ctypes.data, out.ctypes.data)
        assert rc == b.size, rc
        return out

    # ---- slot-level OFDM front end
    def ofdm_geometry(self, N, mu, slot):
        pre = np.zeros(14, np.uint32); cps = np.zeros(14, np.uint32); ss = C.c_uint32(); fl = C.c_uint32()
        self.lib.orc_ofdm_geometry(N, mu, slot, pre.ctypes.data_as(C.c_void_p), cps.ctypes.data_as(C.c_void_p), C.byref(ss), C.byref(fl))
        return pre, cps, ss.value, fl.value

    def symbol_rotation(self, mu, f0):
        out = np.zeros(2 * (14 << mu), np.int16)
        self.lib.orc_symbol_rotation.argtypes = [C.c_int, C.c_double, C.c_void_p]
        self.lib.orc_symbol_rotation(mu, float(f0), out.ctypes.data)
        return out

    def timeshift_rotation(self, N, sample_offset):
        out = np.zeros(2 * N, np.int16)
        self.lib.orc_timeshift_rotation(N, sample_offset, out.ctypes.data_as(C.c_void_p))
        return out

    def ofdm_tx_slot(self, N, mu, nb_rb, slot, nsymb, rot, txdataF):
        pre, cps, _, _ = self.ofdm_geometry(N, mu, slot)
        F = np.ascontiguousarray(txdataF, dtype=np.int16).copy()
        out = np.zeros(2 * int(cps[13] + pre[13] + N), np.int16)
        r = None if rot is None else np.ascontiguousarray(rot, dtype=np.int16)
        self.lib.orc_ofdm_tx_slot(N, mu, nb_rb, slot, nsymb, None if r is None else r.ctypes.data_as(C.c_void_p), F.ctypes.data_as(C.c_void_p),
                                  out.ctypes.data_as(C.c_void_p))
        return out, F

    def ofdm_rx_slot(self, N, mu, nb_rb, slot, divisor, sample_offset, rot, rxdata):
        x = np.ascontiguousarray(rxdata, dtype=np.int16)
        out = np.zeros(2 * 14 * N, np.int16)
        r = None if rot is None else np.ascontiguousarray(rot, dtype=np.int16)
        self.lib.orc_ofdm_rx_slot(N, mu, nb_rb, slot, divisor, sample_offset, None if r is None else r.ctypes.data_as(C.c_void_p),
                                  x.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
        return out

    def dft(self, N, inverse, x, scale=1):
        x = np.ascontiguousarray(x, dtype=np.int16)
        y = np.zeros(2 * N, dtype=np.int16)
        rc = self.lib.orc_dft(N, int(inverse), x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p), scale)
        assert rc == 0, rc
        return y

    def dft4(self, N, x, scale=1):
        """Four-way sizes 12 ... 3240 (DFT-s-OFDM): x = int16[2 * 4 * N], transform l at c16 positions 4 n + l."""
        x = np.ascontiguousarray(x, dtype=np.int16)
        assert x.size == 8 * N
        y = np.zeros(8 * N, dtype=np.int16)
        rc = self.lib.orc_dft4(N, x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p), scale)
        assert rc == 0, rc
        return y

    def segmentation(self, data, B, BG):
        Cc, K, Zo, F = C.c_uint(), C.c_uint(), C.c_uint(), C.c_uint()
        Kb = self.lib.orc_segmentation(None, None, B, C.byref(Cc), C.byref(K), C.byref(Zo), C.byref(F), BG)
        if Kb < 0:
            return Kb, 0, 0, 0, 0, None
        segs = np.zeros((Cc.value, K.value // 8 + 8), dtype=np.uint8)
        if data is not None:
            d = np.ascontiguousarray(data, dtype=np.uint8)
            ptrs = (_u8p * Cc.value)(*[C.cast(segs[r].ctypes.data, _u8p) for r in range(Cc.value)])
            self.lib.orc_segmentation(_ptr(d, _u8p), ptrs, B, C.byref(Cc), C.byref(K), C.byref(Zo), C.byref(F), BG)
        return Kb, Cc.value, K.value, Zo.value, F.value, segs[:, :K.value // 8]


# ------------------------------------------------------------------------------------------------ compiled reference
def have_reference():
    return all(os.path.exists(os.path.join(REFDIR, f)) for f in
               ("libref_ldpc_dec.so", "libref_ldpc_enc.so", "libref_ldpc_enc_orig.so", "libref_coding.so", "libref_dfts.so"))


class Reference:
    """The unmodified OAI sources compiled by oracle/build_ref.sh."""

    def __init__(self, avx512=False):
        if not have_reference():
            raise FileNotFoundError("oracle/_ref not built (run oracle/build_ref.sh where /root/reference exists)")
        self.dec = C.CDLL(os.path.join(REFDIR, "libref_ldpc_dec512.so" if avx512 else "libref_ldpc_dec.so"))
        self.enc = C.CDLL(os.path.join(REFDIR, "libref_ldpc_enc.so"))
        self.enc_orig = C.CDLL(os.path.join(REFDIR, "libref_ldpc_enc_orig.so"))
        self.cod = C.CDLL(os.path.join(REFDIR, "libref_coding.so"))
        self.cod.crcTableInit()
        for n in ("crc24a", "crc24b", "crc24c", "crc16", "crc12", "crc11", "crc8", "crc6"):
            f = getattr(self.cod, n)
            f.restype = C.c_uint32
            f.argtypes = [_u8p, C.c_int]
        self.cod.check_crc.argtypes = [_u8p, C.c_uint32, C.c_uint8]
        self.dec.LDPCdecoder.argtypes = [C.POINTER(DecParams), C.c_uint8, C.c_uint8, C.c_uint8, _i8p, _i8p, C.c_void_p, C.c_void_p]
        self.dec.LDPCinit()
        self._prof = LdpcTimeStats()

    def crc(self, poly_id, data, bitlen):
        d = np.ascontiguousarray(data, dtype=np.uint8)
        d = np.concatenate([d, np.zeros(8, np.uint8)])
        name = ("crc24a", "crc24b", "crc24c", "crc16", "crc12", "crc11", "crc8", "crc6")[poly_id]
        return int(getattr(self.cod, name)(_ptr(d, _u8p), bitlen))

    def decode(self, BG, Z, R, max_iter, llr, out_mode=0, use_crc=0, crc_len_bits=0, crc_type=0, abort_in=0):
        n = ncols_for_rate(BG, R) * Z
        # the reference reads whole 32-byte vectors: give it NR_LDPC_MAX_NUM_LLR-sized aligned buffers like its callers do
        buf = np.zeros(27000 + 64, dtype=np.int8)
        off = (-buf.ctypes.data) % 64
        a = buf[off:off + 27000]
        a[:n] = np.asarray(llr, dtype=np.int8)[:n]
        out = np.zeros(27000 + 64, dtype=np.int8)
        offo = (-out.ctypes.data) % 64
        o = out[offo:offo + 27000]
        p = DecParams()
        p.BG, p.Z, p.R, p.numMaxIter, p.outMode, p.E, p.crc_type = BG, Z, R, max_iter, out_mode, crc_len_bits, crc_type
        p.check_crc = C.cast(self.cod.check_crc, C.c_void_p).value if use_crc else None
        ab = DecodeAbort()
        ab.failed = bool(abort_in)
        it = self.dec.LDPCdecoder(C.byref(p), 0, 0, 0, _ptr(a, _i8p), _ptr(o, _i8p), C.cast(C.byref(self._prof), C.c_void_p),
                                  C.cast(C.byref(ab), C.c_void_p))
        self.last_abort = bool(ab.failed)
        return it, (o[:(n + 7) // 8].view(np.uint8).copy() if out_mode == 0 else o[:n].copy())

    def decode_raw_fn(self):
        return self.dec.LDPCdecoder

    def encode(self, BG, Z, K, payloads, orig=False):
        """payloads: (n_seg, K/8) uint8; returns (n_seg, 66Z|50Z) uint8 of 0/1 -- LDPCencoder in groups of 8."""
        payloads = np.ascontiguousarray(payloads, dtype=np.uint8)
        nseg = payloads.shape[0]
        nout = (66 if BG == 1 else 50) * Z
        pad_in = np.zeros((nseg, K // 8 + 64), dtype=np.uint8)
        pad_in[:, :K // 8] = payloads
        outs = np.zeros((nseg, 68 * 384 + 64), dtype=np.uint8)
        inp = (_u8p * nseg)(*[C.cast(pad_in[i].ctypes.data, _u8p) for i in range(nseg)])
        oup = (_u8p * nseg)(*[C.cast(outs[i].ctypes.data + ((-outs[i].ctypes.data) % 32), _u8p) for i in range(nseg)])
        ip = EncParams()
        ip.n_segments, ip.Kb, ip.Zc, ip.BG, ip.K, ip.gen_code = nseg, (22 if BG == 1 else 10), Z, BG, K, 0
        if orig:
            for j in range(nseg):
                one_in = (_u8p * 1)(inp[j])
                one_out = (_u8p * 1)(oup[j])
                ip.n_segments = 1
                rc = self.enc_orig.LDPCencoder(one_in, one_out, C.byref(ip))
                assert rc == nout, rc  # ldpc_encoder.c:259 returns the coded length
        else:
            for m in range((nseg + 7) // 8):
                ip.macro_num = m
                rc = self.enc.LDPCencoder(inp, oup, C.byref(ip))
                assert rc == 0
        res = np.zeros((nseg, nout), dtype=np.uint8)
        for i in range(nseg):
            o = (-outs[i].ctypes.data) % 32
            res[i] = outs[i, o:o + nout]
        return res

    # ---- coding helpers of the compiled reference (libref_coding.so)
    def check_crc(self, data, n, crc_type):
        d = np.ascontiguousarray(data, dtype=np.uint8)
        return int(self.cod.check_crc(_ptr(d, _u8p), n, crc_type))

    def rate_matching_tx(self, Tbslbrm, BG, Z, w, C_, F, Foffset, rv, E):
        w = np.ascontiguousarray(w, dtype=np.uint8)
        e = np.zeros(E + 64, dtype=np.uint8)
        f = self.cod.nr_rate_matching_ldpc
        f.argtypes = [C.c_uint32, C.c_uint8, C.c_uint16, _u8p, _u8p, C.c_uint8, C.c_uint32, C.c_uint32, C.c_uint8, C.c_uint32]
        rc = f(Tbslbrm, BG, Z, _ptr(w, _u8p), _ptr(e, _u8p), C_, F, Foffset, rv, E)
        return rc, e[:E]

    def rate_matching_rx(self, Tbslbrm, BG, Z, w, soft, C_, rv, clear, E, F, Foffset):
        soft = np.ascontiguousarray(soft, dtype=np.int16)
        assert w.dtype == np.int16 and w.flags.c_contiguous
        f = self.cod.nr_rate_matching_ldpc_rx
        f.argtypes = [C.c_uint32, C.c_uint8, C.c_uint16, _i16p, _i16p, C.c_uint8, C.c_uint8, C.c_uint8, C.c_uint32, C.c_uint32, C.c_uint32]
        return f(Tbslbrm, BG, Z, _ptr(w, _i16p), _ptr(soft, _i16p), C_, rv, clear, E, F, Foffset)

    def interleave(self, E, Qm, e):
        e = np.ascontiguousarray(e, dtype=np.uint8)
        f = np.zeros(E + 64, dtype=np.uint8)
        fn = self.cod.nr_interleaving_ldpc
        fn.argtypes = [C.c_uint32, C.c_uint8, _u8p, _u8p]
        fn.restype = None
        fn(E, Qm, _ptr(e, _u8p), _ptr(f, _u8p))
        return f[:E]

    def deinterleave(self, E, Qm, f):
        f = np.ascontiguousarray(f, dtype=np.int16)
        e = np.zeros(E + 64, dtype=np.int16)
        fn = self.cod.nr_deinterleaving_ldpc
        fn.argtypes = [C.c_uint32, C.c_uint8, _i16p, _i16p]
        fn.restype = None
        fn(E, Qm, _ptr(e, _i16p), _ptr(f, _i16p))
        return e[:E]

    def get_R(self, rv, E, BG, Z, llrLen, rnd):
        ll = C.c_int(llrLen)
        fn = self.cod.nr_get_R_ldpc_decoder
        fn.argtypes = [C.c_int] * 4 + [C.POINTER(C.c_int), C.c_int]
        r = fn(rv, E, BG, Z, C.byref(ll), rnd)
        return r, ll.value

    def segmentation(self, data, B, BG):
        Cc, K, Zo, F = C.c_uint(), C.c_uint(), C.c_uint(), C.c_uint()
        fn = self.cod.nr_segmentation
        fn.argtypes = [_u8p, C.POINTER(_u8p), C.c_uint, C.POINTER(C.c_uint), C.POINTER(C.c_uint), C.POINTER(C.c_uint), C.POINTER(C.c_uint), C.c_uint8]
        fn.restype = C.c_int32
        Kb = fn(None, None, B, C.byref(Cc), C.byref(K), C.byref(Zo), C.byref(F), BG)
        if Kb < 0:
            return Kb, 0, 0, 0, 0, None
        segs = np.zeros((Cc.value, K.value // 8 + 8), dtype=np.uint8)
        if data is not None:
            d = np.ascontiguousarray(data, dtype=np.uint8)
            ptrs = (_u8p * Cc.value)(*[C.cast(segs[r].ctypes.data, _u8p) for r in range(Cc.value)])
            fn(_ptr(d, _u8p), ptrs, B, C.byref(Cc), C.byref(K), C.byref(Zo), C.byref(F), BG)
        return Kb, Cc.value, K.value, Zo.value, F.value, segs[:, :K.value // 8]

    # ---- DFT library of the reference (libref_dfts.so): the per-size entry points and the dft()/idft() dispatchers
    def dft(self, N, inverse, x, scale=1):
        if not hasattr(self, "_dfts"):
            self._dfts = C.CDLL(os.path.join(REFDIR, "libref_dfts.so"))
            self._dfts.dfts_autoinit()
        buf = np.zeros(2 * N + 64, dtype=np.int16)
        o = ((-buf.ctypes.data) % 32) // 2
        xin = buf[o:o + 2 * N]
        xin[:] = np.asarray(x, dtype=np.int16)
        out = np.zeros(2 * N + 64, dtype=np.int16)
        oo = ((-out.ctypes.data) % 32) // 2
        y = out[oo:oo + 2 * N]
        getattr(self._dfts, ("idft" if inverse else "dft") + str(N))(xin.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p), C.c_ubyte(scale))
        return y.copy()

    def dft4(self, N, x, scale=1, name=None):
        """The four-way entry points dft12 ... dft3240 (4 N c16 in and out)."""
        self.dft(64, False, np.zeros(128, np.int16))                                   # loads the library
        buf = np.zeros(8 * N + 64, dtype=np.int16)
        o = ((-buf.ctypes.data) % 32) // 2
        xin = buf[o:o + 8 * N]
        xin[:] = np.asarray(x, dtype=np.int16)
        out = np.zeros(8 * N + 64, dtype=np.int16)
        oo = ((-out.ctypes.data) % 32) // 2
        y = out[oo:oo + 8 * N]
        getattr(self._dfts, name or ("dft" + str(N)))(xin.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p), C.c_ubyte(scale))
        return y.copy()

    # ---- PUSCH LLR computation of the reference (libref_llr.so: nr_ulsch_compute_llr)
    def ulsch_llr(self, Qm, rxF, maga, magb, magc):
        if not hasattr(self, "_llr"):
            self._llr = C.CDLL(os.path.join(REFDIR, "libref_llr.so"))
        n = np.asarray(rxF).size // 2
        def al(a):
            buf = np.zeros(2 * n + 64, dtype=np.int16)
            o = ((-buf.ctypes.data) % 32) // 2
            v = buf[o:o + 2 * n + 32]
            if a is not None:
                v[:2 * n] = np.asarray(a, dtype=np.int16)
            return v
        x, a, b, c = al(rxF), al(maga), al(magb), al(magc)
        out = np.zeros(n * Qm + 128, dtype=np.int16)
        oo = ((-out.ctypes.data) % 32) // 2
        o = out[oo:oo + n * Qm + 64]
        p = lambda v: v.ctypes.data_as(C.c_void_p)
        self._llr.nr_ulsch_compute_llr(p(x), p(a), p(b), p(c), p(o), C.c_uint32(n), C.c_uint8(0), C.c_uint8(Qm))
        return o[:n * Qm].copy()

    # ---- scrambling + QAM mapper of the reference (libref_mod.so: nr_scrambling.c, nr_modulation.c, nr_gen_mod_table.c)
    def _chest(self):
        if not hasattr(self, "_chestlib"):
            self._chestlib = C.CDLL(os.path.join(REFDIR, "libref_chest.so"))
            assert self._chestlib.refh_chest_init(os.path.join(REFDIR, "libref_dfts.so").encode()) == 0
        return self._chestlib

    def lowpapr_seq(self, u, v, n_re):
        """gNB_dmrs_lowpaprtype1_sequence[u][v][index(n_re)] of the compiled reference (ul_ref_seq_nr.c), None when n_re is not 6 * 2^a 3^b 5^c."""
        out = np.zeros(2 * n_re, np.int16)
        return out if self._chest().refh_lowpapr_seq(u, v, n_re, out.ctypes.data_as(C.c_void_p)) >= 0 else None

    def chest_set_transform_precoding(self, on, u=0, v=0):
        self._chest().refh_chest_set_transform_precoding(int(on), u, v)

    def pusch_set_transform_precoding(self, on):
        assert self._pusch().refh_pusch_set_transform_precoding(int(on), os.path.join(REFDIR, "libref_dfts.so").encode()) == 0

    def pusch_channel_estimation(self, P, rxdataF, n_rb_ul, chest_freq=0, dmrs_type=0):
        self._chest()
        prm = np.array([P.fft_size, P.nb_rx, n_rb_ul, P.slot, P.symbol, P.port, P.rb_start, P.bwp_start, P.rb_size, P.first_carrier_offset, P.scid,
                        P.dmrs_scrambling_id, dmrs_type, chest_freq], dtype=np.int32)
        x = np.ascontiguousarray(rxdataF, dtype=np.int16).copy()
        est = np.zeros(P.nb_rx * 14 * P.fft_size * 2, np.int16)
        out = np.zeros(5, np.int32)
        pil = np.zeros(2 * 6 * P.rb_size, np.int16)
        self._chestlib.refh_pusch_chest(prm.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p), est.ctypes.data_as(C.c_void_p),
                                        out.ctypes.data_as(C.c_void_p), pil.ctypes.data_as(C.c_void_p))
        return est.reshape(P.nb_rx, 14, P.fft_size, 2), out, pil

    def pdsch_channel_estimation(self, P, rxdataF, n_rb_dl, chest_freq=0, dmrs_type=0):
        if not hasattr(self, "_uechestlib"):
            self._uechestlib = C.CDLL(os.path.join(REFDIR, "libref_uechest.so"))
            assert self._uechestlib.refh_uechest_init(os.path.join(REFDIR, "libref_dfts.so").encode()) == 0
        prm = np.array([P.fft_size, P.nb_rx, n_rb_dl, P.slot, P.symbol, P.port, P.rb_start, P.bwp_start, P.rb_size, P.first_carrier_offset, P.scid,
                        P.dmrs_scrambling_id, dmrs_type, chest_freq], dtype=np.int32)
        x = np.ascontiguousarray(rxdataF, dtype=np.int16).copy()
        est = np.zeros(P.nb_rx * 14 * P.fft_size * 2, np.int16)
        self._uechestlib.refh_pdsch_chest(prm.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p), est.ctypes.data_as(C.c_void_p))
        return est.reshape(P.nb_rx, 14, P.fft_size, 2)

    def ue_slot_fep(self, N, mu, nb_rb, nrx, slot, divisor, rot_dl224, timeshift, rxdata):
        """the UE's nr_slot_fep for the 14 symbols of a slot.  rxdata [nrx][samples][2]; returns rxdataF [nrx][14 N 2]."""
        if not hasattr(self, "_uechestlib"):
            self._uechestlib = C.CDLL(os.path.join(REFDIR, "libref_uechest.so"))
            assert self._uechestlib.refh_uechest_init(os.path.join(REFDIR, "libref_dfts.so").encode()) == 0
        x = np.ascontiguousarray(rxdata, dtype=np.int16).reshape(nrx, -1)
        out = np.zeros((nrx, 2 * 14 * N), np.int16)
        r = np.ascontiguousarray(rot_dl224, dtype=np.int16); t = np.ascontiguousarray(timeshift, dtype=np.int16)
        self._uechestlib.refh_ue_slot_fep.argtypes = [C.c_int] * 6 + [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
        self._uechestlib.refh_ue_slot_fep(N, mu, nb_rb, nrx, slot, divisor, r.ctypes.data, t.ctypes.data, x.ctypes.data, x.shape[1] // 2, out.ctypes.data)
        return out

    def chest_time_domain_avg(self, est, num_symbols, start_symbol, dmrs_bitmap, num_rbs):
        if not hasattr(self, "_pdschlib"):
            self._pdschlib = C.CDLL(os.path.join(REFDIR, "libref_pdsch.so"))
        e = np.ascontiguousarray(est, dtype=np.int16).copy()
        self._pdschlib.refh_chest_time_avg(e.shape[2], e.shape[0], num_symbols, start_symbol, dmrs_bitmap, num_rbs, e.ctypes.data_as(C.c_void_p))
        return e

    def pdsch_rx_slot(self, P, start_symbol, nr_symbols, rxdataF, dl_ch_est, G, nl=1):
        if not hasattr(self, "_pdschlib"):
            self._pdschlib = C.CDLL(os.path.join(REFDIR, "libref_pdsch.so"))
        prm = np.array([P.fft_size, P.nb_rx, P.rb_start, P.bwp_start, P.rb_size, P.first_carrier_offset, P.Qm, start_symbol, nr_symbols, P.ul_dmrs_symb_pos,
                        P.dmrs_config_type, P.num_dmrs_cdm_grps_no_data, G, nl], dtype=np.int32)
        x = np.ascontiguousarray(rxdataF, dtype=np.int16).copy(); h = np.ascontiguousarray(dl_ch_est, dtype=np.int16).copy()
        llr = np.zeros(G + 64, np.int16); valid = np.zeros(14, np.int32)
        sh = self._pdschlib.refh_pdsch_rx_slot(prm.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p), h.ctypes.data_as(C.c_void_p), llr.ctypes.data_as(C.c_void_p),
                                               valid.ctypes.data_as(C.c_void_p), None)
        return llr[:G].copy(), int(sh), valid

    def rfsim_rx_add_input(self, nb_tx, nb_rx, L, offset, pl_dB, noise_dB, ch, sig, out, rx_ant, TS, cir, noise=None):
        """The real rxAddInput (oracle/_ref/libref_rfsim.so); gaussZiggurat returns `noise` in call order."""
        if not hasattr(self, "_rfsimlib"):
            self._rfsimlib = C.CDLL(os.path.join(REFDIR, "libref_rfsim.so"))
        return _rfsim_call(self._rfsimlib.refh_rfsim_rx_add_input, nb_tx, nb_rx, L, offset, pl_dB, noise_dB, ch, sig, out, rx_ant, TS, cir, noise)

    def _prach(self):
        if not hasattr(self, "_prachlib"):
            self._prachlib = C.CDLL(os.path.join(REFDIR, "libref_prach.so"))
            assert self._prachlib.refh_prach_init(os.path.join(REFDIR, "libref_dfts.so").encode()) == 0
        return self._prachlib

    def prach_seq(self, short_sequence, num_sequences, root_index):
        """compute_nr_prach_seq: gNB->X_u [64][839][2] int16."""
        xu = np.zeros((64, 839, 2), np.int16)
        self._prach().refh_prach_seq(short_sequence, num_sequences, root_index, xu.ctypes.data_as(C.c_void_p))
        return xu

    def db_fixed_times10(self, x):
        return int(self._prach().refh_db_fixed_times10(C.c_uint32(x)))

    def rx_nr_prach(self, nb_rx, short_sequence, root_index, num_roots, NCS, prach_fmt, mu, xu, rxsigF):
        """The real rx_nr_prach (unrestricted set).  Returns (max_preamble, max_preamble_energy, max_preamble_delay)."""
        x = np.ascontiguousarray(xu, dtype=np.int16); r = np.ascontiguousarray(rxsigF, dtype=np.int16)
        p = np.array([nb_rx, short_sequence, root_index, num_roots, NCS, prach_fmt, mu], np.int32)
        out = np.zeros(3, np.int32)
        self._prach().refh_rx_nr_prach(p.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p), r.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
        return tuple(int(v) for v in out)

    def pdsch_rx_slot_ptrs(self, P, T, start_symbol, nr_symbols, rxdataF, dl_ch_est, G, n_rb_dl=273):
        """The real nr_rx_pdsch + nr_pdsch_ptrs_processing (libref_pdsch_ptrs.so).  Returns (llr, log2_maxh, valid[14], phase[14][2], ptrs_re[14])."""
        if not hasattr(self, "_pdschptrslib"):
            self._pdschptrslib = C.CDLL(os.path.join(REFDIR, "libref_pdsch_ptrs.so"))
        L = self._pdschptrslib
        q = np.array([T.on, T.L, T.K, T.re_offset, T.rnti, T.slot, T.nscid, T.nid, n_rb_dl], dtype=np.int32)
        L.refh_pdsch_set_ptrs(q.ctypes.data_as(C.c_void_p))
        prm = np.array([P.fft_size, P.nb_rx, P.rb_start, P.bwp_start, P.rb_size, P.first_carrier_offset, P.Qm, start_symbol, nr_symbols, P.ul_dmrs_symb_pos,
                        P.dmrs_config_type, P.num_dmrs_cdm_grps_no_data, G, 1], dtype=np.int32)
        x = np.ascontiguousarray(rxdataF, dtype=np.int16).copy(); h = np.ascontiguousarray(dl_ch_est, dtype=np.int16).copy()
        llr = np.zeros(G + 64, np.int16); valid = np.zeros(14, np.int32)
        sh = L.refh_pdsch_rx_slot(prm.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p), h.ctypes.data_as(C.c_void_p), llr.ctypes.data_as(C.c_void_p),
                                  valid.ctypes.data_as(C.c_void_p), None)
        ph = np.zeros((14, 2), np.int16); nre = np.zeros(14, np.int32)
        L.refh_pdsch_get_ptrs(ph.ctypes.data_as(C.c_void_p), nre.ctypes.data_as(C.c_void_p))
        L.refh_pdsch_set_ptrs(None)
        return llr[:G].copy(), int(sh), valid, ph, nre

    def pdsch_tx_slot(self, P, bits, n_rb_dl):
        if not hasattr(self, "_pdschtxlib"):
            self._pdschtxlib = C.CDLL(os.path.join(REFDIR, "libref_pdschtx.so"))
        prm = np.array([P.fft_size, n_rb_dl, P.nb_tx, P.slot, P.rb_start, P.bwp_start, P.rb_size, P.first_carrier_offset, P.Qm, P.nrOfLayers, P.start_symbol,
                        P.nr_of_symbols, P.dl_dmrs_symb_pos, P.dmrs_config_type, P.num_dmrs_cdm_grps_no_data, P.dmrs_ports, P.scid, P.dl_dmrs_scrambling_id,
                        P.data_scrambling_id, P.rnti, P.amp], dtype=np.int32)
        b = np.ascontiguousarray(bits, dtype=np.uint8).copy()
        out = np.zeros((P.nb_tx, 14, P.fft_size, 2), np.int16)
        self._pdschtxlib.refh_pdsch_tx_slot.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
        self._pdschtxlib.refh_pdschtx_set_precoding.argtypes = [C.c_int, C.c_void_p]
        w = np.array(list(P.pm_weights), dtype=np.int16)
        self._pdschtxlib.refh_pdschtx_set_precoding(int(P.pm_idx), w.ctypes.data)
        self._pdschtxlib.refh_pdschtx_set_ptrs(int(P.ptrs_on), int(P.ptrs_L), int(P.ptrs_K), int(P.ptrs_re_offset))
        self._pdschtxlib.refh_pdsch_tx_slot(prm.ctypes.data, b.ctypes.data, b.size, out.ctypes.data)
        self._pdschtxlib.refh_pdschtx_set_precoding(0, None)
        self._pdschtxlib.refh_pdschtx_set_ptrs(0, 0, 0, 0)
        return out

    def _pusch(self):
        if not hasattr(self, "_puschlib"):
            self._puschlib = C.CDLL(os.path.join(REFDIR, "libref_pusch.so"))
        return self._puschlib

    @staticmethod
    def _pusch_params(P, nb_layer, symbol, ch_symbol, shift, nvar, valid):
        return np.array([P.fft_size, P.nb_rx, nb_layer, P.rb_start, P.bwp_start, P.rb_size, P.first_carrier_offset, P.Qm, symbol, ch_symbol, P.ul_dmrs_symb_pos,
                         P.num_dmrs_cdm_grps_no_data, P.dmrs_config_type, shift, nvar, valid], dtype=np.int32)

    def pusch_inner_rx_symbol(self, P, symbol, ch_symbol, shift, rxdataF, ch_est, valid, nb_layer=1, nvar=0):
        L = self._pusch()
        prm = self._pusch_params(P, nb_layer, symbol, ch_symbol, shift, nvar, valid)
        x = np.ascontiguousarray(rxdataF, dtype=np.int16).copy(); h = np.ascontiguousarray(ch_est, dtype=np.int16).copy()
        blen = (P.rb_size * 12 + 15) & ~15
        llr = np.zeros(nb_layer * valid * P.Qm + 64, np.int16); comp = np.zeros(nb_layer * 2 * blen, np.int16)
        L.refh_pusch_inner_rx(prm.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p), h.ctypes.data_as(C.c_void_p), llr.ctypes.data_as(C.c_void_p),
                              comp.ctypes.data_as(C.c_void_p))
        if nb_layer == 2:
            return llr[:2 * valid * P.Qm].reshape(2, -1), comp.reshape(2, -1)
        return llr[:valid * P.Qm], comp

    def pusch_log2_maxh(self, P, meas_symbol, ch_symbol, rxdataF, ch_est, nb_layer=1, max_ch=0):
        L = self._pusch()
        prm = self._pusch_params(P, nb_layer, meas_symbol, ch_symbol, 0, 0, 0)
        x = np.ascontiguousarray(rxdataF, dtype=np.int16).copy(); h = np.ascontiguousarray(ch_est, dtype=np.int16).copy()
        avg = np.zeros(8, np.int32)
        raw = L.refh_pusch_log2_maxh(prm.ctypes.data_as(C.c_void_p), max_ch, x.ctypes.data_as(C.c_void_p), h.ctypes.data_as(C.c_void_p), avg.ctypes.data_as(C.c_void_p))
        # the final rule of nr_rx_pusch_tp for one layer (:1642-1646): + 1 + log2_approx(nb_rx >> 2), floored at 0
        if nb_layer == 2:                                         # - 3 for the MMSE receiver only (:1640-1641)
            return max(0, int(raw) - (3 if P.Qm >= 6 else 0)), avg[:2 * P.nb_rx].copy()
        return max(0, int(raw) + 1 + int(P.nb_rx >> 2).bit_length()), avg[:P.nb_rx].copy()

    def _ofdm(self):
        if not hasattr(self, "_ofdmlib"):
            self._ofdmlib = C.CDLL(os.path.join(REFDIR, "libref_ofdm.so"))
            rc = self._ofdmlib.refh_ofdm_init(os.path.join(REFDIR, "libref_dfts.so").encode())
            assert rc == 0, rc
            self._ofdmlib.refh_rotation_tables.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
        return self._ofdmlib

    def rotation_tables(self, N, mu, nb_rb, divisor, dl_freq, ul_freq):
        L = self._ofdm()
        dl = np.zeros(448, np.int16); ul = np.zeros(448, np.int16); ts = np.zeros(2 * N, np.int16)
        L.refh_rotation_tables(N, mu, nb_rb, divisor, float(dl_freq), float(ul_freq), dl.ctypes.data, ul.ctypes.data, ts.ctypes.data)
        return dl, ul, ts

    def ofdm_tx_slot(self, N, mu, nb_rb, slot, nsymb, rot224, txdataF, out_len):
        L = self._ofdm()
        F = np.ascontiguousarray(txdataF, dtype=np.int16).copy()
        out = np.zeros(2 * out_len + 64, np.int16)
        r = None if rot224 is None else np.ascontiguousarray(rot224, dtype=np.int16)
        L.refh_ofdm_tx_slot(N, mu, nb_rb, slot, nsymb, None if r is None else r.ctypes.data_as(C.c_void_p), F.ctypes.data_as(C.c_void_p),
                            out.ctypes.data_as(C.c_void_p))
        return out[:2 * out_len], F

    def ofdm_rx_slot(self, N, mu, nb_rb, slot, divisor, sample_offset, rot224, rxdata):
        L = self._ofdm()
        x = np.ascontiguousarray(rxdata, dtype=np.int16).copy()
        buf = np.zeros(2 * 14 * N + 32, np.int16)
        off = ((-buf.ctypes.data) % 32) // 2           # dft() asserts a 32-byte aligned output
        out = buf[off:off + 2 * 14 * N]
        r = None if rot224 is None else np.ascontiguousarray(rot224, dtype=np.int16)
        L.refh_ofdm_rx_slot(N, mu, nb_rb, slot, divisor, sample_offset, None if r is None else r.ctypes.data_as(C.c_void_p),
                            x.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
        return out.copy()

    def _mod(self):
        if not hasattr(self, "_modlib"):
            self._modlib = C.CDLL(os.path.join(REFDIR, "libref_mod.so"))
            self._modlib.nr_generate_modulation_table()
            self._modlib.init_byte2m128i()
        return self._modlib

    def scramble(self, in_bits, q, Nid, rnti):
        L = self._mod()
        x = np.asarray(in_bits, dtype=np.uint8)
        buf = np.zeros(x.size + 128, dtype=np.uint8)
        o = (-buf.ctypes.data) % 32
        v = buf[o:o + x.size + 64]
        v[:x.size] = x
        out = np.zeros((x.size + 31) // 32 + 8, dtype=np.uint32)
        L.nr_codeword_scrambling(v.ctypes.data_as(C.c_void_p), C.c_uint32(x.size), C.c_uint8(q), C.c_uint32(Nid), C.c_uint32(rnti), out.ctypes.data_as(C.c_void_p))
        return out[:(x.size + 31) // 32].copy()

    def unscramble_llr(self, llr, q, Nid, rnti):
        L = self._mod()
        x = np.asarray(llr, dtype=np.int16)
        buf = np.zeros(x.size + 128, dtype=np.int16)
        o = ((-buf.ctypes.data) % 32) // 2
        v = buf[o:o + x.size + 64]
        v[:x.size] = x
        L.nr_codeword_unscrambling(v.ctypes.data_as(C.c_void_p), C.c_uint32(x.size), C.c_uint8(q), C.c_uint32(Nid), C.c_uint32(rnti))
        return v[:x.size].copy()

    def modulate(self, packed_bits, length, Qm):
        L = self._mod()
        x = np.ascontiguousarray(packed_bits).view(np.uint8)
        buf = np.zeros(x.size + 64, dtype=np.uint8)
        buf[:x.size] = x
        out = np.zeros(2 * (length // Qm) + 64, dtype=np.int16)
        oo = ((-out.ctypes.data) % 32) // 2
        o = out[oo:oo + 2 * (length // Qm) + 16]
        L.nr_modulation(buf.ctypes.data_as(C.c_void_p), C.c_uint32(length), C.c_uint16(Qm), o.ctypes.data_as(C.c_void_p))
        return o[:2 * (length // Qm)].copy()
