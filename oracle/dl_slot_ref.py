"""TEST INFRASTRUCTURE ONLY -- the nr_dlsim-shaped PDSCH slot (gNB transmit + UE receive) run through the UNMODIFIED reference functions compiled under
oracle/_ref, in the order nr_generate_pdsch / nr_feptx0 and nr_ue_pdsch_procedures / nr_ue_dlsch_procedures call them.  Two uses: (1) the CPU baseline of the
slot metric (bench.py cpu_baseline / --impl reference, tools/bench_dl_slot.py): every stage is timed around the reference call itself -- inside the C harness for
the functions that need one (allocation and copying of the harness excluded), around the bare ctypes call with pre-built buffers for the library functions --
on ONE host thread, the way nr_dlsim runs a slot; (2) an end-to-end parity check of the CUDA chain (tests/test_gpu_dl_slot_chain.py): same payload in, the same
time-domain samples out, the same LLRs, iteration counts and transport block back.  Only tests/, smoke() and bench.py may import this."""
import ctypes as C
import time

import numpy as np

from . import bindings as ob
from openairinterface5g_b200 import transport as T
from openairinterface5g_b200.ofdm import NrOfdmParms

_u8p, _i8p, _i16p = C.POINTER(C.c_uint8), C.POINTER(C.c_int8), C.POINTER(C.c_int16)


def _aligned(n, dtype, align=64):
    item = np.dtype(dtype).itemsize
    buf = np.zeros(n + align // item + 8, dtype=dtype)
    off = ((-buf.ctypes.data) % align) // item
    return buf[off:off + n]


class RefDlSlot:
    def __init__(self, A=434280, N=4096, mu=1, carrier_rb=273, rb_start=0, rb_size=273, nb_ant=2, Qm=6, slot=1, rnti=0x1234, nid=77, dl_freq=3619200000.0,
                 max_iter=8, dmrs_id=55, n_layers=2, tx_amp=512, start_symbol=1, nr_symbols=13):
        self.ref = ob.Reference()
        self.P = NrOfdmParms(N, mu, carrier_rb)
        self.N, self.mu, self.carrier_rb, self.nb, self.Qm, self.slot, self.rnti, self.nid, self.max_iter, self.nl = N, mu, carrier_rb, nb_ant, Qm, slot, rnti, nid, max_iter, n_layers
        self.rb_start, self.rb_size, self.A, self.dmrs_id, self.start_symbol, self.nr_symbols = rb_start, rb_size, A, dmrs_id, start_symbol, nr_symbols
        self.dl_freq = dl_freq
        seg = self.seg = T.nr_segmentation(A + 24, 1)
        self.C_, self.K, self.Z, self.F = seg["C"], seg["K"], seg["Z"], seg["F"]
        self.txP = ob.PdschTxParms(N, nb_ant, slot, rb_start, 0, rb_size, self.P.first_carrier_offset, Qm, n_layers, start_symbol, nr_symbols, 1 << 2, 0, 2,
                                   (1 << n_layers) - 1, 0, dmrs_id, nid, rnti, tx_amp)
        self.G = self.txP.G()
        self.E = [T.nr_get_E(self.G, self.C_, Qm, n_layers, r) for r in range(self.C_)]
        self.R = T.nr_get_R_ldpc_decoder(0, self.E[0], 1, self.Z)[0]
        rot = self.P.symbol_rotation(dl_freq)
        self.rot224 = np.zeros(448, np.int16); self.rot224[:rot.size] = rot.reshape(-1)
        self.ts = self.ref.rotation_tables(N, mu, carrier_rb, 8, dl_freq, dl_freq)[2]
        self.t = {}

    def _tic(self, name, dt):
        self.t[name] = self.t.get(name, 0.0) + dt

    # ---------------------------------------------------------------- gNB
    def transmit(self, payload):
        ref, cod, A, Z, K, F, Cn, Qm = self.ref, self.ref.cod, self.A, self.Z, self.K, self.F, self.C_, self.Qm
        a = np.zeros(A // 8 + 8, np.uint8); a[:A // 8] = payload
        t0 = time.perf_counter()
        crc = int(cod.crc24a(ob._ptr(a, _u8p), A)) >> 8
        a[A // 8:A // 8 + 3] = [(crc >> 16) & 255, (crc >> 8) & 255, crc & 255]
        self._tic("tb_crc", time.perf_counter() - t0)
        # nr_segmentation (bytes + CRC24B)
        segs = np.zeros((Cn, K // 8 + 64), dtype=np.uint8)
        ptrs = (_u8p * Cn)(*[C.cast(segs[r].ctypes.data, _u8p) for r in range(Cn)])
        fn = cod.nr_segmentation
        fn.argtypes = [_u8p, C.POINTER(_u8p), C.c_uint, C.POINTER(C.c_uint), C.POINTER(C.c_uint), C.POINTER(C.c_uint), C.POINTER(C.c_uint), C.c_uint8]
        fn.restype = C.c_int32
        c1, c2, c3, c4 = C.c_uint(), C.c_uint(), C.c_uint(), C.c_uint()
        t0 = time.perf_counter()
        fn(ob._ptr(a, _u8p), ptrs, A + 24, C.byref(c1), C.byref(c2), C.byref(c3), C.byref(c4), 1)
        self._tic("segmentation", time.perf_counter() - t0)
        assert (c1.value, c2.value, c3.value, c4.value) == (Cn, K, Z, F)
        # LDPCencoder, 8 segments per call (nr_dlsch_coding.c:167-171)
        nout = 66 * Z
        outs = [_aligned(68 * 384, np.uint8, 64) for _ in range(Cn)]
        oup = (_u8p * Cn)(*[C.cast(o.ctypes.data, _u8p) for o in outs])
        ip = ob.EncParams()
        ip.n_segments, ip.Kb, ip.Zc, ip.BG, ip.K, ip.gen_code = Cn, 22, Z, 1, K, 0
        t0 = time.perf_counter()
        for m in range((Cn + 7) // 8):
            ip.macro_num = m
            assert ref.enc.LDPCencoder(ptrs, oup, C.byref(ip)) == 0
        self._tic("ldpc_encode", time.perf_counter() - t0)
        # rate matching + interleaving per segment (nr_dlsch_coding.c:175-260)
        rm = cod.nr_rate_matching_ldpc
        rm.argtypes = [C.c_uint32, C.c_uint8, C.c_uint16, _u8p, _u8p, C.c_uint8, C.c_uint32, C.c_uint32, C.c_uint8, C.c_uint32]
        il = cod.nr_interleaving_ldpc
        il.argtypes = [C.c_uint32, C.c_uint8, _u8p, _u8p]; il.restype = None
        f = np.zeros(self.G + 64, np.uint8)
        e = np.zeros(max(self.E) + 64, np.uint8)
        off = 0
        dt_rm = 0.0
        for r in range(Cn):
            d = outs[r]
            d[K - F - 2 * Z:K - 2 * Z] = 2                                   # NR_NULL filler marks (nr_dlsch_coding.c)
            t0 = time.perf_counter()
            rc = rm(0, 1, Z, ob._ptr(d, _u8p), ob._ptr(e, _u8p), Cn, F, K - F - 2 * Z, 0, self.E[r])
            il(self.E[r], Qm, ob._ptr(e, _u8p), C.cast(f.ctypes.data + off, _u8p))
            dt_rm += time.perf_counter() - t0
            assert rc == 0
            off += self.E[r]
        self._tic("rate_match+interleave", dt_rm)
        self.f = f[:self.G].copy()
        # nr_generate_pdsch after the encoder (timed inside the harness)
        txF = ref.pdsch_tx_slot(self.txP, self.f, self.carrier_rb)
        ref._pdschtxlib.refh_pdschtx_last_seconds.restype = C.c_double
        self._tic("scramble..precoding", ref._pdschtxlib.refh_pdschtx_last_seconds())
        self.txF = txF
        # nr_feptx0: rotation + IDFT + CP per antenna
        prefix, start = self.P.slot_geometry(self.slot)
        out_len = start[13] + prefix[13] + self.N
        txdata = np.zeros((self.nb, 2 * out_len), np.int16)
        L = ref._ofdm()
        L.refh_ofdm_last_seconds.restype = C.c_double
        for ant in range(self.nb):
            txdata[ant], _ = ref.ofdm_tx_slot(self.N, self.mu, self.carrier_rb, self.slot, 14, self.rot224, txF[ant].reshape(-1), out_len)
            self._tic("ofdm_mod", L.refh_ofdm_last_seconds())
        return txdata

    # ---------------------------------------------------------------- the simulator's channel (not timed, not on the path)
    def channel(self, txdata, seed=1, snr_db=35.0, gain=3.0, coupling=0.15):
        rng = np.random.default_rng(seed)
        nb = self.nb
        x = txdata.reshape(nb, -1, 2).astype(np.float64)
        x = x[..., 0] + 1j * x[..., 1]
        ph = rng.uniform(0, 2 * np.pi, (nb, nb))
        H = (coupling + (1.0 - coupling) * np.eye(nb)) * np.exp(1j * ph) * gain
        y = H @ x
        sig = np.sqrt(np.mean(np.abs(y) ** 2)) * 10.0 ** (-snr_db / 20.0) * 0.70711
        y = y + sig * (rng.standard_normal(y.shape) + 1j * rng.standard_normal(y.shape))
        frame = np.zeros((nb, self.P.samples_per_frame, 2), np.int16)
        ss = self.P.slot_timestamp(self.slot)
        frame[:, ss:ss + y.shape[1], 0] = np.clip(np.round(y.real), -32768, 32767)
        frame[:, ss:ss + y.shape[1], 1] = np.clip(np.round(y.imag), -32768, 32767)
        return frame

    # ---------------------------------------------------------------- UE
    def receive(self, frame):
        ref, cod, N, nb, nl, Qm, Z, K, F, Cn = self.ref, self.ref.cod, self.N, self.nb, self.nl, self.Qm, self.Z, self.K, self.F, self.C_
        rxF = ref.ue_slot_fep(N, self.mu, self.carrier_rb, nb, self.slot, 8, self.rot224, self.ts, frame.reshape(nb, -1)).reshape(nb, 14, N, 2)
        ref._uechestlib.refh_uechest_last_seconds.restype = C.c_double
        self._tic("ofdm_demod", ref._uechestlib.refh_uechest_last_seconds())
        est = np.zeros((nl * nb, 14, N, 2), np.int16)
        for p in range(nl):
            CP = ob.ChestParms(N, nb, self.slot, 2, p, self.rb_start, 0, self.rb_size, self.P.first_carrier_offset, 0, self.dmrs_id)
            est[p * nb:(p + 1) * nb] = ref.pdsch_channel_estimation(CP, rxF, self.carrier_rb)
            self._tic("channel_estimation", ref._uechestlib.refh_uechest_last_seconds())
        PP = ob.PuschParms(N, nb, self.rb_start, 0, self.rb_size, self.P.first_carrier_offset, Qm, 1 << 2, 0, 2)
        llr, shift, _ = ref.pdsch_rx_slot(PP, self.start_symbol, self.nr_symbols, rxF, est, self.G, nl=nl)
        ref._pdschlib.refh_pdsch_last_seconds.restype = C.c_double
        self._tic("nr_rx_pdsch", ref._pdschlib.refh_pdsch_last_seconds())
        self.rxF, self.est, self.shift = rxF, est, shift
        # nr_dlsch_unscrambling
        Lm = ref._mod()
        v = _aligned(self.G + 64, np.int16, 32)
        v[:self.G] = llr
        t0 = time.perf_counter()
        Lm.nr_codeword_unscrambling(v.ctypes.data_as(C.c_void_p), C.c_uint32(self.G), C.c_uint8(0), C.c_uint32(self.nid), C.c_uint32(self.rnti))
        self._tic("unscrambling", time.perf_counter() - t0)
        self.llr = v[:self.G].copy()
        # nr_dlsch_decoding per segment: de-interleave, rate recovery, int8 packing, LDPCdecoder with CRC24B stop
        di = cod.nr_deinterleaving_ldpc
        di.argtypes = [C.c_uint32, C.c_uint8, _i16p, _i16p]; di.restype = None
        rr = cod.nr_rate_matching_ldpc_rx
        rr.argtypes = [C.c_uint32, C.c_uint8, C.c_uint16, _i16p, _i16p, C.c_uint8, C.c_uint8, C.c_uint8, C.c_uint32, C.c_uint32, C.c_uint32]
        p = ob.DecParams()
        p.BG, p.Z, p.R, p.numMaxIter, p.outMode, p.E, p.crc_type = 1, Z, self.R, self.max_iter, 0, K - F, 1
        p.check_crc = C.cast(cod.check_crc, C.c_void_p).value
        ab = ob.DecodeAbort()
        prof = ob.LdpcTimeStats()
        e16 = np.zeros(max(self.E) + 64, np.int16)
        w = np.zeros(66 * Z + 64, np.int16)
        zin = _aligned(27000, np.int8, 64)
        zout = _aligned(27000, np.int8, 64)
        off, its, out = 0, [], []
        nbytes = (self.seg["Kprime"] - self.seg["L"]) // 8
        for r in range(Cn):
            E = self.E[r]
            seg_llr = v[off:off + E]
            t0 = time.perf_counter()
            di(E, Qm, ob._ptr(e16, _i16p), C.cast(seg_llr.ctypes.data, _i16p))
            rc = rr(0, 1, Z, ob._ptr(w, _i16p), ob._ptr(e16, _i16p), Cn, 0, 1, E, F, K - F - 2 * Z)
            self._tic("deinterleave+rate_recovery", time.perf_counter() - t0)
            assert rc == 0
            t0 = time.perf_counter()
            z = np.zeros(68 * Z, np.int16)                                    # the packing glue of nr_dlsch_decoding.c:235-250 (restated with numpy)
            z[2 * Z:K - F] = w[:K - F - 2 * Z]; z[K - F:K] = 127; z[K:] = w[K - 2 * Z:66 * Z]
            zin[:68 * Z] = np.clip(z, -128, 127)
            self._tic("llr_packing (numpy)", time.perf_counter() - t0)
            ab.failed = False
            t0 = time.perf_counter()
            it = ref.dec.LDPCdecoder(C.byref(p), 0, 0, 0, ob._ptr(zin, _i8p), ob._ptr(zout, _i8p), C.cast(C.byref(prof), C.c_void_p), C.cast(C.byref(ab), C.c_void_p))
            self._tic("ldpc_decode", time.perf_counter() - t0)
            its.append(it); out.append(zout[:nbytes].view(np.uint8).copy())
            off += E
        tb = np.concatenate(out)
        t0 = time.perf_counter()
        a = np.zeros(tb.size + 8, np.uint8); a[:tb.size] = tb
        crc = int(cod.crc24a(ob._ptr(a, _u8p), self.A + 24))
        self._tic("tb_crc", time.perf_counter() - t0)
        return tb, np.array(its), crc


def time_slot(seconds=15.0, **kw):
    """Run whole slots (transmit + receive) on one host thread for about `seconds`.  Returns dict(slots_per_s, slots, seconds, stages_us, decoded_ok, mean_iterations)."""
    ch = RefDlSlot(**kw)
    payload = np.random.default_rng(5).integers(0, 256, size=ch.A // 8, dtype=np.uint8)
    tx = ch.transmit(payload)
    frame = ch.channel(tx, seed=3)
    ch.t = {}
    n, ok, its = 0, True, []
    t_end = time.perf_counter() + seconds
    while n == 0 or time.perf_counter() < t_end:
        ch.transmit(payload)
        tb, it, crc = ch.receive(frame)
        ok = ok and crc == 0 and np.array_equal(tb[:payload.size], payload)
        its.append(float(np.mean(it)))
        n += 1
    total = sum(v for k, v in ch.t.items() if "numpy" not in k)        # the numpy restatement of the caller's packing loop is not reference code: not counted
    return {"slots_per_s": n / total, "slots": n, "seconds_in_reference_code": total, "stages_us": {k: 1e6 * v / n for k, v in ch.t.items()}, "decoded_ok": bool(ok),
            "mean_iterations": float(np.mean(its))}


if __name__ == "__main__":
    import json
    import sys
    print(json.dumps(time_slot(float(sys.argv[1]) if len(sys.argv) > 1 else 10.0)))
