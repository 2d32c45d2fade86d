"""Drop-in by symbol interposition (channel estimation has no plug-in boundary in OAI, so the build creates one).

integration/oai_shim_pusch_chest.c defines OAI's own `nr_pusch_channel_estimation` -- same prototype, compiled against OAI's headers -- and forwards to
libldpc_b200.so.  Here the reference-side CALLER (oracle/ref_harness_chest.c: it fills PHY_VARS_gNB / nfapi_nr_pusch_pdu_t the way nr_rx_pusch_tp's
callers do and calls the function by name) is linked against that interposer instead of the reference's nr_ul_channel_estimation.c
(integration/build_shims.sh -> oracle/_ref/libshimtest_chest.so, prebuilt where /root/reference exists, travels to the GPU box).  What the unchanged host C
gets back through OAI's own structures -- ul_ch_estimates, *max_ch, *nvar, delay_t -- must be what the pinned oracle computes."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle.bindings import ChestParms

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIMTEST = os.path.join(ROOT, "oracle", "_ref", "libshimtest_chest.so")

CASES = [  # N, nb_rx, slot, symbol, port, rb_start, rb_size, carrier PRBs, scid, dmrs id, dmrs_type, chest_freq
    (4096, 4, 1, 2, 0, 0, 273, 273, 0, 77, 0, 0), (2048, 2, 7, 3, 1, 10, 50, 106, 1, 1007, 0, 0), (1024, 2, 6, 11, 2, 20, 32, 52, 0, 300, 0, 0),
    (1024, 3, 0, 11, 2, 20, 32, 52, 0, 300, 1, 0), (2048, 2, 9, 3, 1, 10, 50, 106, 1, 1007, 0, 1), (512, 2, 8, 4, 1, 0, 25, 25, 1, 303, 1, 1),
]


def test_oai_caller_reaches_the_gpu_through_the_interposed_symbol(oracle):
    if not os.path.exists(SHIMTEST):
        pytest.fail(f"{SHIMTEST} missing: run integration/build_shims.sh where /root/reference exists (the file travels with the repo snapshot)")
    lib = C.CDLL(SHIMTEST)
    assert lib.refh_chest_init(os.path.join(ROOT, "oracle", "_ref", "libref_dfts.so").encode()) == 0
    rng = np.random.default_rng(91)
    for N, nb_rx, slot, symbol, port, rb_start, rb_size, carrier, scid, nid, dmrs_type, chest_freq in CASES:
        fco = N - carrier * 6
        P = ChestParms(N, nb_rx, slot, symbol, port, rb_start, 0, rb_size, fco, scid, nid, dmrs_type, chest_freq)
        rx = rng.integers(-3000, 3001, size=(nb_rx, 14, N, 2)).astype(np.int16)
        prm = np.array([N, nb_rx, carrier, slot, symbol, port, rb_start, 0, rb_size, fco, scid, nid, dmrs_type, chest_freq], dtype=np.int32)
        est = np.zeros(nb_rx * 14 * N * 2, np.int16)
        out = np.zeros(5, np.int32)
        pil = np.zeros(2 * 6 * rb_size, np.int16)
        assert lib.refh_pusch_chest(prm.ctypes.data_as(C.c_void_p), rx.ctypes.data_as(C.c_void_p), est.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p),
                                    pil.ctypes.data_as(C.c_void_p)) == 0
        est_o, out_o = oracle.pusch_channel_estimation(P, rx)
        assert np.array_equal(out, out_o), (N, nb_rx, slot, symbol, port, dmrs_type, chest_freq, out, out_o)
        assert np.array_equal(est.reshape(nb_rx, 14, N, 2)[:, symbol], est_o[:, symbol]), (N, nb_rx, slot, symbol, port, dmrs_type, chest_freq)


@pytest.mark.parametrize("so", ["libldpc_b200.so"])
def test_reference_loader_side_with_oai_types(oracle, so):
    """The LDPC plug-in boundary from OAI's side: oracle/ref_harness_loader.c is compiled against OAI's own nrLDPC_defs.h / nrLDPC_types.h, looks the four
    symbols up like load_LDPClib (dlopen RTLD_LAZY | RTLD_NODELETE | RTLD_GLOBAL, dlclose after the lookup), and calls them like ldpctest -- LDPCinit again
    before every segment, one blocking LDPCdecoder call per segment, encoder in groups of 8 segments via macro_num.  Code words, decoded bits and returned
    iteration counts must be the oracle's."""
    from common import make_case
    path = os.path.join(ROOT, "oracle", "_ref", "libref_loader.so")
    if not os.path.exists(path):
        pytest.fail(f"{path} missing: run oracle/build_ref.sh where /root/reference exists")
    h = C.CDLL(path)
    assert h.refh_loader_open(os.path.join(ROOT, "openairinterface5g_b200", so).encode()) == 0
    for BG, Z, R, n_seg, ebn0 in ((1, 384, 13, 11, 2.3), (2, 96, 15, 3, 1.5), (1, 176, 23, 9, 4.5)):
        K, P, llr = make_case(oracle, BG, Z, R, n_seg, ebn0, seed=Z + n_seg)
        Kb = 22 if BG == 1 else 10
        nout = (66 if BG == 1 else 50) * Z
        cw = np.zeros((n_seg, nout), np.uint8)
        pay = np.ascontiguousarray(P, dtype=np.uint8)
        assert h.refh_loader_encode(BG, Z, Kb, K, n_seg, pay.ctypes.data_as(C.c_void_p), cw.ctypes.data_as(C.c_void_p)) == 0
        for j in range(n_seg):
            assert np.array_equal(cw[j], oracle.encode(BG, Z, K, P[j])), (BG, Z, j)
        ncol = llr.shape[1] // Z
        out = np.zeros((n_seg, ncol * Z // 8), np.uint8)
        iters = np.zeros(n_seg, np.int32)
        llr = np.ascontiguousarray(llr, dtype=np.int8)
        h.refh_loader_decode(BG, Z, R, 8, K, n_seg, llr.shape[1], llr.ctypes.data_as(C.c_void_p), out.shape[1], out.ctypes.data_as(C.c_void_p),
                             iters.ctypes.data_as(C.c_void_p))
        for j in range(n_seg):
            it_o, out_o = oracle.decode(BG, Z, R, 8, llr[j], 0)
            assert iters[j] == it_o and np.array_equal(out[j], np.asarray(out_o).view(np.uint8)), (BG, Z, R, j, iters[j], it_o)
    # LDPCshutdown (free_LDPClib) is exercised in its own process below: this one shares the library instance with the other tests' fixtures


def test_reference_loader_lifecycle_in_its_own_process():
    """load -> decode -> free_LDPClib -> load again -> decode, as two physim runs in one process would (ldpctest.c:503-504, dlsim.c:1306)."""
    import subprocess
    import sys
    code = (
        "import ctypes as C, numpy as np, os, sys; sys.path.insert(0, 'tests'); from common import make_case; from oracle.bindings import Oracle\n"
        "orc = Oracle(); h = C.CDLL('oracle/_ref/libref_loader.so'); so = os.path.abspath('openairinterface5g_b200/libldpc_b200.so').encode()\n"
        "K, P, llr = make_case(orc, 1, 384, 13, 2, 2.4, 5); llr = np.ascontiguousarray(llr, dtype=np.int8)\n"
        "for rnd in range(2):\n"
        "    assert h.refh_loader_open(so) == 0\n"
        "    out = np.zeros((2, 68 * 384 // 8), np.uint8); it = np.zeros(2, np.int32)\n"
        "    h.refh_loader_decode(1, 384, 13, 8, K, 2, llr.shape[1], llr.ctypes.data_as(C.c_void_p), out.shape[1], out.ctypes.data_as(C.c_void_p), it.ctypes.data_as(C.c_void_p))\n"
        "    for j in range(2):\n"
        "        io, oo = orc.decode(1, 384, 13, 8, llr[j], 0)\n"
        "        assert it[j] == io and np.array_equal(out[j], np.asarray(oo).view(np.uint8))\n"
        "    assert h.refh_loader_close() == 0\n"
        "print('LIFECYCLE_OK')\n")
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert "LIFECYCLE_OK" in r.stdout, r.stdout[-500:] + r.stderr[-1500:]


def test_ue_caller_reaches_the_gpu_through_the_interposed_symbol(oracle):
    """The UE-side twin: oracle/ref_harness_uechest.c (the caller, as nr_ue_pdsch_procedures passes the arguments) linked against
    integration/oai_shim_pdsch_chest.c instead of nr_dl_channel_estimation.c -- DMRS types 1 / 2, chest_freq 0 / 1."""
    path = os.path.join(ROOT, "oracle", "_ref", "libshimtest_uechest.so")
    if not os.path.exists(path):
        pytest.fail(f"{path} missing: run integration/build_shims.sh where /root/reference exists")
    lib = C.CDLL(path)
    assert lib.refh_uechest_init(os.path.join(ROOT, "oracle", "_ref", "libref_dfts.so").encode()) == 0
    rng = np.random.default_rng(92)
    for N, nb_rx, slot, symbol, port, rb_start, rb_size, carrier, scid, nid, dmrs_type, chest_freq in (
            (4096, 2, 1, 2, 0, 0, 273, 273, 0, 77, 0, 0), (2048, 2, 7, 3, 1, 10, 50, 106, 1, 1007, 0, 0), (1024, 4, 6, 11, 2, 20, 32, 52, 0, 300, 1, 0),
            (1024, 2, 3, 4, 4, 0, 52, 52, 0, 21, 1, 0), (2048, 2, 9, 3, 1, 10, 50, 106, 1, 1007, 0, 1), (512, 2, 8, 4, 5, 0, 25, 25, 1, 303, 1, 1)):
        fco = N - carrier * 6
        P = ChestParms(N, nb_rx, slot, symbol, port, rb_start, 0, rb_size, fco, scid, nid, dmrs_type, chest_freq)
        rx = rng.integers(-3000, 3001, size=(nb_rx, 14, N, 2)).astype(np.int16)
        prm = np.array([N, nb_rx, carrier, slot, symbol, port, rb_start, 0, rb_size, fco, scid, nid, dmrs_type, chest_freq], dtype=np.int32)
        est = np.zeros(nb_rx * 14 * N * 2, np.int16)
        assert lib.refh_pdsch_chest(prm.ctypes.data_as(C.c_void_p), rx.ctypes.data_as(C.c_void_p), est.ctypes.data_as(C.c_void_p)) == 0
        est_o = oracle.pdsch_channel_estimation(P, rx)
        assert np.array_equal(est.reshape(nb_rx, 14, N, 2)[:, symbol], est_o[:, symbol]), (N, nb_rx, slot, symbol, port, dmrs_type, chest_freq)


def test_oai_ue_caller_reaches_the_gpu_through_nr_rx_pdsch(oracle):
    """integration/oai_shim_rx_pdsch.c defines OAI's `nr_rx_pdsch`; the reference-side caller (oracle/ref_harness_pdsch.c: fills PHY_VARS_NR_UE / NR_UE_DLSCH_t and
    calls the function symbol by symbol like nr_ue_pdsch_procedures) is linked against it instead of nr_dlsch_demodulation.c (oracle/_ref/libshimtest_pdsch.so).
    LLRs, log2_maxh and dl_valid_re that the unchanged host C gets back must be the pinned oracle's, one to four layers."""
    from oracle.bindings import PuschParms
    so = os.path.join(ROOT, "oracle", "_ref", "libshimtest_pdsch.so")
    if not os.path.exists(so):
        pytest.fail(f"{so} missing: run integration/build_shims.sh where /root/reference exists (the file travels with the repo snapshot)")
    lib = C.CDLL(so)
    rng = np.random.default_rng(17)
    cases = [  # N, nb_rx, rb_start, rb_size, Qm, dmrs_pos, dmrs_type, cdm groups, carrier PRBs, start_symbol, nr_symbols, layers, amplitudes
        (4096, 2, 0, 273, 6, 1 << 2, 0, 2, 273, 1, 13, 1, (2000, 1500)), (2048, 1, 10, 50, 4, (1 << 2) | (1 << 11), 0, 1, 106, 1, 13, 1, (2000, 1500)),
        (1024, 4, 0, 52, 6, 1 << 2, 1, 1, 52, 1, 13, 1, (2000, 1500)), (4096, 2, 0, 273, 6, 1 << 2, 0, 1, 273, 1, 13, 2, (2000, 1500)),
        (2048, 2, 30, 76, 2, 1 << 3, 0, 2, 106, 2, 10, 2, (300, 200)), (1024, 2, 20, 32, 4, 1 << 2, 1, 2, 52, 2, 12, 2, (12000, 9000)),
        (2048, 4, 10, 50, 4, (1 << 2) | (1 << 11), 0, 1, 106, 1, 13, 3, (4000, 6000)), (1024, 4, 0, 52, 6, 1 << 2, 1, 1, 52, 1, 13, 4, (2000, 1500))]
    for N, nb_rx, rb_start, rb_size, Qm, dpos, dtype_, cdm, carrier, start, nsym, nl, (ay, ah) in cases:
        rx = rng.integers(-ay, ay + 1, size=(nb_rx, 14, N, 2)).astype(np.int16)
        h = rng.integers(-ah, ah + 1, size=(nl * nb_rx, 14, N, 2)).astype(np.int16)
        P = PuschParms(N, nb_rx, rb_start, 0, rb_size, N - carrier * 6, Qm, dpos, dtype_, cdm)
        llr_o, sh_o = oracle.pdsch_rx_slot(P, start, nsym, rx, h, nl=nl)
        G = llr_o.size
        prm = np.array([N, nb_rx, rb_start, 0, rb_size, N - carrier * 6, Qm, start, nsym, dpos, dtype_, cdm, G, nl], dtype=np.int32)
        llr = np.zeros(G + 64, np.int16)
        valid = np.zeros(14, np.int32)
        sh = lib.refh_pdsch_rx_slot(prm.ctypes.data_as(C.c_void_p), rx.ctypes.data_as(C.c_void_p), h.ctypes.data_as(C.c_void_p), llr.ctypes.data_as(C.c_void_p),
                                    valid.ctypes.data_as(C.c_void_p), None)
        assert sh == sh_o, (N, nb_rx, Qm, nl, sh, sh_o)
        assert np.array_equal(llr[:G], llr_o), (N, nb_rx, rb_size, Qm, nl, np.nonzero(llr[:G] != llr_o)[0][:5])
        per = [(rb_size * ((12 - 6 * cdm) if dtype_ == 0 else (12 - 4 * cdm)) if (dpos >> s) & 1 else rb_size * 12) for s in range(start, start + nsym)]
        assert [int(v) for v in valid[start:start + nsym]] == per and int(valid.sum()) * Qm * nl == G


def test_oai_ue_caller_with_ptrs_reaches_the_gpu_through_nr_rx_pdsch(oracle):
    """The same caller with PT-RS switched on in the PDU (pduBitmap bit 0, C-RNTI, PTRSTimeDensity / PTRSFreqDensity / PTRSReOffset): the interposer keeps
    dl_valid_re / ptrs_re_per_slot / dlsch->ptrs_symbols per symbol like nr_pdsch_ptrs_processing and the library estimates, interpolates and rotates inside the
    slot receiver.  LLRs, log2_maxh, dl_valid_re and the PT-RS RE counts must be the pinned oracle's."""
    from oracle.bindings import PuschParms, PtrsParms
    from common import PTRS_CASES, ptrs_inputs
    so = os.path.join(ROOT, "oracle", "_ref", "libshimtest_pdsch.so")
    if not os.path.exists(so):
        pytest.fail(f"{so} missing: run integration/build_shims.sh where /root/reference exists (the file travels with the repo snapshot)")
    lib = C.CDLL(so)
    rng = np.random.default_rng(19)
    for case in (PTRS_CASES[0], PTRS_CASES[2], PTRS_CASES[5], PTRS_CASES[8]):
        N, nb_rx, rb_start, rb_size, Qm, dpos, dtype_, cdm, carrier, start, nsym, L, K, reoff, rnti, slot, nscid, nid = case
        for kind, a, b in (("random", 2000, 1500), ("coherent", 30, 0.05)):
            rx, h = ptrs_inputs(oracle, rng, case, kind, a, b)
            P = PuschParms(N, nb_rx, rb_start, 0, rb_size, N - carrier * 6, Qm, dpos, dtype_, cdm)
            llr_o, sh_o, ph_o, nre_o = oracle.pdsch_rx_slot_ptrs(P, PtrsParms(1, L, K, reoff, rnti, slot, nscid, nid), start, nsym, rx, h)
            G = llr_o.size
            q = np.array([1, L, K, reoff, rnti, slot, nscid, nid, carrier], dtype=np.int32)
            lib.refh_pdsch_set_ptrs(q.ctypes.data_as(C.c_void_p))
            try:
                prm = np.array([N, nb_rx, rb_start, 0, rb_size, N - carrier * 6, Qm, start, nsym, dpos, dtype_, cdm, G, 1], dtype=np.int32)
                llr = np.zeros(G + 64, np.int16)
                valid = np.zeros(14, np.int32)
                sh = lib.refh_pdsch_rx_slot(prm.ctypes.data_as(C.c_void_p), rx.ctypes.data_as(C.c_void_p), h.ctypes.data_as(C.c_void_p), llr.ctypes.data_as(C.c_void_p),
                                            valid.ctypes.data_as(C.c_void_p), None)
                ph = np.zeros((14, 2), np.int16); nre = np.zeros(14, np.int32)
                lib.refh_pdsch_get_ptrs(ph.ctypes.data_as(C.c_void_p), nre.ctypes.data_as(C.c_void_p))
            finally:
                lib.refh_pdsch_set_ptrs(None)
            assert sh == sh_o and np.array_equal(nre, nre_o), (case, kind, sh, sh_o, nre, nre_o)
            assert int(valid.sum()) * Qm == G and np.array_equal(llr[:G], llr_o), (case, kind, np.nonzero(llr[:G] != llr_o)[0][:5])


def test_oai_gnb_caller_reaches_the_gpu_through_nr_rx_pusch_tp(oracle):
    """integration/oai_shim_rx_pusch.c defines OAI's `nr_rx_pusch_tp`; the reference-side caller (oracle/ref_harness_rxpusch.c: PHY_VARS_gNB with the rxdataF ring,
    pusch_vars and the ULSCH PDU as phy_init_nr_gNB / the scheduler leave them) is linked against it and against the interposed channel estimator
    (oracle/_ref/libshimtest_rxpusch.so).  What the unchanged host C reads afterwards -- pusch_vars->llr (layer de-mapped, unscrambled), ul_ch_estimates, log2_maxh,
    dmrs_symbol, ul_valid_re_per_slot, llr_offset -- must be what the pinned oracle functions give when chained the way the reference chains them
    (nr_ulsch_demodulation.c:1447-1700), one layer and two (MMSE with the estimator's max_ch / nvar)."""
    from oracle.bindings import ChestParms, PuschParms
    so = os.path.join(ROOT, "oracle", "_ref", "libshimtest_rxpusch.so")
    if not os.path.exists(so):
        pytest.fail(f"{so} missing: run integration/build_shims.sh where /root/reference exists (the file travels with the repo snapshot)")
    lib = C.CDLL(so)
    rng = np.random.default_rng(23)
    cases = [  # N, nb_rx, carrier PRBs, slot, rb_start, rb_size, Qm, dmrs_pos, cdm groups, layers, dmrs id, rnti, data scrambling id
        (4096, 4, 273, 1, 0, 273, 6, 1 << 2, 2, 1, 55, 0x1234, 77), (2048, 2, 106, 7, 20, 50, 4, 1 << 2, 2, 1, 1007, 0x4321, 99),
        (1024, 2, 52, 6, 3, 32, 2, (1 << 2) | (1 << 11), 1, 1, 300, 0x1001, 5),
        (4096, 4, 273, 1, 0, 273, 6, 1 << 2, 2, 2, 55, 0x1234, 77), (2048, 2, 106, 9, 10, 50, 8, 1 << 3, 2, 2, 1007, 0x2222, 512)]
    for N, nb_rx, carrier, slot, rb_start, rb_size, Qm, dpos, cdm, nl, dmrs_id, rnti, nid in cases:
        fco = N - carrier * 6
        rx = rng.integers(-1500, 1501, size=(nb_rx, 14, N, 2)).astype(np.int16)
        P = PuschParms(N, nb_rx, rb_start, 0, rb_size, fco, Qm, dpos, 0, cdm)
        dms = [s for s in range(14) if (dpos >> s) & 1]
        # ---- expected: the oracle functions chained like the reference chains them
        est = np.zeros((nl * nb_rx, 14, N, 2), np.int16)
        max_ch, nvar = 0, 0
        for s in dms:
            for p in range(nl):
                e, st = oracle.pusch_channel_estimation(ChestParms(N, nb_rx, slot, s, p, rb_start, 0, rb_size, fco, 0, dmrs_id, 0, 0), rx)
                est[p * nb_rx:(p + 1) * nb_rx, s] = e[:, s]
                max_ch = max(max_ch, int(st[0])); nvar += int(st[1])
        nvar //= 14 * nl * nb_rx
        valid = [oracle.pusch_nb_re(P, s) for s in range(14)]
        meas = [s for s in range(14) if valid[s] > 0][0]
        if nl == 1:
            sh_o, _ = oracle.pusch_log2_maxh(P, meas, dms[0], rx, est)
        else:
            sh_o, _ = oracle.pusch_log2_maxh_2l(P, meas, dms[0], max_ch, rx, est)
        cur, out = dms[0], []
        for s in range(14):
            if (dpos >> s) & 1:
                cur = s
            if valid[s] == 0:
                continue
            if nl == 1:
                out.append(oracle.pusch_inner_rx_symbol(P, s, cur, sh_o, rx, est)[0])
            else:
                l2, _ = oracle.pusch_inner_rx_symbol_2l(P, s, cur, sh_o, nvar, rx, est)
                out.append(np.stack([l2[0].reshape(valid[s], Qm), l2[1].reshape(valid[s], Qm)], axis=1).reshape(-1))
        want = oracle.unscramble_llr(np.concatenate(out), 0, nid, rnti)
        G = want.size
        # ---- the unchanged caller
        prm = np.array([N, nb_rx, carrier, slot, rb_start, 0, rb_size, fco, Qm, 0, 14, dpos, 0, cdm, nl, (1 << nl) - 1, 0, dmrs_id, rnti, nid, 0, 0], dtype=np.int32)
        llr = np.zeros(G, np.int16)
        est_out = np.zeros((nl * nb_rx, 14, N, 2), np.int16)
        info = np.zeros(40, np.int32)
        assert lib.refh_rx_pusch(prm.ctypes.data_as(C.c_void_p), rx.ctypes.data_as(C.c_void_p), G, llr.ctypes.data_as(C.c_void_p), est_out.ctypes.data_as(C.c_void_p),
                                 info.ctypes.data_as(C.c_void_p)) == 0
        for s in dms:
            assert np.array_equal(est_out[:, s], est[:, s]), (N, nl, "estimates", s)
        assert info[0] == sh_o and info[1] == dms[0], (N, nl, info[:2], sh_o)
        assert [int(v) for v in info[2:16]] == valid
        assert [int(v) for v in info[16:30]] == [int(x) for x in np.concatenate([[0], np.cumsum(np.array(valid) * Qm)[:-1]])]
        assert np.array_equal(llr, want), (N, nb_rx, Qm, nl, np.nonzero(llr != want)[0][:5])
        assert all(int(v) > 0 for v in info[30:30 + nb_rx])


def test_oai_gnb_caller_with_transform_precoding_through_nr_rx_pusch_tp(oracle):
    """The same caller with pusch_pdu->transform_precoding = enabled (DFT-s-OFDM): the interposed estimator takes OAI's own low-PAPR sequence table
    (gNB_dmrs_lowpaprtype1_sequence[u][v][index], built by the caller like nr_init.c:249 does) and the interposed receiver runs equalisation + nr_idft inside
    the library call.  Expected values: the pinned oracle chained like nr_rx_pusch_tp chains the reference functions."""
    from oracle.bindings import ChestParms, PuschParms
    so = os.path.join(ROOT, "oracle", "_ref", "libshimtest_rxpusch.so")
    if not os.path.exists(so):
        pytest.fail(f"{so} missing: run integration/build_shims.sh where /root/reference exists (the file travels with the repo snapshot)")
    lib = C.CDLL(so)
    gold = np.load(os.path.join(ROOT, "tests", "golden", "transform_precoding.npz"))
    rng = np.random.default_rng(37)
    for N, nb_rx, carrier, slot, rb_start, rb_size, Qm, u, dmrs_id, rnti, nid in ((4096, 4, 273, 1, 0, 270, 6, 0, 55, 0x1234, 77), (2048, 2, 106, 7, 20, 50, 4, 17, 1007, 0x4321, 99),
                                                                               (1024, 2, 52, 6, 3, 2, 2, 29, 300, 0x1001, 5), (1024, 1, 52, 2, 0, 1, 6, 3, 9, 0x77, 1)):
        fco = N - carrier * 6
        dpos, cdm = 1 << 2, 2
        rx = rng.integers(-1500, 1501, size=(nb_rx, 14, N, 2)).astype(np.int16)
        P = PuschParms(N, nb_rx, rb_start, 0, rb_size, fco, Qm, dpos, 0, cdm)
        seq = oracle.lowpapr_seq(u, 0, 6 * rb_size)
        if seq is None:
            seq = gold[f"seq_{6 * rb_size}"][u].copy()
        oracle.chest_set_lowpapr(seq)
        oracle.pusch_set_transform_precoding(1)
        try:
            e, st = oracle.pusch_channel_estimation(ChestParms(N, nb_rx, slot, 2, 0, rb_start, 0, rb_size, fco, 0, dmrs_id, 0, 0), rx)
            est = np.zeros((nb_rx, 14, N, 2), np.int16)
            est[:, 2] = e[:, 2]
            sh_o, _ = oracle.pusch_log2_maxh(P, 0, 2, rx, est)
            out = [oracle.pusch_inner_rx_symbol(P, s, 2, sh_o, rx, est)[0] for s in range(14) if s != 2]
        finally:
            oracle.chest_set_lowpapr(None)
            oracle.pusch_set_transform_precoding(0)
        want = oracle.unscramble_llr(np.concatenate(out), 0, nid, rnti)
        G = want.size
        prm = np.array([N, nb_rx, carrier, slot, rb_start, 0, rb_size, fco, Qm, 0, 14, dpos, 0, cdm, 1, 1, 0, dmrs_id, rnti, nid, 0, 0], dtype=np.int32)
        llr = np.zeros(G, np.int16)
        est_out = np.zeros((nb_rx, 14, N, 2), np.int16)
        info = np.zeros(40, np.int32)
        lib.refh_rxpusch_set_transform_precoding(1, u, 0)
        try:
            assert lib.refh_rx_pusch(prm.ctypes.data_as(C.c_void_p), rx.ctypes.data_as(C.c_void_p), G, llr.ctypes.data_as(C.c_void_p), est_out.ctypes.data_as(C.c_void_p),
                                     info.ctypes.data_as(C.c_void_p)) == 0
        finally:
            lib.refh_rxpusch_set_transform_precoding(0, 0, 0)
        assert np.array_equal(est_out[:, 2], est[:, 2]), (N, rb_size, "estimates")
        assert info[0] == sh_o and info[1] == 2
        assert np.array_equal(llr, want), (N, nb_rx, Qm, rb_size, np.nonzero(llr != want)[0][:5])


def test_oai_ru_callers_reach_the_gpu_through_nr_feptx0_and_nr_fep_full(oracle):
    """integration/oai_shim_ru_ofdm.c defines OAI's RU front-end functions `nr_feptx0` (IDFT + cyclic prefix of a tx antenna's symbols) and `nr_fep_full` (the 14 DFTs
    of every rx antenna of a slot); the reference-side caller (oracle/ref_harness_ru.c: an RU_t with frame parameters and buffers as init_nr_ru leaves them) is linked
    against it (oracle/_ref/libshimtest_ru.so).  The samples / sub-carriers the unchanged host C finds in ru->common afterwards must be the pinned oracle's (which is
    pinned against the real PHY_ofdm_mod / nr_slot_fep_ul)."""
    so = os.path.join(ROOT, "oracle", "_ref", "libshimtest_ru.so")
    if not os.path.exists(so):
        pytest.fail(f"{so} missing: run integration/build_shims.sh where /root/reference exists (the file travels with the repo snapshot)")
    lib = C.CDLL(so)
    rng = np.random.default_rng(29)
    for N, mu, nb_rb, slot, nant, chunks, n_ta, divisor in ((4096, 1, 273, 1, 2, 1, 0, 8), (4096, 1, 273, 4, 2, 2, 800, 8), (2048, 1, 106, 19, 1, 1, 0, 8),
                                                             (1024, 0, 52, 3, 2, 2, 0, 8), (2048, 2, 66, 7, 1, 1, 400, 16)):
        spf = lib.refh_ru_fep_full(N, mu, nb_rb, slot, nant, divisor, n_ta, None, None)
        pre, cps, ss, _ = oracle.ofdm_geometry(N, mu, slot)
        # ---- transmit: every antenna's slot, in `chunks` nr_feptx0 calls per antenna
        F = rng.integers(-2000, 2001, size=(nant, 14 * N * 2)).astype(np.int16)
        txdata = np.zeros((nant, spf * 2), np.int16)
        assert lib.refh_ru_feptx(N, mu, nb_rb, slot, nant, chunks, F.ctypes.data_as(C.c_void_p), txdata.ctypes.data_as(C.c_void_p)) == spf
        for a in range(nant):
            t_o, _ = oracle.ofdm_tx_slot(N, mu, nb_rb, slot, 14, None, F[a])
            assert np.array_equal(txdata[a, 2 * ss:2 * ss + t_o.size], t_o), (N, mu, slot, a, "feptx0")
            assert not txdata[a, :2 * ss].any() and not txdata[a, 2 * ss + t_o.size:].any()
        # ---- receive: a frame of samples, timing advance offset (wraps the frame ring when the slot is early in the frame)
        rx = rng.integers(-3000, 3001, size=(nant, spf * 2)).astype(np.int16)
        rxF = np.zeros((nant, 14 * N * 2), np.int16)
        assert lib.refh_ru_fep_full(N, mu, nb_rb, slot, nant, divisor, n_ta, rx.ctypes.data_as(C.c_void_p), rxF.ctypes.data_as(C.c_void_p)) == spf
        for a in range(nant):
            assert np.array_equal(rxF[a], oracle.ofdm_rx_slot(N, mu, nb_rb, slot, divisor, n_ta, None, rx[a])), (N, mu, slot, a, "fep_full")


def _ulsch_lib(name, bind=None):
    so = os.path.join(ROOT, "oracle", "_ref", name)
    if not os.path.exists(so):
        pytest.fail(f"{so} missing: run oracle/build_ref.sh and integration/build_shims.sh where /root/reference exists (the files travel with the repo snapshot)")
    lib = C.CDLL(so)
    lib.refh_ulsch_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 5
    if bind:
        assert lib.refh_ulsch_bind_ldpc(os.path.join(ROOT, "oracle", "_ref", bind).encode()) == 0
    return lib


def _ulsch_call(lib, prm, llr, G, Cn, K, Z, tbs, ncb):
    inf = np.zeros(8, np.int32); it = np.zeros(Cn, np.int32); c = np.zeros(Cn * K // 8, np.uint8); tb = np.zeros(tbs + 3, np.uint8); d = np.zeros(Cn * ncb, np.int16)
    rc = lib.refh_ulsch_decode(prm.ctypes.data, llr.ctypes.data, G, inf.ctypes.data, it.ctypes.data, c.ctypes.data, tb.ctypes.data, d.ctypes.data)
    return rc, inf, it, c, tb, d


def test_oai_gnb_caller_reaches_the_gpu_through_nr_ulsch_decoding(oracle, monkeypatch):
    """integration/oai_shim_ulsch_decoding.c defines OAI's `nr_ulsch_decoding`; the reference-side caller (oracle/ref_harness_ulsch.c: gNB thread pool + response
    FIFO, a ULSCH from the reference's own new_gNB_ulsch, the collection loop of phy_procedures_gNB_uespec_RX with nr_postDecode's copy) is built twice: around
    the reference's own function with the compiled CPU decoder (libref_ulsch.so), and with the interposer linked ahead of it (libshimtest_ulsch.so), where the whole
    transport block is one library call.  Same LLRs into both: C / K / Z / F / llrLen, every segment's iteration count, harq_process->c[r], the assembled
    transport block and -- with NRB200_SHIM_MIRROR_HARQ=1 -- the combined soft buffers harq_process->d[r] must be identical, for a new transmission (rv 0), for
    a retransmission that combines (rv 2 after an undecodable rv 0), with fillers, BG2, a single segment and 2 layers x 56 segments."""
    from common import make_tb_llrs
    monkeypatch.setenv("NRB200_SHIM_MIRROR_HARQ", "1")
    ref = _ulsch_lib("libref_ulsch.so", "libref_ldpc_dec.so")
    shim = _ulsch_lib("libshimtest_ulsch.so")
    cases = [  # A bits, Qm, layers, PRBs, BG, snr of the first transmission (dB), N_RB_UL
        (33640, 6, 1, 52, 1, 6.0, 106), (235624, 6, 1, 273, 1, 10.0, 273), (471272, 6, 2, 273, 1, 10.0, 273), (3752, 2, 1, 20, 2, 4.0, 106), (1032, 4, 1, 4, 2, 6.0, 52),
        (33640, 6, 1, 52, 1, 0.0, 106)]
    for A, Qm, nl, rb, BG, snr, nrb in cases:
        pay, llr0, info = make_tb_llrs(oracle, A, Qm, nl, rb, 0, seed=A % 97, snr_db=snr, BG=BG)
        _, llr2, _ = make_tb_llrs(oracle, A, Qm, nl, rb, 2, seed=A % 97, snr_db=snr + 3.0, BG=BG)        # the same payload, redundancy version 2
        Cn, K, Z, G = info["C"], info["K"], info["Z"], info["G"]
        ncb = (66 if BG == 1 else 50) * Z
        for rnd, (rv, llr, new) in enumerate(((0, llr0, 1), (2, llr2, 0))):
            prm = np.array([nrb, rb, Qm, nl, A // 8, rv, BG, 0, 8, new, rnd, 2], np.int32)
            a = _ulsch_call(ref, prm, llr, G, Cn, K, Z, A // 8, ncb)
            b = _ulsch_call(shim, prm, llr, G, Cn, K, Z, A // 8, ncb)
            assert a[0] == b[0] == Cn, (A, rnd, a[0], b[0])
            assert np.array_equal(a[1][:5], b[1][:5]), (A, rnd, a[1], b[1])
            assert np.array_equal(a[5], b[5]), (A, rnd, "soft buffers")
            ok = a[2] <= 8
            if ok.all():                                                    # a lost block: the reference's abort flag cuts siblings short (not reproduced)
                assert np.array_equal(a[2], b[2]), (A, rnd, a[2], b[2])
                assert np.array_equal(a[3], b[3]) and np.array_equal(a[4], b[4]), (A, rnd)
                assert np.array_equal(b[4][:A // 8], pay)
            else:
                assert (b[2] > 8).any(), (A, rnd, a[2], b[2])
            if snr > 0.5:
                assert ok.all(), (A, rnd, a[2])
        if snr < 0.5:                                                         # rv 0 alone was hopeless; combined with rv 2 the block decodes, on both sides
            assert (a[2] <= 8).all() and np.array_equal(b[4][:A // 8], pay), (a[2], b[2])
