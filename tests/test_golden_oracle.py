"""The CPU oracle against the committed golden vectors (tests/golden/*.npz, produced from the compiled reference by
tools/gen_golden.py).  Runs anywhere -- this is what pins the oracle on machines without /root/reference."""
import os
import numpy as np
import pytest

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    return np.load(os.path.join(G, name))


def test_decoder_golden(oracle):
    d = _load("ldpc_decoder.npz")
    for ci in range(int(d["ncases"][0])):
        BG, Z, R, n, mi, om, uc, K = [int(x) for x in d[f"c{ci}_par"]]
        llr = d[f"c{ci}_llr"]
        for i in range(n):
            it, out = oracle.decode(BG, Z, R, mi, llr[i], om)
            # intended arithmetic == the AVX512 build everywhere
            assert it == d[f"c{ci}_iters_avx512"][i] and np.array_equal(np.asarray(out).view(np.uint8), d[f"c{ci}_out_avx512"][i].view(np.uint8)), (ci, i)
            oracle.lib.orc_set_quirks(1)   # + the AVX2 generator defect (only changes BG2 R15)
            it2, out2 = oracle.decode(BG, Z, R, mi, llr[i], om)
            oracle.lib.orc_set_quirks(0)
            assert it2 == d[f"c{ci}_iters_avx2"][i] and np.array_equal(np.asarray(out2).view(np.uint8), d[f"c{ci}_out_avx2"][i].view(np.uint8)), (ci, i)
            if (BG, R) != (2, 15):
                assert it == it2


def test_decoder_crc_mode_golden(oracle):
    d = _load("ldpc_decoder.npz")
    BG, Z, R, n, K = [int(x) for x in d["crc_par"]]
    for mi in (2, 3, 8):
        for i in range(n):
            it, out = oracle.decode(BG, Z, R, mi, d["crc_llr"][i], 0, 1, K, 1)
            assert it == d[f"crc{mi}_iters"][i] and np.array_equal(out, d[f"crc{mi}_out"][i]), (mi, i)


def test_encoder_golden(oracle):
    d = _load("ldpc_encoder.npz")
    for ci in range(int(d["ncases"][0])):
        BG, Z, K = [int(x) for x in d[f"e{ci}_par"]]
        P = d[f"e{ci}_in"]
        nout = (66 if BG == 1 else 50) * Z
        orig = np.unpackbits(d[f"e{ci}_orig"], axis=1)[:, :nout]
        optim = np.unpackbits(d[f"e{ci}_optim"], axis=1)[:, :nout]
        for i in range(P.shape[0]):
            assert np.array_equal(oracle.encode(BG, Z, K, P[i]), orig[i]), (BG, Z, i)
        assert np.array_equal(orig, optim) == ((BG, Z) != (2, 64))   # the reference's BG2 Z=64 default-encoder defect


def test_coding_golden(oracle):
    d = _load("coding.npz")
    for i, n in enumerate(d["crc_lens"]):
        for p in range(8):
            assert oracle.crc(p, d["crc_data"][i], int(n)) == int(d["crc_vals"][i][p])
    for ci, (BG, Z, F, E, rv, Tb, Cs, Qm) in enumerate(d["rm_cases"].tolist()):
        K = (22 if BG == 1 else 10) * Z
        Fo = K - F - 2 * Z
        N = (66 if BG == 1 else 50) * Z
        rc, e = oracle.rate_matching_tx(Tb, BG, Z, d[f"rm{ci}_w"], Cs, F, Fo, rv, E)
        assert rc == 0 and np.array_equal(e, d[f"rm{ci}_e"])
        assert np.array_equal(oracle.interleave(E, Qm, e), d[f"rm{ci}_f"])
        dei = oracle.deinterleave(E, Qm, d[f"rm{ci}_soft"])
        assert np.array_equal(dei, d[f"rm{ci}_dei"])
        w = np.zeros(N + 16, dtype=np.int16)
        oracle.rate_matching_rx(Tb, BG, Z, w, dei, Cs, rv, 1, E, F, Fo)
        assert np.array_equal(w[:N], d[f"rm{ci}_w1"])
        oracle.rate_matching_rx(Tb, BG, Z, w, dei, Cs, rv, 0, E, F, Fo)
        assert np.array_equal(w[:N], d[f"rm{ci}_w2"])
        assert list(oracle.get_R(rv, E, BG, Z, 0, 0)) == d[f"rm{ci}_getR"].tolist()
    for ci, (BG, B) in enumerate(d["seg_cases"].tolist()):
        Kb, Cc, K, Zc, F, s = oracle.segmentation(d[f"seg{ci}_in"], B, BG)
        assert [Kb, Cc, K, Zc, F] == d[f"seg{ci}_par"].tolist()
        assert np.array_equal(s, d[f"seg{ci}_out"])


def test_dft_golden(oracle):
    d = _load("dft.npz")
    for N in (64, 128, 256, 512, 768, 1024, 1536, 2048, 3072, 4096, 6144, 8192):
        for i in range(d[f"x{N}"].shape[0]):
            assert np.array_equal(oracle.dft(N, False, d[f"x{N}"][i], 1), d[f"dft{N}"][i]), N
            assert np.array_equal(oracle.dft(N, True, d[f"x{N}"][i], 1), d[f"idft{N}"][i]), N


def test_ofdm_parms_mirror(oracle):
    """openairinterface5g_b200/ofdm.py (product host code) derives the same slot geometry and rotation tables as the oracle restatement
    (itself pinned to nr_parms.c / nr_modulation.c through the compiled reference)."""
    from openairinterface5g_b200.ofdm import NrOfdmParms
    for N, mu, nb in ((4096, 1, 273), (2048, 2, 66), (1024, 0, 52), (1536, 1, 78)):
        P = NrOfdmParms(N, mu, nb)
        for slot in range(10 << mu):
            pre, cps, ss, fl = oracle.ofdm_geometry(N, mu, slot)
            p2, s2 = P.slot_geometry(slot)
            assert list(pre) == p2 and list(cps) == s2 and ss == P.slot_timestamp(slot) and fl == P.samples_per_frame
        assert np.array_equal(oracle.symbol_rotation(mu, 3.6192e9).reshape(-1, 2), P.symbol_rotation(3.6192e9))
        assert np.array_equal(oracle.timeshift_rotation(N, P.nb_prefix_samples // 8).reshape(-1, 2), P.timeshift_rotation())


def test_transport_arithmetic_mirror(oracle):
    """openairinterface5g_b200/transport.py (product host code) against the oracle's nr_segmentation / nr_get_R_ldpc_decoder restatements."""
    import ctypes as C
    from openairinterface5g_b200 import transport as T
    L = oracle.lib
    for BG in (1, 2):
        for B in list(range(24, 9000, 137)) + [8448, 8449, 3840, 3841, 100000, 235848, 471696, 1000000]:
            c, k, z, f = C.c_uint(), C.c_uint(), C.c_uint(), C.c_uint()
            kb = L.orc_segmentation(None, None, B, C.byref(c), C.byref(k), C.byref(z), C.byref(f), BG)
            if kb < 0:
                with pytest.raises(ValueError):
                    T.nr_segmentation(B, BG)
                continue
            s = T.nr_segmentation(B, BG)
            assert (s["C"], s["K"], s["Z"], s["F"], s["Kb"]) == (c.value, k.value, z.value, f.value, kb), (BG, B)
        for Z in (384, 208, 64, 22):
            for rv in range(4):
                for E in (500, 3000, 9072, 9126, 12000, 26000, 40000):
                    ll = C.c_int(0)
                    r = L.orc_get_R_ldpc_decoder(rv, E, BG, Z, C.byref(ll), 0)
                    assert T.nr_get_R_ldpc_decoder(rv, E, BG, Z) == (r, ll.value), (BG, Z, rv, E)
    G = T.nr_get_G(273, 14, 12, 1, 0, 6, 1)
    assert G == 255528 and sum(T.nr_get_E(G, 28, 6, 1, r) for r in range(28)) == G


def test_phy_golden(oracle):
    """Oracle restatements against tests/golden/phy.npz (outputs of the compiled reference, tools/gen_golden_phy.py): scrambling, mapper, slot-level OFDM,
    channel estimation, one- and two-layer PUSCH receivers.  This is what pins the oracle on a machine without /root/reference."""
    from oracle.bindings import ChestParms, PuschParms
    g = _load("phy.npz")
    q, nid, rnti = [int(x) for x in g["scr_par"]]
    sc = oracle.scramble(g["scr_bits"], q, nid, rnti)
    assert np.array_equal(sc, g["scr_out"])
    for Qm in (2, 4, 6, 8):
        assert np.array_equal(oracle.modulate(sc, (3001 // Qm) * Qm, Qm), g[f"mod{Qm}"])
    assert np.array_equal(oracle.unscramble_llr(g["unscr_in"], q, nid, rnti), g["unscr_out"])
    N, mu, nb_rb, slot, div, ta = [int(x) for x in g["ofdm_par"]]
    assert np.array_equal(oracle.symbol_rotation(mu, 3619200000.0), g["ofdm_rot_dl"]) and np.array_equal(oracle.symbol_rotation(mu, 3609200000.0), g["ofdm_rot_ul"])
    assert np.array_equal(oracle.timeshift_rotation(N, (N // 128 * 9) // div), g["ofdm_timeshift"])
    y, Frot = oracle.ofdm_tx_slot(N, mu, nb_rb, slot, 14, g["ofdm_rot_dl"], g["ofdm_txF"])
    assert np.array_equal(y, g["ofdm_tx_out"]) and np.array_equal(Frot.reshape(-1), g["ofdm_tx_rotated"].reshape(-1))
    assert np.array_equal(oracle.ofdm_rx_slot(N, mu, nb_rb, slot, div, ta, g["ofdm_rot_ul"], g["ofdm_rx_in"]), g["ofdm_rx_out"])
    P = ChestParms(*[int(x) for x in g["chest_par"]])
    assert np.array_equal(oracle.pusch_dmrs_pilots(P), g["chest_pilots"])
    est, st = oracle.pusch_channel_estimation(P, g["chest_rx"])
    assert np.array_equal(st, g["chest_state"]) and np.array_equal(est[:, P.symbol], g["chest_est"])
    PP = PuschParms(*[int(x) for x in g["rx1_par"]])
    sh, avg = oracle.pusch_log2_maxh(PP, 2, 2, g["rx1_rx"], g["rx1_h"])
    assert sh == int(g["rx1_shift"][0]) and np.array_equal(avg, g["rx1_avg"])
    for s in (2, 5):
        l, c = oracle.pusch_inner_rx_symbol(PP, s, 2, sh, g["rx1_rx"], g["rx1_h"])
        assert np.array_equal(l, g[f"rx1_llr{s}"]) and np.array_equal(c, g[f"rx1_comp{s}"])
    par2 = [int(x) for x in g["rx2_par"]]
    PP2 = PuschParms(*par2[:10])
    l, c = oracle.pusch_inner_rx_symbol_2l(PP2, 4, 2, par2[11], par2[10], g["rx1_rx"], g["rx2_h"])
    assert np.array_equal(l, g["rx2_llr"]) and np.array_equal(c, g["rx2_comp"])
    sh2, avg2 = oracle.pusch_log2_maxh_2l(PP2, 0, 2, par2[12], g["rx1_rx"], g["rx2_h"])
    assert sh2 == int(g["rx2_shift"][0]) and np.array_equal(avg2, g["rx2_avg"])
    PU = ChestParms(*[int(x) for x in g["uechest_par"]])
    assert np.array_equal(oracle.pdsch_channel_estimation(PU, g["chest_rx"])[:, PU.symbol], g["uechest_est"])
    pd = [int(x) for x in g["pdsch_par"]]
    l, sh = oracle.pdsch_rx_slot(PuschParms(*pd[:10]), pd[10], pd[11], g["rx1_rx"], g["rx1_h"])
    assert sh == int(g["pdsch_shift"][0]) and np.array_equal(l, g["pdsch_llr"])


def test_dft_fourway_golden(oracle):
    """The DFT-s-OFDM entry points (four interleaved transforms per call) against fixtures produced by the compiled reference (tools/gen_golden_dft4.py)."""
    d = _load("dft4.npz")
    for N in d["sizes"]:
        N = int(N)
        assert np.array_equal(oracle.dft4(N, d[f"x{N}"], 1), d[f"y{N}_s1"]), N
        assert np.array_equal(oracle.dft4(N, d[f"x{N}"], 0), d[f"y{N}_s0"]), N


def test_chest_variants_golden(oracle):
    """DMRS type 2 and chest_freq = 1 estimators against vectors of the compiled reference (tools/gen_golden_chest_variants.py)."""
    from oracle.bindings import ChestParms
    g = _load("chest_variants.npz")
    for i in range(int(g["n_cases"][0])):
        P = ChestParms(*[int(x) for x in g[f"par{i}"]])
        npil = g[f"pilots{i}"].size
        assert np.array_equal(oracle.pusch_dmrs_pilots(P)[:npil], g[f"pilots{i}"]), i
        est, st = oracle.pusch_channel_estimation(P, g["rx"])
        assert np.array_equal(st, g[f"state{i}"]) and np.array_equal(est[:, P.symbol], g[f"est{i}"]), i


def test_chest_time_avg_golden(oracle):
    g = _load("chest_variants.npz")
    for i in range(3):
        nsym, start, bitmap, nrb = [int(x) for x in g[f"tavg_par{i}"]]
        out = oracle.chest_time_domain_avg(g["tavg_in"], nsym, start, bitmap, nrb)
        first = min(s for s in range(start, start + nsym) if (bitmap >> s) & 1)
        assert np.array_equal(out[:, first], g[f"tavg_out{i}"]), i
        assert np.array_equal(np.delete(out, first, axis=1), np.delete(g["tavg_in"], first, axis=1)), i


def test_ue_chest_variants_golden(oracle):
    from oracle.bindings import ChestParms
    g = _load("chest_variants.npz")
    for i in range(4):
        P = ChestParms(*[int(x) for x in g[f"ue_par{i}"]])
        assert np.array_equal(oracle.pdsch_channel_estimation(P, g["rx"])[:, P.symbol], g[f"ue_est{i}"]), i


def test_pdsch_tx_precoding_golden(oracle):
    """nr_generate_pdsch after the encoder, identity and wideband non-identity precoding, against vectors of the compiled reference."""
    from oracle.bindings import PdschTxParms
    g = _load("chest_variants.npz")
    for i in range(2):
        P = PdschTxParms(*[int(x) for x in g[f"tx_par{i}"]])
        pm = int(g[f"tx_pm{i}"][0])
        P.set_precoding(pm, g[f"tx_w{i}"] if pm else None)
        assert np.array_equal(oracle.pdsch_tx_slot(P, g[f"tx_bits{i}"]), g[f"tx_out{i}"]), i


def test_transform_precoding_golden(oracle):
    """DFT-s-OFDM: computed low-PAPR sequences, the estimator with low-PAPR pilots and the one-layer inner receiver with nr_freq_equalization + nr_idft,
    against vectors of the compiled reference (tools/gen_golden_tp.py)."""
    from oracle.bindings import ChestParms, PuschParms
    g = _load("transform_precoding.npz")
    for M, u in ((30, 4), (36, 0), (150, 17), (1620, 29)):
        assert np.array_equal(oracle.lowpapr_seq(u, 0, M), g[f"seq_{M}_u{u}"]), (M, u)
    assert oracle.lowpapr_seq(0, 0, 24) is None                             # table-driven lengths are the caller's (see the header)
    for i in range(3):
        P = ChestParms(*[int(x) for x in g[f"chest_par{i}"]])
        oracle.chest_set_lowpapr(g[f"chest_seq{i}"])
        try:
            est, st = oracle.pusch_channel_estimation(P, g["rx"])
        finally:
            oracle.chest_set_lowpapr(None)
        assert np.array_equal(st, g[f"chest_state{i}"]) and np.array_equal(est[:, P.symbol], g[f"chest_est{i}"]), i
    oracle.pusch_set_transform_precoding(1)
    try:
        for i in range(5):
            P = PuschParms(*[int(x) for x in g[f"rx_par{i}"]])
            symbol, shift = [int(x) for x in g[f"rx_sym{i}"]]
            llr, comp = oracle.pusch_inner_rx_symbol(P, symbol, 2, shift, g["rx"], g["h"])
            assert np.array_equal(llr, g[f"rx_llr{i}"]) and np.array_equal(comp[:24 * P.rb_size], g[f"rx_comp{i}"]), i
    finally:
        oracle.pusch_set_transform_precoding(0)


def _ue_layers_inputs(g, i):
    import zlib
    N, nb_rx, _, _, _, _, _, _, _, _, _, nl, ay, ah = [int(x) for x in g[f"case{i}"]]
    rng = np.random.default_rng(3000 + i)
    rx = rng.integers(-ay, ay + 1, size=(nb_rx, 14, N, 2)).astype(np.int16)
    h = rng.integers(-ah, ah + 1, size=(nl * nb_rx, 14, N, 2)).astype(np.int16)
    assert zlib.crc32(rx.tobytes() + h.tobytes()) == int(g[f"crc{i}"]), "numpy's seeded stream changed: regenerate with tools/gen_golden_ue_layers.py"
    return rx, h


def test_ue_3_4_layers_golden(oracle):
    """nr_rx_pdsch with three and four layers against vectors of the compiled reference (tools/gen_golden_ue_layers.py; seeded inputs guarded by a checksum)."""
    from oracle.bindings import PuschParms
    g = _load("ue_layers.npz")
    for i in range(int(g["n"])):
        N, nb_rx, rb_start, rb_size, Qm, dpos, dtype_, cdm, carrier, start, nsym, nl, ay, ah = [int(x) for x in g[f"case{i}"]]
        rx, h = _ue_layers_inputs(g, i)
        llr, sh = oracle.pdsch_rx_slot(PuschParms(N, nb_rx, rb_start, 0, rb_size, N - carrier * 6, Qm, dpos, dtype_, cdm), start, nsym, rx, h, nl=nl)
        assert sh == int(g[f"sh{i}"]) and np.array_equal(llr, g[f"llr{i}"]), i


def test_ptrs_ue_golden(oracle):
    """PT-RS at the UE (nr_pdsch_ptrs_processing inside nr_rx_pdsch): LLRs, log2_maxh, per-symbol phase estimates and PT-RS RE counts against vectors of the
    compiled reference (tools/gen_golden_ptrs.py)."""
    from oracle.bindings import PuschParms, PtrsParms
    g = _load("ptrs.npz")
    for i in range(int(g["n"])):
        N, nb_rx, rb_start, rb_size, Qm, dpos, dtype_, cdm, carrier, start, nsym, L, K, reoff, rnti, slot, nscid, nid = [int(x) for x in g[f"case{i}"]]
        P = PuschParms(N, nb_rx, rb_start, 0, rb_size, N - carrier * 6, Qm, dpos, dtype_, cdm)
        llr, sh, ph, nre = oracle.pdsch_rx_slot_ptrs(P, PtrsParms(1, L, K, reoff, rnti, slot, nscid, nid), start, nsym, g[f"rx{i}"], g[f"h{i}"])
        assert sh == int(g[f"sh{i}"]) and np.array_equal(nre, g[f"nre{i}"]) and np.array_equal(ph, g[f"phase{i}"]), i
        assert np.array_equal(llr, g[f"llr{i}"]), i


def test_ptrs_gnb_tx_golden(oracle):
    """PT-RS insertion in nr_generate_pdsch (pduBitmap & 1), with and without wideband precoding, against vectors of the compiled reference."""
    from oracle.bindings import PdschTxParms
    g = _load("ptrs.npz")
    for j in range(int(g["n_tx"])):
        N, carrier, ntx, slot, rb0, nrb, Qm, nl, s0, ns, dpos, dtype_, cdm, ports, scid, amp, L, K, reoff, pm = [int(x) for x in g[f"tx_case{j}"]]
        P = PdschTxParms(N, ntx, slot, rb0, 0, nrb, N - carrier * 6, Qm, nl, s0, ns, dpos, dtype_, cdm, ports, scid, 40 + slot, 501, 0x1234 + slot, amp).set_ptrs(L, K, reoff)
        if pm:
            P.set_precoding(pm, g[f"tx_w{j}"])
        assert np.array_equal(oracle.pdsch_tx_slot(P, g[f"tx_bits{j}"]), g[f"tx_out{j}"]), j


def test_transform_precoding_64qam_cannot_be_demapped(oracle):
    """A property of the reference, kept visible: after nr_freq_equalization the compensated symbols sit at 128 k (k = 1, 3, 5, 7 for 64QAM) while the constant
    thresholds it installs are 316 / 158 (nr_freq_equalization.c:63-67), so a NOISELESS DFT-s-OFDM 64QAM symbol is demapped with bit errors; QPSK and 16QAM are
    clean.  The oracle (pinned bit-exactly to the reference for this path) and the library reproduce it; the closed-loop slot test therefore runs Qm 2 and 4."""
    from oracle.bindings import PuschParms
    N, nb, fco = 1024, 25, 1024 - 6 * 52
    M = 12 * nb
    rng = np.random.default_rng(1)
    errors = {}
    for Qm in (2, 4, 6):
        bits = rng.integers(0, 2, size=M * Qm).astype(np.uint8)
        words = np.zeros((M * Qm + 31) // 32 + 1, np.uint32)
        for i, b in enumerate(bits):
            words[i >> 5] |= np.uint32(int(b) << (i & 31))
        sym = oracle.modulate(words, M * Qm, Qm).reshape(-1, 2).astype(np.float64)
        x = (sym[:, 0] + 1j * sym[:, 1]) * 724 / 32768.0
        h = 0.6 + 0.37j
        y = h * np.fft.fft(x) / np.sqrt(M)                                  # the UE's transform precoder, a flat channel, no noise
        rx = np.zeros((1, 14, N, 2), np.int16); hh = np.zeros((1, 14, N, 2), np.int16)
        sc = (fco + np.arange(M)) % N
        rx[0, 0, sc, 0] = np.round(y.real); rx[0, 0, sc, 1] = np.round(y.imag)
        unit = 23170.0 * 724 / 32768.0
        hh[0, 2, :M, 0] = round(h.real * unit); hh[0, 2, :M, 1] = round(h.imag * unit)
        P = PuschParms(N, 1, 0, 0, nb, fco, Qm, 1 << 2, 0, 2)
        oracle.pusch_set_transform_precoding(1)
        try:
            sh, _ = oracle.pusch_log2_maxh(P, 0, 2, rx, hh)
            llr, _ = oracle.pusch_inner_rx_symbol(P, 0, 2, sh, rx, hh)
        finally:
            oracle.pusch_set_transform_precoding(0)
        errors[Qm] = int(((llr < 0).astype(np.uint8) != bits).sum())
    assert errors[2] == 0 and errors[4] == 0 and errors[6] > 100, errors
