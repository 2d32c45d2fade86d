"""GPU parity of the UE-side single-layer PDSCH receiver (pdsch_ue = 1 in nrb200_pusch_rx_t) against the CPU oracle, which
tests/test_oracle_vs_reference.py pins to the reference's own nr_rx_pdsch symbol loop."""
import numpy as np
import pytest

from oracle.bindings import PuschParms
from openairinterface5g_b200.ldpc import PuschRxDesc

pytestmark = pytest.mark.gpu

CASES = [  # N, nb_rx, rb_start, rb_size, Qm, dmrs_pos, dmrs_type, cdm groups, carrier PRBs, start_symbol, nr_symbols
    (4096, 2, 0, 273, 6, 1 << 2, 0, 2, 273, 1, 13), (4096, 4, 0, 273, 8, 1 << 2, 0, 1, 273, 1, 13), (2048, 1, 10, 50, 4, (1 << 2) | (1 << 11), 0, 1, 106, 1, 13),
    (2048, 2, 30, 76, 2, 1 << 3, 0, 2, 106, 2, 10), (1024, 4, 0, 52, 6, 1 << 2, 1, 1, 52, 1, 13), (1024, 2, 20, 32, 4, 1 << 2, 1, 2, 52, 2, 12), (512, 4, 3, 11, 8, 1 << 1, 0, 1, 25, 1, 6),
    (2048, 2, 0, 106, 6, (1 << 2) | (1 << 13), 0, 1, 106, 1, 13),       # the last symbol is a DMRS symbol with data: its (shorter) magnitude buffers serve every symbol
]


def test_pdsch_rx_ue_vs_oracle(ldpc, oracle):
    rng = np.random.default_rng(66)
    for N, nb_rx, rb_start, rb_size, Qm, dpos, dtype_, cdm, carrier, start, nsym in CASES:
        big = N == 512
        ay, ah = (32767, 32767) if big else (2000, 1500)
        rx = rng.integers(-ay, ay + 1, size=(nb_rx, 14, N, 2)).astype(np.int16)
        h = rng.integers(-ah, ah + 1, size=(nb_rx, 14, N, 2)).astype(np.int16)
        fco = N - carrier * 6
        P = PuschParms(N, nb_rx, rb_start, 0, rb_size, fco, Qm, dpos, dtype_, cdm)
        llr_o, sh_o = oracle.pdsch_rx_slot(P, start, nsym, rx, h)
        for unscr in (None, (0x2345, 501)):
            d = PuschRxDesc(N, nb_rx, rb_start, 0, rb_size, fco, Qm, start, nsym, dpos, dtype_, cdm, 0xFFFFFFFF, 0, 0, 0 if unscr is None else 1,
                            0 if unscr is None else unscr[0], 0 if unscr is None else unscr[1], 1, 0, 0, 1)
            llr, sh = ldpc.pusch_inner_rx_host(d, rx, h)
            assert sh == sh_o, (N, nb_rx, Qm, sh, sh_o)
            ref = llr_o if unscr is None else oracle.unscramble_llr(llr_o, 0, unscr[1], unscr[0])
            assert llr.size == ref.size and np.array_equal(llr, ref), (N, nb_rx, rb_start, rb_size, Qm, dpos, dtype_, cdm, unscr)


CASES_2L = [  # N, nb_rx, rb_start, rb_size, Qm, dmrs_pos, dmrs_type, cdm groups, carrier PRBs, start_symbol, nr_symbols, amplitude (rx, h)
    (4096, 2, 0, 273, 6, 1 << 2, 0, 1, 273, 1, 13, (2000, 1500)), (4096, 4, 0, 273, 8, 1 << 2, 0, 2, 273, 1, 13, (900, 700)),
    (2048, 2, 10, 50, 4, (1 << 2) | (1 << 11), 0, 1, 106, 1, 13, (4000, 6000)), (2048, 3, 30, 76, 2, 1 << 3, 0, 2, 106, 2, 10, (300, 200)),
    (1024, 4, 0, 52, 6, 1 << 2, 1, 1, 52, 1, 13, (2000, 1500)), (1024, 2, 20, 32, 4, 1 << 2, 1, 2, 52, 2, 12, (12000, 9000)),
    (512, 4, 3, 11, 8, 1 << 1, 0, 1, 25, 1, 6, (32767, 32767)), (512, 2, 0, 25, 6, 1 << 2, 0, 1, 25, 0, 14, (60, 40)),
    (2048, 2, 0, 106, 6, (1 << 2) | (1 << 13), 0, 1, 106, 1, 13, (2000, 1500)),      # last symbol = DMRS symbol with data
]


def test_pdsch_rx_ue_2layers_vs_oracle(ldpc, oracle):
    """Two layers: per-layer MRC, zero forcing (nr_zero_forcing_rx), determinant thresholds, layer de-mapping, descrambling -- one launch per slot."""
    rng = np.random.default_rng(67)
    for N, nb_rx, rb_start, rb_size, Qm, dpos, dtype_, cdm, carrier, start, nsym, (ay, ah) in CASES_2L:
        rx = rng.integers(-ay, ay + 1, size=(nb_rx, 14, N, 2)).astype(np.int16)
        h = rng.integers(-ah, ah + 1, size=(2 * nb_rx, 14, N, 2)).astype(np.int16)
        fco = N - carrier * 6
        P = PuschParms(N, nb_rx, rb_start, 0, rb_size, fco, Qm, dpos, dtype_, cdm)
        llr_o, sh_o = oracle.pdsch_rx_slot(P, start, nsym, rx, h, nl=2)
        for unscr in (None, (0x2345, 501)):
            d = PuschRxDesc(N, nb_rx, rb_start, 0, rb_size, fco, Qm, start, nsym, dpos, dtype_, cdm, 0xFFFFFFFF, 0, 0, 0 if unscr is None else 1,
                            0 if unscr is None else unscr[0], 0 if unscr is None else unscr[1], 2, 0, 0, 1)
            llr, sh = ldpc.pusch_inner_rx_host(d, rx, h)
            assert sh == sh_o, (N, nb_rx, Qm, sh, sh_o)
            ref = llr_o if unscr is None else oracle.unscramble_llr(llr_o, 0, unscr[1], unscr[0])
            assert llr.size == ref.size and np.array_equal(llr, ref), (N, nb_rx, rb_start, rb_size, Qm, dpos, dtype_, cdm, unscr, np.nonzero(llr != ref)[0][:6])


def test_pdsch_rx_ue_ptrs_vs_oracle(ldpc, oracle):
    """PT-RS at the UE (ptrs = 1): per-symbol common phase error from the PT-RS REs (IEEE double arithmetic on the device), PT-RS REs removed from the LLR stream,
    interpolation over the symbols without PT-RS, rotation of every non-DMRS symbol -- two launches per slot, bit-exact against the oracle that
    tests/test_oracle_vs_reference.py pins to the real nr_rx_pdsch + nr_pdsch_ptrs_processing.  Random full-scale slots and coherent ones with a phase ramp."""
    from oracle.bindings import PtrsParms
    from common import PTRS_CASES, PTRS_SIGNALS, ptrs_inputs
    rng = np.random.default_rng(72)
    for case in PTRS_CASES:
        N, nb_rx, rb_start, rb_size, Qm, dpos, dtype_, cdm, carrier, start, nsym, L, K, reoff, rnti, slot, nscid, nid = case
        fco = N - carrier * 6
        for kind, a, b in PTRS_SIGNALS:
            rx, h = ptrs_inputs(oracle, rng, case, kind, a, b)
            P = PuschParms(N, nb_rx, rb_start, 0, rb_size, fco, Qm, dpos, dtype_, cdm)
            llr_o, sh_o, ph_o, nre_o = oracle.pdsch_rx_slot_ptrs(P, PtrsParms(1, L, K, reoff, rnti, slot, nscid, nid), start, nsym, rx, h)
            for unscr in (0, 1):
                d = PuschRxDesc(N, nb_rx, rb_start, 0, rb_size, fco, Qm, start, nsym, dpos, dtype_, cdm, 0xFFFFFFFF, 0, 0, unscr, rnti, 501, 1, 0, 0, 1)
                d.set_ptrs(L, K, reoff, slot, nscid, nid)
                mask, n_re = ldpc.pdsch_ptrs_layout(d)
                assert mask == oracle.ptrs_symbols(start, nsym, L, dpos) and [n_re if (mask >> s) & 1 else 0 for s in range(14)] == nre_o.tolist(), (case, mask, n_re)
                llr, sh = ldpc.pusch_inner_rx_host(d, rx, h)
                assert sh == sh_o, (case, kind, sh, sh_o)
                ref = llr_o if not unscr else oracle.unscramble_llr(llr_o, 0, 501, rnti)
                assert llr.size == ref.size and np.array_equal(llr, ref), (case, kind, a, b, unscr, np.nonzero(llr != ref)[0][:6])


def test_pdsch_rx_ue_ptrs_golden(ldpc):
    """The same path against the committed vectors of the compiled reference (tests/golden/ptrs.npz, tools/gen_golden_ptrs.py)."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ptrs.npz"))
    for i in range(int(g["n"])):
        N, nb_rx, rb_start, rb_size, Qm, dpos, dtype_, cdm, carrier, start, nsym, L, K, reoff, rnti, slot, nscid, nid = [int(x) for x in g[f"case{i}"]]
        d = PuschRxDesc(N, nb_rx, rb_start, 0, rb_size, N - carrier * 6, Qm, start, nsym, dpos, dtype_, cdm, 0xFFFFFFFF, 0, 0, 0, rnti, 0, 1, 0, 0, 1)
        d.set_ptrs(L, K, reoff, slot, nscid, nid)
        llr, sh = ldpc.pusch_inner_rx_host(d, g[f"rx{i}"], g[f"h{i}"])
        assert sh == int(g[f"sh{i}"]) and np.array_equal(llr, g[f"llr{i}"]), (i, np.nonzero(llr != g[f"llr{i}"])[0][:6])


def test_pdsch_rx_ptrs_unsupported_combinations_fail_loudly(ldpc):
    """PT-RS with two layers (the reference squeezes and rotates layer 0 only) and on the gNB side (DESIGN.md defect 19) return -4: no silent fallback."""
    N, nb_rx = 512, 2
    rx = np.zeros((nb_rx, 14, N, 2), np.int16); h2 = np.zeros((2 * nb_rx, 14, N, 2), np.int16)
    for nl, ue in ((2, 1), (1, 0)):
        d = PuschRxDesc(N, nb_rx, 0, 0, 25, N - 150, 4, 1, 13, 1 << 2, 0, 1, 5, 0, 0, 0, 7, 0, nl, 0, 0, ue)
        d.set_ptrs(1, 2, 0, 0, 0, 0)
        assert ldpc.pusch_num_llr(d) == 0
        with pytest.raises(Exception):
            ldpc.pusch_inner_rx_host(d, rx, h2[:nl * nb_rx])
    d = PuschRxDesc(N, nb_rx, 0, 0, 25, N - 150, 4, 1, 13, 1 << 2, 0, 1, 5, 0, 0, 0, 7, 0, 1, 0, 0, 1)
    d.set_ptrs(1, 3, 0, 0, 0, 0)                                           # K_PTRS must be 2 or 4
    assert ldpc.pusch_num_llr(d) == 0


CASES_NL = [  # N, nb_rx, rb_start, rb_size, Qm, dmrs_pos, dmrs_type, cdm groups, carrier PRBs, start_symbol, nr_symbols, layers, amplitude (rx, h)
    (4096, 4, 0, 273, 6, 1 << 2, 0, 2, 273, 1, 13, 4, (2000, 1500)), (4096, 4, 0, 273, 8, 1 << 2, 0, 2, 273, 1, 13, 3, (900, 700)),
    (2048, 4, 10, 50, 4, (1 << 2) | (1 << 11), 0, 1, 106, 1, 13, 3, (4000, 6000)), (2048, 3, 30, 76, 2, 1 << 3, 0, 2, 106, 2, 10, 3, (300, 200)),
    (1024, 4, 0, 52, 6, 1 << 2, 1, 1, 52, 1, 13, 4, (2000, 1500)), (1024, 4, 20, 32, 4, 1 << 2, 1, 2, 52, 2, 12, 4, (12000, 9000)),
    (512, 4, 3, 11, 8, 1 << 1, 0, 1, 25, 1, 6, 4, (32767, 32767)), (512, 4, 0, 25, 6, 1 << 2, 0, 1, 25, 0, 14, 3, (60, 40)),
    (2048, 4, 0, 106, 6, (1 << 2) | (1 << 13), 0, 1, 106, 1, 13, 4, (2000, 1500)), (1024, 2, 0, 52, 6, 1 << 2, 0, 2, 52, 1, 13, 3, (2000, 1500)),
    (2048, 4, 4, 100, 2, 1 << 2, 0, 2, 106, 1, 13, 4, (2500, 2500)), (2048, 4, 4, 100, 8, 1 << 2, 0, 2, 106, 1, 13, 4, (2500, 2500)),
]


def test_pdsch_rx_ue_3_4_layers_vs_oracle(ldpc, oracle):
    """Three and four layers at the UE: per-layer MRC, zero forcing with the reference's recursive fixed-point determinant / adjugate (unrolled at compile time in
    pdsch_rxn_kernel<QM, NL>), determinant thresholds of the last symbol, layer de-mapping, descrambling -- one launch per slot; up to 16 (layer, antenna) planes in
    the level measurement.  The oracle is pinned to the real nr_rx_pdsch (tests/test_oracle_vs_reference.py::test_pdsch_rx_slot_ue_3_4_layers)."""
    rng = np.random.default_rng(76)
    for N, nb_rx, rb_start, rb_size, Qm, dpos, dtype_, cdm, carrier, start, nsym, nl, (ay, ah) in CASES_NL:
        rx = rng.integers(-ay, ay + 1, size=(nb_rx, 14, N, 2)).astype(np.int16)
        h = rng.integers(-ah, ah + 1, size=(nl * nb_rx, 14, N, 2)).astype(np.int16)
        fco = N - carrier * 6
        P = PuschParms(N, nb_rx, rb_start, 0, rb_size, fco, Qm, dpos, dtype_, cdm)
        llr_o, sh_o = oracle.pdsch_rx_slot(P, start, nsym, rx, h, nl=nl)
        for unscr in (None, (0x2345, 501)):
            d = PuschRxDesc(N, nb_rx, rb_start, 0, rb_size, fco, Qm, start, nsym, dpos, dtype_, cdm, 0xFFFFFFFF, 0, 0, 0 if unscr is None else 1,
                            0 if unscr is None else unscr[0], 0 if unscr is None else unscr[1], nl, 0, 0, 1)
            llr, sh = ldpc.pusch_inner_rx_host(d, rx, h)
            assert sh == sh_o, (N, nb_rx, Qm, nl, sh, sh_o)
            ref = llr_o if unscr is None else oracle.unscramble_llr(llr_o, 0, unscr[1], unscr[0])
            assert llr.size == ref.size and np.array_equal(llr, ref), (N, nb_rx, rb_start, rb_size, Qm, nl, dpos, dtype_, cdm, unscr, np.nonzero(llr != ref)[0][:6])


def test_pdsch_rx_ue_3_4_layers_golden(ldpc):
    """The same kernels against the committed vectors of the compiled reference (tests/golden/ue_layers.npz)."""
    import os
    import zlib
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ue_layers.npz"))
    for i in range(int(g["n"])):
        N, nb_rx, rb_start, rb_size, Qm, dpos, dtype_, cdm, carrier, start, nsym, nl, ay, ah = [int(x) for x in g[f"case{i}"]]
        rng = np.random.default_rng(3000 + i)
        rx = rng.integers(-ay, ay + 1, size=(nb_rx, 14, N, 2)).astype(np.int16)
        h = rng.integers(-ah, ah + 1, size=(nl * nb_rx, 14, N, 2)).astype(np.int16)
        assert zlib.crc32(rx.tobytes() + h.tobytes()) == int(g[f"crc{i}"])
        d = PuschRxDesc(N, nb_rx, rb_start, 0, rb_size, N - carrier * 6, Qm, start, nsym, dpos, dtype_, cdm, 0xFFFFFFFF, 0, 0, 0, 0, 0, nl, 0, 0, 1)
        llr, sh = ldpc.pusch_inner_rx_host(d, rx, h)
        assert sh == int(g[f"sh{i}"]) and np.array_equal(llr, g[f"llr{i}"]), (i, np.nonzero(llr != g[f"llr{i}"])[0][:6])


def test_pdsch_rx_ue_ptrs_fuzz(ldpc, oracle):
    """Random PT-RS configurations (tests/common.py:ptrs_fuzz_cases; the CPU suite runs 300 of the tail-safe kind against the real nr_rx_pdsch): 120 that the reference
    can run and 60 whose last symbol has an RE count off the SIMD grid (where the reference itself overruns its LLR buffer, DESIGN.md defect 22) against the oracle."""
    from oracle.bindings import PtrsParms
    from common import ptrs_fuzz_cases, ptrs_inputs
    rng = np.random.default_rng(84)
    cases = ptrs_fuzz_cases(rng, 120) + ptrs_fuzz_cases(rng, 60, safe_tail=False)
    done = 0
    for n, case in enumerate(cases):
        N, nb_rx, rb_start, rb_size, Qm, dpos, dtype_, cdm, carrier, start, nsym, L, K, reoff, rnti, slot, nscid, nid = case
        kind, a, b = (("random", 2000, 1500), ("coherent", 30, 0.05), ("coherent", 0, 0.2))[n % 3]
        rx, h = ptrs_inputs(oracle, rng, case, kind, a, b)
        fco = N - carrier * 6
        llr_o, sh_o, ph_o, nre_o = oracle.pdsch_rx_slot_ptrs(PuschParms(N, nb_rx, rb_start, 0, rb_size, fco, Qm, dpos, dtype_, cdm), PtrsParms(1, L, K, reoff, rnti, slot, nscid, nid),
                                                             start, nsym, rx, h)
        d = PuschRxDesc(N, nb_rx, rb_start, 0, rb_size, fco, Qm, start, nsym, dpos, dtype_, cdm, 0xFFFFFFFF, 0, 0, n & 1, rnti, 77, 1, 0, 0, 1)
        d.set_ptrs(L, K, reoff, slot, nscid, nid)
        assert ldpc.pusch_num_llr(d) == llr_o.size, case
        if llr_o.size == 0:
            continue                                                          # an allocation of DMRS symbols without data
        llr, sh = ldpc.pusch_inner_rx_host(d, rx, h)
        ref = llr_o if not (n & 1) else oracle.unscramble_llr(llr_o, 0, 77, rnti)
        assert sh == sh_o and np.array_equal(llr, ref), (case, kind, sh, sh_o, np.nonzero(llr != ref)[0][:6])
        done += 1
    assert done > 150


def test_pdsch_rx_ue_layers_fuzz(ldpc, oracle):
    """Random allocations / DMRS layouts with 1 ... 4 layers (no PT-RS) against the oracle: 240 configurations, incl. DMRS symbols with data as the slot's last symbol."""
    from common import ptrs_fuzz_cases
    rng = np.random.default_rng(86)
    done = 0
    for n, case in enumerate(ptrs_fuzz_cases(rng, 240, safe_tail=False)):
        N, nb_rx, rb_start, rb_size, Qm, dpos, dtype_, cdm, carrier, start, nsym = case[:11]
        nl = 1 + n % 4
        nb_rx = max(nb_rx, 2) if nl > 1 else nb_rx
        ay, ah = ((2000, 1500), (600, 900), (32767, 32767))[n % 3]
        rx = rng.integers(-ay, ay + 1, size=(nb_rx, 14, N, 2)).astype(np.int16)
        h = rng.integers(-ah, ah + 1, size=(nl * nb_rx, 14, N, 2)).astype(np.int16)
        fco = N - carrier * 6
        llr_o, sh_o = oracle.pdsch_rx_slot(PuschParms(N, nb_rx, rb_start, 0, rb_size, fco, Qm, dpos, dtype_, cdm), start, nsym, rx, h, nl=nl)
        d = PuschRxDesc(N, nb_rx, rb_start, 0, rb_size, fco, Qm, start, nsym, dpos, dtype_, cdm, 0xFFFFFFFF, 0, 0, n & 1, 0x2345, 501, nl, 0, 0, 1)
        assert ldpc.pusch_num_llr(d) == llr_o.size, case
        if llr_o.size == 0:
            continue
        llr, sh = ldpc.pusch_inner_rx_host(d, rx, h)
        ref = llr_o if not (n & 1) else oracle.unscramble_llr(llr_o, 0, 501, 0x2345)
        assert sh == sh_o and np.array_equal(llr, ref), (case[:11], nl, sh, sh_o, np.nonzero(llr != ref)[0][:6])
        done += 1
    assert done > 200
