"""GPU parity of the fused single-layer PUSCH inner receiver (extract + MRC compensation + LLR + descrambling, one launch per slot)
against the CPU oracle, which tests/test_oracle_vs_reference.py pins symbol by symbol to the compiled reference inner_rx."""
import numpy as np
import pytest

from oracle.bindings import PuschParms
from openairinterface5g_b200.ldpc import PuschRxDesc

pytestmark = pytest.mark.gpu

CASES = [  # N, nb_rx, rb_start, rb_size, Qm, dmrs_pos, dmrs_type, cdm_no_data, carrier PRBs, start_symbol, nr_of_symbols
    (4096, 4, 0, 273, 6, 1 << 2, 0, 2, 273, 0, 14), (4096, 2, 0, 273, 8, 1 << 2, 0, 1, 273, 0, 14), (2048, 1, 10, 50, 4, (1 << 2) | (1 << 11), 0, 1, 106, 0, 14),
    (2048, 2, 30, 76, 2, 1 << 3, 0, 1, 106, 2, 10), (1024, 4, 0, 52, 6, 1 << 2, 1, 1, 52, 0, 14), (1024, 2, 20, 32, 4, 1 << 2, 1, 2, 52, 1, 12),
    (512, 8, 3, 11, 8, 1 << 0, 0, 1, 25, 0, 7), (4096, 4, 100, 173, 6, (1 << 2) | (1 << 7) | (1 << 11), 1, 1, 273, 0, 14),
    (2048, 4, 56, 50, 6, 1 << 2, 0, 1, 106, 0, 14),   # allocation wraps around DC (start_re + nb_re > N)
]


def _oracle_slot(oracle, P, start, nsym, rx, h, shift, unscr=None):
    """Compose the per-symbol oracle the way nr_rx_pusch_tp / nr_pusch_symbol_processing do."""
    dm = [s for s in range(start, start + nsym) if (P.ul_dmrs_symb_pos >> s) & 1]
    cur, out = dm[0], []
    for s in range(start, start + nsym):
        if (P.ul_dmrs_symb_pos >> s) & 1:
            cur = s
        if oracle.pusch_nb_re(P, s) == 0:
            continue
        out.append(oracle.pusch_inner_rx_symbol(P, s, cur, shift, rx, h)[0])
    llr = np.concatenate(out)
    if unscr is not None:
        llr = oracle.unscramble_llr(llr, 0, unscr[1], unscr[0])
    return llr


def test_pusch_inner_rx_vs_oracle(ldpc, oracle):
    rng = np.random.default_rng(50)
    for N, nb_rx, rb_start, rb_size, Qm, dpos, dtype_, cdm, carrier, start, nsym in CASES:
        big = nb_rx == 8
        ay, ah = (32767, 32767) if big else (2000, 1500)
        rx = rng.integers(-ay, ay + 1, size=(nb_rx, 14, N, 2)).astype(np.int16)
        h = rng.integers(-ah, ah + 1, size=(nb_rx, 14, N, 2)).astype(np.int16)
        fco = N - carrier * 6
        P = PuschParms(N, nb_rx, rb_start, 0, rb_size, fco, Qm, dpos, dtype_, cdm)
        dm0 = [s for s in range(start, start + nsym) if (dpos >> s) & 1][0]
        meas = [s for s in range(start, start + nsym) if oracle.pusch_nb_re(P, s) > 0][0]
        cur = dm0 if meas < dm0 else max(s for s in range(start, meas + 1) if (dpos >> s) & 1)
        sh_o, _ = oracle.pusch_log2_maxh(P, meas, cur, rx, h)
        for unscr in (None, (0x1234, 77)):
            d = PuschRxDesc(N, nb_rx, rb_start, 0, rb_size, fco, Qm, start, nsym, dpos, dtype_, cdm, 0xFFFFFFFF, 0, 0,
                            0 if unscr is None else 1, 0 if unscr is None else unscr[0], 0 if unscr is None else unscr[1])
            llr, sh = ldpc.pusch_inner_rx_host(d, rx, h)
            assert sh == sh_o, (N, nb_rx, Qm, sh, sh_o)
            ref = _oracle_slot(oracle, P, start, nsym, rx, h, sh_o, unscr)
            assert llr.size == ref.size and np.array_equal(llr, ref), (N, nb_rx, rb_start, rb_size, Qm, unscr)
        # explicit shift (the caller's own log2_maxh)
        d = PuschRxDesc(N, nb_rx, rb_start, 0, rb_size, fco, Qm, start, nsym, dpos, dtype_, cdm, 3, 0, 0, 0, 0, 0)
        llr, sh = ldpc.pusch_inner_rx_host(d, rx, h)
        assert sh == 3 and np.array_equal(llr, _oracle_slot(oracle, P, start, nsym, rx, h, 3))


def test_pusch_inner_rx_two_layers_vs_oracle(ldpc, oracle):
    """nrOfLayers = 2: matched filter per layer + MMSE (Qm >= 6) or joint max-log ML detector (Qm < 6) + layer de-mapping + descrambling in one launch."""
    rng = np.random.default_rng(51)
    for N, nb_rx, rb_start, rb_size, Qm, carrier, nvar, max_ch, dpos, cdm in ((4096, 4, 0, 273, 6, 273, 40, 1400, 1 << 2, 2), (2048, 2, 10, 50, 8, 106, 7, 30000, 1 << 2, 2),
                                                                             (1024, 4, 20, 32, 6, 52, 1, 0, 1 << 3, 2), (2048, 2, 30, 75, 8, 106, 1000, 9000, 1 << 2, 1),
                                                                             (4096, 4, 3, 11, 6, 273, 90000, 200000, (1 << 2) | (1 << 11), 2),
                                                                             # Qm < 6: the joint max-log ML detector (QPSK-QPSK, 16QAM-16QAM), any number of rx antennas;
                                                                             # rb_size = 3 mod 4 leaves the last 4 REs of a symbol without LLRs in the reference
                                                                             (4096, 4, 0, 273, 4, 273, 0, 1400, 1 << 2, 2), (2048, 2, 10, 51, 2, 106, 0, 30000, 1 << 2, 2),
                                                                             (1024, 3, 20, 31, 4, 52, 0, 0, 1 << 3, 1), (2048, 1, 30, 75, 2, 106, 0, 9000, (1 << 2) | (1 << 11), 2),
                                                                             (1024, 2, 0, 52, 4, 52, 0, 200000, 1 << 2, 2)):
        fco = N - carrier * 6
        P = PuschParms(N, nb_rx, rb_start, 0, rb_size, fco, Qm, dpos, 0, cdm)
        rx = rng.integers(-2000, 2001, size=(nb_rx, 14, N, 2)).astype(np.int16)
        h = rng.integers(-1500, 1501, size=(2 * nb_rx, 14, N, 2)).astype(np.int16)
        dms = [s for s in range(14) if (dpos >> s) & 1]
        meas = [s for s in range(14) if oracle.pusch_nb_re(P, s) > 0][0]
        sh_o, _ = oracle.pusch_log2_maxh_2l(P, meas, dms[0], max_ch, rx, h)
        for shift in (0xFFFFFFFF, 7):
            for unscr in (None, (0x4321, 99)):
                d = PuschRxDesc(N, nb_rx, rb_start, 0, rb_size, fco, Qm, 0, 14, dpos, 0, cdm, shift, 0, 0, 0 if unscr is None else 1,
                                0 if unscr is None else unscr[0], 0 if unscr is None else unscr[1], 2, nvar, max_ch)
                llr, sh = ldpc.pusch_inner_rx_host(d, rx, h)
                use = sh_o if shift == 0xFFFFFFFF else 7
                assert sh == use, (N, nb_rx, Qm, sh, use)
                cur, out = dms[0], []
                for s in range(14):
                    if (dpos >> s) & 1:
                        cur = s
                    v = oracle.pusch_nb_re(P, s)
                    if v == 0:
                        continue
                    l2, _ = oracle.pusch_inner_rx_symbol_2l(P, s, cur, use, nvar, rx, h)
                    out.append(np.stack([l2[0].reshape(v, Qm), l2[1].reshape(v, Qm)], axis=1).reshape(-1))       # layer de-mapping (:1422-1428)
                ref = np.concatenate(out)
                if unscr is not None:
                    ref = oracle.unscramble_llr(ref, 0, unscr[1], unscr[0])
                assert llr.size == ref.size and np.array_equal(llr, ref), (N, nb_rx, rb_size, Qm, shift, unscr)


def test_pusch_inner_rx_fuzz(ldpc, oracle):
    """150 random single-layer PUSCH allocations / DMRS layouts on a 25-PRB carrier (the per-symbol oracle is swept against the reference's inner_rx on the same
    generator by tests/test_oracle_vs_reference.py::test_pusch_inner_rx_fuzz): whole-slot LLRs and the measured log2_maxh."""
    from common import ptrs_fuzz_cases
    rng = np.random.default_rng(88)
    done = 0
    for n, case in enumerate(ptrs_fuzz_cases(rng, 150, safe_tail=False)):
        N, nb_rx, rb_start, rb_size, Qm, dpos, dtype_, cdm, carrier, start, nsym = case[:11]
        if (dpos & (dpos << 1)) or ((dpos >> 13) & dpos & 1):
            continue                                   # "Double DMRS configuration is not yet supported" (get_nb_re_pusch)
        ay, ah = ((2000, 1500), (600, 900), (32767, 32767))[n % 3]
        rx = rng.integers(-ay, ay + 1, size=(nb_rx, 14, N, 2)).astype(np.int16)
        h = rng.integers(-ah, ah + 1, size=(nb_rx, 14, N, 2)).astype(np.int16)
        fco = N - carrier * 6
        P = PuschParms(N, nb_rx, rb_start, 0, rb_size, fco, Qm, dpos, dtype_, cdm)
        dm = [s for s in range(start, start + nsym) if (dpos >> s) & 1]
        with_data = [s for s in range(start, start + nsym) if oracle.pusch_nb_re(P, s) > 0]
        d = PuschRxDesc(N, nb_rx, rb_start, 0, rb_size, fco, Qm, start, nsym, dpos, dtype_, cdm, 0xFFFFFFFF, 0, 0, n & 1, 0x1234, 77)
        if not with_data:
            assert ldpc.pusch_num_llr(d) == 0
            continue
        meas = with_data[0]
        cur = dm[0] if meas < dm[0] else max(s for s in dm if s <= meas)
        sh_o, _ = oracle.pusch_log2_maxh(P, meas, cur, rx, h)
        llr, sh = ldpc.pusch_inner_rx_host(d, rx, h)
        ref = _oracle_slot(oracle, P, start, nsym, rx, h, sh_o, (0x1234, 77) if n & 1 else None)
        assert sh == sh_o and llr.size == ref.size and np.array_equal(llr, ref), (case[:11], sh, sh_o, np.nonzero(llr != ref)[0][:6] if llr.size == ref.size else (llr.size, ref.size))
        done += 1
    assert done > 100
