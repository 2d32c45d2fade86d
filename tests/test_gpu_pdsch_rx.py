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
