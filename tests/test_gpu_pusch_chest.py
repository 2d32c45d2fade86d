"""GPU parity of the PUSCH channel estimator (DMRS type 1, frequency-domain interpolation) against the CPU oracle, which
tests/test_oracle_vs_reference.py pins to the compiled reference nr_pusch_channel_estimation."""
import numpy as np
import pytest

from oracle.bindings import ChestParms
from openairinterface5g_b200.ldpc import PuschChestDesc

pytestmark = pytest.mark.gpu

CASES = [  # N, nb_rx, slot, symbol, port, rb_start, rb_size, carrier PRBs, scid, dmrs id, delay
    (4096, 4, 1, 2, 0, 0, 273, 273, 0, 77, 0), (4096, 2, 8, 2, 0, 0, 273, 273, 1, 1007, 3), (2048, 2, 3, 3, 0, 10, 50, 106, 0, 5, -2),
    (2048, 1, 19, 11, 1, 30, 76, 106, 0, 65535, 1), (1024, 4, 0, 2, 2, 0, 52, 52, 1, 0, 7), (1024, 2, 5, 0, 3, 20, 32, 52, 0, 300, -30), (512, 8, 2, 2, 0, 3, 11, 25, 0, 9, 0),
    (4096, 4, 11, 7, 0, 100, 173, 273, 0, 500, 12), (1536, 2, 4, 2, 0, 0, 78, 78, 0, 33, -5),
]


def test_pusch_chest_vs_oracle(ldpc, oracle):
    rng = np.random.default_rng(61)
    for N, nb_rx, slot, symbol, port, rb_start, rb_size, carrier, scid, nid, delay in CASES:
        fco = N - carrier * 6
        P = ChestParms(N, nb_rx, slot, symbol, port, rb_start, 0, rb_size, fco, scid, nid)
        d = PuschChestDesc(N, nb_rx, slot, symbol, port, rb_start, 0, rb_size, fco, scid, nid, 14 * N, 14 * N, 1)
        big = nb_rx == 8
        if big:
            rx = rng.integers(-32768, 32768, size=(nb_rx, 14, N, 2)).astype(np.int16)
        else:
            rx = rng.integers(-300, 301, size=(nb_rx, 14, N, 2)).astype(np.int16)
            pil = oracle.pusch_dmrs_pilots(P).reshape(-1, 2).astype(np.float64)
            k0 = (rb_start * 12 + fco) % N
            idx = (k0 + 2 * np.arange(6 * rb_size) + ((port >> 1) & 1)) % N
            for a in range(nb_rx):
                h = (900 + 100 * a) * np.exp(1j * (0.3 * a - 2 * np.pi * delay * np.arange(6 * rb_size) * 2 / N))
                y = h * (pil[:, 0] - 1j * pil[:, 1]) / 32767.0
                rx[a, symbol, idx, 0] += np.round(y.real).astype(np.int16); rx[a, symbol, idx, 1] += np.round(y.imag).astype(np.int16)
        est_o, out_o = oracle.pusch_channel_estimation(P, rx)
        prev = rng.integers(-5, 6, size=rx.shape).astype(np.int16)                     # stale estimates: the DMRS symbol must be fully rewritten
        est, st = ldpc.pusch_chest_host(d, rx, prev.copy())
        assert np.array_equal(st, out_o), (N, nb_rx, slot, symbol, port, st, out_o)
        assert np.array_equal(est[:, symbol], est_o[:, symbol]), (N, nb_rx, slot, symbol, port)
        other = [s for s in range(14) if s != symbol]
        assert np.array_equal(est[:, other], prev[:, other])                          # nothing else is touched


def test_pusch_chest_two_ports_in_one_call(ldpc, oracle):
    """n_ports = 2: both DMRS ports of a 2-layer PUSCH in the same launches == two single-port estimations."""
    rng = np.random.default_rng(62)
    for N, nb_rx, slot, symbol, port, rb_start, rb_size, carrier, scid, nid in ((4096, 4, 1, 2, 0, 0, 273, 273, 0, 77), (2048, 2, 3, 3, 2, 10, 50, 106, 1, 5), (1024, 8, 9, 2, 0, 4, 40, 52, 0, 900)):
        fco = N - carrier * 6
        rx = rng.integers(-2500, 2501, size=(nb_rx, 14, N, 2)).astype(np.int16)
        d = PuschChestDesc(N, nb_rx, slot, symbol, port, rb_start, 0, rb_size, fco, scid, nid, 14 * N, 14 * N, 2)
        est, st = ldpc.pusch_chest_host(d, rx)
        assert est.shape[0] == 2 * nb_rx and st.size == 10
        for q in range(2):
            est_o, out_o = oracle.pusch_channel_estimation(ChestParms(N, nb_rx, slot, symbol, port + q, rb_start, 0, rb_size, fco, scid, nid), rx)
            assert np.array_equal(st[5 * q:5 * q + 5], out_o), (N, nb_rx, q, st, out_o)
            assert np.array_equal(est[q * nb_rx:(q + 1) * nb_rx, symbol], est_o[:, symbol]), (N, nb_rx, q)


def test_pdsch_chest_ue_side_vs_oracle(ldpc, oracle):
    """pdsch_ue = 1: the UE's PDSCH estimator (same kernels, the UE's least-squares arithmetic)."""
    rng = np.random.default_rng(64)
    for N, nb_rx, slot, symbol, port, rb_start, rb_size, carrier, scid, nid, delay in CASES:
        fco = N - carrier * 6
        P = ChestParms(N, nb_rx, slot, symbol, port, rb_start, 0, rb_size, fco, scid, nid)
        rx = rng.integers(-32768, 32768, size=(nb_rx, 14, N, 2)).astype(np.int16) if nb_rx == 8 else rng.integers(-3000, 3001, size=(nb_rx, 14, N, 2)).astype(np.int16)
        d = PuschChestDesc(N, nb_rx, slot, symbol, port, rb_start, 0, rb_size, fco, scid, nid, 14 * N, 14 * N, 1, 1)
        est, st = ldpc.pusch_chest_host(d, rx)
        est_o = oracle.pdsch_channel_estimation(P, rx)
        assert np.array_equal(est[:, symbol], est_o[:, symbol]), (N, nb_rx, slot, symbol, port)
        assert st[0] == 0 and st[1] == 0
