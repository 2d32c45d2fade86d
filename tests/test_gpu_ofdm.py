"""GPU parity of the slot-level OFDM front end (nrb200_ofdm_{mod,demod}_slot_*) against the CPU oracle, which is pinned bit-exactly to
the compiled reference (apply_nr_rotation_TX + PHY_ofdm_mod; nr_slot_fep_ul + apply_nr_rotation_RX) in test_oracle_vs_reference.py."""
import numpy as np
import pytest

from openairinterface5g_b200.dfts import load_dftslib
from openairinterface5g_b200.ofdm import NrOfdmParms

pytestmark = pytest.mark.gpu

CASES = [(4096, 1, 273, 0), (4096, 1, 273, 3), (2048, 1, 106, 1), (1024, 0, 52, 2), (1536, 1, 78, 4), (512, 0, 25, 0), (3072, 1, 162, 2), (2048, 2, 66, 5),
         (6144, 1, 273, 1), (8192, 2, 264, 7), (256, 0, 11, 1)]


def _txF(rng, N, nb_rb, na, amp):
    F = np.zeros((na, 14, N, 2), np.int16)
    F[:, :, :nb_rb * 6] = rng.integers(-amp, amp + 1, size=(na, 14, nb_rb * 6, 2))
    F[:, :, N - nb_rb * 6:] = rng.integers(-amp, amp + 1, size=(na, 14, nb_rb * 6, 2))
    return F


def test_ofdm_mod_slot_vs_oracle(oracle):
    dl = load_dftslib()
    rng = np.random.default_rng(30)
    for N, mu, nb_rb, slot in CASES:
        P = NrOfdmParms(N, mu, nb_rb)
        rot = P.symbol_rotation(3619200000.0)
        na = 2
        F = _txF(rng, N, nb_rb, na, 32767 if slot == 3 else 4000)
        for use_rot in (True, False):
            y = dl.ofdm_mod_slot_host(P, slot, F.reshape(na, -1), rot if use_rot else None)
            for a in range(na):
                y_o, _ = oracle.ofdm_tx_slot(N, mu, nb_rb, slot, 14, rot.reshape(-1) if use_rot else None, F[a].reshape(-1))
                assert np.array_equal(y[a], y_o), (N, mu, nb_rb, slot, use_rot, a)


def test_ofdm_demod_slot_vs_oracle(oracle):
    dl = load_dftslib()
    rng = np.random.default_rng(31)
    for N, mu, nb_rb, slot in CASES:
        P = NrOfdmParms(N, mu, nb_rb)
        rot = P.symbol_rotation(3609200000.0)
        na = 2
        amp = 32767 if slot == 3 else 3000
        rx = rng.integers(-amp, amp + 1, size=(na, 2 * P.samples_per_frame)).astype(np.int16)
        for div, ta in ((8, 0), (8, N // 8), (4, N // 32 + 3)):
            Pd = NrOfdmParms(N, mu, nb_rb, div)
            for use_rot in (True, False):
                y = dl.ofdm_demod_slot_host(Pd, slot, rx, rot if use_rot else None, sample_offset=ta)
                for a in range(na):
                    y_o = oracle.ofdm_rx_slot(N, mu, nb_rb, slot, div, ta, rot.reshape(-1) if use_rot else None, rx[a])
                    assert np.array_equal(y[a], y_o), (N, mu, nb_rb, slot, div, ta, use_rot, a)


def test_ofdm_loopback_property():
    """TX slot -> place in a frame -> RX slot recovers the sub-carriers up to the Q15 rounding of two transforms (size-independent
    property at the full 100 MHz configuration; exact equality is covered by the oracle comparisons above)."""
    dl = load_dftslib()
    rng = np.random.default_rng(32)
    N, mu, nb_rb, slot = 4096, 1, 273, 4
    P = NrOfdmParms(N, mu, nb_rb)
    F = _txF(rng, N, nb_rb, 1, 2000)          # small enough that no intermediate stage of the Q15 transforms saturates
    y = dl.ofdm_mod_slot_host(P, slot, F.reshape(1, -1), None)
    frame = np.zeros((1, 2 * P.samples_per_frame), np.int16)
    ss = P.slot_timestamp(slot)
    frame[0, 2 * ss:2 * ss + y.shape[1]] = y[0]
    # window start exactly at the useful part: divisor large enough that the back-off is zero samples
    Pd = NrOfdmParms(N, mu, nb_rb, ofdm_offset_divisor=10 ** 6)
    G = dl.ofdm_demod_slot_host(Pd, slot, frame, None).reshape(14, N, 2).astype(np.int32)
    # idft4096 and dft4096 each scale by 1/64: G ~ F / 4096 * N / ... -> overall F/1 * (1/64 * 1/64 * N) = F
    err = np.abs(G - F[0].astype(np.int32))
    assert err.max() <= 128, err.max()
