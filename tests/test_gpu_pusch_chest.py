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


VARIANT_CASES = [  # N, nb_rx, slot, symbol, port, rb_start, rb_size, carrier PRBs, scid, dmrs id, delay
    (4096, 4, 4, 2, 0, 0, 273, 273, 0, 77, 2), (2048, 2, 8, 3, 1, 10, 50, 106, 1, 1007, -3), (1024, 3, 0, 11, 2, 20, 32, 52, 0, 300, 5),
    (1024, 2, 12, 2, 3, 0, 52, 52, 1, 0, 0), (512, 8, 16, 5, 0, 3, 11, 25, 0, 9, 1), (2048, 1, 4, 0, 0, 30, 2, 106, 0, 65535, 0),
    (1024, 2, 8, 13, 3, 0, 52, 52, 0, 41, 1),      # last symbol of the slot, ports 2 / 3: the shifted pointer reaches the symbol's end
]


def _variant_input(rng, oracle, P, N, nb_rx, symbol, port, rb_start, fco, dmrs_type, delay):
    if nb_rx == 8:
        return rng.integers(-32768, 32768, size=(nb_rx, 14, N, 2)).astype(np.int16)
    rx = rng.integers(-300, 301, size=(nb_rx, 14, N, 2)).astype(np.int16)
    npil = (4 if dmrs_type else 6) * P.rb_size
    pil = oracle.pusch_dmrs_pilots(P).reshape(-1, 2).astype(np.float64)[:npil]
    k0 = (rb_start * 12 + fco) % N
    n = np.arange(npil)
    idx = ((k0 + 2 * n) % N if dmrs_type == 0 else (k0 + 6 * (n // 2) + (n & 1)) % N) + ((port >> 1) & 1)
    keep = idx < N
    for a in range(nb_rx):
        h = (2000 + 300 * a) * np.exp(1j * (0.4 * a - 2 * np.pi * delay * (idx - k0) / N))
        y = h * (pil[:, 0] - 1j * pil[:, 1]) / 23170.0 / np.sqrt(2)
        rx[a, symbol, idx[keep], 0] += np.round(y.real).astype(np.int16)[keep]; rx[a, symbol, idx[keep], 1] += np.round(y.imag).astype(np.int16)[keep]
    return rx


@pytest.mark.parametrize("dmrs_type,chest_freq", [(1, 0), (0, 1), (1, 1)])
def test_pusch_chest_variants_vs_oracle(ldpc, oracle, dmrs_type, chest_freq):
    """DMRS type 2 (frequency-domain) and the per-PRB averages of both DMRS types, host entry point and device-resident entry point."""
    import torch
    rng = np.random.default_rng(70 + 2 * dmrs_type + chest_freq)
    for N, nb_rx, slot, symbol, port, rb_start, rb_size, carrier, scid, nid, delay in VARIANT_CASES:
        fco = N - carrier * 6
        P = ChestParms(N, nb_rx, slot, symbol, port, rb_start, 0, rb_size, fco, scid, nid, dmrs_type, chest_freq)
        d = PuschChestDesc(N, nb_rx, slot, symbol, port, rb_start, 0, rb_size, fco, scid, nid, 14 * N, 14 * N, 1, 0, dmrs_type, chest_freq)
        rx = _variant_input(rng, oracle, P, N, nb_rx, symbol, port, rb_start, fco, dmrs_type, delay)
        est_o, out_o = oracle.pusch_channel_estimation(P, rx)
        prev = rng.integers(-5, 6, size=rx.shape).astype(np.int16)
        est, st = ldpc.pusch_chest_host(d, rx, prev.copy())
        assert np.array_equal(st, out_o), (N, nb_rx, slot, symbol, port, st, out_o)
        assert np.array_equal(est[:, symbol], est_o[:, symbol]), (N, nb_rx, slot, symbol, port)
        other = [s for s in range(14) if s != symbol]
        assert np.array_equal(est[:, other], prev[:, other])
        # device-resident slot buffers
        t_rx = torch.from_numpy(rx).cuda()
        t_est = torch.from_numpy(prev.copy()).cuda()
        scratch = torch.empty(ldpc.pusch_chest_scratch_bytes(d), dtype=torch.uint8, device="cuda")
        state = torch.zeros(18, dtype=torch.int32, device="cuda")
        ldpc.pusch_chest_torch(d, t_rx, t_est, scratch, state)
        torch.cuda.synchronize()
        assert np.array_equal(state[:5].cpu().numpy(), out_o) and np.array_equal(t_est.cpu().numpy()[:, symbol], est_o[:, symbol]), (N, nb_rx, slot, symbol, "dev")


def test_pusch_chest_variants_refuse_what_the_reference_cannot_do(ldpc):
    from openairinterface5g_b200.ldpc import Nrb200Error
    rx = np.zeros((2, 14, 512, 2), np.int16)
    for kw in (dict(rb_size=1, chest_freq=1), dict(slot=5, dmrs_config_type=1, chest_freq=1), dict(n_ports=2, dmrs_config_type=1), dict(pdsch_ue=1, chest_freq=1, rb_size=1),
               dict(port=4, dmrs_config_type=1), dict(pdsch_ue=1, port=4)):
        f = dict(fft_size=512, nb_rx=2, slot=4, symbol=2, port=0, rb_start=0, bwp_start=0, rb_size=20, first_carrier_offset=362, scid=0, ul_dmrs_scrambling_id=7,
                 rx_stride=14 * 512, ch_stride=14 * 512, n_ports=1, pdsch_ue=0, dmrs_config_type=0, chest_freq=0)
        f.update(kw)
        with pytest.raises(Nrb200Error):
            ldpc.pusch_chest_host(PuschChestDesc(**f), rx)


def test_chest_time_domain_avg_vs_oracle(ldpc, oracle):
    """nr_chest_time_domain_avg: host and device entry points, 1-4 DMRS symbols, full-scale inputs (saturating sums)."""
    import torch
    rng = np.random.default_rng(78)
    for N, nb_rx, start, nsym, bitmap, nrb in ((512, 2, 0, 14, 0b00000000000100, 25), (512, 3, 0, 14, 0b00100000000100, 20), (1024, 2, 2, 12, 0b00101000001000, 52),
                                               (1024, 4, 0, 14, 0b00100100100100, 40), (4096, 4, 0, 14, 0b00100000000100, 273), (512, 2, 4, 8, 0b11000000110000, 11)):
        est = rng.integers(-32768, 32768, size=(nb_rx, 14, N, 2)).astype(np.int16)
        est[:, :, ::5] //= 50
        want = oracle.chest_time_domain_avg(est, nsym, start, bitmap, nrb)
        got, first = ldpc.chest_time_avg_host(est, nsym, start, bitmap, nrb)
        assert first == min(s for s in range(start, start + nsym) if (bitmap >> s) & 1)
        assert np.array_equal(got, want), (N, nb_rx, start, nsym, bin(bitmap), nrb)
        t = torch.from_numpy(est.copy()).cuda()
        assert ldpc.chest_time_avg_torch(t, nsym, start, bitmap, nrb) == first
        torch.cuda.synchronize()
        assert np.array_equal(t.cpu().numpy(), want)
    from openairinterface5g_b200.ldpc import Nrb200Error
    with pytest.raises(Nrb200Error):
        ldpc.chest_time_avg_host(est, 14, 0, 0, 11)             # no DMRS symbol: AssertFatal in the reference
    with pytest.raises(Nrb200Error):
        ldpc.chest_time_avg_host(est, 14, 0, 0b11111, 11)       # five DMRS symbols


@pytest.mark.parametrize("dmrs_type,chest_freq", [(1, 0), (0, 1), (1, 1)])
def test_pdsch_chest_ue_variants_vs_oracle(ldpc, oracle, dmrs_type, chest_freq):
    """pdsch_ue = 1 with DMRS type 2 (ports 0-5) and with the per-PRB averages: the UE's own walks over the pilots (oracle pinned to nr_pdsch_channel_estimation)."""
    rng = np.random.default_rng(80 + 2 * dmrs_type + chest_freq)
    cases = [(4096, 2, 4, 2, 0, 0, 273, 273, 0, 77), (2048, 2, 8, 3, 1, 10, 50, 106, 1, 1007), (1024, 3, 0, 11, 2, 20, 32, 52, 0, 300), (1024, 2, 12, 2, 3, 0, 52, 52, 1, 0),
             (512, 4, 16, 5, 0, 3, 11, 25, 0, 9), (2048, 1, 5, 0, 1, 30, 2, 106, 0, 65535), (1024, 2, 9, 13, 3, 0, 52, 52, 0, 41)]
    if dmrs_type == 1:
        cases += [(1024, 2, 3, 4, 4, 0, 52, 52, 0, 21), (1024, 2, 7, 6, 5, 8, 30, 52, 1, 22), (1024, 2, 7, 13, 5, 0, 52, 52, 1, 23)]
    for N, nb_rx, slot, symbol, port, rb_start, rb_size, carrier, scid, nid in cases:
        fco = N - carrier * 6
        P = ChestParms(N, nb_rx, slot, symbol, port, rb_start, 0, rb_size, fco, scid, nid, dmrs_type, chest_freq)
        d = PuschChestDesc(N, nb_rx, slot, symbol, port, rb_start, 0, rb_size, fco, scid, nid, 14 * N, 14 * N, 1, 1, dmrs_type, chest_freq)
        rx = rng.integers(-32768, 32768, size=(nb_rx, 14, N, 2)).astype(np.int16) if N == 512 else rng.integers(-3000, 3001, size=(nb_rx, 14, N, 2)).astype(np.int16)
        est, st = ldpc.pusch_chest_host(d, rx)
        est_o = oracle.pdsch_channel_estimation(P, rx)
        assert np.array_equal(est[:, symbol], est_o[:, symbol]), (N, nb_rx, slot, symbol, port, dmrs_type, chest_freq)
        assert st[0] == 0 and st[1] == 0


def test_channel_estimation_fuzz(ldpc, oracle):
    """150 random DMRS type 1 configurations (tests/common.py:chest_fuzz_cases; the CPU suite sweeps the oracle against the real estimators with the same generator):
    the gNB estimator (estimates + state) and the UE estimator, every third case with full-scale noise."""
    from common import chest_fuzz_cases, chest_inputs
    rng = np.random.default_rng(92)
    for n, (N, nb_rx, slot, symbol, port, rb_start, rb_size, carrier, scid, nid, delay) in enumerate(chest_fuzz_cases(rng, 150)):
        fco = N - carrier * 6
        P = ChestParms(N, nb_rx, slot, symbol, port, rb_start, 0, rb_size, fco, scid, nid)
        rx = chest_inputs(oracle, rng, P, port, delay, None if n % 3 == 2 else 300)
        est_o, out_o = oracle.pusch_channel_estimation(P, rx)
        est, st = ldpc.pusch_chest_host(PuschChestDesc(N, nb_rx, slot, symbol, port, rb_start, 0, rb_size, fco, scid, nid, 14 * N, 14 * N, 1), rx)
        assert np.array_equal(st, out_o) and np.array_equal(est[:, symbol], est_o[:, symbol]), ("gNB", N, nb_rx, slot, symbol, port, rb_start, rb_size, st, out_o)
        est, st = ldpc.pusch_chest_host(PuschChestDesc(N, nb_rx, slot, symbol, port, rb_start, 0, rb_size, fco, scid, nid, 14 * N, 14 * N, 1, 1), rx)
        assert np.array_equal(est[:, symbol], oracle.pdsch_channel_estimation(P, rx)[:, symbol]), ("UE", N, nb_rx, slot, symbol, port, rb_start, rb_size)
