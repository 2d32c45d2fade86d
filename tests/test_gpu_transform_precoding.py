"""GPU parity of the DFT-s-OFDM (transform precoding) variants of the gNB receiver against the CPU oracle, which tests/test_oracle_vs_reference.py pins to the
compiled reference: nr_pusch_channel_estimation with low-PAPR type-1 pilots (nr_ul_channel_estimation.c:122-133) and the one-layer inner receiver with
nr_freq_equalization + nr_idft between compensation and LLRs (nr_ulsch_demodulation.c:1326-1336, :16-265)."""
import os
import numpy as np
import pytest

from oracle.bindings import ChestParms, PuschParms
from openairinterface5g_b200.ldpc import PuschChestDesc, PuschRxDesc

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "transform_precoding.npz")
CARRIER = {512: 25, 1024: 52, 2048: 106, 4096: 273}


def _seq(ldpc, u, n_re):
    """the caller's sequence: computed for n_re = 30 and >= 36, from the reference-generated fixture for the table-driven lengths"""
    s = ldpc.lowpapr_sequence(u, 0, n_re)
    return s if s is not None else np.load(GOLD)[f"seq_{n_re}"][u].copy()


def test_lowpapr_sequence_matches_reference_vectors_and_oracle(ldpc, oracle):
    g = np.load(GOLD)
    for M, u in ((30, 4), (36, 0), (150, 17), (1620, 29)):
        assert np.array_equal(ldpc.lowpapr_sequence(u, 0, M), g[f"seq_{M}_u{u}"]), (M, u)
    for M in (30, 36, 48, 300, 3240):
        for u in (0, 13, 29):
            for v in (0, 1):
                assert np.array_equal(ldpc.lowpapr_sequence(u, v, M), oracle.lowpapr_seq(u, v, M)), (M, u, v)
    assert ldpc.lowpapr_sequence(0, 0, 24) is None and ldpc.lowpapr_sequence(30, 0, 36) is None


@pytest.mark.parametrize("chest_freq", [0, 1])
def test_chest_with_lowpapr_pilots_vs_oracle(ldpc, oracle, chest_freq):
    import torch
    rng = np.random.default_rng(71 + chest_freq)
    for N, nb_rx, slot, symbol, port, rb_start, rb_size, u in ((4096, 4, 1, 2, 0, 0, 270, 0), (2048, 2, 3, 3, 0, 10, 25, 5), (1024, 1, 0, 2, 1, 7, 6, 29), (2048, 2, 2, 2, 0, 0, 5, 11),
                                                               (1024, 2, 4, 2, 0, 3, 2, 3), (512, 3, 7, 11, 2, 1, 4, 20)):
        fco = N - CARRIER[N] * 6
        seq = _seq(ldpc, u, 6 * rb_size)
        rx = rng.integers(-3000, 3001, size=(nb_rx, 14, N, 2)).astype(np.int16)
        P = ChestParms(N, nb_rx, slot, symbol, port, rb_start, 0, rb_size, fco, 0, 55, 0, chest_freq)
        oracle.chest_set_lowpapr(seq)
        try:
            est_o, out_o = oracle.pusch_channel_estimation(P, rx)
        finally:
            oracle.chest_set_lowpapr(None)
        d = PuschChestDesc(N, nb_rx, slot, symbol, port, rb_start, 0, rb_size, fco, 0, 55, 14 * N, 14 * N, 1, 0, 0, chest_freq).set_lowpapr(seq)
        assert np.array_equal(ldpc.pusch_dmrs_pilots(d), np.stack([seq[0::2], -seq[1::2]], axis=1).reshape(-1))
        est, st = ldpc.pusch_chest_host(d, rx)
        assert np.array_equal(st, out_o), (N, rb_size, st, out_o)
        assert np.array_equal(est[:, symbol], est_o[:, symbol]), (N, rb_size)
        # device entry point: the sequence is the caller's device buffer
        dev = torch.device("cuda", 0)
        dseq = torch.from_numpy(seq).to(dev)
        dd = PuschChestDesc(N, nb_rx, slot, symbol, port, rb_start, 0, rb_size, fco, 0, 55, 14 * N, 14 * N, 1, 0, 0, chest_freq).set_lowpapr(dseq)
        t_rx = torch.from_numpy(rx).to(dev)
        t_est = torch.zeros((nb_rx, 14, N, 2), dtype=torch.int16, device=dev)
        scratch = torch.empty(ldpc.pusch_chest_scratch_bytes(dd), dtype=torch.uint8, device=dev)
        state = torch.zeros(18, dtype=torch.int32, device=dev)
        ldpc.pusch_chest_torch(dd, t_rx, t_est, scratch, state)
        torch.cuda.synchronize()
        assert np.array_equal(t_est.cpu().numpy()[:, symbol], est_o[:, symbol]) and np.array_equal(state.cpu().numpy()[:5], out_o)


def _oracle_slot(oracle, P, rx, h, shift, unscr):
    out = [oracle.pusch_inner_rx_symbol(P, s, 2, shift, rx, h)[0] for s in range(14) if oracle.pusch_nb_re(P, s) > 0]
    llr = np.concatenate(out)
    return llr if unscr is None else oracle.unscramble_llr(llr, 0, unscr[1], unscr[0])


@pytest.mark.parametrize("N,nb_rx,rb_start,rb_size,Qm", [(2048, 2, 10, 25, 6), (4096, 4, 0, 270, 4), (1024, 1, 7, 6, 2), (2048, 2, 0, 5, 6), (1024, 2, 3, 1, 4), (4096, 2, 5, 128, 6),
                                                        (4096, 1, 0, 256, 2), (1024, 2, 3, 2, 6), (4096, 2, 0, 135, 4), (1024, 4, 0, 50, 6), (2048, 8, 6, 100, 4)])
def test_inner_rx_with_transform_precoding_vs_oracle(ldpc, oracle, N, nb_rx, rb_start, rb_size, Qm):
    import torch
    rng = np.random.default_rng(N + rb_size + Qm)
    fco = N - CARRIER[N] * 6
    P = PuschParms(N, nb_rx, rb_start, 0, rb_size, fco, Qm, 1 << 2, 0, 2)
    rx = rng.integers(-2000, 2001, size=(nb_rx, 14, N, 2)).astype(np.int16)
    h = rng.integers(-1500, 1501, size=(nb_rx, 14, N, 2)).astype(np.int16)
    h[:, :, ::7] //= 40                                                      # some weak groups: amp = 0 / small amp in the equaliser
    oracle.pusch_set_transform_precoding(1)
    try:
        sh_o, _ = oracle.pusch_log2_maxh(P, 0, 2, rx, h)
        for shift, unscr in ((0xFFFFFFFF, None), (7, (0x4321, 99))):
            d = PuschRxDesc(N, nb_rx, rb_start, 0, rb_size, fco, Qm, 0, 14, 1 << 2, 0, 2, shift, 0, 0, 0 if unscr is None else 1, 0 if unscr is None else unscr[0],
                            0 if unscr is None else unscr[1], 1, 0, 0, 0, 0, 0, 1, 0)
            llr, sh = ldpc.pusch_inner_rx_host(d, rx, h)
            use = sh_o if shift == 0xFFFFFFFF else 7
            assert sh == use
            ref = _oracle_slot(oracle, P, rx, h, use, unscr)
            assert llr.size == ref.size and np.array_equal(llr, ref), (N, rb_size, Qm, shift)
        # device entry point with the caller's scratch
        dev = torch.device("cuda", 0)
        d = PuschRxDesc(N, nb_rx, rb_start, 0, rb_size, fco, Qm, 0, 14, 1 << 2, 0, 2, 8, 14 * N, 14 * N, 1, 0x77, 5, 1, 0, 0, 0, 0, 0, 1, 0)
        scratch = torch.empty(ldpc.pusch_tp_scratch_bytes(d), dtype=torch.uint8, device=dev)
        d.d_tp_scratch = scratch.data_ptr()
        out = torch.empty(ldpc.pusch_num_llr(d), dtype=torch.int16, device=dev)
        ldpc.pusch_inner_rx_torch(d, torch.from_numpy(rx).to(dev), torch.from_numpy(h).to(dev), out)
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy(), _oracle_slot(oracle, P, rx, h, 8, (0x77, 5)))
    finally:
        oracle.pusch_set_transform_precoding(0)


def test_transform_precoding_flag_is_ignored_where_the_reference_ignores_it(ldpc, oracle):
    """256QAM and two layers never reach the equalisation / nr_idft step (inner_rx :1326): same LLRs as with the flag clear."""
    rng = np.random.default_rng(5)
    N, nb_rx, rb_size = 1024, 2, 24
    rx = rng.integers(-2000, 2001, size=(nb_rx, 14, N, 2)).astype(np.int16)
    h = rng.integers(-1500, 1501, size=(2 * nb_rx, 14, N, 2)).astype(np.int16)
    for Qm, nl in ((8, 1), (6, 2)):
        mk = lambda tp: PuschRxDesc(N, nb_rx, 4, 0, rb_size, N - 52 * 6, Qm, 0, 14, 1 << 2, 0, 2, 7, 0, 0, 0, 0, 0, nl, 50, 3000, 0, 0, 0, tp, 0)
        a, _ = ldpc.pusch_inner_rx_host(mk(0), rx, h[:nl * nb_rx])
        b, _ = ldpc.pusch_inner_rx_host(mk(1), rx, h[:nl * nb_rx])
        assert np.array_equal(a, b)


def test_transform_precoding_refuses_sizes_the_reference_cannot_do(ldpc):
    """12 * rb_size must be one of nr_idft's sizes; 768 and 2304 are excluded because the reference's own result is not reproducible there; data on
    the DMRS symbol (num_dmrs_cdm_grps_no_data = 1) has no transform size either.  The library never falls back: -4."""
    from openairinterface5g_b200.ldpc import Nrb200Error
    N = 4096
    rx = np.zeros((1, 14, N, 2), np.int16)
    for rb_size, cdm in ((7, 2), (64, 2), (192, 2), (11, 2), (25, 1)):
        d = PuschRxDesc(N, 1, 0, 0, rb_size, N - 273 * 6, 4, 0, 14, 1 << 2, 0, cdm, 7, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 1, 0)
        with pytest.raises(Nrb200Error):
            ldpc.pusch_inner_rx_host(d, rx, rx)
