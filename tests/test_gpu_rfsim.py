"""GPU parity of the rfsimulator channel application (rxAddInput, radio/rfsimulator/apply_channelmod.c) against the CPU oracle, which
tests/test_oracle_vs_reference.py pins to the compiled reference.  Double-precision arithmetic, bit-exact: every output sample of every receive antenna."""
import numpy as np
import pytest

from common import RFSIM_CASES, rfsim_inputs

pytestmark = pytest.mark.gpu


def test_rfsim_rx_add_input_vs_oracle(ldpc, oracle):
    rng = np.random.default_rng(78)
    for case in RFSIM_CASES:
        nb_tx, nb_rx, L, offset, pl, npw, n, TS, cf, amp = case
        cir, ch, sig, out, noise = rfsim_inputs(rng, case)
        out_all = np.stack([np.roll(out, 7 * a, axis=0) for a in range(nb_rx)])
        noise_all = rng.normal(size=(nb_rx, n, 2))
        for nz in (noise_all, None):
            want = np.stack([oracle.rfsim_rx_add_input(nb_tx, nb_rx, L, offset, pl, npw, ch, sig, out_all[a], a, TS, cir, None if nz is None else nz[a]) for a in range(nb_rx)])
            got = ldpc.rfsim_rx_add_input_host(nb_tx, nb_rx, offset, pl, npw, ch, sig, out_all, TS, nz)
            assert np.array_equal(got, want), (case, nz is None, np.argwhere(got != want)[:5])
            assert not np.array_equal(got, out_all)


def test_rfsim_device_resident_accumulates_two_peers(ldpc, oracle):
    """Two connected peers accumulate into the same output like the loop over sockets in simulator.c:960-981; device-resident entry point on torch tensors."""
    import torch
    from openairinterface5g_b200.ldpc import RfsimChan
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(79)
    nb_tx, nb_rx, L, n, TS = 2, 2, 40, 30720, 614400
    cir = 2 * (n + L + 16)
    out = np.zeros((nb_rx, n, 2), np.int16)
    want = out.copy()
    d_out = torch.zeros((nb_rx, n, 2), dtype=torch.int16, device=dev)
    for peer in range(2):
        ch = rng.normal(size=(nb_tx * nb_rx, L, 2)) * 0.1
        sig = rng.integers(-6000, 6001, size=(cir, 2)).astype(np.int16)
        noise = rng.normal(size=(nb_rx, n, 2))
        want = np.stack([oracle.rfsim_rx_add_input(nb_tx, nb_rx, L, peer, -2.0, -30.0, ch, sig, want[a], a, TS, cir, noise[a]) for a in range(nb_rx)])
        ldpc.rfsim_rx_add_input_torch(RfsimChan(nb_tx, nb_rx, L, peer, -2.0, -30.0, 0), torch.from_numpy(ch).to(dev), torch.from_numpy(sig).to(dev), d_out, TS,
                                      torch.from_numpy(noise).to(dev))
    torch.cuda.synchronize()
    assert np.array_equal(d_out.cpu().numpy(), want)


def test_rfsim_rejects_bad_descriptors(ldpc):
    z = np.zeros((4, 2), np.int16)
    with pytest.raises(Exception):
        ldpc.rfsim_rx_add_input_host(9, 1, 0, 0.0, 0.0, np.zeros((9, 1, 2)), z, np.zeros((1, 2, 2), np.int16), 10)
