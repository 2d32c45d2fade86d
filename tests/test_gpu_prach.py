"""GPU parity of the gNB PRACH detector (rx_nr_prach, NR_TRANSPORT/nr_prach.c) against the CPU oracle, which tests/test_oracle_vs_reference.py pins
to the compiled reference.  The root sequences are the oracle-side reference's own (compute_nr_prach_seq) where /root/reference was present at build time, committed as a
golden fixture otherwise."""
import os
import numpy as np
import pytest

from common import PRACH_CASES, prach_inputs, prach_num_roots
from openairinterface5g_b200.ldpc import PrachDesc

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "prach.npz")


def test_rx_nr_prach_vs_oracle_and_golden(ldpc, oracle):
    g = np.load(GOLD)
    rng = np.random.default_rng(81)
    for i, case in enumerate(PRACH_CASES):
        nb_rx, short, root, NCS, fmt, mu, pre, delay, amp, sigma = case
        xu = g[f"xu{i}"]
        rx = prach_inputs(rng, case, xu)
        d = PrachDesc(nb_rx, short, NCS, fmt, mu, 0, 0, 0)
        assert ldpc.prach_num_roots(d) == prach_num_roots(short, NCS) == int(np.count_nonzero(np.abs(xu).sum(axis=(1, 2))))
        want = oracle.rx_nr_prach(nb_rx, short, NCS, fmt, mu, xu, rx)
        got = ldpc.rx_nr_prach_host(d, xu, rx)
        assert got == want, (case, got, want)
        assert np.array_equal(rx, g[f"rx{i}"]) and got == tuple(int(v) for v in g[f"out{i}"]), (case, got, g[f"out{i}"])     # the compiled reference's own answer
        if pre >= 0 and sigma * 3 < amp <= 20000:
            assert got[0] == pre, (case, got)


def test_rx_nr_prach_device_resident_many_occasions(ldpc, oracle):
    """Device-resident entry point on torch tensors: 32 occasions of random noise + preambles back to back on one stream, every answer the oracle's."""
    import torch
    g = np.load(GOLD)
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(82)
    case0 = PRACH_CASES[1]
    xu = g["xu1"]
    d = PrachDesc(case0[0], case0[1], case0[3], case0[4], case0[5], 0, 839, 0)
    d_xu = torch.from_numpy(xu).to(dev)
    scratch = torch.empty(ldpc.prach_scratch_bytes(d), dtype=torch.uint8, device=dev)
    outs, wants = [], []
    for k in range(32):
        case = case0[:6] + (int(rng.integers(0, 64)), int(rng.integers(0, 10)), 900, 250)
        rx = prach_inputs(rng, case, xu)
        o = torch.zeros(3, dtype=torch.int32, device=dev)
        ldpc.rx_nr_prach_torch(d, d_xu, torch.from_numpy(rx).to(dev), o, scratch)
        outs.append(o); wants.append(oracle.rx_nr_prach(case[0], case[1], case[3], case[4], case[5], xu, rx))
        assert wants[-1][0] == case[6]
    torch.cuda.synchronize()
    assert [tuple(int(v) for v in o.cpu()) for o in outs] == wants


def test_rx_nr_prach_rejects_restricted_sets(ldpc):
    d = PrachDesc(2, 0, 13, 0, 1, 1, 0, 0)
    assert ldpc.prach_num_roots(d) == 0
    with pytest.raises(Exception):
        ldpc.rx_nr_prach_host(d, np.zeros((64, 839, 2), np.int16), np.zeros((2, 839, 2), np.int16))


def test_rx_nr_prach_fuzz(ldpc, oracle):
    """Random occasions on the committed root-sequence sets (X_u depends on sequence length, root index and N_CS only): antennas, formats, numerologies, sent preamble,
    delay, amplitude and noise drawn at random, 12 per set, against the oracle (which the CPU suite sweeps against the real rx_nr_prach on 150 random occasions)."""
    g = np.load(GOLD)
    rng = np.random.default_rng(94)
    for i, base in enumerate(PRACH_CASES):
        xu = g[f"xu{i}"]
        short, root, NCS = base[1], base[2], base[3]
        for _ in range(12):
            fmt = int(rng.integers(4, 13)) if short else int(rng.integers(0, 4))
            amp = int(rng.choice([400, 1500, 6000, 32767]))
            case = (int(rng.integers(1, 5)), short, root, NCS, fmt, int(rng.integers(0, 4)), int(rng.integers(-1, 64)), int(rng.integers(0, 8)), amp,
                    int(rng.choice([0, amp // 8, amp // 2])))
            rx = prach_inputs(rng, case, xu)
            want = oracle.rx_nr_prach(case[0], short, NCS, fmt, case[5], xu, rx)
            got = ldpc.rx_nr_prach_host(PrachDesc(case[0], short, NCS, fmt, case[5], 0, 0, 0), xu, rx)
            assert got == want, (case, got, want)
