"""GPU parity of the fused PDSCH transmitter kernel (scrambling + QAM mapping + layer mapping + DMRS + resource mapping + identity precoding, one launch)
against the CPU oracle, which tests/test_oracle_vs_reference.py pins to the reference's own nr_generate_pdsch."""
import numpy as np
import pytest

from oracle.bindings import PdschTxParms
from openairinterface5g_b200.ldpc import PdschTxDesc

pytestmark = pytest.mark.gpu

CASES = [  # N, carrier PRBs, nb_tx, slot, rb_start, rb_size, Qm, layers, start_symbol, nr_symbols, dmrs_pos, dmrs_type, cdm groups, dmrs_ports, scid, amp
    (4096, 273, 2, 1, 0, 273, 6, 2, 1, 13, 1 << 2, 0, 1, 0b0011, 0, 512), (4096, 273, 4, 7, 0, 273, 8, 1, 1, 13, 1 << 2, 0, 2, 0b0001, 0, 512),
    (2048, 106, 2, 3, 10, 50, 4, 2, 1, 13, (1 << 2) | (1 << 11), 0, 2, 0b1100, 1, 700), (2048, 106, 4, 19, 30, 76, 2, 4, 2, 10, 1 << 3, 0, 2, 0b1111, 0, 1000),
    (1024, 52, 2, 5, 0, 52, 6, 2, 1, 13, 1 << 2, 1, 1, 0b000011, 0, 512), (2048, 106, 4, 0, 20, 31, 4, 3, 2, 12, 1 << 2, 1, 2, 0b001101, 1, 300),
    (512, 25, 1, 9, 3, 11, 8, 1, 1, 6, 1 << 1, 0, 1, 0, 0, 512), (512, 25, 2, 11, 0, 25, 6, 2, 0, 14, (1 << 2) | (1 << 3), 0, 2, 0b0101, 0, 2047),
    (1536, 79, 2, 2, 0, 79, 6, 2, 1, 13, (1 << 2) | (1 << 7) | (1 << 11), 1, 3, 0b110000, 0, 512), (4096, 273, 4, 4, 0, 273, 8, 4, 1, 13, 1 << 2, 0, 2, 0b1111, 0, 512),
]


def test_pdsch_tx_vs_oracle(ldpc, oracle):
    rng = np.random.default_rng(71)
    for N, carrier, ntx, slot, rb0, nrb, Qm, nl, s0, ns, dpos, dtype_, cdm, ports, scid, amp in CASES:
        fco = N - carrier * 6
        P = PdschTxParms(N, ntx, slot, rb0, 0, nrb, fco, Qm, nl, s0, ns, dpos, dtype_, cdm, ports, scid, 40 + slot, 501, 0x1234, amp)
        d = PdschTxDesc(N, ntx, slot, rb0, 0, nrb, fco, Qm, nl, s0, ns, dpos, dtype_, cdm, ports, scid, 40 + slot, 501, 0x1234, amp, 0)
        assert ldpc.pdsch_tx_num_bits(d) == P.G()
        bits = rng.integers(0, 2, size=P.G(), dtype=np.uint8)
        want = oracle.pdsch_tx_slot(P, bits)
        got = ldpc.pdsch_tx_slot_host(d, bits)
        assert np.array_equal(got, want), (N, nrb, Qm, nl, dpos, dtype_, cdm, ports, [tuple(x) for x in np.argwhere(got != want)[:6]])
        # REs outside the allocation keep what the caller had there
        pre = rng.integers(-100, 100, size=want.shape).astype(np.int16)
        got2 = ldpc.pdsch_tx_slot_host(d, bits, pre)
        mask = np.zeros(want.shape[1:3], bool)
        ks = (fco + rb0 * 12 + np.arange(nrb * 12)) % N
        mask[s0:s0 + ns, :][:, ks] = True
        assert np.array_equal(got2[:, mask], want[:, mask]) and np.array_equal(got2[:, ~mask], pre[:, ~mask])


def test_pdsch_tx_rejects_unsupported(ldpc):
    bad = PdschTxDesc(1024, 4, 0, 20, 0, 31, 1024 - 52 * 6, 4, 3, 2, 12, 1 << 2, 1, 2, 0b001101, 1, 40, 501, 0x1234, 300, 0)   # the reference's over-mapping case
    assert ldpc.pdsch_tx_num_bits(bad) == 0


def test_pdsch_tx_wideband_precoding_vs_oracle(ldpc, oracle):
    """pm_idx > 0: the precoding stage of the fused kernel (saturating accumulation for RB pairs inside the symbol, wrapping for the pair that reaches or crosses
    its last sub-carrier), DMRS types 1 and 2, 1-4 layers on 2-4 antennas, odd and even rb_size; then back to the identity with the same descriptor."""
    rng = np.random.default_rng(72)
    cases = [  # N, carrier, ntx, slot, rb0, nrb, Qm, nl, dmrs_type, cdm, ports, amp
        (4096, 273, 4, 1, 0, 273, 6, 2, 0, 2, 0b0011, 512), (4096, 273, 2, 3, 0, 272, 8, 2, 0, 2, 0b0011, 30000), (2048, 106, 4, 5, 10, 51, 4, 1, 0, 2, 0b0001, 700),
        (1024, 52, 4, 0, 3, 40, 6, 4, 0, 2, 0b1111, 30000), (1024, 52, 2, 7, 0, 26, 2, 2, 1, 1, 0b000011, 512), (512, 25, 4, 2, 2, 21, 6, 3, 1, 2, 0b001101, 20000),
        (512, 25, 4, 2, 0, 12, 8, 2, 0, 1, 0b0011, 30000),
    ]
    for N, carrier, ntx, slot, rb0, nrb, Qm, nl, dtype_, cdm, ports, amp in cases:
        fco = N - carrier * 6
        P = PdschTxParms(N, ntx, slot, rb0, 0, nrb, fco, Qm, nl, 1, 13, 1 << 2, dtype_, cdm, ports, 0, 40 + slot, 501, 0x1234, amp)
        d = PdschTxDesc(N, ntx, slot, rb0, 0, nrb, fco, Qm, nl, 1, 13, 1 << 2, dtype_, cdm, ports, 0, 40 + slot, 501, 0x1234, amp, 0)
        w = rng.integers(-32767, 32768, size=(4, 4, 2)).astype(np.int16)
        if amp < 1000:
            w //= 3
        P.set_precoding(2, w); d.set_precoding(2, w)
        bits = rng.integers(0, 2, size=P.G(), dtype=np.uint8)
        want = oracle.pdsch_tx_slot(P, bits)
        got = ldpc.pdsch_tx_slot_host(d, bits)
        assert np.array_equal(got, want), (N, nrb, Qm, nl, ntx, [tuple(x) for x in np.argwhere(got != want)[:6]])
        assert np.count_nonzero(got[ntx - 1]) > 0
        P.set_precoding(0, None); d.set_precoding(0, None)
        assert np.array_equal(ldpc.pdsch_tx_slot_host(d, bits), oracle.pdsch_tx_slot(P, bits))
    from openairinterface5g_b200.ldpc import Nrb200Error
    d1 = PdschTxDesc(512, 1, 2, 0, 0, 12, 362, 6, 1, 1, 13, 1 << 2, 0, 1, 0b0001, 0, 42, 501, 0x1234, 512, 0).set_precoding(1, np.ones((1, 1, 2), np.int16))
    with pytest.raises((Nrb200Error, ValueError)):                  # rejected when the descriptor is validated
        ldpc.pdsch_tx_slot_host(d1, np.zeros(10, np.uint8))          # "No precoding can be done with a single antenna port"


PTRS = [(0, 2, 0), (1, 4, 2), (2, 2, 5), (1, 2, 11), (2, 4, 1), (0, 4, 0), (1, 2, 3), (2, 2, 0), (1, 4, 7), (1, 2, 1)]     # per CASES entry: L (log2), K, PTRSReOffset


def test_pdsch_tx_ptrs_vs_oracle(ldpc, oracle):
    """PT-RS insertion (ptrs = 1): pilots on every layer of the PT-RS symbols, data skipping them with the truncating scaling, fewer bits consumed -- still one launch;
    identity and wideband precoding.  The oracle is pinned to nr_generate_pdsch with pduBitmap & 1 (tests/test_oracle_vs_reference.py::test_pdsch_tx_slot_ptrs)."""
    rng = np.random.default_rng(74)
    for (N, carrier, ntx, slot, rb0, nrb, Qm, nl, s0, ns, dpos, dtype_, cdm, ports, scid, amp), (L, K, reoff) in zip(CASES, PTRS):
        fco = N - carrier * 6
        for pm in (0, 1):
            if pm and (ntx < 2 or ntx > 4):
                continue
            P = PdschTxParms(N, ntx, slot, rb0, 0, nrb, fco, Qm, nl, s0, ns, dpos, dtype_, cdm, ports, scid, 40 + slot, 501, 0x1234 + slot, amp).set_ptrs(L, K, reoff)
            d = PdschTxDesc(N, ntx, slot, rb0, 0, nrb, fco, Qm, nl, s0, ns, dpos, dtype_, cdm, ports, scid, 40 + slot, 501, 0x1234 + slot, amp, 0).set_ptrs(L, K, reoff)
            if pm:
                w = rng.integers(-12000, 12001, size=(4, 4, 2)).astype(np.int16)
                P.set_precoding(2, w); d.set_precoding(2, w)
            assert ldpc.pdsch_tx_num_bits(d) == P.G()
            bits = rng.integers(0, 2, size=P.G(), dtype=np.uint8)
            want = oracle.pdsch_tx_slot(P, bits)
            got = ldpc.pdsch_tx_slot_host(d, bits)
            assert np.array_equal(got, want), (N, nrb, Qm, nl, dpos, L, K, reoff, pm, [tuple(x) for x in np.argwhere(got != want)[:6]])
    bad = PdschTxDesc(512, 1, 2, 0, 0, 12, 362, 6, 1, 1, 13, 1 << 2, 0, 1, 0b0001, 0, 42, 501, 0x1234, 512, 0).set_ptrs(1, 3, 0)
    assert ldpc.pdsch_tx_num_bits(bad) == 0


def test_pdsch_tx_ptrs_golden(ldpc):
    """The same path against the committed vectors of the compiled reference (tests/golden/ptrs.npz, tools/gen_golden_ptrs.py)."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ptrs.npz"))
    for j in range(int(g["n_tx"])):
        N, carrier, ntx, slot, rb0, nrb, Qm, nl, s0, ns, dpos, dtype_, cdm, ports, scid, amp, L, K, reoff, pm = [int(x) for x in g[f"tx_case{j}"]]
        d = PdschTxDesc(N, ntx, slot, rb0, 0, nrb, N - carrier * 6, Qm, nl, s0, ns, dpos, dtype_, cdm, ports, scid, 40 + slot, 501, 0x1234 + slot, amp, 0).set_ptrs(L, K, reoff)
        if pm:
            d.set_precoding(pm, g[f"tx_w{j}"])
        got = ldpc.pdsch_tx_slot_host(d, g[f"tx_bits{j}"])
        assert np.array_equal(got, g[f"tx_out{j}"]), (j, [tuple(x) for x in np.argwhere(got != g[f"tx_out{j}"])[:6]])


def test_pdsch_tx_fuzz(ldpc, oracle):
    """250 random transmitter configurations (1-4 layers, PT-RS on / off, wideband precoding on / off; the same generator sweeps the oracle against the real
    nr_generate_pdsch on the CPU): txdataF bit for bit, and the library declines exactly the configurations the oracle declines."""
    import ctypes as C
    from common import pdsch_tx_fuzz_cases
    rng = np.random.default_rng(90)
    oracle.lib.orc_pdsch_tx_slot.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    done = 0
    for N, carrier, ntx, slot, rb0, nrb, Qm, nl, s0, ns, dpos, dtype_, cdm, ports, scid, amp, ptrs, pm in pdsch_tx_fuzz_cases(rng, 250):
        fco = N - carrier * 6
        P = PdschTxParms(N, ntx, slot, rb0, 0, nrb, fco, Qm, nl, s0, ns, dpos, dtype_, cdm, ports, scid, 40 + slot, 501, 0x1234 + slot, amp)
        d = PdschTxDesc(N, ntx, slot, rb0, 0, nrb, fco, Qm, nl, s0, ns, dpos, dtype_, cdm, ports, scid, 40 + slot, 501, 0x1234 + slot, amp, 0)
        if ptrs:
            P.set_ptrs(*ptrs); d.set_ptrs(*ptrs)
        if pm:
            w = rng.integers(-12000, 12001, size=(4, 4, 2)).astype(np.int16)
            P.set_precoding(pm, w); d.set_precoding(pm, w)
        if P.G() <= 0:
            continue
        bits = rng.integers(0, 2, size=P.G(), dtype=np.uint8)
        want = np.zeros((ntx, 14, N, 2), np.int16)
        rc = oracle.lib.orc_pdsch_tx_slot(C.addressof(P), bits.ctypes.data, want.ctypes.data)
        if rc < 0:
            assert ldpc.pdsch_tx_num_bits(d) == 0, (N, nrb, nl, dpos, dtype_, cdm)
            continue
        assert ldpc.pdsch_tx_num_bits(d) == P.G() == rc
        got = ldpc.pdsch_tx_slot_host(d, bits)
        assert np.array_equal(got, want), (N, nrb, Qm, nl, dpos, dtype_, cdm, ptrs, pm, [tuple(x) for x in np.argwhere(got != want)[:6]])
        done += 1
    assert done > 150
