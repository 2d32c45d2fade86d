"""GPU parity of rate matching / interleaving (TX) and de-interleaving / rate recovery / HARQ combine / decoder-input packing (RX)
against the CPU oracle (pinned to nr_rate_matching.c by tests/test_oracle_vs_reference.py) and the golden fixtures."""
import os
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

CASES = [  # BG, Z, F, E list, rv, Tbslbrm, C, Qm
    (1, 384, 0, [9072, 9072, 9078], 0, 0, 3, 6), (1, 384, 88, [9072, 9080], 2, 0, 2, 2), (1, 96, 40, [30000], 1, 0, 1, 4),
    (2, 128, 16, [5000, 5008], 3, 0, 2, 8), (1, 384, 0, [20000] * 4, 0, 200000, 20, 2), (2, 52, 0, [1200], 0, 0, 1, 4),
    (1, 208, 120, [4104, 4098], 3, 90000, 4, 6), (2, 384, 200, [19008], 2, 0, 2, 8), (1, 384, 0, [60000], 1, 0, 1, 2), (2, 16, 8, [400], 0, 0, 1, 2)]


def _oracle_tx(oracle, BG, Z, F, E, rv, Tb, C_, Qm, w):
    K = (22 if BG == 1 else 10) * Z
    rc, e = oracle.rate_matching_tx(Tb, BG, Z, w, C_, F, K - F - 2 * Z, rv, E)
    assert rc == 0
    return oracle.interleave(E, Qm, e)


@pytest.mark.parametrize("case", CASES)
def test_rm_tx_vs_oracle(ldpc, oracle, case):
    BG, Z, F, Es, rv, Tb, C_, Qm = case
    rng = np.random.default_rng(Z + F)
    N = (66 if BG == 1 else 50) * Z
    K = (22 if BG == 1 else 10) * Z
    Fo = K - F - 2 * Z
    d = rng.integers(0, 2, size=(len(Es), N), dtype=np.uint8)
    d_marked = d.copy()
    d_marked[:, Fo:Fo + F] = 2                                   # NR_NULL marks as nr_dlsch_coding.c:179 sets them
    d[:, Fo:Fo + F] = 0
    got = ldpc.rm_tx_host(BG, Z, Qm, rv, C_, Tb, F, d, Es)
    want = np.concatenate([_oracle_tx(oracle, BG, Z, F, E, rv, Tb, C_, Qm, d_marked[r]) for r, E in enumerate(Es)])
    assert np.array_equal(got, want)


@pytest.mark.parametrize("case", CASES)
def test_rm_rx_vs_oracle(ldpc, oracle, case):
    BG, Z, F, Es, rv, Tb, C_, Qm = case
    Es = [E - E % Qm for E in Es]
    rng = np.random.default_rng(Z + F + 1)
    N = (66 if BG == 1 else 50) * Z
    K = (22 if BG == 1 else 10) * Z
    kcz = (68 if BG == 1 else 52) * Z
    Fo = K - F - 2 * Z
    n = len(Es)
    soft = rng.integers(-200, 200, size=sum(Es), dtype=np.int16)
    harq_gpu = rng.integers(-3000, 3000, size=(n, N + 16), dtype=np.int16)
    harq_cpu = harq_gpu.copy()
    for clear in (1, 0, 0):                                      # first transmission then two soft-combining rounds
        llr = ldpc.rm_rx_host(BG, Z, Qm, rv, C_, Tb, F, soft, Es, harq_gpu, clear)
        off = 0
        for r, E in enumerate(Es):
            e = oracle.deinterleave(E, Qm, soft[off:off + E])
            assert oracle.rate_matching_rx(Tb, BG, Z, harq_cpu[r], e, C_, rv, clear, E, F, Fo) == 0
            off += E
            z = np.zeros(kcz, dtype=np.int16)                     # nr_ulsch_decoding.c:195-210
            z[K - F:K] = 127
            z[2 * Z:K - F] = harq_cpu[r][:K - F - 2 * Z]
            z[K:] = harq_cpu[r][K - 2 * Z:kcz - 2 * Z]
            assert np.array_equal(llr[r], np.clip(z, -128, 127).astype(np.int8)), (case, clear, r)
        assert np.array_equal(harq_gpu, harq_cpu), (case, clear)
        soft = rng.integers(-200, 200, size=sum(Es), dtype=np.int16)


def test_rm_golden(ldpc):
    d = np.load(os.path.join(G, "coding.npz"))
    for ci, (BG, Z, F, E, rv, Tb, Cs, Qm) in enumerate(d["rm_cases"].tolist()):
        K = (22 if BG == 1 else 10) * Z
        N = (66 if BG == 1 else 50) * Z
        w = d[f"rm{ci}_w"].copy()
        w[w == 2] = 0
        f = ldpc.rm_tx_host(BG, Z, Qm, rv, Cs, Tb, F, w[None, :], [E])
        assert np.array_equal(f, d[f"rm{ci}_f"]), ci
        harq = np.zeros((1, N + 16), dtype=np.int16)
        soft_il = np.zeros(E, dtype=np.int16)                     # fixtures hold the de-interleaved stream: re-interleave it
        EQm = E // Qm
        dei = d[f"rm{ci}_dei"]
        for i in range(Qm):
            soft_il[i::Qm][:EQm] = dei[i * EQm:(i + 1) * EQm]
        ldpc.rm_rx_host(BG, Z, Qm, rv, Cs, Tb, F, soft_il, [E], harq, 1)
        assert np.array_equal(harq[0, :N], d[f"rm{ci}_w1"]), ci
        ldpc.rm_rx_host(BG, Z, Qm, rv, Cs, Tb, F, soft_il, [E], harq, 0)
        assert np.array_equal(harq[0, :N], d[f"rm{ci}_w2"]), ci


def test_dl_ul_chain_roundtrip(ldpc, oracle):
    """TB-level property test at a BASELINE-like shape: segmentation CRC -> encode -> rate match -> interleave -> BPSK/AWGN ->
    de-interleave -> rate recover -> decode (CRC stop) recovers every segment; all codec steps on the GPU."""
    BG, Z, Qm, rv, C_ = 1, 384, 2, 0, 4
    K = 22 * Z
    rng = np.random.default_rng(12)
    P = rng.integers(0, 256, size=(C_, K // 8), dtype=np.uint8)
    for r in range(C_):
        crc = oracle.crc(1, P[r], K - 24) >> 8
        P[r, -3:] = [(crc >> 16) & 0xFF, (crc >> 8) & 0xFF, crc & 0xFF]
    cw = ldpc.encode_batch_host(BG, Z, K, P)
    Es = [20000] * C_
    f = ldpc.rm_tx_host(BG, Z, Qm, rv, C_, 0, 0, cw, Es)
    sigma = 0.55
    y = (1.0 - 2.0 * f.astype(np.float64)) + sigma * rng.standard_normal(f.size)
    soft = np.clip(np.floor(y * 8 / sigma / sigma), -127, 127).astype(np.int16)
    harq = np.zeros((C_, 66 * Z), dtype=np.int16)
    llr = ldpc.rm_rx_host(BG, Z, Qm, rv, C_, 0, 0, soft, Es, harq, 1)
    R = oracle.get_R(rv, Es[0], BG, Z, 0, 0)[0]
    iters, out = ldpc.decode_batch_host(BG, Z, R, 8, llr, use_crc=1, crc_len_bits=K, crc_type=1)
    assert (iters <= 8).all()
    assert np.array_equal(out[:, :K // 8], P)


def test_rm_fuzz(ldpc, oracle):
    """100 random rate-matching configurations (tests/common.py:rm_fuzz_cases; the CPU suite sweeps the oracle against nr_rate_matching.c with the same generator):
    bit selection + interleaving, and de-interleaving + rate recovery + soft combining + decoder-input packing over two rounds."""
    from common import rm_fuzz_cases
    rng = np.random.default_rng(97)
    done = 0
    for BG, Z, F, Es, rv, Tb, C_, Qm in rm_fuzz_cases(rng, 100):
        N = (66 if BG == 1 else 50) * Z
        K = (22 if BG == 1 else 10) * Z
        kcz = (68 if BG == 1 else 52) * Z
        Fo = K - F - 2 * Z
        n = len(Es)
        d = rng.integers(0, 2, size=(n, N), dtype=np.uint8)
        d_marked = d.copy()
        d_marked[:, Fo:Fo + F] = 2
        d[:, Fo:Fo + F] = 0
        rcs = [oracle.rate_matching_tx(Tb, BG, Z, d_marked[r], C_, F, Fo, rv, E)[0] for r, E in enumerate(Es)]
        if any(rcs):
            continue                                             # a combination nr_rate_matching_ldpc itself refuses
        got = ldpc.rm_tx_host(BG, Z, Qm, rv, C_, Tb, F, d, Es)
        want = np.concatenate([_oracle_tx(oracle, BG, Z, F, E, rv, Tb, C_, Qm, d_marked[r]) for r, E in enumerate(Es)])
        assert np.array_equal(got, want), (BG, Z, F, Es, rv, Tb, C_, Qm)
        soft = rng.integers(-200, 200, size=sum(Es), dtype=np.int16)
        harq_gpu = rng.integers(-3000, 3000, size=(n, N + 16), dtype=np.int16)
        harq_cpu = harq_gpu.copy()
        for clear in (1, 0):
            llr = ldpc.rm_rx_host(BG, Z, Qm, rv, C_, Tb, F, soft, Es, harq_gpu, clear)
            off = 0
            for r, E in enumerate(Es):
                e = oracle.deinterleave(E, Qm, soft[off:off + E])
                assert oracle.rate_matching_rx(Tb, BG, Z, harq_cpu[r], e, C_, rv, clear, E, F, Fo) == 0
                off += E
                z = np.zeros(kcz, dtype=np.int16)
                z[K - F:K] = 127
                z[2 * Z:K - F] = harq_cpu[r][:K - F - 2 * Z]
                z[K:] = harq_cpu[r][K - 2 * Z:kcz - 2 * Z]
                assert np.array_equal(llr[r], np.clip(z, -128, 127).astype(np.int8)), (BG, Z, F, Es, rv, Tb, C_, Qm, clear, r)
            assert np.array_equal(harq_gpu, harq_cpu), (BG, Z, F, Es, rv, Tb, C_, Qm, clear)
            soft = rng.integers(-200, 200, size=sum(Es), dtype=np.int16)
        done += 1
    assert done > 60
