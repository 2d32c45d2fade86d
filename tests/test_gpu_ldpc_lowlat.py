"""GPU parity of the low-latency decode path: the cluster kernel (one code block on 2 / 4 / 8 CTAs, ldpc_decoder_cluster.cuh) and the combining
per-call path behind LDPCdecoder (nrb200_ll.cu) against the CPU oracle -- output bytes and returned iteration counts, bit exact -- plus the
reference's abort semantics (nrLDPC_decoder.c:557-560, 190-193)."""
import ctypes as C
import os
import subprocess
import sys
import threading
import time
import numpy as np
import pytest
from common import NCOLS, make_case

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _check_batch(ldpc, oracle, BG, Z, R, n, ebn0, seed, max_iter=8, out_mode=0, crc=False):
    K, P, llr = make_case(oracle, BG, Z, R, n, ebn0, seed)
    kw = dict(use_crc=1, crc_len_bits=K, crc_type=1) if crc else {}
    if crc:                                                               # valid CRC24B in half of the blocks
        for i in range(0, n, 2):
            c = oracle.crc(1, P[i], K - 24) >> 8
            P[i, -3:] = [(c >> 16) & 255, (c >> 8) & 255, c & 255]
        from openairinterface5g_b200.synth import awgn_llr
        cw = np.stack([oracle.encode(BG, Z, K, P[i]) for i in range(n)])
        llr = awgn_llr(cw, Z, NCOLS[(BG, R)], ebn0, (22 if BG == 1 else 10) / (NCOLS[(BG, R)] - 2), seed)
    iters, out = ldpc.decode_batch_host(BG, Z, R, max_iter, llr, outMode=out_mode, latency_mode=1, **kw)
    for i in range(n):
        it_o, out_o = oracle.decode(BG, Z, R, max_iter, llr[i], out_mode, *( (1, K, 1) if crc else ()))
        assert iters[i] == it_o, (BG, Z, R, n, i, iters[i], it_o)
        assert np.array_equal(out[i].view(np.uint8), np.asarray(out_o).view(np.uint8)), (BG, Z, R, n, i)


@pytest.mark.parametrize("n", [1, 5, 16, 17, 30, 40, 74])
def test_cluster_sizes_headline_graph(ldpc, oracle, n):
    """latency_mode = 1: launch_decode picks 8 / 4 / 2 CTAs per block from the batch size (<= 15 / <= 33 / <= 74); every size against the oracle at the
    waterfall."""
    _check_batch(ldpc, oracle, 1, 384, 13, n, 2.2, seed=100 + n)


@pytest.mark.parametrize("BG,Z,R,ebn0", [(1, 384, 23, 4.0), (1, 384, 89, 7.0), (2, 384, 15, 1.0), (2, 384, 13, 2.5), (2, 384, 23, 5.0),
                                         (1, 256, 13, 2.4), (2, 256, 15, 1.2), (1, 128, 13, 2.6), (2, 128, 13, 3.0), (1, 128, 89, 7.0)])
def test_cluster_all_graphs(ldpc, oracle, BG, Z, R, ebn0):
    """Every decoder LUT at the three lifting sizes the cluster kernel serves (Z / 4 a multiple of 32), 8- and 4-CTA clusters."""
    _check_batch(ldpc, oracle, BG, Z, R, 4, ebn0, seed=Z + R)
    _check_batch(ldpc, oracle, BG, Z, R, 20, ebn0, seed=Z + R + 1)


@pytest.mark.parametrize("max_iter", [0, 1, 2, 3, 20])
def test_cluster_iteration_caps_and_modes(ldpc, oracle, max_iter):
    _check_batch(ldpc, oracle, 1, 384, 13, 3, 2.3, seed=max_iter, max_iter=max_iter)
    _check_batch(ldpc, oracle, 1, 128, 13, 3, 2.3, seed=max_iter, max_iter=max_iter, out_mode=1 + max_iter % 2)
    _check_batch(ldpc, oracle, 1, 384, 23, 6, 4.2, seed=max_iter, max_iter=max_iter, crc=True)


def test_single_cta_kernel_still_covered_for_small_batches(oracle):
    """NRB200_CLUSTER=0 routes small batches to the one-CTA-per-block kernel again; =2 / =4 force small clusters on a single block."""
    code = ("import numpy as np, sys; sys.path.insert(0, 'tests'); from common import make_case; from oracle.bindings import Oracle;"
            "from openairinterface5g_b200.ldpc import load_LDPClib; lib = load_LDPClib(); orc = Oracle();"
            "ok = True\n"
            "for (BG, Z, R, e) in ((1, 384, 13, 2.3), (2, 384, 15, 1.0), (1, 256, 23, 4.0)):\n"
            "    K, P, llr = make_case(orc, BG, Z, R, 3, e, 77); it, out = lib.decode_batch_host(BG, Z, R, 8, llr, latency_mode=1)\n"
            "    for i in range(3):\n"
            "        a = orc.decode(BG, Z, R, 8, llr[i], 0); ok = ok and a[0] == it[i] and np.array_equal(np.asarray(a[1]).view(np.uint8), out[i])\n"
            "print('VARIANT_OK' if ok else 'VARIANT_BAD')")
    for v in ("0", "2", "4"):
        r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=dict(os.environ, NRB200_CLUSTER=v), capture_output=True, text=True, timeout=300)
        assert "VARIANT_OK" in r.stdout, (v, r.stdout[-500:] + r.stderr[-1500:])


def test_per_call_abi_many_threads_combined(ldpc, oracle):
    """16 caller threads x 12 blocking LDPCdecoder calls each, tpool style (nr_ulsch_decoding.c:435-468): every result against the oracle, and
    the combining queue must have carried more than one block per launch on average."""
    n_thr, per = 16, 12
    K, P, llr = make_case(oracle, 1, 384, 13, n_thr * per, 2.4, seed=31)
    want = [oracle.decode(1, 384, 13, 8, llr[i]) for i in range(n_thr * per)]
    got = [None] * (n_thr * per)
    l0, b0 = C.c_uint64(), C.c_uint64()
    ldpc.lib.nrb200_ll_stats(C.byref(l0), C.byref(b0))
    start = threading.Barrier(n_thr)

    def work(t):
        start.wait()
        for j in range(per):
            i = t * per + j
            got[i] = ldpc.LDPCdecoder(1, 384, 13, 8, llr[i])
    ths = [threading.Thread(target=work, args=(t,)) for t in range(n_thr)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    for i in range(n_thr * per):
        assert got[i][0] == want[i][0] and np.array_equal(got[i][1], want[i][1]), i
    l1, b1 = C.c_uint64(), C.c_uint64()
    ldpc.lib.nrb200_ll_stats(C.byref(l1), C.byref(b1))
    assert b1.value - b0.value == n_thr * per
    assert l1.value - l0.value <= b1.value - b0.value


def test_per_call_abi_mixed_configurations_in_flight(ldpc, oracle):
    """Callers with different graphs / stop modes at the same time: the queue groups by configuration, nobody gets somebody else's kernel."""
    cfgs = [(1, 384, 13, 2.4, False), (1, 384, 23, 4.2, True), (2, 128, 15, 1.5, False), (1, 24, 13, 3.5, False), (2, 384, 13, 2.8, False), (1, 10, 13, 4.0, False)]
    cases = []
    for k, (BG, Z, R, e, crc) in enumerate(cfgs):
        K, P, llr = make_case(oracle, BG, Z, R, 4, e, seed=50 + k)
        cases.append((BG, Z, R, K, crc, llr))
    got = {}

    def work(k):
        BG, Z, R, K, crc, llr = cases[k]
        for i in range(4):
            got[(k, i)] = ldpc.LDPCdecoder(BG, Z, R, 8, llr[i], E=K if crc else 0, crc_type=1, check_crc=crc)
    ths = [threading.Thread(target=work, args=(k,)) for k in range(len(cfgs))]
    [t.start() for t in ths]
    [t.join() for t in ths]
    for k, (BG, Z, R, K, crc, llr) in enumerate(cases):
        for i in range(4):
            it_o, out_o = oracle.decode(BG, Z, R, 8, llr[i], 0, *((1, K, 1) if crc else ()))
            assert got[(k, i)][0] == it_o and np.array_equal(got[(k, i)][1], np.asarray(out_o).view(np.uint8)), (k, i)


def test_abort_flag_polled_while_the_call_runs(ldpc, oracle):
    """decode_abort_t set by ANOTHER thread while LDPCdecoder is in flight (a sibling segment failed): the reference polls the flag at the top
    of every iteration (nrLDPC_decoder.c:557-560) and returns numMaxIter + 2.  Whether a given call sees the flag in time is a race in the
    reference too; with numMaxIter = 200 on an undecodable block (~1 ms of iterations) the flag, set 50 us into the call, must be seen."""
    from openairinterface5g_b200.ldpc import DecodeAbort
    K, P, llr = make_case(oracle, 1, 384, 13, 1, 0.0, seed=5)                # far below the waterfall: never converges
    it_full, _ = ldpc.LDPCdecoder(1, 384, 13, 200, llr[0])
    assert it_full == 201
    seen = 0
    for rep in range(5):
        ab = DecodeAbort()
        res = {}

        def call():
            res["r"] = ldpc.LDPCdecoder(1, 384, 13, 200, llr[0], abort=ab)
        t = threading.Thread(target=call)
        t.start()
        time.sleep(50e-6)
        ab.failed = True
        t.join()
        assert res["r"][0] in (201, 202)
        seen += res["r"][0] == 202
        assert bool(ab.failed)
    assert seen >= 1


def test_internal_failure_is_reported_as_decode_failure(ldpc):
    """OAI's callers only test `decodeIterations <= numMaxIter`; an invalid configuration must come back as numMaxIter + 1 with the
    abort flag set, never as a negative value that reads as success."""
    from openairinterface5g_b200.ldpc import DecodeAbort, DecParams
    p = DecParams()
    p.BG, p.Z, p.R, p.numMaxIter, p.outMode = 1, 385, 13, 8, 0                 # 385 is not an NR lifting size
    ab = DecodeAbort()
    buf = np.zeros(27000, np.int8)
    out = np.zeros(27000, np.int8)
    i8p = C.POINTER(C.c_int8)
    it = ldpc.lib.LDPCdecoder(C.byref(p), 0, 0, 0, buf.ctypes.data_as(i8p), out.ctypes.data_as(i8p), None, C.cast(C.byref(ab), C.c_void_p))
    assert it == 9 and bool(ab.failed)
    assert "LDPCdecoder" in ldpc.last_error()
