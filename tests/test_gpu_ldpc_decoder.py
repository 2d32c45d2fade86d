"""GPU parity: libldpc_b200.so decoder vs the CPU oracle (bit exact: output bytes and returned iteration counts)."""
import numpy as np
import pytest
from common import ALL_Z, RATES, NCOLS, make_case

pytestmark = pytest.mark.gpu


def _check(ldpc, oracle, BG, Z, R, n, ebn0, seed, max_iter=8, out_mode=0):
    K, P, llr = make_case(oracle, BG, Z, R, n, ebn0, seed)
    iters, out = ldpc.decode_batch_host(BG, Z, R, max_iter, llr, outMode=out_mode)
    for i in range(n):
        it_o, out_o = oracle.decode(BG, Z, R, max_iter, llr[i], out_mode)
        assert iters[i] == it_o, (BG, Z, R, i, iters[i], it_o)
        assert np.array_equal(out[i].view(np.uint8), np.asarray(out_o).view(np.uint8)), (BG, Z, R, i)
    return iters


@pytest.mark.parametrize("ebn0", [1.0, 2.2, 3.0, 5.0])
def test_bg1_z384_r13_headline(ldpc, oracle, ebn0):
    _check(ldpc, oracle, 1, 384, 13, 6, ebn0, seed=int(ebn0 * 10))


@pytest.mark.parametrize("BG,R,ebn0", [(1, 23, 4.0), (1, 89, 7.0), (2, 15, 1.0), (2, 13, 2.5), (2, 23, 5.0)])
def test_all_rates_z384(ldpc, oracle, BG, R, ebn0):
    _check(ldpc, oracle, BG, 384, R, 4, ebn0, seed=R)


@pytest.mark.parametrize("Z", ALL_Z)
def test_all_lifting_sizes(ldpc, oracle, Z):
    for BG in (1, 2):
        R = RATES[BG][0]
        _check(ldpc, oracle, BG, Z, R, 2, 3.5 if BG == 1 else 2.5, seed=Z)


@pytest.mark.parametrize("out_mode", [1, 2])
def test_output_modes(ldpc, oracle, out_mode):
    _check(ldpc, oracle, 1, 96, 13, 3, 3.0, seed=5, out_mode=out_mode)


@pytest.mark.parametrize("max_iter", [1, 2, 3, 5, 20])
def test_iteration_caps(ldpc, oracle, max_iter):
    _check(ldpc, oracle, 1, 128, 13, 4, 2.0, seed=max_iter, max_iter=max_iter)
    _check(ldpc, oracle, 2, 64, 13, 4, 3.0, seed=max_iter, max_iter=max_iter)


def test_submit_wait_batches_in_flight(ldpc, oracle):
    """Three batches enqueued before the first is dequeued (nrb200_ldpc_decode_batch_host_submit / _wait, chunked over two streams each:
    200 blocks > one wave) give what the blocking call and the oracle give, whatever order the tickets are waited for in."""
    cases = [make_case(oracle, 1, 384, 13, 200, ebn0, seed)[2] for ebn0, seed in ((1.0, 11), (2.4, 12), (4.0, 13))]
    n = 68 * 384
    outs = [np.zeros((200, n // 8), dtype=np.uint8) for _ in cases]
    its = [np.zeros(200, dtype=np.int32) for _ in cases]
    tickets = [ldpc.decode_batch_host_submit(1, 384, 13, 8, llr, outs[i], its[i]) for i, llr in enumerate(cases)]
    for i in (1, 0, 2):
        ldpc.decode_batch_host_wait(tickets[i])
    for i, llr in enumerate(cases):
        it_b, out_b = ldpc.decode_batch_host(1, 384, 13, 8, llr)
        assert np.array_equal(its[i], it_b) and np.array_equal(outs[i], out_b)
        for j in (0, 57, 199):
            it_o, out_o = oracle.decode(1, 384, 13, 8, llr[j], 0)
            assert its[i][j] == it_o and np.array_equal(outs[i][j], np.asarray(out_o).view(np.uint8))


@pytest.mark.parametrize("threads", [864, 960])
def test_z384_thread_variants(oracle, threads):
    """The 9- and 10-bin instantiations of the Z = 384 kernel (NRB200_PACKED_THREADS, read when the graph tables are first built: own process)."""
    import os, subprocess, sys
    code = ("import numpy as np, sys; sys.path.insert(0, 'tests'); from common import make_case; from oracle.bindings import Oracle;"
            "from openairinterface5g_b200.ldpc import load_LDPClib; lib = load_LDPClib(); orc = Oracle();"
            "K, P, llr = make_case(orc, 1, 384, 13, 5, 2.3, 77); it, out = lib.decode_batch_host(1, 384, 13, 8, llr);"
            "ok = all(orc.decode(1, 384, 13, 8, llr[i], 0)[0] == it[i] and np.array_equal(np.asarray(orc.decode(1, 384, 13, 8, llr[i], 0)[1]).view(np.uint8), out[i]) for i in range(5));"
            "print('VARIANT_OK' if ok else 'VARIANT_BAD')")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], cwd=root, env=dict(os.environ, NRB200_PACKED_THREADS=str(threads)), capture_output=True, text=True, timeout=300)
    assert "VARIANT_OK" in r.stdout, r.stdout[-500:] + r.stderr[-1500:]
