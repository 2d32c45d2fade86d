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
