"""Drop-in by symbol interposition (channel estimation has no plug-in boundary in OAI, so the build creates one).

integration/oai_shim_pusch_chest.c defines OAI's own `nr_pusch_channel_estimation` -- same prototype, compiled against OAI's headers -- and forwards to
libldpc_b200.so.  Here the reference-side CALLER (oracle/ref_harness_chest.c: it fills PHY_VARS_gNB / nfapi_nr_pusch_pdu_t the way nr_rx_pusch_tp's
callers do and calls the function by name) is linked against that interposer instead of the reference's nr_ul_channel_estimation.c
(integration/build_shims.sh -> oracle/_ref/libshimtest_chest.so, prebuilt where /root/reference exists, travels to the GPU box).  What the unchanged host C
gets back through OAI's own structures -- ul_ch_estimates, *max_ch, *nvar, delay_t -- must be what the pinned oracle computes."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle.bindings import ChestParms

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIMTEST = os.path.join(ROOT, "oracle", "_ref", "libshimtest_chest.so")

CASES = [  # N, nb_rx, slot, symbol, port, rb_start, rb_size, carrier PRBs, scid, dmrs id, dmrs_type, chest_freq
    (4096, 4, 1, 2, 0, 0, 273, 273, 0, 77, 0, 0), (2048, 2, 7, 3, 1, 10, 50, 106, 1, 1007, 0, 0), (1024, 2, 6, 11, 2, 20, 32, 52, 0, 300, 0, 0),
    (1024, 3, 0, 11, 2, 20, 32, 52, 0, 300, 1, 0), (2048, 2, 9, 3, 1, 10, 50, 106, 1, 1007, 0, 1), (512, 2, 8, 4, 1, 0, 25, 25, 1, 303, 1, 1),
]


def test_oai_caller_reaches_the_gpu_through_the_interposed_symbol(oracle):
    if not os.path.exists(SHIMTEST):
        pytest.fail(f"{SHIMTEST} missing: run integration/build_shims.sh where /root/reference exists (the file travels with the repo snapshot)")
    lib = C.CDLL(SHIMTEST)
    assert lib.refh_chest_init(os.path.join(ROOT, "oracle", "_ref", "libref_dfts.so").encode()) == 0
    rng = np.random.default_rng(91)
    for N, nb_rx, slot, symbol, port, rb_start, rb_size, carrier, scid, nid, dmrs_type, chest_freq in CASES:
        fco = N - carrier * 6
        P = ChestParms(N, nb_rx, slot, symbol, port, rb_start, 0, rb_size, fco, scid, nid, dmrs_type, chest_freq)
        rx = rng.integers(-3000, 3001, size=(nb_rx, 14, N, 2)).astype(np.int16)
        prm = np.array([N, nb_rx, carrier, slot, symbol, port, rb_start, 0, rb_size, fco, scid, nid, dmrs_type, chest_freq], dtype=np.int32)
        est = np.zeros(nb_rx * 14 * N * 2, np.int16)
        out = np.zeros(5, np.int32)
        pil = np.zeros(2 * 6 * rb_size, np.int16)
        assert lib.refh_pusch_chest(prm.ctypes.data_as(C.c_void_p), rx.ctypes.data_as(C.c_void_p), est.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p),
                                    pil.ctypes.data_as(C.c_void_p)) == 0
        est_o, out_o = oracle.pusch_channel_estimation(P, rx)
        assert np.array_equal(out, out_o), (N, nb_rx, slot, symbol, port, dmrs_type, chest_freq, out, out_o)
        assert np.array_equal(est.reshape(nb_rx, 14, N, 2)[:, symbol], est_o[:, symbol]), (N, nb_rx, slot, symbol, port, dmrs_type, chest_freq)
