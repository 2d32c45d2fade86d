"""The oracle restatements compose into a working PUSCH slot: what the CPU chain transmits it also receives.  This is the CPU twin of
tests/test_gpu_slot_chain.py (same order of operations, oracle functions instead of kernels) and guards the test infrastructure itself."""
import numpy as np

from common import oracle_pusch_receive, oracle_pusch_transmit
from openairinterface5g_b200.ofdm import NrOfdmParms


def test_oracle_pusch_slot_roundtrip(oracle):
    P = NrOfdmParms(1024, 0, 52)
    rot = P.symbol_rotation(3609200000.0)
    A, Qm, rb_start, rb_size, nb_rx, slot, rnti, nid = 15976, 4, 0, 52, 2, 1, 0x1234, 77
    payload, frame, est, info = oracle_pusch_transmit(oracle, P, A, Qm, rb_start, rb_size, nb_rx, slot, rnti, nid, rot, seed=3)
    for use_est in (est, None):                                             # genie estimates, then the estimator on the DMRS symbol
        tb, its, llr, shift = oracle_pusch_receive(oracle, P, info, Qm, rb_start, rb_size, nb_rx, slot, rnti, nid, rot, frame, use_est)
        assert (its <= 8).all(), its
        assert np.array_equal(tb[:payload.size], payload)
        assert oracle.crc(0, tb, A + 24) == 0
