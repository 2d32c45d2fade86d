"""GPU parity of Gold-sequence scrambling / unscrambling and the QAM mapper against the CPU oracle (pinned to nr_scrambling.c,
nr_modulation.c, nr_gen_mod_table.c)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_scramble_modulate_vs_oracle(ldpc, oracle):
    rng = np.random.default_rng(6)
    for size, q, Nid, rnti in ((64, 0, 0, 1), (33, 0, 5, 77), (1000, 0, 123, 0x1234), (9072 * 6, 1, 1007, 65535), (471744, 0, 42, 4660)):
        bits = rng.integers(0, 2, size=size, dtype=np.uint8)
        sc = ldpc.scramble_host(bits, q, Nid, rnti)
        assert np.array_equal(sc, oracle.scramble(bits, q, Nid, rnti)), (size, q, Nid, rnti)
        for Qm in (2, 4, 6, 8):
            length = (size // Qm) * Qm
            if length:
                assert np.array_equal(ldpc.modulate_host(sc, length, Qm), oracle.modulate(sc, length, Qm)), (size, Qm)


def test_unscramble_llr_vs_oracle(ldpc, oracle):
    rng = np.random.default_rng(8)
    for size, q, Nid, rnti in ((64, 0, 0, 1), (9071, 1, 1007, 65535), (12 * 273 * 6 * 13, 0, 500, 4660)):
        llr = rng.integers(-32768, 32768, size=size).astype(np.int16)
        un = ldpc.unscramble_llr_host(llr, q, Nid, rnti)
        assert np.array_equal(un, oracle.unscramble_llr(llr, q, Nid, rnti))
        # involution (-32768 maps onto itself both times)
        assert np.array_equal(ldpc.unscramble_llr_host(un, q, Nid, rnti), llr)
