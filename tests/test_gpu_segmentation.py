"""Transport-block CRC attachment + nr_segmentation on the device (nrb200_tb_segment_*) against the oracle restatement of crc_byte.c / nr_segmentation.c
(pinned to the compiled reference in tests/test_oracle_vs_reference.py).  The scalar part runs without a GPU."""
import numpy as np
import pytest

SIZES = [(1, 434280), (1, 235624), (1, 8424), (1, 8400), (1, 3848), (1, 3824), (2, 3824), (2, 3752), (2, 640), (2, 552), (2, 184), (2, 24), (1, 1277992), (2, 19464), (1, 33816)]


def test_tb_segment_parms_vs_oracle(oracle):
    from openairinterface5g_b200.ldpc import LdpcLib
    lib = LdpcLib()
    for BG, A in SIZES + [(1, a) for a in range(8, 30000, 1016)] + [(2, a) for a in range(8, 9000, 376)]:
        B = A + (24 if A > 3824 else 16)
        Kb, Cc, K, Z, F, _ = oracle.segmentation(None, B, BG)
        q = lib.tb_segment_parms(BG, A)
        assert (q["Kb"], q["C"], q["K"], q["Z"], q["F"]) == (Kb, Cc, K, Z, F), (BG, A)


@pytest.mark.gpu
def test_tb_segment_vs_oracle(ldpc, oracle):
    rng = np.random.default_rng(12)
    for BG, A in SIZES:
        payload = rng.integers(0, 256, size=A // 8, dtype=np.uint8)
        if A > 3824:
            crc = oracle.crc(0, payload, A) >> 8
            tb = np.concatenate([payload, np.array([(crc >> 16) & 255, (crc >> 8) & 255, crc & 255], np.uint8)])
        else:
            crc = oracle.crc(3, payload, A) >> 16
            tb = np.concatenate([payload, np.array([(crc >> 8) & 255, crc & 255], np.uint8)])
        q = ldpc.tb_segment_parms(BG, A)
        if (q["Kprime"] - q["L"]) % 8:
            continue                                            # not a byte-aligned split: neither the reference's byte loop nor the kernel handles it
        _, Cc, K, Z, F, segs_o = oracle.segmentation(tb, tb.size * 8, BG)
        segs = ldpc.tb_segment_host(BG, payload)
        assert segs.shape == segs_o.shape and np.array_equal(segs, segs_o), (BG, A, np.argwhere(segs != segs_o)[:4])
