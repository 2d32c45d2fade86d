"""Shared helpers for the test-suite: ldpctest-style case generation (reference ldpctest.c:269-357)."""
import numpy as np
from openairinterface5g_b200.synth import awgn_llr

ALL_Z = [2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 18, 20, 22, 24, 26, 28, 30, 32, 36, 40, 44, 48, 52, 56, 60, 64, 72, 80, 88,
         96, 104, 112, 120, 128, 144, 160, 176, 192, 208, 224, 240, 256, 288, 320, 352, 384]
RATES = {1: (13, 23, 89), 2: (15, 13, 23)}
NCOLS = {(1, 13): 68, (1, 23): 35, (1, 89): 27, (2, 15): 52, (2, 13): 32, (2, 23): 17}


def payloads(BG, Z, n, seed):
    K = (22 if BG == 1 else 10) * Z
    rng = np.random.default_rng(seed)
    P = rng.integers(0, 256, size=(n, (K + 7) // 8), dtype=np.uint8)
    if K % 8:
        P[:, -1] &= (0xFF << (8 - K % 8)) & 0xFF
    return K, P


def make_case(oracle, BG, Z, R, n, ebn0_db, seed):
    """payload -> oracle encoder -> BPSK/AWGN -> int8 LLRs.  Returns (K, payload bytes, llr[n, ncols*Z])."""
    K, P = payloads(BG, Z, n, seed)
    nc = NCOLS[(BG, R)]
    rate = (22 if BG == 1 else 10) / (nc - 2)
    cw = np.stack([oracle.encode(BG, Z, K, P[i]) for i in range(n)])
    return K, P, awgn_llr(cw, Z, nc, ebn0_db, rate, seed)
