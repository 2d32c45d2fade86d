"""Synthetic ldpctest-style inputs (reference openair1/PHY/CODING/TESTBENCH/ldpctest.c:294-313, 119, 522):
BPSK (+1 <-> bit 0) + AWGN, sigma = 1/sqrt(2*SNR_lin), SNR_lin = 10^(EbN0/10) * rate, LLR = clamp(floor(y/(sigma/16)), -128, 127)
(coding_unitary_defs.h:37-49, qbits = 8); the first 2Z (punctured) LLRs are 0.  numpy only -- this is input generation, not the hot path.
"""
import numpy as np


def awgn_llr(coded_bits, Z, ncols, ebn0_db, rate, seed, qbits=8):
    """coded_bits: (n_cb, >= (ncols-2)*Z) array of 0/1 (encoder output: K-2Z systematic + parity).
    Returns int8 (n_cb, ncols*Z) decoder input."""
    coded_bits = np.asarray(coded_bits)
    n_cb = coded_bits.shape[0]
    n_tx = (ncols - 2) * Z
    snr_lin = 10.0 ** (ebn0_db / 10.0) * rate
    sigma = 1.0 / np.sqrt(2.0 * snr_lin)
    rng = np.random.default_rng(seed)
    y = (1.0 - 2.0 * coded_bits[:, :n_tx].astype(np.float64)) + sigma * rng.standard_normal((n_cb, n_tx))
    q = np.floor(y / (sigma / 16.0))
    maxlev = 1 << (qbits - 1)
    q = np.clip(q, -maxlev, maxlev - 1).astype(np.int8)
    llr = np.zeros((n_cb, ncols * Z), dtype=np.int8)
    llr[:, 2 * Z:] = q
    return llr


def random_payloads(n_cb, K, seed):
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, size=(n_cb, K // 8), dtype=np.uint8)
