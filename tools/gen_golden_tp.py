#!/usr/bin/env python3
"""Generate tests/golden/transform_precoding.npz from the UNMODIFIED reference (oracle/_ref/libref_chest.so, libref_pusch.so): low-PAPR type-1 DMRS sequences
(ul_ref_seq_nr.c), nr_pusch_channel_estimation with transform precoding, and inner_rx with nr_freq_equalization + nr_idft.  Seeded inputs, outputs ONLY
from the reference.  Run where /root/reference exists; tests/test_golden_oracle.py pins the oracle to these vectors where it does not."""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.bindings import Reference, ChestParms, PuschParms  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "transform_precoding.npz")


def main():
    ref = Reference()
    rng = np.random.default_rng(2031)
    g = {}
    # sequences: the four table-driven lengths for every group (the product takes these from its caller), a few computed ones
    for M in (6, 12, 18, 24):
        g[f"seq_{M}"] = np.stack([ref.lowpapr_seq(u, 0, M) for u in range(30)])
    for M, u in ((30, 4), (36, 0), (150, 17), (1620, 29)):
        g[f"seq_{M}_u{u}"] = ref.lowpapr_seq(u, 0, M)
    N, nrx, carrier = 512, 2, 25
    rx = rng.integers(-4000, 4001, size=(nrx, 14, N, 2)).astype(np.int16)
    h = rng.integers(-1500, 1501, size=(nrx, 14, N, 2)).astype(np.int16)
    g["rx"], g["h"] = rx, h
    for i, (slot, symbol, rb_start, rb_size, u) in enumerate(((3, 2, 2, 20, 5), (6, 3, 0, 25, 29), (1, 2, 4, 3, 11))):
        par = [N, nrx, slot, symbol, 0, rb_start, 0, rb_size, N - carrier * 6, 0, 77 + i, 0, 0]
        ref.chest_set_transform_precoding(1, u, 0)
        est, out, _ = ref.pusch_channel_estimation(ChestParms(*par), rx, carrier)
        ref.chest_set_transform_precoding(0)
        g[f"chest_par{i}"], g[f"chest_u{i}"], g[f"chest_est{i}"], g[f"chest_state{i}"] = np.array(par, np.int32), np.int32(u), est[:, symbol], out
        g[f"chest_seq{i}"] = ref.lowpapr_seq(u, 0, 6 * rb_size)
    ref.pusch_set_transform_precoding(1)
    for i, (rb_start, rb_size, Qm, symbol, shift) in enumerate(((2, 20, 6, 0, 8), (0, 25, 4, 7, 7), (4, 1, 6, 13, 8), (1, 16, 2, 4, 9), (0, 5, 4, 1, 6))):
        par = [N, nrx, rb_start, 0, rb_size, N - carrier * 6, Qm, 1 << 2, 0, 2]
        llr, comp = ref.pusch_inner_rx_symbol(PuschParms(*par), symbol, 2, shift, rx, h, 12 * rb_size)
        g[f"rx_par{i}"], g[f"rx_sym{i}"], g[f"rx_llr{i}"], g[f"rx_comp{i}"] = np.array(par, np.int32), np.array([symbol, shift], np.int32), llr, comp[:24 * rb_size]
    ref.pusch_set_transform_precoding(0)
    np.savez_compressed(OUT, **g)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
