#!/usr/bin/env python3
"""Generate tests/golden/ue_layers.npz from the UNMODIFIED reference (oracle/_ref/libref_pdsch.so: nr_rx_pdsch with Nl = 3 and 4).  The inputs are seeded
(numpy PCG64, reproduced by the tests and guarded by a checksum stored here), the outputs come ONLY from the reference.  Run where /root/reference exists."""
import os
import sys
import zlib
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.bindings import Reference, PuschParms  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "ue_layers.npz")
CASES = [  # N, nb_rx, rb_start, rb_size, Qm, dmrs_pos, dmrs_type, cdm groups, carrier PRBs, start_symbol, nr_symbols, layers, amplitude rx, amplitude h
    (512, 4, 3, 11, 8, 1 << 1, 0, 1, 25, 1, 6, 4, 32767, 32767), (512, 4, 0, 25, 6, 1 << 2, 0, 1, 25, 0, 14, 3, 60, 40), (512, 4, 2, 20, 4, 1 << 2, 1, 2, 25, 1, 13, 4, 2500, 2500),
    (512, 3, 0, 12, 2, 1 << 3, 0, 2, 25, 2, 10, 3, 300, 200),
]


def inputs(case, seed):
    N, nb_rx, _, _, _, _, _, _, _, _, _, nl, ay, ah = case
    rng = np.random.default_rng(seed)
    rx = rng.integers(-ay, ay + 1, size=(nb_rx, 14, N, 2)).astype(np.int16)
    h = rng.integers(-ah, ah + 1, size=(nl * nb_rx, 14, N, 2)).astype(np.int16)
    return rx, h


def main():
    ref = Reference()
    g = {"n": np.int32(len(CASES))}
    for i, c in enumerate(CASES):
        N, nb_rx, rb_start, rb_size, Qm, dpos, dtype_, cdm, carrier, start, nsym, nl, ay, ah = c
        rx, h = inputs(c, 3000 + i)
        P = PuschParms(N, nb_rx, rb_start, 0, rb_size, N - carrier * 6, Qm, dpos, dtype_, cdm)
        per = [(rb_size * ((12 - 6 * cdm) if dtype_ == 0 else (12 - 4 * cdm)) if (dpos >> s) & 1 else rb_size * 12) for s in range(start, start + nsym)]
        G = sum(per) * Qm * nl
        llr, sh, valid = ref.pdsch_rx_slot(P, start, nsym, rx, h, G, nl=nl)
        g[f"case{i}"], g[f"llr{i}"], g[f"sh{i}"] = np.array(c, np.int32), llr, np.int32(sh)
        g[f"crc{i}"] = np.uint32(zlib.crc32(rx.tobytes() + h.tobytes()))
    np.savez_compressed(OUT, **g)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
