#!/usr/bin/env python3
"""Generate tests/golden/dft4.npz from the UNMODIFIED reference compiled by oracle/build_ref.sh (oracle/_ref/libref_dfts.so): the four-way DFT-s-OFDM entry
points dft12 ... dft3240 (oai_dfts.c:4352-7706), one call per size (4 N c16 in, 4 N c16 out), scale_flag 1 and 0.  Run in the container that has /root/reference."""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.bindings import Reference   # noqa: E402

SIZES = [12, 24, 36, 48, 60, 72, 96, 108, 120, 144, 180, 192, 216, 240, 288, 300, 324, 360, 384, 600, 900, 1200, 1500, 1920, 3000, 3240]


def main():
    ref = Reference()
    g = {"sizes": np.array(SIZES)}
    for N in SIZES:
        rng = np.random.default_rng(7000 + N)
        x = rng.integers(-2500, 2501, size=8 * N).astype(np.int16)
        g[f"x{N}"] = x
        g[f"y{N}_s1"] = ref.dft4(N, x, 1)
        g[f"y{N}_s0"] = ref.dft4(N, x, 0)
    out = os.path.join(ROOT, "tests", "golden", "dft4.npz")
    np.savez_compressed(out, **g)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
