#!/usr/bin/env python3
"""Generate tests/golden/prach.npz from the UNMODIFIED reference (oracle/_ref/libref_prach.so): the root sequences of compute_nr_prach_seq for every case of
tests/common.py:PRACH_CASES (only the rows in use), the seeded inputs, and rx_nr_prach's answer.  Run where /root/reference exists."""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle.bindings import Reference  # noqa: E402
from common import PRACH_CASES, prach_inputs, prach_num_roots  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "prach.npz")


def main():
    ref = Reference()
    rng = np.random.default_rng(81)
    g = {}
    for i, case in enumerate(PRACH_CASES):
        nb_rx, short, root, NCS, fmt, mu, pre, delay, amp, sigma = case
        nroots = prach_num_roots(short, NCS)
        xu = ref.prach_seq(short, nroots, root)
        rx = prach_inputs(rng, case, xu)
        g[f"xu{i}"], g[f"rx{i}"] = xu, rx
        g[f"out{i}"] = np.array(ref.rx_nr_prach(nb_rx, short, root, nroots, NCS, fmt, mu, xu, rx), np.int32)
    np.savez_compressed(OUT, **g)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
