#!/usr/bin/env python3
"""Time (CUDA events) the rfsimulator channel kernel alone: python tools/rfsim_time.py [nb_ant] [taps] [samples].  Also the launch ncu captures
(ncu --set full -k regex:rfsim -s 3 -c 1 python tools/rfsim_time.py)."""
import json
import os
import sys
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from openairinterface5g_b200.ldpc import load_LDPClib, RfsimChan  # noqa: E402


def main():
    nb = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    L = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 614400
    lib = load_LDPClib()
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(5)
    cir = 2 * (n + L + 16)
    ch = torch.from_numpy(rng.normal(size=(nb * nb, L, 2)) * 0.1).to(dev)
    sig = torch.from_numpy(rng.integers(-6000, 6001, size=(cir, 2)).astype(np.int16)).to(dev)
    nz = torch.randn((nb, n, 2), dtype=torch.float64, device=dev)
    out = torch.zeros((nb, n, 2), dtype=torch.int16, device=dev)
    d = RfsimChan(nb, nb, L, 0, -2.0, -30.0, 0)
    for _ in range(5):
        lib.rfsim_rx_add_input_torch(d, ch, sig, out, 614400, nz)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        lib.rfsim_rx_add_input_torch(d, ch, sig, out, 614400, nz)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(json.dumps({"what": f"rfsim_channel_kernel {nb}x{nb} {L} taps {n} samples", "ms": ms, "antenna_samples_per_s": nb * n / ms * 1e3,
                      "fp64_ops_per_s": 8.0 * nb * L * nb * n / ms * 1e3, "bytes_algorithmic": nb * n * 4 + nb * n * 4 * 2 + nb * n * 16}))


if __name__ == "__main__":
    main()
