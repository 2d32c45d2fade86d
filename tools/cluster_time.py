#!/usr/bin/env python3
"""Device-resident time of ONE small decode launch (n code blocks, BG1 Z=384 R13, Eb/N0 given: 1.0 dB = all 8+1 passes) -- the quantity the
cluster kernel exists for.  Back-to-back launches on one stream between CUDA events; NRB200_CLUSTER / NRB200_CLUSTER_WARPS select the variant
(read once per process).  Usage: python tools/cluster_time.py [ebn0=1.0] [n ...]"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openairinterface5g_b200.ldpc import load_LDPClib

def main():
    ebn0 = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
    ns = [int(x) for x in sys.argv[2:]] or [1, 8, 16, 32, 52, 74]
    lib = load_LDPClib()
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev); g.manual_seed(1)
    B, Z, K = max(ns), 384, 8448
    payload = torch.randint(0, 256, (B, K // 8), dtype=torch.uint8, device=dev, generator=g)
    cw = lib.encode_batch_torch(1, Z, K, payload)
    sigma = 1.0 / np.sqrt(2.0 * 10 ** (ebn0 / 10) / 3.0)
    y = (1.0 - 2.0 * cw.float()) + sigma * torch.randn(cw.shape, device=dev, generator=g)
    llr = torch.zeros((B, 68 * Z), dtype=torch.int8, device=dev)
    llr[:, 2 * Z:] = torch.clamp(torch.floor(y / (sigma / 16)), -128, 127).to(torch.int8)
    for n in ns:
        l = llr[:n].contiguous()
        out = torch.empty((n, 68 * Z // 8), dtype=torch.uint8, device=dev); it = torch.empty(n, dtype=torch.int32, device=dev)
        for _ in range(5): lib.decode_batch_torch(1, Z, 13, 8, l, out=out, iters=it, latency_mode=1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 50
        e0.record()
        for _ in range(reps): lib.decode_batch_torch(1, Z, 13, 8, l, out=out, iters=it, latency_mode=1)
        e1.record(); torch.cuda.synchronize()
        us = 1e3 * e0.elapsed_time(e1) / reps
        print(f"cluster={os.environ.get('NRB200_CLUSTER','auto')} warps={os.environ.get('NRB200_CLUSTER_WARPS','default')} n={n} ebn0={ebn0} "
              f"us_per_launch={us:.1f} us_per_block={us / n:.2f} mean_iters={it.float().mean().item():.2f}", flush=True)

if __name__ == "__main__":
    main()
