#!/usr/bin/env python3
"""Device-resident decode kernel time for the headline shape (BG1 Z=384 R13, batch 1024); prints ms and CB/s.
Used for quick A/B runs of kernel variants (env NRB200_PACKED_THREADS, NRB200_FORCE_GENERIC)."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openairinterface5g_b200.ldpc import load_LDPClib

def main():
    ebn0 = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
    lib = load_LDPClib()
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev); g.manual_seed(1)
    B, Z, K = 1024, 384, 8448
    payload = torch.randint(0, 256, (B, K // 8), dtype=torch.uint8, device=dev, generator=g)
    cw = lib.encode_batch_torch(1, Z, K, payload)
    sigma = 1.0 / np.sqrt(2.0 * 10 ** (ebn0 / 10) / 3.0)
    y = (1.0 - 2.0 * cw.float()) + sigma * torch.randn(cw.shape, device=dev, generator=g)
    llr = torch.zeros((B, 68 * Z), dtype=torch.int8, device=dev)
    llr[:, 2 * Z:] = torch.clamp(torch.floor(y / (sigma / 16)), -128, 127).to(torch.int8)
    out = torch.empty((B, 68 * Z // 8), dtype=torch.uint8, device=dev); it = torch.empty(B, dtype=torch.int32, device=dev)
    for _ in range(5): lib.decode_batch_torch(1, Z, 13, 8, llr, out=out, iters=it)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 30
    for _ in range(n): lib.decode_batch_torch(1, Z, 13, 8, llr, out=out, iters=it)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"threads={os.environ.get('NRB200_PACKED_THREADS','default')} generic={os.environ.get('NRB200_FORCE_GENERIC','0')} ebn0={ebn0} ms={ms:.4f} CB/s={B/ms*1e3:.0f} mean_iters={it.float().mean().item():.3f}")

if __name__ == "__main__":
    main()
