#!/usr/bin/env python3
"""Phase timing of the cluster decoder from its own clock64() marks (NRB200_CLUSTER_TIMERS=1 is set here): prologue, then per iteration
CN compute | CTA barrier + cluster barrier 1 | BN compute | cluster barrier 2, per CTA of code block 0.  Usage: python tools/cluster_phases.py [ebn0=1.0] [n=1]"""
import ctypes as C
import os, sys
os.environ["NRB200_CLUSTER_TIMERS"] = "1"
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openairinterface5g_b200.ldpc import load_LDPClib

def main():
    ebn0 = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    lib = load_LDPClib()
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev); g.manual_seed(1)
    Z, K = 384, 8448
    payload = torch.randint(0, 256, (n, K // 8), dtype=torch.uint8, device=dev, generator=g)
    cw = lib.encode_batch_torch(1, Z, K, payload)
    sigma = 1.0 / np.sqrt(2.0 * 10 ** (ebn0 / 10) / 3.0)
    y = (1.0 - 2.0 * cw.float()) + sigma * torch.randn(cw.shape, device=dev, generator=g)
    llr = torch.zeros((n, 68 * Z), dtype=torch.int8, device=dev)
    llr[:, 2 * Z:] = torch.clamp(torch.floor(y / (sigma / 16)), -128, 127).to(torch.int8)
    for _ in range(5):
        it, out = lib.decode_batch_torch(1, Z, 13, 8, llr, latency_mode=1)
    torch.cuda.synchronize()
    marks = np.zeros(8 * 64, np.int64)
    assert lib.lib.nrb200_debug_cluster_marks(marks.ctypes.data_as(C.c_void_p)) == 0
    m = marks.reshape(8, 64)
    ranks = [r for r in range(8) if m[r, 0] != 0]
    print(f"iters={it.tolist()[:4]} CTAs with marks={len(ranks)}")
    names = ["tables loaded", "R init + first cluster sync", "LLR fetch + broadcast + sync", "A / P fill"]
    for r in ranks:
        d = np.diff(m[r])
        pro = d[:3]
        body = d[3:]
        k = 0
        rows = []
        while k + 4 <= len(body) and m[r, 3 + k + 4] != 0:
            rows.append(body[k:k + 4]); k += 4
        rows = np.array(rows)
        tot = m[r][np.nonzero(m[r])[0][-1]] - m[r, 0]
        mean = rows.mean(axis=0) if len(rows) else np.zeros(4)
        print(f"rank {r}: prologue {pro.tolist()} | per iteration mean: CN {mean[0]:.0f}  sync1 {mean[1]:.0f}  BN {mean[2]:.0f}  sync2 {mean[3]:.0f}  = {mean.sum():.0f} cycles x {len(rows)} | total {tot} cycles")

if __name__ == "__main__":
    main()
