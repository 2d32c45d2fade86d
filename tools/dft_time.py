#!/usr/bin/env python3
"""Time (and, under ncu, profile) the batched Q15 transform kernel: python tools/dft_time.py [N=4096] [inverse=1] [batch=1792]."""
import os
import sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from openairinterface5g_b200.dfts import DftsLib   # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
inv = int(sys.argv[2]) if len(sys.argv) > 2 else 1
nb = int(sys.argv[3]) if len(sys.argv) > 3 else 1792
dev = torch.device("cuda", 0)
dl = DftsLib(os.environ["NRB200_DFTS_SO"]) if os.environ.get("NRB200_DFTS_SO") else DftsLib()
dl.autoinit()
g = torch.Generator(device=dev); g.manual_seed(1)
bufs = [torch.randint(-3000, 3000, (nb, 2 * N), dtype=torch.int16, device=dev, generator=g) for _ in range(5)]
out = torch.empty_like(bufs[0])
for i in range(3):
    dl.batch_torch(N, inv, bufs[i % 5], 1, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(40):
    dl.batch_torch(N, inv, bufs[i % 5], 1, out=out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 40
print(f"N={N} inverse={inv} batch={nb} ms={ms:.4f} transforms/s={nb / ms * 1e3:.0f} GB/s={nb * N * 8 / ms / 1e6:.1f}")
