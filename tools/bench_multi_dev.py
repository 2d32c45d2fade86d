#!/usr/bin/env python3
"""ONE process, all visible GPUs (BASELINE config 5's shape: OAI is one process): nrb200_ldpc_decode_batch_host_multi spreads a batch of
1024 x n_dev code blocks (BG1 Z=384 R13, 8 iterations, Eb/N0 1 dB: all 9 passes) over n_dev devices from pinned host buffers -- H2D of the LLRs and
D2H of the hard bits inside the timed region -- and, beside it, the PLATFORM CEILING for that traffic: the same bytes moved by plain
cudaMemcpyAsync on all n_dev devices at once with no kernel at all.  e2e / ceiling says how much of what the host's PCIe / memory system can
deliver the decode path uses.  Usage: python tools/bench_multi_dev.py [steps=20]   -> one JSON line per device count"""
import json
import os
import sys
import time
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from openairinterface5g_b200.ldpc import load_LDPClib   # noqa: E402

BG, Z, R, K, NUM_LLR = 1, 384, 13, 8448, 68 * 384


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    lib = load_LDPClib()
    ndev = lib.device_count()
    B = 1024
    dev0 = torch.device("cuda", 0)
    g = torch.Generator(device=dev0); g.manual_seed(7)
    sigma = 1.0 / np.sqrt(2.0 * 10 ** 0.1 / 3.0)
    payload = torch.randint(0, 256, (B, K // 8), dtype=torch.uint8, device=dev0, generator=g)
    cw = lib.encode_batch_torch(BG, Z, K, payload)
    y = (1.0 - 2.0 * cw.float()) + sigma * torch.randn(cw.shape, device=dev0, generator=g)
    llr1 = torch.zeros((B, NUM_LLR), dtype=torch.int8, device=dev0)
    llr1[:, 2 * Z:] = torch.clamp(torch.floor(y / (sigma / 16)), -128, 127).to(torch.int8)
    it_ref, out_ref = lib.decode_batch_torch(BG, Z, R, 8, llr1)
    torch.cuda.synchronize()
    it_ref, out_ref = it_ref.cpu().numpy(), out_ref.cpu().numpy()
    counts = [n for n in (1, 2, 4, 8) if n <= ndev]
    for n in counts:
        h_llr = torch.empty((B * n, NUM_LLR), dtype=torch.int8).pin_memory()
        for d in range(n):
            h_llr[d * B:(d + 1) * B].copy_(llr1)
        h_out = torch.empty((B * n, NUM_LLR // 8), dtype=torch.uint8).pin_memory()
        h_it = np.zeros(B * n, np.int32)
        np_llr, np_out = h_llr.numpy(), h_out.numpy()
        for _ in range(4):
            lib.decode_batch_host_multi(BG, Z, R, 8, np_llr, n, out=np_out, iters=h_it)
        t0 = time.perf_counter()
        for _ in range(steps):
            lib.decode_batch_host_multi(BG, Z, R, 8, np_llr, n, out=np_out, iters=h_it)
        dt = time.perf_counter() - t0
        ok = all(np.array_equal(h_it[d * B:(d + 1) * B], it_ref) and np.array_equal(np_out[d * B:(d + 1) * B], out_ref) for d in range(n))
        # platform ceiling: the same bytes, plain copies on all devices at once, no kernel
        d_in = [torch.empty((B, NUM_LLR), dtype=torch.int8, device=f"cuda:{d}") for d in range(n)]
        d_out = [torch.empty((B, NUM_LLR // 8), dtype=torch.uint8, device=f"cuda:{d}") for d in range(n)]
        streams = [torch.cuda.Stream(device=d) for d in range(n)]

        def copies():
            for d in range(n):
                with torch.cuda.stream(streams[d]):
                    d_in[d].copy_(h_llr[d * B:(d + 1) * B], non_blocking=True)
                    h_out[d * B:(d + 1) * B].copy_(d_out[d], non_blocking=True)
            for d in range(n):
                streams[d].synchronize()
        for _ in range(3):
            copies()
        t0 = time.perf_counter()
        for _ in range(steps):
            copies()
        dtc = time.perf_counter() - t0
        bytes_step = n * B * (NUM_LLR + NUM_LLR // 8)
        print(json.dumps({"what": "one process, batch spread over n_dev GPUs (nrb200_ldpc_decode_batch_host_multi), pinned host buffers", "n_dev": n,
                          "value": n * B * steps / dt, "unit": "CB/s", "ms_per_step": 1e3 * dt / steps, "bit_exact_vs_one_device": bool(ok),
                          "pcie_gbs": bytes_step * steps / dt / 1e9,
                          "platform_ceiling": {"what": "same bytes by plain cudaMemcpyAsync on all devices at once, no kernel", "gbs": bytes_step * steps / dtc / 1e9,
                                               "equivalent_cb_per_s": n * B * steps / dtc},
                          "frac_of_platform_ceiling": dtc / dt}), flush=True)


if __name__ == "__main__":
    main()
