#!/usr/bin/env python3
"""Per-step wall times of the host decode API, blocking vs submit / wait with 1-3 batches in flight (diagnostic)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openairinterface5g_b200.ldpc import load_LDPClib
lib = load_LDPClib()
B, N = 1024, 68 * 384
rng = np.random.default_rng(1)
llr = [torch.from_numpy(rng.integers(-20, 21, size=(B, N), dtype=np.int8)).pin_memory().numpy() for _ in range(3)]
for depth in (0, 1, 2, 3):
    D = max(depth, 1)
    outs = [torch.empty((B, N // 8), dtype=torch.uint8).pin_memory().numpy() for _ in range(D)]
    its = [np.zeros(B, dtype=np.int32) for _ in range(D)]
    ts, tickets = [], []
    t_all = time.perf_counter()
    for i in range(80):
        if i == 20: t_all = time.perf_counter()
        t0 = time.perf_counter()
        if depth == 0:
            lib.decode_batch_host(1, 384, 13, 8, llr[i % 3], out=outs[0], iters=its[0])
        else:
            tickets.append(lib.decode_batch_host_submit(1, 384, 13, 8, llr[i % 3], outs[i % D], its[i % D]))
            if len(tickets) == D: lib.decode_batch_host_wait(tickets.pop(0))
        ts.append(time.perf_counter() - t0)
    while tickets: lib.decode_batch_host_wait(tickets.pop(0))
    total = time.perf_counter() - t_all
    ts = np.array(ts[20:]) * 1e3
    print(f"depth={depth} CB/s={60 * B / total:.0f} step ms: median {np.median(ts):.3f} min {ts.min():.3f} max {ts.max():.3f}", flush=True)
