#!/usr/bin/env python3
"""BLER versus Eb/N0 of the B200 decoder at the ldpctest operating point (BG1, Z=384, K=8448, R=1/3 LUT, 8 iterations,
parity-check stop; BPSK, LLR quantisation of coding_unitary_defs.h:37-49).  The decoder is bit exact with the reference, so this curve IS
the reference's curve; the sweep documents it and the mean iteration count that the early-stop throughput (operating point B) depends on.
Everything stays on the device: encode (our encoder kernel), noise + quantisation (torch, plumbing), decode, compare.
Usage: python tools/bler_sweep.py [n_cb_per_point=10240] [out.json]"""
import json
import os
import sys
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from openairinterface5g_b200.ldpc import load_LDPClib   # noqa: E402


def main():
    n_cb = int(sys.argv[1]) if len(sys.argv) > 1 else 10240
    out_path = sys.argv[2] if len(sys.argv) > 2 else None
    dev = torch.device("cuda", 0)
    lib = load_LDPClib()
    BG, Z, K, R, iters_max = 1, 384, 8448, 13, 8
    g = torch.Generator(device=dev); g.manual_seed(1)
    rows = []
    B = 1024
    for step in range(0, 25):
        ebn0 = -2.0 + 0.25 * step
        snr = 10.0 ** (ebn0 / 10.0) * (1.0 / 3.0)
        sigma = (1.0 / (2.0 * snr)) ** 0.5
        blk_err = bit_err = it_sum = 0
        for _ in range(n_cb // B):
            payload = torch.randint(0, 256, (B, K // 8), dtype=torch.uint8, device=dev, generator=g)
            cw = lib.encode_batch_torch(BG, Z, K, payload)                       # (B, 66 Z) one bit per byte
            y = (1.0 - 2.0 * cw.to(torch.float32)) + sigma * torch.randn(cw.shape, device=dev, generator=g)
            q = torch.clamp(torch.floor(y / (sigma / 16.0)), -128, 127).to(torch.int8)
            llr = torch.zeros((B, 68 * Z), dtype=torch.int8, device=dev)
            llr[:, 2 * Z:] = q
            it, hard = lib.decode_batch_torch(BG, Z, R, iters_max, llr)
            diff = (hard[:, :K // 8] ^ payload)
            nerr = torch.bitwise_count(diff).sum(dim=1) if hasattr(torch, "bitwise_count") else sum(((diff >> b) & 1).sum(dim=1) for b in range(8))
            blk_err += int((nerr > 0).sum()); bit_err += int(nerr.sum()); it_sum += int(it.sum())
        n = (n_cb // B) * B
        rows.append({"ebn0_db": ebn0, "blocks": n, "bler": blk_err / n, "ber": bit_err / (n * K), "mean_returned_iterations": it_sum / n})
        print(json.dumps(rows[-1]), flush=True)
    if out_path:
        json.dump({"config": {"BG": BG, "Z": Z, "K": K, "R_lut": R, "numMaxIter": iters_max, "stop": "parity check", "blocks_per_point": n_cb // B * B},
                   "points": rows}, open(out_path, "w"), indent=1)


if __name__ == "__main__":
    main()
