"""Large-sample differential parity at the BASELINE configurations: every block of >= 1e5 random code blocks decoded by the CUDA path at
batch = 1024 is compared -- output bytes AND returned iteration count -- with the UNMODIFIED reference decoder (oracle/_ref/libref_ldpc_dec.so,
built by oracle/build_ref.sh from nrLDPC_decoder.c) run on all host cores (oracle/cpu_bench.c:orc_decode_all), across the ldpctest Eb/N0 sweep
-2 ... +4 dB (ldpctest.c:269-357).  Both arms' BLER curves are written side by side to gpurun_out/bler_cpu_gpu_*.json.
Inputs are generated on the device (our encoder kernel + torch noise: input generation only, the encoder has its own parity tests)."""
import json
import os
import numpy as np
import pytest
from common import decode_all_reference

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BG, Z, K = 1, 384, 8448


def _decode_all(oracle, reference, llr, R, max_iter, crc=None):
    return decode_all_reference(oracle, reference, llr, BG, Z, R, max_iter, crc)


def _gen(ldpc, torch, g, n, ncols, ebn0, rate, with_crc24b=False, E=None):
    """n code blocks: random payload (optionally with a valid CRC24B in the last 24 bits) -> encoder -> BPSK + AWGN -> int8 LLRs (ldpctest.c:294-313)."""
    dev = torch.device("cuda", 0)
    payload = torch.randint(0, 256, (n, K // 8), dtype=torch.uint8, device=dev, generator=g)
    if with_crc24b:
        crc = ldpc.crc_batch_torch(1, payload, K - 24).to(torch.int64) >> 8          # poly 1 = CRC24B, left-aligned result
        payload[:, -3] = ((crc >> 16) & 255).to(torch.uint8); payload[:, -2] = ((crc >> 8) & 255).to(torch.uint8); payload[:, -1] = (crc & 255).to(torch.uint8)
    cw = ldpc.encode_batch_torch(BG, Z, K, payload)[:, :(ncols - 2) * Z]
    sigma = (1.0 / (2.0 * 10.0 ** (ebn0 / 10.0) * rate)) ** 0.5
    y = (1.0 - 2.0 * cw.to(torch.float32)) + sigma * torch.randn(cw.shape, device=dev, generator=g)
    q = torch.clamp(torch.floor(y / (sigma / 16.0)), -128, 127).to(torch.int8)
    llr = torch.zeros((n, ncols * Z), dtype=torch.int8, device=dev)
    llr[:, 2 * Z:] = q
    if E is not None:
        llr[:, 2 * Z + E:] = 0                                                        # beyond the rate-matched length: never transmitted
    return payload, llr


def _bler(out, payload):
    return float((out[:, :K // 8] != payload).any(axis=1).mean())


def test_headline_1e5_blocks_every_block_vs_reference(ldpc, oracle, reference):
    """BASELINE config 2: BG1 Z=384 K=8448, R=1/3 LUT, numMaxIter 8, parity-check stop, batch 1024: 25 Eb/N0 points x 4096 blocks = 102 400 blocks."""
    import torch
    g = torch.Generator(device="cuda:0"); g.manual_seed(20261017)
    rows, total = [], 0
    for step in range(25):
        ebn0 = -2.0 + 0.25 * step
        e_gpu = e_cpu = n_pt = 0
        it_sum = 0
        for _ in range(4):
            payload, llr = _gen(ldpc, torch, g, 1024, 68, ebn0, 1.0 / 3.0)
            it_g, out_g = ldpc.decode_batch_torch(BG, Z, 13, 8, llr)
            h_llr = llr.cpu().numpy()
            it_g, out_g, pay = it_g.cpu().numpy(), out_g.cpu().numpy(), payload.cpu().numpy()
            it_c, out_c = _decode_all(oracle, reference, h_llr, 13, 8)
            bad = np.nonzero(it_g != it_c)[0]
            assert bad.size == 0, (ebn0, "iteration counts differ", bad[:8], it_g[bad[:8]], it_c[bad[:8]])
            assert np.array_equal(out_g, out_c), (ebn0, "output bytes differ", np.nonzero((out_g != out_c).any(axis=1))[0][:8])
            e_gpu += int((out_g[:, :K // 8] != pay).any(axis=1).sum()); e_cpu += int((out_c[:, :K // 8] != pay).any(axis=1).sum())
            it_sum += int(it_c.sum()); n_pt += 1024
        total += n_pt
        rows.append({"ebn0_db": ebn0, "blocks": n_pt, "bler_gpu": e_gpu / n_pt, "bler_cpu_reference": e_cpu / n_pt, "mean_returned_iterations": it_sum / n_pt})
    assert total >= 100000
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump({"config": {"BG": BG, "Z": Z, "K": K, "R_lut": 13, "numMaxIter": 8, "stop": "parity check", "batch": 1024, "blocks_total": total,
                          "cpu_arm": "oracle/_ref/libref_ldpc_dec.so (unmodified nrLDPC_decoder.c, AVX2), all host cores",
                          "compared": "every block: output bytes and returned iteration count, bit exact"}, "points": rows},
              open(os.path.join(ROOT, "gpurun_out", "bler_cpu_gpu_r13.json"), "w"), indent=1)


@pytest.mark.parametrize("E", [None, 9072])
def test_r23_crc_stop_every_block_vs_reference(ldpc, oracle, reference, E):
    """The graph the slot chains decode with (K=8448, decoder LUT R=2/3: 35 columns; E=9072 rate-matched bits at MCS 28, the LLRs beyond E are 0)
    with the CRC24B stop (check_crc called from inside the loop, nrLDPC_decoder.c:850-862): 8 Eb/N0 points x 2048 blocks."""
    import torch
    g = torch.Generator(device="cuda:0"); g.manual_seed(23 + (E or 0))
    rate = 22.0 / 33.0 if E is None else K / float(E)
    lo = 2.0 if E is None else 4.5
    rows = []
    for step in range(8):
        ebn0 = lo + 0.5 * step
        e_gpu = e_cpu = 0
        for _ in range(2):
            payload, llr = _gen(ldpc, torch, g, 1024, 35, ebn0, rate, with_crc24b=True, E=E)
            it_g, out_g = ldpc.decode_batch_torch(BG, Z, 23, 8, llr, use_crc=1, crc_len_bits=K, crc_type=1,
                                                  out=torch.zeros((1024, 35 * Z // 8), dtype=torch.uint8, device=llr.device))
            it_g, out_g, pay = it_g.cpu().numpy(), out_g.cpu().numpy(), payload.cpu().numpy()
            it_c, out_c = _decode_all(oracle, reference, llr.cpu().numpy(), 23, 8, crc=(1, K))
            assert np.array_equal(it_g, it_c), (ebn0, np.nonzero(it_g != it_c)[0][:8])
            assert np.array_equal(out_g, out_c), (ebn0, np.nonzero((out_g != out_c).any(axis=1))[0][:8])
            e_gpu += int((out_g[:, :K // 8] != pay).any(axis=1).sum()); e_cpu += int((out_c[:, :K // 8] != pay).any(axis=1).sum())
        rows.append({"ebn0_db": ebn0, "blocks": 2048, "bler_gpu": e_gpu / 2048, "bler_cpu_reference": e_cpu / 2048})
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump({"config": {"BG": BG, "Z": Z, "K": K, "R_lut": 23, "E": E, "numMaxIter": 8, "stop": "CRC24B (check_crc)", "batch": 1024}, "points": rows},
              open(os.path.join(ROOT, "gpurun_out", "bler_cpu_gpu_r23_E%s.json" % (E or "full")), "w"), indent=1)
