#!/usr/bin/env python3
"""PUSCH slot receive chain throughput (SURVEY 8d "Metric 2" analogue for the uplink, hot-path stages only): one full-band 100 MHz slot
(273 PRB, 64QAM, 4 rx antennas, 28 code blocks) through OFDM demod -> channel estimation -> level -> compensation/LLR/descrambling -> rate recovery -> LDPC decode
-> TB CRC, device resident, eager launches and replayed from a CUDA graph, plus the same with the slot's samples copied from pinned host
memory and the transport block copied back.  Prints JSON lines; summarised under profiles/."""
import json
import os
import sys
import time
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from openairinterface5g_b200.ldpc import load_LDPClib          # noqa: E402
from openairinterface5g_b200.dfts import load_dftslib           # noqa: E402
from openairinterface5g_b200.slot_chain import PuschSlotChain   # noqa: E402


def timed(fn, n, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    dev = torch.device("cuda", 0)
    lib, dl = load_LDPClib(), load_dftslib()
    for nl in (1, 2):
        run(lib, dl, dev, nl)
    # throughput: K independent slots in flight (one stream + one CUDA graph per slot), device resident and with the samples / transport block crossing PCIe
    from openairinterface5g_b200.slot_chain import PuschSlotPipeline
    for K, nl in ((4, 1), (8, 1), (16, 1), (4, 2), (8, 2), (16, 2)):
        pipe = PuschSlotPipeline(lib, dl, dev, K) if nl == 1 else PuschSlotPipeline(lib, dl, dev, K, A=471272, n_layers=2)
        rounds = max(4, 256 // K)
        ms = pipe.timed_rounds(rounds) / (rounds * K)
        ok = pipe.check()
        ms_e = pipe.timed_rounds(rounds, e2e=True) / (rounds * K)
        ok_e = pipe.check(host=True)
        print(json.dumps({"workload": f"PUSCH slot rx 100MHz 273PRB 64QAM 4rx {nl} layer(s), {pipe.chains[0].C} CB K=8448 (TB {pipe.chains[0].A} bit)", "mode": f"{K} slots in flight (one stream + CUDA graph per slot)",
                          "slots_per_s": 1e3 / ms, "decoded_ok": f"{sum(ok)}/{K}", "e2e_slots_per_s": 1e3 / ms_e, "e2e_decoded_ok": f"{sum(ok_e)}/{K}",
                          "h2d_bytes_per_slot": int(pipe.h_rx[0].numel() * 2), "d2h_bytes_per_slot": int(pipe.h_tb[0].numel()),
                          "realtime_factor_vs_2000_slots_per_s": 1e3 / ms / 2000.0}), flush=True)
        del pipe
        torch.cuda.empty_cache()


def run(lib, dl, dev, nl):
    ch = PuschSlotChain(lib, dl, dev) if nl == 1 else PuschSlotChain(lib, dl, dev, A=471272, n_layers=2)
    payload, rxdata, est = ch.synthesize(seed=3, snr_db=30.0)
    tb, iters, crc = ch.receive(rxdata)
    torch.cuda.synchronize()
    ok = bool((iters <= ch.max_iter).all()) and int(crc[0]) == 0 and bool((tb.view(-1)[:payload.size].cpu() == torch.from_numpy(payload)).all())
    base = {"workload": f"PUSCH slot rx 100MHz 273PRB 64QAM 4rx {nl} layer(s), {ch.C} CB K=8448 (TB {ch.A} bit)", "decoded_ok": ok,
            "mean_iterations": float(iters.float().mean())}
    l0 = lib.launch_count() + dl.launch_count()
    ms = timed(lambda: ch.receive(rxdata), 200)
    launches = (lib.launch_count() + dl.launch_count() - l0) / 205
    print(json.dumps(dict(base, mode="device-resident, eager", ms_per_slot=ms, slots_per_s=1e3 / ms, kernels_per_slot=launches,
                          realtime_factor_vs_2000_slots_per_s=1e3 / ms / 2000.0)), flush=True)
    # per-stage times (each stage alone, back to back 200x)
    st = {
        "ofdm_demod": lambda: dl.ofdm_demod_slot_torch(ch.drx, rxdata, ch.ts, ch.rxF),
        "channel_estimation": lambda: lib.pusch_chest_torch(ch.cdesc, ch.rxF, ch.est, ch.chest_scratch, ch.chest_state),
        "level+inner_rx": lambda: lib.pusch_inner_rx_torch(ch.desc, ch.rxF, ch.est, ch.llr16, level=ch.level),
        "rm_rx": lambda: lib.rm_rx_torch(1, ch.Z, ch.Qm, 0, ch.C, 0, ch.F, ch.llr16, ch.E, ch.Eoff, ch.harq, ch.llr8, clear=1),
        "ldpc_decode": lambda: lib.decode_batch_torch(1, ch.Z, ch.R, ch.max_iter, ch.llr8, use_crc=1, crc_len_bits=ch.K - ch.F, crc_type=1, out=ch.hard, iters=ch.iters),
    }
    print(json.dumps(dict(base, mode="per-stage us", **{k: 1e3 * timed(f, 200) for k, f in st.items()})), flush=True)
    try:
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            ch.receive(rxdata)
            torch.cuda.synchronize()
            with torch.cuda.graph(g, stream=s):
                ch.receive(rxdata)
        ms_g = timed(g.replay, 500)
        print(json.dumps(dict(base, mode="device-resident, CUDA graph replay", ms_per_slot=ms_g, slots_per_s=1e3 / ms_g)), flush=True)
    except Exception as e:                                           # graph capture is an optimisation, not a requirement
        print(json.dumps(dict(base, mode="CUDA graph", unavailable=str(e)[:200])), flush=True)
    # end to end: the slot's time-domain samples come from pinned host memory, the transport block goes back
    ss = ch.P.slot_timestamp(ch.slot)
    nsamp = ch.P.samples_per_slot(ch.slot)
    h_slot = rxdata[:, ss:ss + nsamp].contiguous().cpu().pin_memory()
    h_tb = torch.empty_like(tb, device="cpu").pin_memory()

    def e2e():
        rxdata[:, ss:ss + nsamp].copy_(h_slot, non_blocking=True)
        t, _, _ = ch.receive(rxdata)
        h_tb.copy_(t, non_blocking=True)
    t0 = time.perf_counter()
    ms_e = timed(e2e, 200)
    print(json.dumps(dict(base, mode="e2e (H2D slot samples, D2H transport block)", ms_per_slot=ms_e, slots_per_s=1e3 / ms_e,
                          h2d_bytes_per_slot=h_slot.numel() * 2, d2h_bytes_per_slot=h_tb.numel())), flush=True)


if __name__ == "__main__":
    main()
