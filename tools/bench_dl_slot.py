#!/usr/bin/env python3
"""nr_dlsim-shaped PDSCH slot chain (BASELINE config 3 / SURVEY 8d "Metric 2", hot-path stages only): one 100 MHz slot, 273 PRB, 2 x 2, two layers, 64QAM,
52 code blocks of K = 8448 through the gNB transmit chain (TB CRC, segmentation, LDPC encode, rate match + interleave, scrambling ... resource mapping, OFDM
modulation) and the UE receive chain (OFDM demod, channel estimation on both DMRS ports, zero-forcing receiver, rate recovery, LDPC decode, TB CRC), device
resident, eager launches and replayed from a CUDA graph, per-stage times, and end to end with the payload coming from pinned host memory and the decoded
transport block going back.  `sweep` prints decoder iteration statistics over channel gain / SNR instead.  Prints JSON lines; summarised under profiles/."""
import json
import os
import sys
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from openairinterface5g_b200.ldpc import load_LDPClib          # noqa: E402
from openairinterface5g_b200.dfts import load_dftslib           # noqa: E402
from openairinterface5g_b200.dl_slot_chain import PdschSlotChain, PdschSlotPipeline   # noqa: E402


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


def sweep(lib, dl, dev, **kw):
    ch = PdschSlotChain(lib, dl, dev, **kw)
    for pseed in (5, 100, 101, 102):
        payload = torch.from_numpy(np.random.default_rng(pseed).integers(0, 256, size=ch.A // 8, dtype=np.uint8)).to(dev)
        tx = ch.transmit(payload).clone()
        for gain in (0.25, 0.5, 1.0, 2.0, 4.0, 8.0):
            for snr in (30.0, 35.0, 45.0, 60.0):
                rx = ch.channel(tx, seed=3 + pseed, snr_db=snr, gain=gain)
                tb, iters, crc = ch.receive(rx)
                torch.cuda.synchronize()
                it = iters.cpu().numpy()
                sat = float((ch.llr16.abs() >= 127).float().mean())
                print(json.dumps({"payload_seed": pseed, "gain": gain, "snr_db": snr, "log2_maxh": int(ch.level.cpu()[8]), "failed_cb": int((it > ch.max_iter).sum()),
                                  "mean_iter": round(float(it.mean()), 2), "max_iter": int(it.max()), "tb_ok": int(crc[0]) == 0, "llr_frac_at_int8_rail": round(sat, 3)}), flush=True)


def throughput(lib, dl, dev, K, n_rounds=60, e2e=False):
    """K slots in flight (PdschSlotPipeline: own buffers + stream + CUDA graph per slot).  Returns (slots/s, number of slots decoded correctly)."""
    pipe = PdschSlotPipeline(lib, dl, dev, K)
    ms = pipe.timed_rounds(n_rounds, e2e=e2e)
    return K * n_rounds / (ms * 1e-3), sum(pipe.check(host=e2e))


def stage_saturation(lib, dl, dev, K=16, n=20):
    """Where the GPU time of a slot goes when the device is full: every stage alone, K chains each on its own stream, us per slot at saturation."""
    pipe = PdschSlotPipeline(lib, dl, dev, K, use_graphs=False)
    st = {
        "tb_crc+segmentation": lambda ch, p, rx: ch.lib.tb_segment_torch(1, ch.A, p, ch.segs, ch.crc1),
        "ldpc_encode": lambda ch, p, rx: lib.encode_batch_torch(1, ch.Z, ch.K, ch.segs, out=ch.cw),
        "rm_tx": lambda ch, p, rx: lib.rm_tx_torch(1, ch.Z, ch.Qm, 0, ch.C, 0, ch.F, ch.cw, ch.E, ch.Eoff, ch.f),
        "scramble..map (pdsch_tx)": lambda ch, p, rx: lib.pdsch_tx_slot_torch(ch.txd, ch.f, ch.txF),
        "ofdm_mod": lambda ch, p, rx: dl.ofdm_mod_slot_torch(ch.dtx, ch.txF, ch.txdata),
        "ofdm_demod": lambda ch, p, rx: dl.ofdm_demod_slot_torch(ch.drx, rx, ch.ts, ch.rxF),
        "channel_estimation": lambda ch, p, rx: lib.pusch_chest_torch(ch.cdesc, ch.rxF, ch.est, ch.chest_scratch, ch.chest_state),
        "level+zf_rx": lambda ch, p, rx: lib.pusch_inner_rx_torch(ch.rxd, ch.rxF, ch.est, ch.llr16, level=ch.level),
        "rm_rx": lambda ch, p, rx: lib.rm_rx_torch(1, ch.Z, ch.Qm, 0, ch.C, 0, ch.F, ch.llr16, ch.E, ch.Eoff, ch.harq, ch.llr8, clear=1),
        "ldpc_decode": lambda ch, p, rx: lib.decode_batch_torch(1, ch.Z, ch.R, ch.max_iter, ch.llr8, use_crc=1, crc_len_bits=ch.K - ch.F, crc_type=1, out=ch.hard, iters=ch.iters),
        "tb copy + crc": lambda ch, p, rx: (ch.tb.view(-1).copy_(ch.hard[:, :ch.nbytes].reshape(-1)), lib.crc_batch_torch(0, ch.tb, ch.A + 24, out=ch.tbcrc)),
    }
    out = {}
    cur = torch.cuda.current_stream(dev)
    REP = 8                                     # each stage 8x per CUDA graph: one replay per chain and round keeps the host out of the measurement
    for name, f in st.items():
        graphs = []
        for k in range(K):
            s = pipe.streams[k]
            with torch.cuda.stream(s):
                f(pipe.chains[k], pipe.payload[k], pipe.rx[k])
                s.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=s):
                    for _ in range(REP):
                        f(pipe.chains[k], pipe.payload[k], pipe.rx[k])
            graphs.append(g)

        def rnd():
            for k in range(K):
                with torch.cuda.stream(pipe.streams[k]):
                    graphs[k].replay()
        for _ in range(3):
            rnd()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(cur); pipe.fork(e0)
        for _ in range(n):
            rnd()
        pipe.join(cur); e1.record(cur)
        torch.cuda.synchronize()
        out[name] = round(1e3 * e0.elapsed_time(e1) / (n * K * REP), 2)
    out["sum"] = round(sum(out.values()), 2)
    return out


def main():
    dev = torch.device("cuda", 0)
    lib, dl = load_LDPClib(), load_dftslib()
    if len(sys.argv) > 1 and sys.argv[1] == "sweep":
        sweep(lib, dl, dev)
        return
    ch = PdschSlotChain(lib, dl, dev)
    if len(sys.argv) > 1 and sys.argv[1] == "once":               # two slots and out: the launch list for ncu
        p0 = torch.from_numpy(np.random.default_rng(5).integers(0, 256, size=ch.A // 8, dtype=np.uint8)).to(dev)
        rx0 = ch.channel(ch.transmit(p0), seed=3)
        for _ in range(2):
            ch.transmit(p0); ch.receive(rx0)
        torch.cuda.synchronize()
        return
    h_payload = torch.from_numpy(np.random.default_rng(5).integers(0, 256, size=ch.A // 8, dtype=np.uint8)).pin_memory()
    payload = h_payload.to(dev)
    tx = ch.transmit(payload)
    rx = ch.channel(tx, seed=3)
    tb, iters, crc = ch.receive(rx)
    torch.cuda.synchronize()
    ok = bool((iters <= ch.max_iter).all()) and int(crc[0]) == 0 and bool((tb.view(-1)[:payload.numel()] == payload).all())
    base = {"workload": f"nr_dlsim-shaped PDSCH slot 100MHz 273PRB 64QAM 2x2 {ch.nl} layers, {ch.C} CB K={ch.K} (TB {ch.A} bit, G {ch.G})", "decoded_ok": ok,
            "mean_iterations": float(iters.float().mean())}
    ss = ch.P.slot_timestamp(ch.slot)

    def slot():                                                   # gNB transmit + UE receive of one slot; the channel is the simulator's and stays outside
        ch.transmit(payload)
        ch.receive(rx)
    l0 = lib.launch_count() + dl.launch_count()
    ms_tx = timed(lambda: ch.transmit(payload), 200)
    k_tx = (lib.launch_count() + dl.launch_count() - l0) / 205
    l0 = lib.launch_count() + dl.launch_count()
    ms_rx = timed(lambda: ch.receive(rx), 200)
    k_rx = (lib.launch_count() + dl.launch_count() - l0) / 205
    ms = timed(slot, 200)
    print(json.dumps(dict(base, mode="device-resident, eager", ms_per_slot=ms, slots_per_s=1e3 / ms, ms_gnb_tx=ms_tx, ms_ue_rx=ms_rx, kernels_tx=k_tx, kernels_rx=k_rx,
                          realtime_factor_vs_2000_slots_per_s=1e3 / ms / 2000.0)), flush=True)
    st = {
        "tb_crc+segmentation": lambda: seg_only(ch, payload),
        "ldpc_encode": lambda: lib.encode_batch_torch(1, ch.Z, ch.K, ch.segs, out=ch.cw),
        "rm_tx": lambda: lib.rm_tx_torch(1, ch.Z, ch.Qm, 0, ch.C, 0, ch.F, ch.cw, ch.E, ch.Eoff, ch.f),
        "scramble..map (pdsch_tx)": lambda: lib.pdsch_tx_slot_torch(ch.txd, ch.f, ch.txF),
        "ofdm_mod": lambda: dl.ofdm_mod_slot_torch(ch.dtx, ch.txF, ch.txdata),
        "ofdm_demod": lambda: dl.ofdm_demod_slot_torch(ch.drx, rx, ch.ts, ch.rxF),
        "channel_estimation": lambda: lib.pusch_chest_torch(ch.cdesc, ch.rxF, ch.est, ch.chest_scratch, ch.chest_state),
        "level+zf_rx": lambda: lib.pusch_inner_rx_torch(ch.rxd, ch.rxF, ch.est, ch.llr16, level=ch.level),
        "rm_rx": lambda: lib.rm_rx_torch(1, ch.Z, ch.Qm, 0, ch.C, 0, ch.F, ch.llr16, ch.E, ch.Eoff, ch.harq, ch.llr8, clear=1),
        "ldpc_decode": lambda: lib.decode_batch_torch(1, ch.Z, ch.R, ch.max_iter, ch.llr8, use_crc=1, crc_len_bits=ch.K - ch.F, crc_type=1, out=ch.hard, iters=ch.iters),
    }
    print(json.dumps(dict(base, mode="per-stage us", **{k: 1e3 * timed(f, 200) for k, f in st.items()})), flush=True)
    try:
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            slot()
            torch.cuda.synchronize()
            with torch.cuda.graph(g, stream=s):
                slot()
        ms_g = timed(g.replay, 500)
        print(json.dumps(dict(base, mode="device-resident, CUDA graph replay", ms_per_slot=ms_g, slots_per_s=1e3 / ms_g)), flush=True)
    except Exception as e:                                           # graph capture is an optimisation, not a requirement
        print(json.dumps(dict(base, mode="CUDA graph", unavailable=str(e)[:200])), flush=True)
    for K in (4, 8, 16, 32):
        try:
            v, okk = throughput(lib, dl, dev, K)
            ve, oke = throughput(lib, dl, dev, K, e2e=True)
            print(json.dumps(dict(base, mode=f"{K} slots in flight (one stream + CUDA graph per slot)", slots_per_s=v, decoded_ok=f"{okk}/{K}", e2e_slots_per_s=ve,
                                  e2e_decoded_ok=f"{oke}/{K}")), flush=True)
        except Exception as e:
            print(json.dumps(dict(base, mode=f"{K} slots in flight", unavailable=str(e)[:300])), flush=True)
    try:
        print(json.dumps(dict(base, mode="per-stage us per slot at saturation (16 chains, one stream each, that stage alone)", **stage_saturation(lib, dl, dev))), flush=True)
    except Exception as e:
        print(json.dumps(dict(base, mode="stage saturation", unavailable=str(e)[:300])), flush=True)
    # end to end: payload from pinned host memory, transport block back to the host (what the MAC hands over / gets back)
    h_tb = torch.empty_like(tb, device="cpu").pin_memory()

    def e2e():
        payload.copy_(h_payload, non_blocking=True)
        ch.transmit(payload)
        t, _, _ = ch.receive(rx)
        h_tb.copy_(t, non_blocking=True)
    ms_e = timed(e2e, 200)
    print(json.dumps(dict(base, mode="e2e (H2D payload, D2H transport block)", ms_per_slot=ms_e, slots_per_s=1e3 / ms_e, h2d_bytes_per_slot=h_payload.numel(),
                          d2h_bytes_per_slot=h_tb.numel())), flush=True)


def seg_only(ch, payload):
    ch.lib.tb_segment_torch(1, ch.A, payload, ch.segs, ch.crc1)


if __name__ == "__main__":
    main()
