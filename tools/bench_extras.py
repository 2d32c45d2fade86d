#!/usr/bin/env python3
"""Secondary measurements (not the headline bench line): Q15 IDFT/DFT 4096 throughput (one 100 MHz slot = 28 transforms for 2 antennas),
LDPC encoder and rate-matching throughput, each with the reference's CPU path timed on the host cores where a compiled reference
exists.  Prints one JSON object per line; used by tools/gpu_round.sh and summarised under profiles/."""
import ctypes as C
import json
import os
import sys
import time
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from openairinterface5g_b200.ldpc import load_LDPClib          # noqa: E402
from openairinterface5g_b200.dfts import load_dftslib           # noqa: E402


def peaks():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def timeit(fn, n=20, warm=3):
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


def cpu_dft(N, inverse, seconds=5.0):
    from oracle import bindings as ob
    if not ob.have_reference():
        return None
    ref = C.CDLL(os.path.join(ob.REFDIR, "libref_dfts.so"))
    ref.dfts_autoinit()
    orc = ob.Oracle()
    f = orc.lib.orc_bench_dft
    f.restype = C.c_long
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.POINTER(C.c_double)]
    x = np.random.default_rng(0).integers(-3000, 3000, size=(64, 2 * N)).astype(np.int16)
    el = C.c_double()
    cores = os.cpu_count() or 1
    n = f(C.cast(getattr(ref, ("idft" if inverse else "dft") + str(N)), C.c_void_p), x.ctypes.data, N, 64, cores, seconds, C.byref(el))
    return {"value": n / el.value, "unit": "transforms/s", "cores": cores, "kind": "reference", "sample": f"{n} {N}-point transforms in {el.value:.1f} s"}


def main():
    dev = torch.device("cuda", 0)
    lib, dl = load_LDPClib(), load_dftslib()
    hbm = peaks()
    g = torch.Generator(device=dev); g.manual_seed(1)
    # ---- DFT 4096: 1024 slots' worth? keep it at 28 transforms x 64 slots = 1792 transforms (29 MB in, 29 MB out)
    for N, inverse, nb in ((4096, True, 1792), (4096, False, 1792), (2048, True, 3584), (1536, False, 3584)):
        bufs = [torch.randint(-3000, 3000, (nb, 2 * N), dtype=torch.int16, device=dev, generator=g) for _ in range(5)]   # 5 x 29 MB > L2
        out = torch.empty_like(bufs[0])
        i = [0]

        def run():
            dl.batch_torch(N, inverse, bufs[i[0] % 5], 1, out=out); i[0] += 1
        ms = timeit(run, n=40)
        algo = nb * N * 4 * 2
        line = {"what": f"{'idft' if inverse else 'dft'}{N}", "batch": nb, "ms": ms, "value": nb / ms * 1e3, "unit": "transforms/s",
                "roofline": {"bound": "hbm", "achieved": algo / ms / 1e6, "peak": hbm, "unit": "GB/s", "frac": algo / ms / 1e6 / hbm, "algorithmic_bytes_per_transform": N * 8}}
        if N == 4096:
            line["cpu_baseline"] = cpu_dft(N, inverse)
            h = bufs[0].cpu().numpy()
            t0 = time.perf_counter()
            for _ in range(5):
                dl.batch_host(N, inverse, h, 1)
            line["e2e_transforms_per_s"] = 5 * nb / (time.perf_counter() - t0)
        print(json.dumps(line), flush=True)
    # ---- encoder + rate matching (BG1 Z=384, batch 1024)
    B, Z, K = 1024, 384, 8448
    payload = torch.randint(0, 256, (B, K // 8), dtype=torch.uint8, device=dev, generator=g)
    cw = torch.empty((B, 66 * Z), dtype=torch.uint8, device=dev)
    ms = timeit(lambda: lib.encode_batch_torch(1, Z, K, payload, out=cw), n=20)
    algo = B * (K // 8 + 66 * Z)
    print(json.dumps({"what": "ldpc_encode BG1 Z=384", "batch": B, "ms": ms, "value": B / ms * 1e3, "unit": "CB/s",
                      "roofline": {"bound": "hbm", "achieved": algo / ms / 1e6, "peak": hbm, "unit": "GB/s", "frac": algo / ms / 1e6 / hbm, "algorithmic_bytes_per_cb": K // 8 + 66 * Z}}), flush=True)


    # ---- PUSCH max-log LLRs, 64QAM, 64 slots x 13 symbols x 3276 REs (HBM bound: 12 B in + 12 B out per RE)
    for Qm in (6, 8, 2):
        n = 3276 * 13 * 64
        ys = [torch.randint(-8000, 8000, (2 * n,), dtype=torch.int16, device=dev, generator=g) for _ in range(6)]
        ma, mb, mc = (torch.randint(0, 20000, (2 * n,), dtype=torch.int16, device=dev, generator=g) for _ in range(3))
        out = torch.empty(n * Qm, dtype=torch.int16, device=dev)
        i = [0]

        def run():
            lib.pusch_llr_torch(Qm, ys[i[0] % 6], ma, mb, mc, out=out); i[0] += 1
        ms = timeit(run, n=40)
        planes = {2: 1, 4: 2, 6: 3, 8: 4}[Qm]
        algo = n * (4 * planes + 2 * Qm)
        print(json.dumps({"what": f"pusch_llr Qm={Qm}", "re": n, "ms": ms, "value": n / ms * 1e3, "unit": "RE/s",
                          "roofline": {"bound": "hbm", "achieved": algo / ms / 1e6, "peak": hbm, "unit": "GB/s", "frac": algo / ms / 1e6 / hbm, "algorithmic_bytes_per_re": 4 * planes + 2 * Qm}}), flush=True)
    # ---- slot-level OFDM front end at 100 MHz (N=4096, mu=1, 273 PRB): 64 antenna-slots per launch = 896 transforms
    from openairinterface5g_b200.ofdm import NrOfdmParms
    Pm = NrOfdmParms(4096, 1, 273)
    rot = Pm.symbol_rotation(3619200000.0)
    na = int(os.environ.get("NRB200_OFDM_ANTENNA_SLOTS", "64"))
    dtx = Pm.desc(1, na, rot)
    Fs = [torch.randint(-3000, 3000, (na, 14 * 4096 * 2), dtype=torch.int16, device=dev, generator=g) for _ in range(5)]
    tx = torch.empty((na, 2 * dtx.t_stride), dtype=torch.int16, device=dev)
    i = [0]

    def run_tx():
        dl.ofdm_mod_slot_torch(dtx, Fs[i[0] % 5], tx); i[0] += 1
    ms = timeit(run_tx, n=40)
    algo = na * (14 * 4096 * 4 + dtx.t_stride * 4)
    line = {"what": "ofdm_mod_slot 4096/273PRB (rotation + IDFT + CP)", "antenna_slots": na, "ms": ms, "value": na / ms * 1e3, "unit": "antenna-slots/s",
            "roofline": {"bound": "hbm", "achieved": algo / ms / 1e6, "peak": hbm, "unit": "GB/s", "frac": algo / ms / 1e6 / hbm,
                         "algorithmic_bytes_per_antenna_slot": algo // na}}
    print(json.dumps(line), flush=True)
    drx = Pm.desc(1, na, rot, rx=True, t_stride=2 * Pm.samples_per_slot0)
    drx.t_ring = 0
    base = Pm.slot_timestamp(1) - 64
    for l in range(14):
        drx.t_off[l] = drx.t_off[l] - base
    Xs = [torch.randint(-3000, 3000, (na, 2 * 2 * Pm.samples_per_slot0), dtype=torch.int16, device=dev, generator=g) for _ in range(3)]
    ts = torch.from_numpy(Pm.timeshift_rotation()).to(dev)
    rxF = torch.empty((na, 14 * 4096 * 2), dtype=torch.int16, device=dev)

    def run_rx():
        dl.ofdm_demod_slot_torch(drx, Xs[i[0] % 3], ts, rxF); i[0] += 1
    ms = timeit(run_rx, n=40)
    algo = na * (14 * 4096 * 4 * 2)
    print(json.dumps({"what": "ofdm_demod_slot 4096/273PRB (window + DFT + rotation/timeshift)", "antenna_slots": na, "ms": ms, "value": na / ms * 1e3,
                      "unit": "antenna-slots/s", "roofline": {"bound": "hbm", "achieved": algo / ms / 1e6, "peak": hbm, "unit": "GB/s",
                                                               "frac": algo / ms / 1e6 / hbm, "algorithmic_bytes_per_antenna_slot": algo // na}}), flush=True)
    try:
        from oracle import bindings as ob
        if ob.have_reference():
            ref = ob.Reference()
            Fh = Fs[0][0].cpu().numpy()
            rot224 = np.zeros(448, np.int16); rot224[:rot.size] = rot.reshape(-1)
            t0 = time.perf_counter(); k = 0
            while time.perf_counter() - t0 < 2.0:
                ref.ofdm_tx_slot(4096, 1, 273, 1, 14, rot224, Fh, dtx.t_stride); k += 1
            print(json.dumps({"what": "reference apply_nr_rotation_TX + nr_normal_prefix_mod (1 host thread)", "value": k / (time.perf_counter() - t0),
                              "unit": "antenna-slots/s", "kind": "reference"}), flush=True)
    except Exception as e:                                    # the CPU comparison is optional
        print(json.dumps({"what": "reference ofdm", "unavailable": str(e)}), flush=True)
    # ---- scrambling + QAM mapper on a full-band 2-layer 64QAM PDSCH codeword (G = 471744 bits), 64 codewords per timing loop
    G = 471744
    bits = torch.randint(0, 2, (G,), dtype=torch.uint8, device=dev, generator=g)
    words = torch.empty((G + 31) // 32 + 1, dtype=torch.int32, device=dev)
    sym = torch.empty(2 * (G // 6), dtype=torch.int16, device=dev)
    llrs = torch.randint(-3000, 3000, (G,), dtype=torch.int16, device=dev, generator=g)
    ms = timeit(lambda: lib.scramble_torch(bits, 0, 42, 4660, words), n=50)
    print(json.dumps({"what": "nr_codeword_scrambling G=471744", "us": ms * 1e3, "value": G / ms * 1e3, "unit": "bits/s"}), flush=True)
    ms = timeit(lambda: lib.modulate_torch(words, G, 6, sym), n=50)
    print(json.dumps({"what": "nr_modulation 64QAM G=471744", "us": ms * 1e3, "value": G / 6 / ms * 1e3, "unit": "symbols/s"}), flush=True)
    ms = timeit(lambda: lib.unscramble_llr_torch(llrs, 0, 42, 4660), n=50)
    print(json.dumps({"what": "nr_codeword_unscrambling G=471744", "us": ms * 1e3, "value": G / ms * 1e3, "unit": "LLR/s"}), flush=True)
    # ---- rfsimulator channel application (rxAddInput): one 10 ms frame at 61.44 Msps, 2 x 2 and 4 x 4, 40 taps, with noise; FP64, 8 operations per tap and tx antenna
    from openairinterface5g_b200.ldpc import RfsimChan
    for nb in (2, 4):
        L, n = 40, 614400
        cir = 2 * (n + L + 16)
        rngn = np.random.default_rng(5)
        chn = rngn.normal(size=(nb * nb, L, 2)) * 0.1
        sign = rngn.integers(-6000, 6001, size=(cir, 2)).astype(np.int16)
        d_ch, d_sig = torch.from_numpy(chn).to(dev), torch.from_numpy(sign).to(dev)
        d_nz = torch.randn((nb, n, 2), dtype=torch.float64, device=dev, generator=g)
        d_o = torch.zeros((nb, n, 2), dtype=torch.int16, device=dev)
        dsc = RfsimChan(nb, nb, L, 0, -2.0, -30.0, 0)
        ms = timeit(lambda: lib.rfsim_rx_add_input_torch(dsc, d_ch, d_sig, d_o, 614400, d_nz), n=20)
        line = {"what": f"rfsim rxAddInput {nb}x{nb}, {L} taps, {n} samples per antenna", "ms": ms, "value": nb * n / ms * 1e3, "unit": "rx-antenna samples/s",
                "fp64_gops": 8.0 * nb * L * nb * n / ms / 1e6, "realtime_factor_61.44Msps": n / ms * 1e3 / 61.44e6}
        try:
            from oracle import bindings as ob
            ref = ob.Reference()
            k = 30720
            t0 = time.perf_counter()
            ref.rfsim_rx_add_input(nb, nb, L, 0, -2.0, -30.0, chn, sign, np.zeros((k, 2), np.int16), 0, 614400, cir, rngn.normal(size=(k, 2)))
            line["reference_1_thread"] = k / (time.perf_counter() - t0)
        except Exception as e:      # the compiled reference is absent
            line["reference_1_thread"] = str(e)
        print(json.dumps(line), flush=True)
    # ---- gNB PRACH detector (rx_nr_prach): long sequences, N_CS 13 (64 roots... 1 root per preamble group), 4 rx antennas; occasions back to back on one stream
    try:
        from openairinterface5g_b200.ldpc import PrachDesc
        gold = np.load(os.path.join(ROOT, "tests", "golden", "prach.npz"))
        for ci, (nrx, short, ncs, fmt, mu) in ((1, (2, 0, 13, 0, 1)), (7, (4, 1, 0, 7, 3))):
            xu = gold[f"xu{ci}"]
            nzc = 139 if short else 839
            dsc = PrachDesc(nrx, short, ncs, fmt, mu, 0, nzc, 0)
            d_xu = torch.from_numpy(xu).to(dev)
            rxs = torch.randint(-800, 801, (nrx, nzc, 2), dtype=torch.int16, device=dev, generator=g)
            o3 = torch.zeros(3, dtype=torch.int32, device=dev)
            scr = torch.empty(lib.prach_scratch_bytes(dsc), dtype=torch.uint8, device=dev)
            ms = timeit(lambda: lib.rx_nr_prach_torch(dsc, d_xu, rxs, o3, scr), n=50)
            line = {"what": f"rx_nr_prach N_ZC={nzc} N_CS={ncs} {nrx} rx ({lib.prach_num_roots(dsc)} roots)", "us": ms * 1e3, "value": 1e3 / ms, "unit": "occasions/s"}
            try:
                from oracle import bindings as ob
                ref = ob.Reference()
                hx = rxs.cpu().numpy()
                t0 = time.perf_counter(); k = 0
                while time.perf_counter() - t0 < 2.0:
                    ref.rx_nr_prach(nrx, short, 22, lib.prach_num_roots(dsc), ncs, fmt, mu, xu, hx); k += 1
                line["reference_1_thread"] = k / (time.perf_counter() - t0)
            except Exception as e:
                line["reference_1_thread"] = str(e)
            print(json.dumps(line), flush=True)
    except Exception as e:
        print(json.dumps({"what": "rx_nr_prach", "error": str(e)}), flush=True)
    # ---- UE slot receiver, 273 PRB, 64QAM, 4 rx: level + receiver launches of one slot, device resident; 1 layer with and without PT-RS, 2, 3 and 4 layers
    from openairinterface5g_b200.ldpc import PuschRxDesc
    N, nrx, nbr = 4096, 4, 273
    rxF = torch.randint(-2000, 2001, (nrx, 14 * N, 2), dtype=torch.int16, device=dev, generator=g)
    est = torch.randint(-1500, 1501, (4 * nrx, 14 * N, 2), dtype=torch.int16, device=dev, generator=g)
    lvl = torch.zeros(9, dtype=torch.int32, device=dev)
    pst = torch.zeros(32, dtype=torch.int32, device=dev)
    for nl, ptrs in ((1, False), (1, True), (2, False), (3, False), (4, False)):
        dsc = PuschRxDesc(N, nrx, 0, 0, nbr, N - nbr * 6, 6, 1, 13, 1 << 2, 0, 2, 0, 14 * N, 14 * N, 1, 0x1234, 77, nl, 0, 0, 1)
        if ptrs:
            dsc.set_ptrs(1, 2, 0, 3, 0, 55, pst.data_ptr())
        nllr = lib.pusch_num_llr(dsc)
        llr16 = torch.empty(nllr, dtype=torch.int16, device=dev)
        ms = timeit(lambda: lib.pusch_inner_rx_torch(dsc, rxF, est, llr16, level=lvl), n=50)
        in_b = 12 * nbr * 12 * (nrx + nl * nrx) * 4
        print(json.dumps({"what": f"UE nr_rx_pdsch slot 273PRB 64QAM 4rx, {nl} layer(s){', PT-RS L=2 K=2' if ptrs else ''} (level + receiver)", "us": ms * 1e3,
                          "value": 1e3 / ms, "unit": "slots/s", "llr": nllr, "algorithmic_GBps": (in_b + 2 * nllr) / ms / 1e6}), flush=True)
    # (the per-call LDPCdecoder / LDPCencoder ABI is measured by tools/bench_abi.py through a C harness: Python caller threads would measure the interpreter lock)


if __name__ == "__main__":
    main()
