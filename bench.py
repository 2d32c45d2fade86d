#!/usr/bin/env python3
"""Headline benchmark: NR LDPC decode throughput, code-blocks/s (BASELINE.json metric 1, configs[1]):
ldpctest BG1 Z=384 K=8448 R=1/3 LUT, numMaxIter=8, outMode=BIT, parity-check early stop, batch=1024 int8-LLR code blocks
per GPU, inputs below the waterfall (Eb/N0 1.0 dB: every block runs all 8+1 passes = fixed work).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one process per GPU under torchrun)
  python bench.py --impl reference --gpus N --steps K ...   # the reference's own AVX2 CPU decoder on the host cores

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for what each key means.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BG, Z, R, K, NCOLS, MAX_ITER = 1, 384, 13, 8448, 68, 8
NUM_LLR = NCOLS * Z                      # 26112 int8 in
ALGO_BYTES_PER_CB = NUM_LLR + K // 8     # 27168 B: LLRs in + K/8 hard bits out
EDGE_UPDATES_PER_PASS = 2 * 316 * Z      # 242688 message updates per flooding iteration
METRIC = "LDPC code-blocks/sec (BG1 Z=384 K=8448 8-iter)"
WORKLOAD = "ldpctest BG1 Z=384 K=8448 R=1/3 8-iter batch=1024 int8 LLR, Eb/N0 1.0 dB (all blocks run 9 passes)"



def _ncu_capture():
    """profiles/ncu_decode_traffic.json (written by tools/ncu_traffic_json.py from an `ncu --set full` capture of the same 1024-block launch), but only
    if it was taken from the kernel sources this run executes: the file carries their SHA-1 and a stale capture is not reported."""
    try:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        from ncu_traffic_json import kernel_source_sha1
        d = json.load(open(os.path.join(ROOT, "profiles", "ncu_decode_traffic.json")))
        return d if d.get("kernel_source_sha1") == kernel_source_sha1() else None
    except Exception:
        return None


def ncu_traffic(n_cb):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the decode kernel from the ncu capture; scaled if this run's batch differs.
    None if there is no capture of the current kernel sources."""
    d = _ncu_capture()
    return int((d["dram_bytes_read"] + d["dram_bytes_write"]) * n_cb / 1024) if d else None


def ncu_on_chip(n_cb, kernel_s, sm_mhz, sm_count=148):
    """The resource that does bound the decoder: the integer ALU pipe (LOP3 / PRMT / IADD3 / SHF / VABSDIFF4; 16 lanes per scheduler = one
    warp instruction per 2 cycles per SM sub-partition).  Warp instructions on that pipe per code block come from the committed ncu capture
    (sm__pipe_alu_cycles_active of the same 1024-block launch, profiles/ncu_decode_traffic.json); rate = this run's blocks / this run's
    kernel time, peak = SMs x 4 schedulers x 0.5 x the SM clock sampled during the timed region."""
    try:
        d = _ncu_capture()
        if d is None:
            return None
        per_cb = float(d["alu_pipe_warp_inst_per_cb"])
        achieved = n_cb * per_cb / kernel_s / 1e9
        peak = sm_count * 4 * 0.5 * float(sm_mhz) * 1e6 / 1e9
        return {"bound": "alu_pipe", "achieved": achieved, "peak": peak, "unit": "G warp-inst/s", "frac": achieved / peak,
                "alu_pipe_warp_inst_per_cb": per_cb, "warp_inst_per_cb": d.get("warp_inst_per_launch", 0) / 1024.0, "ncu": {k: d[k] for k in ("alu_pipe_pct", "issue_active_pct", "fmaheavy_pipe_pct", "lsu_pipe_pct", "source", "kernel_source_sha1") if k in d}}
    except Exception:
        return None


def measured_mix_ceiling():
    """tools/ubench/alu_ceiling (built by __graft_entry__.build()): warp instructions per cycle per scheduler this GPU sustains on the decoder's
    opcode blend with nothing but the pipes in the way (no barriers, branches or dependent descriptor loads), and on LOP3 alone."""
    exe = os.path.join(ROOT, "tools", "ubench", "_bin", "alu_ceiling")
    if not os.path.exists(exe):
        return None
    try:
        return json.loads(subprocess.run([exe], capture_output=True, text=True, timeout=120).stdout.strip().splitlines()[-1])
    except Exception:
        return None


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while a timed region runs."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ reference CPU arm
def cpu_throughput(llr, seconds, threads):
    """Time the CPU decoder on `threads` host threads for ~`seconds` (pthreads inside oracle/cpu_bench.c, one blocking
    LDPCdecoder call per code block like ldpctest.c:329-340).  Returns (kind, CB/s, decodes, elapsed, mean returned iterations)."""
    import ctypes as C
    from oracle import bindings as ob
    orc = ob.Oracle()
    fn, kind = None, "port"
    if ob.have_reference():
        ref = ob.Reference()
        fn, kind = C.cast(ref.dec.LDPCdecoder, C.c_void_p), "reference"
    f = orc.lib.orc_bench_ldpc_decoder
    f.restype = C.c_long
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
                  C.POINTER(C.c_double), C.POINTER(C.c_double)]
    el, mi = C.c_double(), C.c_double()
    n = f(fn, llr.ctypes.data, llr.shape[0], llr.shape[1], BG, Z, R, MAX_ITER, threads, seconds, C.byref(el), C.byref(mi))
    return kind, n / el.value, int(n), el.value, mi.value


def make_llr_numpy(n_cb, ebn0_db, seed):
    """Reference-sized (27000-byte rows, 64-byte aligned) int8 LLR rows generated with the oracle encoder (CPU-only path)."""
    from oracle.bindings import Oracle
    from openairinterface5g_b200.synth import awgn_llr, random_payloads
    orc = Oracle()
    P = random_payloads(n_cb, K, seed)
    cw = np.stack([orc.encode(BG, Z, K, P[i]) for i in range(n_cb)])
    llr = awgn_llr(cw, Z, NCOLS, ebn0_db, 1.0 / 3.0, seed)
    buf = np.zeros((n_cb, 27008), dtype=np.int8)
    buf[:, :NUM_LLR] = llr
    return buf


def run_reference(args, rank):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample = 64
    llr = make_llr_numpy(sample, args.ebn0, 1)
    per_step_s = max(1.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))
    for _ in range(args.warmup):
        cpu_throughput(llr, min(per_step_s, 2.0), cores)
    tot, t = 0, 0.0
    kind = "reference"
    for _ in range(args.steps):
        kind, _, n, dt, mi = cpu_throughput(llr, per_step_s, cores)
        tot += n; t += dt
    v = tot / t
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "CB/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000.0 * t / max(1, args.steps), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int8",
            "data": "synthetic", "config": {"workload": WORKLOAD, "cpu_path": "OAI nrLDPC_decoder.c AVX2 (-O3 -mavx2 -mno-avx512f), one thread per host core"},
            "cpu_baseline": {"value": v, "unit": "CB/s", "cores": cores, "kind": kind,
                             "sample": f"{sample} distinct code blocks decoded round-robin on {cores} threads, {per_step_s:.1f} s per step, mean returned iterations {mi:.2f}"},
            "e2e": {"value": v, "unit": "CB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    if not args.no_slot:
        try:
            cb = slot_cpu_baseline(args.slot_cpu_seconds)
            if cb is not None:
                line["nr_dlsim_slot"] = {"metric": SLOT_METRIC, "value": cb["value"], "unit": "slots/s", "config": {"workload": SLOT_WORKLOAD}, "cpu_baseline": cb}
        except Exception as e:
            line["nr_dlsim_slot"] = {"metric": SLOT_METRIC, "unavailable": repr(e)[:300]}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ second half of BASELINE.json's metric: nr_dlsim slots/s
SLOT_METRIC = "nr_dlsim slots/sec @100MHz (hot-path stages of one slot: gNB PDSCH transmit + UE PDSCH receive)"
SLOT_WORKLOAD = "273 PRB mu=1 2x2, 2 layers, 64QAM, MCS28-shaped TB 434280 bit = 52 code blocks K=8448 E=9072, 1 DMRS symbol, 8-iter CRC-stop decode"


def slot_cpu_baseline(seconds):
    """The same slot through the unmodified reference functions (oracle/dl_slot_ref.py; timers around the reference calls only): one process per host core, each
    running whole slots on its own thread the way nr_dlsim runs them; the aggregate is the sum over processes."""
    from oracle import bindings as ob
    if not ob.have_reference():
        return None
    cores = os.cpu_count() or 1
    procs = [subprocess.Popen([sys.executable, "-m", "oracle.dl_slot_ref", str(seconds)], cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
             for _ in range(cores)]
    res = []
    for p in procs:
        out, _ = p.communicate()
        try:
            res.append(json.loads(out.strip().splitlines()[-1]))
        except Exception:
            pass
    if not res:
        return None
    one = max(r["slots_per_s"] for r in res)
    return {"value": sum(r["slots_per_s"] for r in res), "unit": "slots/s", "cores": len(res), "kind": "reference", "decoded_ok": all(r["decoded_ok"] for r in res),
            "one_thread_slots_per_s": one,
            "sample": f"{sum(r['slots'] for r in res)} slots on {len(res)} processes x 1 thread, {seconds:.0f} s each; per process the stages of a slot run in sequence as in nr_dlsim; "
                      "time counted inside the reference functions only",
            "stages_us": {k: round(v, 1) for k, v in res[0]["stages_us"].items()}}


def slot_b200(lib, dev, cpu_seconds, inflight=16):
    """Device-resident slot chain (openairinterface5g_b200/dl_slot_chain.py).  `value`: slots/s with `inflight` independent slots in flight on one GPU
    (PdschSlotPipeline: one stream + one CUDA graph per slot -- the counterpart of the reference arm's one-process-per-core), CUDA-event timed on the stream all
    slot streams fork from and join into; `e2e`: the same with every slot's payload copied from pinned host memory and its decoded transport block copied back
    inside the timed region; `latency`: one slot alone, eager launches (what a single UE's slot costs)."""
    import torch
    from openairinterface5g_b200.dfts import load_dftslib
    from openairinterface5g_b200.dl_slot_chain import PdschSlotChain, PdschSlotPipeline
    dl = load_dftslib()
    ch = PdschSlotChain(lib, dl, dev)
    h_payload = torch.from_numpy(np.random.default_rng(5).integers(0, 256, size=ch.A // 8, dtype=np.uint8)).pin_memory()
    payload = h_payload.to(dev)
    rx = ch.channel(ch.transmit(payload), seed=3)
    tb, iters, crc = ch.receive(rx)
    torch.cuda.synchronize()
    ok1 = bool((iters <= ch.max_iter).all()) and int(crc[0]) == 0 and bool((tb.view(-1)[:payload.numel()] == payload).all())
    h_tb = torch.empty_like(tb, device="cpu").pin_memory()

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

    def slot():
        ch.transmit(payload)
        ch.receive(rx)

    def e2e():
        payload.copy_(h_payload, non_blocking=True)
        ch.transmit(payload)
        t, _, _ = ch.receive(rx)
        h_tb.copy_(t, non_blocking=True)
    l0 = lib.launch_count() + dl.launch_count()
    ms1 = timed(slot, 200)
    launches = (lib.launch_count() + dl.launch_count() - l0) / 205
    ms1_e = timed(e2e, 200)
    # throughput: `inflight` slots, distinct payloads and channel realisations, every one checked after the timed region
    pipe = PdschSlotPipeline(lib, dl, dev, inflight)
    rounds = 40
    ms = pipe.timed_rounds(rounds) / (rounds * inflight)
    okk = pipe.check()
    ms_e = pipe.timed_rounds(rounds, e2e=True) / (rounds * inflight)
    oke = pipe.check(host=True)
    mean_it = float(np.mean([float(c.iters.float().mean()) for c in pipe.chains]))
    # compulsory HBM traffic of one slot: samples out and in, both grids, estimates, LLRs, soft buffers, code words ("2-3 MB/slot")
    nsamp = ch.txdata.shape[1]
    algo = (2 * ch.nb * nsamp * 4 + 2 * ch.nb * 14 * ch.N * 4 + ch.nl * ch.nb * ch.N * 4 + 2 * ch.G * 2 + ch.G + ch.C * (66 + 68 + 66 * 2) * ch.Z + 2 * ch.A // 8)
    out = {"metric": SLOT_METRIC, "value": 1e3 / ms, "unit": "slots/s", "ms_per_slot": ms,
           "config": {"workload": SLOT_WORKLOAD, "slots_in_flight": inflight,
                      "parallelism": f"{inflight} independent slots (own payload, channel realisation, buffers), one CUDA stream + one CUDA graph per slot"},
           "decoded_ok": bool(all(okk)) and ok1, "slots_decoded": f"{sum(okk)}/{inflight}", "mean_iterations": mean_it, "gpu_launches_per_slot": launches,
           "e2e": {"value": 1e3 / ms_e, "unit": "slots/s", "h2d_bytes_per_step": int(h_payload.numel()), "d2h_bytes_per_step": int(h_tb.numel()),
                   "slots_decoded": f"{sum(oke)}/{inflight}"},
           "latency": {"one_slot_ms": ms1, "one_slot_slots_per_s": 1e3 / ms1, "one_slot_e2e_ms": ms1_e, "decoded_ok": ok1,
                       "note": "one slot alone, eager launches: 16 dependent launches of 10-70 us each on 52 code blocks, latency bound"},
           "roofline": {"bound": "hbm", "achieved": algo / (ms * 1e-3) / 1e9, "peak": _peaks()[0], "unit": "GB/s", "frac": algo / (ms * 1e-3) / 1e9 / _peaks()[0],
                        "algorithmic_bytes_per_slot": int(algo),
                        "note": "per slot the GPU time is decode 12 us, encode 6, channel estimation 5, rate recovery 3, CRCs 6, the rest 11 (profiles/dlslot_*): "
                                "the decoder is ALU bound with its state in shared memory, the small kernels are latency bound"},
           "realtime_factor_vs_2000_slots_per_s": 1e3 / ms / 2000.0,
           "parity": "bit-exact end to end against the unmodified reference functions (tests/test_gpu_dl_slot_chain.py::test_pdsch_slot_vs_reference_chain)"}
    if cpu_seconds > 0:
        cb = slot_cpu_baseline(cpu_seconds)
        if cb is not None:
            out["cpu_baseline"] = cb
    return out


def plugin_abi(with_reference):
    """The drop-in boundary itself: unmodified host code makes ONE blocking LDPCdecoder call per code block through the four dlsym'ed symbols
    (ldpctest.c:329-340 serially, nr_ulsch_decoding.c:435-468 from tpool workers).  tools/abi_bench.c does exactly that from 1 and from `cores` host
    threads against libldpc_b200.so and -- same binary, inputs, threads: the cpu_baseline of this key -- against the compiled reference decoder; every
    call's output and iteration count is checked against the reference's."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench_abi
    cores = os.cpu_count() or 8
    ours, ref = bench_abi.sweep(seconds=1.5, ebn0=1.0, ours_threads=(1, cores), ref_threads=(1, cores), with_reference=with_reference)
    keep = ("host_threads", "value", "us_per_call_mean", "us_per_call_p50", "us_per_call_p99", "mismatches", "checked", "blocks_per_launch", "us_device")
    out = {"metric": "LDPCdecoder calls/s through the OAI loader ABI (one code block per blocking call)", "unit": "CB/s", "workload": WORKLOAD.replace("batch=1024 ", ""),
           "b200": [{k: d.get(k) for k in keep} | ({"error": d["error"]} if "error" in d else {}) for d in ours],
           "path": "mapped pinned staging row read / written by the kernel, callers combined at the launch lock, 8-CTA cluster kernel (csrc/nrb200_ll.cu)"}
    if ref:
        out["cpu_baseline"] = {"kind": "reference", "cores": cores, "rows": [{k: d.get(k) for k in keep[:7]} for d in ref],
                               "sample": "same harness, same 64 blocks, 1.5 s per point"}
        try:
            out["ratio_1_thread"] = ours[0]["value"] / ref[0]["value"]
            out["ratio_all_cores"] = ours[-1]["value"] / ref[-1]["value"]
        except Exception:
            pass
    return out


# ------------------------------------------------------------------------------------------------ this repo's CUDA arm
def run_b200(args, rank, world, local_rank):
    import torch
    os.environ.setdefault("NRB200_DEVICE", str(local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        # NCCL printf()s its version banner to stdout when the first communicator is created; rank 0's stdout carries exactly one JSON line, so file
        # descriptor 1 points at stderr until that has happened (C stdio flushed before it is restored)
        import ctypes
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            ctypes.CDLL(None).fflush(None)
            os.dup2(saved, 1)
            os.close(saved)
    from openairinterface5g_b200.ldpc import load_LDPClib
    lib = load_LDPClib()

    if dist is not None:
        # constant tables are broadcast once at init (north star: "NCCL broadcast of the base-graph matrices only at init"): rank 0's lifted-graph
        # + packed-schedule blob for the benchmark's (BG, Z, R) goes out over NCCL and every rank checks the tables it built itself against it.
        # No collective sits on the data path.
        import ctypes
        size = lib.lib.nrb200_ldpc_graph_blob(BG, Z, R, None, 0)
        assert size > 0
        blob = np.zeros(size, dtype=np.uint8)
        lib.lib.nrb200_ldpc_graph_blob(BG, Z, R, blob.ctypes.data_as(ctypes.c_void_p), size)
        t = torch.from_numpy(blob).to(dev)
        t0 = t.clone()
        dist.broadcast(t0, src=0)
        assert torch.equal(t, t0), "graph tables differ from rank 0"

    B, NB = args.batch, args.nbuf
    gen = torch.Generator(device=dev)
    rate = 1.0 / 3.0
    sigma = 1.0 / np.sqrt(2.0 * (10.0 ** (args.ebn0 / 10.0)) * rate)
    batches = []
    for b in range(NB):
        gen.manual_seed(1000 * (rank + 1) + b)
        payload = torch.randint(0, 256, (B, K // 8), dtype=torch.uint8, device=dev, generator=gen)
        cw = lib.encode_batch_torch(BG, Z, K, payload)                       # (B, 66Z) 0/1, encoded on the GPU
        y = (1.0 - 2.0 * cw.to(torch.float32)) + sigma * torch.randn((B, 66 * Z), device=dev, generator=gen)
        q = torch.clamp(torch.floor(y / (sigma / 16.0)), -128, 127).to(torch.int8)
        llr = torch.zeros((B, NUM_LLR), dtype=torch.int8, device=dev)
        llr[:, 2 * Z:] = q
        batches.append((payload, llr))
    out = torch.empty((B, NUM_LLR // 8), dtype=torch.uint8, device=dev)
    iters = torch.empty(B, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()

    def step(i):
        lib.decode_batch_torch(BG, Z, R, MAX_ITER, batches[i % NB][1], out=out, iters=iters)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    barrier()
    launches0 = lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        barrier()
        e0.record()
        for i in range(args.steps):
            step(i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        launches = lib.launch_count() - launches0          # kernels of this library launched inside the timed region
        if ms < 600.0:   # untimed: keep the same load up long enough for the 200 ms clock sampler to see it
            t_end = time.perf_counter() + 0.8
            j = 0
            while time.perf_counter() < t_end:
                step(j); j += 1
                if j % 8 == 0:
                    torch.cuda.synchronize()
            torch.cuda.synchronize()
    mean_iters = float(iters.float().mean().item())
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * B * args.steps / (ms_max / 1000.0)

    # ---- end to end through the public host API: pinned host LLRs in, hard bits + iteration counts out, copies inside the timed region
    h_llr = [torch.empty((B, NUM_LLR), dtype=torch.int8).pin_memory() for _ in range(min(NB, 3))]
    for j, h in enumerate(h_llr):
        h.copy_(batches[j][1])
    DEPTH = 3   # batches in flight through submit / wait (the enqueue / dequeue halves of decode_batch_host)
    h_outs = [torch.empty((B, NUM_LLR // 8), dtype=torch.uint8).pin_memory() for _ in range(DEPTH)]
    np_outs = [h.numpy() for h in h_outs]
    h_its = [np.zeros(B, dtype=np.int32) for _ in range(DEPTH)]
    np_llr = [h.numpy() for h in h_llr]
    np_out, h_it = np_outs[0], h_its[0]

    def e2e_blocking(n):
        for i in range(n):
            lib.decode_batch_host(BG, Z, R, MAX_ITER, np_llr[i % len(np_llr)], out=np_out, iters=h_it)

    def e2e_pipelined(n):
        # step i: submit (stages nothing: the buffers are pinned; enqueues H2D of 26.7 MB, the kernels, D2H of bits + iteration counts),
        # then wait for step i - DEPTH + 1.  Every step's copies and its result read are inside the timed region.
        tickets = []
        for i in range(n):
            tickets.append(lib.decode_batch_host_submit(BG, Z, R, MAX_ITER, np_llr[i % len(np_llr)], np_outs[i % DEPTH], h_its[i % DEPTH]))
            if len(tickets) == DEPTH:
                lib.decode_batch_host_wait(tickets.pop(0))
        while tickets:
            lib.decode_batch_host_wait(tickets.pop(0))

    def timed_e2e(fn):
        fn(max(6, min(args.warmup, 10)))   # also lets every pooled workspace grow its staging buffers before the clock starts
        barrier()
        t0 = time.perf_counter()
        fn(args.steps)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        te = torch.tensor([dt], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        return world * B * args.steps / float(te.item())

    e2e_blocking_value = timed_e2e(e2e_blocking)
    e2e_value = timed_e2e(e2e_pipelined)
    # what the last pipelined step returned (checked against the oracle below), taken before anything else touches the host buffers
    e2e_last_iters = h_its[(args.steps - 1) % DEPTH].copy()
    e2e_last_out = np_outs[(args.steps - 1) % DEPTH][:3].copy()
    e2e_last_llr = np_llr[(args.steps - 1) % len(np_llr)][:3]

    # ---- platform ceiling for that traffic: the same bytes per step (H2D of the LLRs, D2H of the hard bits) moved by plain cudaMemcpyAsync on all
    # ranks at once, no kernel at all.  e2e / ceiling = how much of what this host's PCIe / memory system delivers to N GPUs the decode path uses.
    d_llr_c = torch.empty((B, NUM_LLR), dtype=torch.int8, device=dev)
    d_out_c = torch.empty((B, NUM_LLR // 8), dtype=torch.uint8, device=dev)
    h_out_c = [torch.empty((B, NUM_LLR // 8), dtype=torch.uint8).pin_memory() for _ in range(2)]
    cs = [torch.cuda.Stream(device=dev) for _ in range(2)]

    def copies_only(n):
        for i in range(n):
            with torch.cuda.stream(cs[i & 1]):
                d_llr_c.copy_(h_llr[i % len(h_llr)], non_blocking=True)
                h_out_c[i & 1].copy_(d_out_c, non_blocking=True)
        for s_ in cs:
            s_.synchronize()
    copy_ceiling = timed_e2e(copies_only)

    # ---- sanity: the timed kernel really decodes (parity of a few blocks of the last batch against the CPU oracle)
    check = e2e_check = None
    if rank == 0 and not args.no_check:
        from oracle.bindings import Oracle
        orc = Oracle()
        last = batches[(args.steps - 1) % NB][1][:3].cpu().numpy()
        lib.decode_batch_torch(BG, Z, R, MAX_ITER, batches[(args.steps - 1) % NB][1], out=out, iters=iters)
        torch.cuda.synchronize()
        o, it = out[:3].cpu().numpy(), iters[:3].cpu().numpy()
        check = all(orc.decode(BG, Z, R, MAX_ITER, last[i])[0] == it[i] and np.array_equal(orc.decode(BG, Z, R, MAX_ITER, last[i])[1], o[i]) for i in range(3))
        e2e_check = all(orc.decode(BG, Z, R, MAX_ITER, e2e_last_llr[i])[0] == e2e_last_iters[i]
                        and np.array_equal(orc.decode(BG, Z, R, MAX_ITER, e2e_last_llr[i])[1], e2e_last_out[i]) for i in range(3))

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        sample = 64
        kind, v, n, dt, mi = cpu_throughput(make_llr_numpy(sample, args.ebn0, 1), args.cpu_seconds, cores)
        cpu = {"value": v, "unit": "CB/s", "cores": cores, "kind": kind,
               "sample": f"{sample} distinct code blocks of the same workload decoded round-robin on {cores} host threads for {dt:.1f} s ({n} decodes, mean returned iterations {mi:.2f})"}

    if rank == 0:
        peak, peak_src = _peaks()
        kernel_s = (ms_max / 1000.0) / args.steps
        achieved = B * ALGO_BYTES_PER_CB / kernel_s / 1e9
        passes = mean_iters  # returned iteration count == CN/BN passes executed when no block stops early
        line = {
            "metric": METRIC, "value": value, "unit": "CB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int8",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_gpu": B, "numMaxIter": MAX_ITER, "ebn0_db": args.ebn0, "mean_returned_iters": mean_iters,
                       "l2": f"inputs rotate over {NB} distinct batches ({NB * B * NUM_LLR / 1e6:.0f} MB > 126 MB L2)", "parallelism": f"cb-shard x{world}"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic(B),
                         "peak_source": peak_src, "kernel": "ldpc_decode_packed_kernel", "kernel_ms": 1000.0 * kernel_s,
                         "note": "on-chip bound: message state lives in shared memory; see edge_updates_per_s",
                         "edge_updates_per_s": value / world * EDGE_UPDATES_PER_PASS * passes},
            "e2e": {"value": e2e_value, "unit": "CB/s", "h2d_bytes_per_step": B * NUM_LLR, "d2h_bytes_per_step": B * (NUM_LLR // 8) + 4 * B,
                    "api": f"nrb200_ldpc_decode_batch_host_submit / _wait on pinned host buffers, {DEPTH} batches in flight",
                    "blocking_call_value": e2e_blocking_value, "parity_check_vs_oracle": e2e_check,
                    "platform_copy_ceiling": {"value": copy_ceiling, "unit": "CB/s", "gbs": copy_ceiling * (NUM_LLR + NUM_LLR // 8) / 1e9,
                                              "what": "the same H2D + D2H bytes per step by plain cudaMemcpyAsync on all ranks at once, no kernel"},
                    "frac_of_platform_copy_ceiling": e2e_value / copy_ceiling},
            "gpu_launches": int(launches), "clocks": clk.summary(), "parity_check_vs_oracle": check,
        }
        oc = ncu_on_chip(B, kernel_s, line["clocks"].get("sm_mhz") or 1965.0)
        if oc is not None:
            ub = measured_mix_ceiling() if (world == 1 and not args.no_ubench) else None
            if ub and "decoder_mix_ipc_per_scheduler" in ub and oc.get("warp_inst_per_cb"):
                ipc = B * oc["warp_inst_per_cb"] / kernel_s / (int(ub.get("sm_count", 148)) * 4 * float(line["clocks"].get("sm_mhz") or 1965.0) * 1e6)
                oc["micro_benchmark"] = dict(ub, decoder_ipc_per_scheduler=ipc, frac_of_measured_mix_ceiling=ipc / ub["decoder_mix_ipc_per_scheduler"])
            line["roofline"]["on_chip"] = oc
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if world == 1 and not args.no_abi:
            try:
                line["e2e_plugin_abi"] = plugin_abi(not args.no_cpu)
            except Exception as e:
                line["e2e_plugin_abi"] = {"unavailable": repr(e)[:300]}
        if world == 1 and not args.no_slot:
            try:
                line["nr_dlsim_slot"] = slot_b200(lib, dev, 0.0 if args.no_cpu else args.slot_cpu_seconds)
            except Exception as e:                                # the second metric must never take the headline line down with it
                line["nr_dlsim_slot"] = {"metric": SLOT_METRIC, "unavailable": repr(e)[:300]}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--nbuf", type=int, default=6)
    ap.add_argument("--ebn0", type=float, default=1.0)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-ubench", action="store_true", help="skip tools/ubench/alu_ceiling (e.g. under ncu, so that the launch list holds the step's kernels only)")
    ap.add_argument("--no-slot", action="store_true", help="skip the nr_dlsim slot chain (second half of the BASELINE metric)")
    ap.add_argument("--slot-cpu-seconds", type=float, default=6.0)
    ap.add_argument("--no-check", action="store_true")
    ap.add_argument("--no-abi", action="store_true", help="skip the per-call loader-ABI measurement (tools/abi_bench)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
    else:
        run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
