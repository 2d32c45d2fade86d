#!/usr/bin/env python3
"""profiles/ncu_decode_traffic.json from the raw page of an `ncu --set full` capture of the decode kernel
(`ncu -i X.ncu-rep --page raw --csv > raw.csv`), stamped with the SHA-1 of the kernel's sources: bench.py only reports
roofline.traffic / on_chip from it when that stamp matches the sources it runs (tools/gpu_round.sh regenerates it each round).
Usage: python tools/ncu_traffic_json.py raw.csv out.json [n_cb=1024] [source note]"""
import csv, hashlib, json, os, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KERNEL_SOURCES = ["ldpc_decoder_packed.cuh", "ldpc_packed_simd.cuh", "ldpc_packed_graph.cc", "ldpc_packed_graph.h", "ldpc_common.cuh", "ldpc_decoder.cu"]


def kernel_source_sha1():
    h = hashlib.sha1()
    for f in KERNEL_SOURCES:
        h.update(open(os.path.join(ROOT, "openairinterface5g_b200", "csrc", f), "rb").read())
    return h.hexdigest()


def main():
    raw = list(csv.reader(open(sys.argv[1])))
    n_cb = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
    d = dict(zip(raw[0], raw[2]))
    unit = dict(zip(raw[0], raw[1]))                 # the raw page prints byte counts in scaled units ("Mbyte")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    f = lambda k: float(d[k].replace(",", "")) * scale.get(unit.get(k, ""), 1.0)
    cyc = f("sm__cycles_active.avg") if "sm__cycles_active.avg" in d else f("sm__cycles_elapsed.max")
    alu = f("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active")
    out = {"kernel": d.get("Kernel Name", "ldpc_decode_packed_kernel"), "launch": f"{n_cb} code blocks, BG1 Z=384 R13, 8 iterations (bench.py under ncu --set full)",
           "dram_bytes_read": int(f("dram__bytes_read.sum")), "dram_bytes_write": int(f("dram__bytes_write.sum")),
           "alu_pipe_pct": alu, "issue_active_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
           "fmaheavy_pipe_pct": f("sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active") if "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active" in d else None,
           "lsu_pipe_pct": f("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
           "sm_cycles_active": cyc, "warp_inst_per_launch": int(f("smsp__inst_executed.sum")),
           "alu_pipe_warp_inst_per_cb": alu / 100.0 * cyc * 148 * 4 * 0.5 / n_cb,
           "gpu_time_us": f("gpu__time_duration.sum") * {"us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}.get(unit.get("gpu__time_duration.sum", "us"), 1.0),
           "kernel_source_sha1": kernel_source_sha1(),
           "source": sys.argv[4] if len(sys.argv) > 4 else os.path.basename(sys.argv[1]),
           "note": "alu_pipe_warp_inst_per_cb = sm__inst_executed_pipe_alu (% of peak, active cycles) x active cycles x 148 SMs x 4 schedulers x 0.5 inst/cycle / blocks"}
    json.dump(out, open(sys.argv[2], "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
