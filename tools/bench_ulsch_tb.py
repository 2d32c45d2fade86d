#!/usr/bin/env python3
"""Time one PUSCH transport block through OAI's nr_ulsch_decoding as its caller sees it (oracle/ref_harness_ulsch.c: call + collection loop):
the reference's own function with the compiled CPU decoder, segments run one after the other on the calling thread (thread pool "n"), against the
interposer (integration/oai_shim_ulsch_decoding.c -> nrb200_ulsch_decode_tb_host).  The reference spreads the C segment jobs over its pool threads, so its
best case on T threads is the single-thread time / min(C, T); both numbers are printed.  python tools/bench_ulsch_tb.py [seconds per arm]"""
import ctypes as C
import json
import os
import sys
import time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle.bindings import Oracle  # noqa: E402
from common import make_tb_llrs     # noqa: E402


_KEEP = []   # LLR buffers the interposer page-locks stay alive for the whole run (OAI's pusch_vars->llr lives as long as the process)


def lib(name, bind=None):
    L = C.CDLL(os.path.join(ROOT, "oracle", "_ref", name))
    L.refh_ulsch_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 5
    if bind:
        assert L.refh_ulsch_bind_ldpc(os.path.join(ROOT, "oracle", "_ref", bind).encode()) == 0
    return L


def many_ues(L, n_thr, n_calls, prm, llrs, G, Cn, K, A):
    """n_thr caller threads, each with its own gNB context and its own transport block (its own LLR buffer), n_calls back-to-back nr_ulsch_decoding calls timed
    inside the harness: what N PUSCH receptions decoded at the same time cost."""
    import threading
    L.refh_ulsch_decode_loop.restype = C.c_double
    L.refh_ulsch_decode_loop.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 4
    bufs = [(np.zeros(8, np.int32), np.zeros(Cn, np.int32), np.zeros(Cn * K // 8, np.uint8), np.zeros(A // 8 + 3, np.uint8)) for _ in range(n_thr)]
    secs = [0.0] * n_thr
    start = threading.Barrier(n_thr)

    def work(t):
        inf, it, c, tb = bufs[t]
        L.refh_ulsch_decode_loop(t, 2, prm.ctypes.data, llrs[t].ctypes.data, G, inf.ctypes.data, it.ctypes.data, c.ctypes.data, tb.ctypes.data)   # warm
        start.wait()
        secs[t] = L.refh_ulsch_decode_loop(t, n_calls, prm.ctypes.data, llrs[t].ctypes.data, G, inf.ctypes.data, it.ctypes.data, c.ctypes.data, tb.ctypes.data)
    ths = [threading.Thread(target=work, args=(t,)) for t in range(n_thr)]
    t0 = time.perf_counter()
    [t.start() for t in ths]
    [t.join() for t in ths]
    assert min(secs) > 0
    return n_thr * n_calls / max(secs), max(secs) / n_calls * 1e6


def main():
    secs = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
    orc = Oracle()
    ref, shim = lib("libref_ulsch.so", "libref_ldpc_dec.so"), lib("libshimtest_ulsch.so")
    cores = os.cpu_count() or 1
    for A, Qm, nl, rb, snr, what in ((235624, 6, 1, 273, 10.0, "100 MHz slot, 1 layer, 64QAM: 28 segments"), (471272, 6, 2, 273, 10.0, "100 MHz slot, 2 layers, 64QAM: 56 segments"),
                                     (33640, 6, 1, 52, 6.0, "52 PRB, 64QAM: 4 segments"), (3752, 2, 1, 20, 4.0, "20 PRB, QPSK: 1 segment (BG2)")):
        BG = 2 if A < 3824 else 1
        pay, llr, info = make_tb_llrs(orc, A, Qm, nl, rb, 0, seed=7, snr_db=snr, BG=BG)
        _KEEP.append(llr)
        Cn, K, Z, G = info["C"], info["K"], info["Z"], info["G"]
        ncb = (66 if BG == 1 else 50) * Z
        prm = np.array([273, rb, Qm, nl, A // 8, 0, BG, 0, 8, 1, 0, 2], np.int32)
        inf = np.zeros(8, np.int32); it = np.zeros(Cn, np.int32); c = np.zeros(Cn * K // 8, np.uint8); tb = np.zeros(A // 8 + 3, np.uint8)
        row = {"what": what, "A_bits": A, "segments": Cn, "G": G}
        for name, L in (("reference_1thread", ref), ("b200_interposer", shim)):
            for _ in range(3):
                assert L.refh_ulsch_decode(prm.ctypes.data, llr.ctypes.data, G, inf.ctypes.data, it.ctypes.data, c.ctypes.data, tb.ctypes.data, None) == Cn
            assert np.array_equal(tb[:A // 8], pay)
            n, t0 = 0, time.perf_counter()
            while time.perf_counter() - t0 < secs:
                L.refh_ulsch_decode(prm.ctypes.data, llr.ctypes.data, G, inf.ctypes.data, it.ctypes.data, c.ctypes.data, tb.ctypes.data, None); n += 1
            us = (time.perf_counter() - t0) / n * 1e6
            row[name + "_us_per_tb"] = round(us, 1)
            row[name + "_mean_iters"] = float(it.mean())
        row["reference_best_case_us_on_%d_threads" % cores] = round(row["reference_1thread_us_per_tb"] / min(Cn, cores), 1)
        row["speedup_vs_1thread"] = round(row["reference_1thread_us_per_tb"] / row["b200_interposer_us_per_tb"], 1)
        row["speedup_vs_best_case"] = round(row["reference_best_case_us_on_%d_threads" % cores] / row["b200_interposer_us_per_tb"], 1)
        print(json.dumps(row), flush=True)
        if Cn == 28:
            # the multi-UE picture: T receptions decoded at the same time, one caller thread each (the reference: one core each, segments one after the other)
            for T in (4, 16, 32) if len(sys.argv) < 3 else (int(sys.argv[2]),):
                llrs = [make_tb_llrs(orc, A, Qm, nl, rb, 0, seed=100 + t, snr_db=snr, BG=BG)[1] for t in range(T)]
                _KEEP.append(llrs)
                out = {"what": f"{T} UEs at once, " + what, "caller_threads": T, "host_cores": cores}
                for name, L, calls in (("reference", ref, 12), ("b200_interposer", shim, 200)):
                    tbs, us = many_ues(L, T, calls, prm, llrs, G, Cn, K, A)
                    out[name + "_tb_per_s"] = round(tbs, 1); out[name + "_us_per_tb_per_caller"] = round(us, 1)
                out["speedup"] = round(out["b200_interposer_tb_per_s"] / out["reference_tb_per_s"], 1)
                print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
