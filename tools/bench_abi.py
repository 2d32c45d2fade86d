#!/usr/bin/env python3
"""The four-symbol drop-in measured per call: tools/abi_bench.c (dlopen + one blocking LDPCdecoder call per code block from T host threads,
like ldpctest.c:329-340 / nr_ulsch_decoding.c:435-468) against libldpc_b200.so and -- same binary, same inputs, same threads -- against the
compiled reference decoder.  Every call's output bytes and iteration count are checked against the reference's (expected.bin).
Usage: python tools/bench_abi.py [seconds=2.0] [ebn0=1.0]   -> one JSON line per (library, thread count)"""
import json
import os
import subprocess
import sys
import tempfile
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
BIN = os.path.join(ROOT, "tools", "ubench", "_bin", "abi_bench")
OURS = os.path.join(ROOT, "openairinterface5g_b200", "libldpc_b200.so")
REF = os.path.join(ROOT, "oracle", "_ref", "libref_ldpc_dec.so")
BG, Z, R, K, MAX_ITER, N = 1, 384, 13, 8448, 8, 64


def make_inputs(tmp, ebn0, seed=1):
    """N ldpctest-style blocks (27008-byte rows) + what the reference decoder returns for each (iterations, first K/8 output bytes)."""
    from oracle.bindings import Oracle, Reference, have_reference
    from openairinterface5g_b200.synth import awgn_llr, random_payloads
    orc = Oracle()
    P = random_payloads(N, K, seed)
    cw = np.stack([orc.encode(BG, Z, K, P[i]) for i in range(N)])
    llr = awgn_llr(cw, Z, 68, ebn0, 1.0 / 3.0, seed)
    rows = np.zeros((N, 27008), np.int8)
    rows[:, :68 * Z] = llr
    rows.tofile(os.path.join(tmp, "llr.bin"))
    dec = Reference().decode if have_reference() else orc.decode
    exp = np.zeros((N, 4 + K // 8), np.uint8)
    for i in range(N):
        it, out = dec(BG, Z, R, MAX_ITER, llr[i])
        exp[i, :4] = np.frombuffer(np.int32(it).tobytes(), np.uint8)
        exp[i, 4:] = np.asarray(out).view(np.uint8)[:K // 8]
    exp.tofile(os.path.join(tmp, "expected.bin"))
    return os.path.join(tmp, "llr.bin"), os.path.join(tmp, "expected.bin")


def run(lib, llr, exp, threads, seconds, env=None):
    r = subprocess.run([BIN, lib, llr, str(N), str(threads), str(seconds), str(BG), str(Z), str(R), str(MAX_ITER), exp], capture_output=True, text=True,
                       timeout=120 + 4 * seconds, env=dict(os.environ, **(env or {})))
    line = [l for l in r.stdout.splitlines() if l.startswith("{")]
    if not line:
        return {"lib": os.path.basename(lib), "host_threads": threads, "error": (r.stderr or r.stdout)[-300:], "rc": r.returncode}
    d = json.loads(line[-1])
    d["rc"] = r.returncode
    return d


def sweep(seconds=2.0, ebn0=1.0, ours_threads=(1, 2, 4, 8, 16, 32), ref_threads=None, with_reference=True):
    """Returns (ours, reference): lists of abi_bench result dicts."""
    cores = os.cpu_count() or 8
    if ref_threads is None:
        ref_threads = sorted({1, cores})
    with tempfile.TemporaryDirectory() as tmp:
        llr, exp = make_inputs(tmp, ebn0)
        ours = [run(OURS, llr, exp, t, seconds) for t in ours_threads]
        ref = [run(REF, llr, exp, t, seconds) for t in ref_threads] if with_reference and os.path.exists(REF) else []
    return ours, ref


def encoder_sweep(seconds=1.0):
    """LDPCencoder per call (1 and 8 segments, 1 and 8 host threads): this library and the reference's default module (ldpc_encoder_optim8segmulti.c)."""
    ref_enc = os.path.join(ROOT, "oracle", "_ref", "libref_ldpc_enc.so")
    rows = []
    for lib in (OURS, ref_enc):
        if not os.path.exists(lib):
            continue
        for thr, nseg in ((1, 1), (1, 8), (8, 8)):
            r = subprocess.run([BIN, "enc", lib, str(thr), str(seconds), str(BG), str(Z), str(nseg)], capture_output=True, text=True, timeout=120)
            line = [l for l in r.stdout.splitlines() if l.startswith("{")]
            rows.append(json.loads(line[-1]) if line else {"lib": os.path.basename(lib), "error": (r.stderr or r.stdout)[-200:]})
    return rows


if __name__ == "__main__":
    secs = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
    ebn0 = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    ours, ref = sweep(secs, ebn0)
    for d in ours + ref:
        d["ebn0_db"] = ebn0
        print(json.dumps(d), flush=True)
    for d in encoder_sweep(min(secs, 1.0)):
        print(json.dumps(d), flush=True)
