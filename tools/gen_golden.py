#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the UNMODIFIED reference compiled by oracle/build_ref.sh (oracle/_ref/libref_*.so).
Run in the container where /root/reference exists; the fixtures are committed so that the GPU box (no reference tree)
can still check the oracle and the CUDA path against reference outputs.  Inputs are seeded; outputs come ONLY from the
reference libraries (AVX2 build; the AVX512 build as well for the one LUT where the two differ)."""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle.bindings import Reference, Oracle  # noqa: E402
from common import make_case, payloads, NCOLS  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main():
    os.makedirs(OUT, exist_ok=True)
    ref, ref512, orc = Reference(), Reference(avx512=True), Oracle()
    # ---- decoder: (BG, Z, R, n, EbN0, maxIter, outMode, use_crc)
    dec = {}
    cases = [(1, 384, 13, 3, 2.3, 8, 0, 0), (1, 384, 13, 2, 1.0, 8, 0, 0), (1, 384, 13, 2, 3.0, 8, 0, 0), (1, 384, 23, 2, 4.0, 8, 0, 0),
             (1, 384, 89, 2, 7.0, 8, 0, 0), (2, 384, 13, 2, 2.5, 8, 0, 0), (2, 384, 23, 2, 5.0, 8, 0, 0), (2, 384, 15, 2, 1.0, 8, 0, 0),
             (1, 96, 13, 2, 3.0, 5, 1, 0), (1, 52, 89, 2, 7.0, 8, 2, 0), (2, 64, 13, 2, 3.0, 20, 0, 0), (2, 20, 23, 2, 5.0, 8, 0, 0),
             (1, 3, 13, 2, 4.0, 8, 0, 0), (2, 11, 15, 2, 3.0, 8, 0, 0), (1, 30, 23, 2, 5.0, 2, 0, 0), (1, 208, 13, 2, 2.5, 8, 0, 0)]
    for ci, (BG, Z, R, n, e, mi, om, uc) in enumerate(cases):
        K, P, llr = make_case(orc, BG, Z, R, n, e, seed=100 + ci)
        its, outs, its5, outs5 = [], [], [], []
        for i in range(n):
            it, o = ref.decode(BG, Z, R, mi, llr[i], om)
            its.append(it); outs.append(o)
            it5, o5 = ref512.decode(BG, Z, R, mi, llr[i], om)
            its5.append(it5); outs5.append(o5)
        dec[f"c{ci}_par"] = np.array([BG, Z, R, n, mi, om, uc, K], dtype=np.int32)
        dec[f"c{ci}_llr"] = llr
        dec[f"c{ci}_iters_avx2"] = np.array(its, np.int32)
        dec[f"c{ci}_out_avx2"] = np.stack(outs)
        dec[f"c{ci}_iters_avx512"] = np.array(its5, np.int32)
        dec[f"c{ci}_out_avx512"] = np.stack(outs5)
    # CRC-stop mode (check_crc = reference crc_byte.c check_crc, CRC24_B appended like nr_segmentation does)
    BG, Z, R = 1, 128, 13
    K = 22 * Z
    rng = np.random.default_rng(77)
    P = rng.integers(0, 256, size=(4, K // 8), dtype=np.uint8)
    for i in range(4):
        crc = ref.crc(1, P[i], K - 24) >> 8
        P[i, -3:] = [(crc >> 16) & 0xFF, (crc >> 8) & 0xFF, crc & 0xFF]
    cw = ref.encode(BG, Z, K, P, orig=True)
    from openairinterface5g_b200.synth import awgn_llr
    llr = awgn_llr(cw, Z, 68, 2.2, 1 / 3, 77)
    for mi in (2, 3, 8):
        its, outs = [], []
        for i in range(4):
            it, o = ref.decode(BG, Z, R, mi, llr[i], 0, 1, K, 1)
            its.append(it); outs.append(o)
        dec[f"crc{mi}_iters"] = np.array(its, np.int32)
        dec[f"crc{mi}_out"] = np.stack(outs)
    dec["crc_par"] = np.array([BG, Z, R, 4, K], dtype=np.int32)
    dec["crc_llr"] = llr
    dec["ncases"] = np.array([len(cases)], np.int32)
    np.savez_compressed(os.path.join(OUT, "ldpc_decoder.npz"), **dec)

    # ---- encoder (ldpc_encoder.c "_orig" == optim8segmulti except the BG2 Z=64 defect; both stored)
    enc = {}
    ecases = [(1, 384), (1, 176), (1, 160), (1, 8), (2, 384), (2, 64), (2, 56), (2, 72), (1, 208), (2, 128), (1, 352), (2, 16)]
    for ci, (BG, Z) in enumerate(ecases):
        K, P = payloads(BG, Z, 9, 300 + ci)
        enc[f"e{ci}_par"] = np.array([BG, Z, K], np.int32)
        enc[f"e{ci}_in"] = P
        enc[f"e{ci}_orig"] = np.packbits(ref.encode(BG, Z, K, P, orig=True), axis=1)
        enc[f"e{ci}_optim"] = np.packbits(ref.encode(BG, Z, K, P), axis=1)
    enc["ncases"] = np.array([len(ecases)], np.int32)
    np.savez_compressed(os.path.join(OUT, "ldpc_encoder.npz"), **enc)

    # ---- CRC / rate matching / interleaving / segmentation
    cod = {}
    rng = np.random.default_rng(5)
    lens = [8, 24, 100, 1001, 3840, 8424, 8448]
    data = rng.integers(0, 256, size=(len(lens), 8448 // 8 + 8), dtype=np.uint8)
    cod["crc_lens"] = np.array(lens, np.int32)
    cod["crc_data"] = data
    cod["crc_vals"] = np.array([[ref.crc(p, data[i], n) for p in range(8)] for i, n in enumerate(lens)], dtype=np.uint64)
    rm = [(1, 384, 0, 9072, 0, 0, 1, 6), (1, 384, 88, 9072, 2, 0, 3, 2), (1, 96, 40, 30000, 1, 0, 2, 4), (2, 128, 16, 5000, 3, 0, 1, 8),
          (1, 384, 0, 20000, 0, 200000, 20, 2), (2, 52, 0, 1200, 0, 0, 1, 4), (1, 208, 120, 4104, 3, 90000, 4, 6), (2, 384, 200, 19008, 2, 0, 2, 8)]
    cod["rm_cases"] = np.array(rm, np.int32)
    for ci, (BG, Z, F, E, rv, Tb, Cs, Qm) in enumerate(rm):
        N = (66 if BG == 1 else 50) * Z
        K = (22 if BG == 1 else 10) * Z
        Fo = K - F - 2 * Z
        w = rng.integers(0, 2, size=N, dtype=np.uint8)
        w[Fo:Fo + F] = 2
        rc, e = ref.rate_matching_tx(Tb, BG, Z, w, Cs, F, Fo, rv, E)
        assert rc == 0
        f = ref.interleave(E, Qm, e)
        soft = rng.integers(-128, 128, size=E, dtype=np.int16)
        dei = ref.deinterleave(E, Qm, soft)
        wrx = np.zeros(N + 16, dtype=np.int16)
        ref.rate_matching_rx(Tb, BG, Z, wrx, dei, Cs, rv, 1, E, F, Fo)
        w1 = wrx.copy()
        ref.rate_matching_rx(Tb, BG, Z, wrx, dei, Cs, rv, 0, E, F, Fo)
        cod[f"rm{ci}_w"] = w; cod[f"rm{ci}_e"] = e; cod[f"rm{ci}_f"] = f; cod[f"rm{ci}_soft"] = soft
        cod[f"rm{ci}_dei"] = dei; cod[f"rm{ci}_w1"] = w1[:N]; cod[f"rm{ci}_w2"] = wrx[:N]
        cod[f"rm{ci}_getR"] = np.array([ref.get_R(rv, E, BG, Z, 0, 0)[0], ref.get_R(rv, E, BG, Z, 0, 0)[1]], np.int32)
    segs = [(1, 8448), (1, 8456), (1, 100000), (1, 424), (2, 3840), (2, 3848), (2, 600), (2, 200), (2, 100), (2, 40000)]
    cod["seg_cases"] = np.array(segs, np.int32)
    for ci, (BG, B) in enumerate(segs):
        d = rng.integers(0, 256, size=B // 8 + 8, dtype=np.uint8)
        Kb, Cc, K, Zc, F, s = ref.segmentation(d, B, BG)
        cod[f"seg{ci}_in"] = d
        cod[f"seg{ci}_par"] = np.array([Kb, Cc, K, Zc, F], np.int32)
        cod[f"seg{ci}_out"] = s
    np.savez_compressed(os.path.join(OUT, "coding.npz"), **cod)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
