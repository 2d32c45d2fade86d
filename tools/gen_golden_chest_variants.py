#!/usr/bin/env python3
"""Generate tests/golden/chest_variants.npz from the UNMODIFIED reference nr_pusch_channel_estimation (oracle/_ref/libref_chest.so): DMRS type 2
with frequency-domain interpolation and the per-PRB averages (chest_freq = 1) of both DMRS types.  Seeded inputs, outputs ONLY from the reference.
Run where /root/reference exists; tests/test_golden_oracle.py pins the oracle to these vectors where it does not (the GPU box)."""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.bindings import Reference, ChestParms, PdschTxParms  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "chest_variants.npz")


def main():
    ref = Reference()
    rng = np.random.default_rng(2027)
    g = {}
    N, nrx, carrier = 512, 3, 25
    rx = rng.integers(-4000, 4001, size=(nrx, 14, N, 2)).astype(np.int16)
    g["rx"] = rx
    cases = []
    for i, (dmrs_type, chest_freq, slot, symbol, port, rb_start, rb_size) in enumerate(((1, 0, 7, 3, 0, 3, 20), (1, 0, 2, 11, 3, 0, 25), (0, 1, 5, 2, 2, 1, 22), (1, 1, 8, 4, 1, 0, 25))):
        par = [N, nrx, slot, symbol, port, rb_start, 0, rb_size, N - carrier * 6, i & 1, 300 + i, dmrs_type, chest_freq]
        est, out, pil = ref.pusch_channel_estimation(ChestParms(*par), rx, carrier, chest_freq=chest_freq, dmrs_type=dmrs_type)
        g[f"par{i}"], g[f"est{i}"], g[f"state{i}"] = np.array(par, np.int32), est[:, symbol], out
        g[f"pilots{i}"] = pil[:2 * (4 if dmrs_type else 6) * rb_size]
        cases.append(i)
    # UE estimator variants (nr_pdsch_channel_estimation): type 2 linear interpolation (ports 1 and 4) and both per-PRB averages
    for i, (dmrs_type, chest_freq, slot, symbol, port, rb_start, rb_size) in enumerate(((1, 0, 7, 3, 1, 3, 20), (1, 0, 2, 11, 4, 0, 25), (0, 1, 5, 2, 2, 1, 22), (1, 1, 9, 4, 5, 0, 25))):
        par = [N, nrx, slot, symbol, port, rb_start, 0, rb_size, N - carrier * 6, i & 1, 400 + i, dmrs_type, chest_freq]
        g[f"ue_par{i}"] = np.array(par, np.int32)
        g[f"ue_est{i}"] = ref.pdsch_channel_estimation(ChestParms(*par), rx, carrier, chest_freq=chest_freq, dmrs_type=dmrs_type)[:, symbol]
    # nr_chest_time_domain_avg: 2, 3 and 4 DMRS symbols on full-scale estimates
    est = rng.integers(-32768, 32768, size=(2, 14, N, 2)).astype(np.int16)
    est[:, :, ::5] //= 50
    g["tavg_in"] = est
    for i, (start, nsym, bitmap, nrb) in enumerate(((0, 14, 0b00100000000100, 25), (2, 12, 0b00101000001000, 20), (0, 14, 0b00100100100100, 22))):
        g[f"tavg_par{i}"] = np.array([nsym, start, bitmap, nrb], np.int32)
        out = ref.chest_time_domain_avg(est, nsym, start, bitmap, nrb)
        first = min(s for s in range(start, start + nsym) if (bitmap >> s) & 1)
        g[f"tavg_out{i}"] = out[:, first]
        assert np.array_equal(np.delete(out, first, axis=1), np.delete(est, first, axis=1))
    # gNB PDSCH transmitter (nr_generate_pdsch after the encoder): identity and wideband non-identity precoding, 2 layers on 4 antennas, allocation ending at the
    # symbol's last sub-carrier region so that both the saturating (SIMD) and the wrapping (scalar) accumulation occur
    for i, (pm, amp) in enumerate(((0, 512), (3, 30000))):
        par = [N, 4, 6, 2, 0, 21, N - carrier * 6, 6, 2, 1, 13, 1 << 2, 0, 2, 0b0011, 0, 46, 501, 0x1234, amp]
        P = PdschTxParms(*par)
        w = np.random.default_rng(77).integers(-32767, 32768, size=(4, 4, 2)).astype(np.int16)
        P.set_precoding(pm, w if pm else None)
        bits = np.random.default_rng(78 + i).integers(0, 2, size=P.G(), dtype=np.uint8)
        g[f"tx_par{i}"], g[f"tx_pm{i}"], g[f"tx_w{i}"], g[f"tx_bits{i}"] = np.array(par, np.int32), np.array([pm], np.int32), w, bits
        g[f"tx_out{i}"] = ref.pdsch_tx_slot(P, bits, carrier)
    g["n_cases"] = np.array([len(cases)], np.int32)
    np.savez_compressed(OUT, **g)
    print(OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
