#!/usr/bin/env python3
"""Generate tests/golden/phy.npz from the UNMODIFIED reference compiled by oracle/build_ref.sh: scrambling / QAM mapper, the slot-level OFDM front
end, PUSCH channel estimation, the single-layer inner receiver and the two-layer MMSE receiver.  Small configurations (the fixture stays < 1 MB);
inputs are seeded, outputs come ONLY from the reference libraries.  Run where /root/reference exists; tests/test_golden_oracle.py pins the oracle to
these vectors on machines without the reference tree (the GPU box)."""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.bindings import Reference, ChestParms, PuschParms  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "phy.npz")


def main():
    ref = Reference()
    rng = np.random.default_rng(2026)
    g = {}
    # ---- scrambling, descrambling, mapper
    bits = rng.integers(0, 2, size=3001, dtype=np.uint8)
    g["scr_bits"] = bits
    g["scr_par"] = np.array([1, 1007, 65535], np.int32)                  # q, Nid, rnti
    sc = ref.scramble(bits, 1, 1007, 65535)
    g["scr_out"] = sc
    for Qm in (2, 4, 6, 8):
        g[f"mod{Qm}"] = ref.modulate(sc, (3001 // Qm) * Qm, Qm)
    llr = rng.integers(-32768, 32768, size=2999).astype(np.int16)
    g["unscr_in"] = llr
    g["unscr_out"] = ref.unscramble_llr(llr, 1, 1007, 65535)
    # ---- OFDM front end: N = 256, mu = 1, 11 PRB (odd), slot 2 -- rotation tables, TX slot, RX slot with a timing offset
    N, mu, nb_rb, slot = 256, 1, 11, 2
    dl, ul, ts = ref.rotation_tables(N, mu, nb_rb, 8, 3619200000.0, 3609200000.0)
    g["ofdm_par"] = np.array([N, mu, nb_rb, slot, 8, 37], np.int32)       # ..., divisor, sample_offset
    g["ofdm_rot_dl"], g["ofdm_rot_ul"], g["ofdm_timeshift"] = dl[:2 * (14 << mu)], ul[:2 * (14 << mu)], ts
    F = np.zeros((14, N, 2), np.int16)
    F[:, :nb_rb * 6] = rng.integers(-6000, 6001, size=(14, nb_rb * 6, 2)); F[:, N - nb_rb * 6:] = rng.integers(-6000, 6001, size=(14, nb_rb * 6, 2))
    g["ofdm_txF"] = F
    p, p0 = N // 128 * 9, N // 128 * (9 + (1 << mu))
    out_len = 14 * N + 14 * p + (p0 - p if (slot * 14) % (7 << mu) == 0 else 0)
    y, Frot = ref.ofdm_tx_slot(N, mu, nb_rb, slot, 14, dl, F, out_len)
    g["ofdm_tx_out"], g["ofdm_tx_rotated"] = y, Frot
    frame_len = 10 * ((p0 + N) * 2 + (p + N) * (14 * (1 << mu) - 2))
    rx = rng.integers(-5000, 5001, size=2 * frame_len).astype(np.int16)
    g["ofdm_rx_in"] = rx
    g["ofdm_rx_out"] = ref.ofdm_rx_slot(N, mu, nb_rb, slot, 8, 37, ul, rx)
    # ---- PUSCH channel estimation: N = 512, 2 rx, port 1, 20 PRB from PRB 3, slot 7 symbol 3
    N, nrx = 512, 2
    P = ChestParms(N, nrx, 7, 3, 1, 3, 0, 20, N - 25 * 6, 1, 321)
    rxF = rng.integers(-1200, 1201, size=(nrx, 14, N, 2)).astype(np.int16)
    est, out, pil = ref.pusch_channel_estimation(P, rxF, 25)
    g["chest_par"] = np.array([N, nrx, 7, 3, 1, 3, 0, 20, N - 25 * 6, 1, 321], np.int32)
    g["chest_rx"], g["chest_est"], g["chest_state"], g["chest_pilots"] = rxF, est[:, 3], out, pil
    # ---- inner receiver, one layer: N = 512, 2 rx, 64QAM, type-1 DMRS at symbol 2 with data (1 CDM group), symbols 2 and 5
    PP = PuschParms(N, nrx, 3, 0, 20, N - 25 * 6, 6, 1 << 2, 0, 1)
    h = rng.integers(-1500, 1501, size=(nrx, 14, N, 2)).astype(np.int16)
    g["rx1_par"] = np.array([N, nrx, 3, 0, 20, N - 25 * 6, 6, 1 << 2, 0, 1], np.int32)
    g["rx1_rx"], g["rx1_h"] = rxF, h
    sh, avg = ref.pusch_log2_maxh(PP, 2, 2, rxF, h)
    g["rx1_shift"], g["rx1_avg"] = np.array([sh], np.int32), avg
    for s in (2, 5):
        valid = 20 * (6 if s == 2 else 12)
        l, c = ref.pusch_inner_rx_symbol(PP, s, 2, sh, rxF, h, valid)
        g[f"rx1_llr{s}"], g[f"rx1_comp{s}"] = l, c
    # ---- two layers, MMSE: 256QAM, nvar 55, shift 7, symbol 4
    PP2 = PuschParms(N, nrx, 3, 0, 20, N - 25 * 6, 8, 1 << 2, 0, 2)
    h2 = rng.integers(-1500, 1501, size=(2 * nrx, 14, N, 2)).astype(np.int16)
    g["rx2_par"] = np.array([N, nrx, 3, 0, 20, N - 25 * 6, 8, 1 << 2, 0, 2, 55, 7, 9000], np.int32)   # ..., nvar, shift, max_ch
    g["rx2_h"] = h2
    l, c = ref.pusch_inner_rx_symbol(PP2, 4, 2, 7, rxF, h2, 240, nb_layer=2, nvar=55)
    g["rx2_llr"], g["rx2_comp"] = l, c
    sh2, avg2 = ref.pusch_log2_maxh(PP2, 0, 2, rxF, h2, nb_layer=2, max_ch=9000)
    g["rx2_shift"], g["rx2_avg"] = np.array([sh2], np.int32), avg2
    # ---- UE side: PDSCH channel estimation (port 2, same geometry) and the single-layer nr_rx_pdsch receiver (16QAM, DMRS with data at symbol 2)
    PU = ChestParms(N, nrx, 7, 3, 2, 3, 0, 20, N - 25 * 6, 1, 321)
    g["uechest_par"] = np.array([N, nrx, 7, 3, 2, 3, 0, 20, N - 25 * 6, 1, 321], np.int32)
    g["uechest_est"] = ref.pdsch_channel_estimation(PU, rxF, 25)[:, 3]
    PD = PuschParms(N, nrx, 3, 0, 20, N - 25 * 6, 4, 1 << 2, 0, 1)
    g["pdsch_par"] = np.array([N, nrx, 3, 0, 20, N - 25 * 6, 4, 1 << 2, 0, 1, 1, 13], np.int32)      # ..., start_symbol, nr_symbols
    G_ = (12 * 12 + 6) * 20 * 4
    l, sh, _ = ref.pdsch_rx_slot(PD, 1, 13, rxF, h, G_)
    g["pdsch_llr"], g["pdsch_shift"] = l, np.array([sh], np.int32)
    np.savez_compressed(OUT, **g)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
