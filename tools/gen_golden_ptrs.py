#!/usr/bin/env python3
"""Generate tests/golden/ptrs.npz from the UNMODIFIED reference (oracle/_ref/libref_pdsch_ptrs.so: nr_rx_pdsch + nr_pdsch_ptrs_processing + ptrs_nr.c): a PDSCH
slot with PT-RS through the UE receiver.  Seeded inputs (stored), outputs ONLY from the reference.  Run where /root/reference exists;
tests/test_golden_oracle.py pins the oracle to these vectors where it does not."""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle.bindings import Oracle, Reference, PuschParms, PtrsParms  # noqa: E402
from common import PTRS_CASES, ptrs_inputs  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "ptrs.npz")


def main():
    ref, orc = Reference(), Oracle()          # the oracle only supplies the Gold words / symbol mask the INPUT generator plants the pilots with
    rng = np.random.default_rng(2032)
    g = {}
    i = 0
    for case in (PTRS_CASES[5], PTRS_CASES[6]):
        N, nb_rx, rb_start, rb_size, Qm, dpos, dtype_, cdm, carrier, start, nsym, L, K, reoff, rnti, slot, nscid, nid = case
        for kind, a, b in (("random", 2000, 1500), ("coherent", 30, 0.05), ("coherent", 0, 0.0)):
            rx, h = ptrs_inputs(orc, rng, case, kind, a, b)
            P = PuschParms(N, nb_rx, rb_start, 0, rb_size, N - carrier * 6, Qm, dpos, dtype_, cdm)
            T = PtrsParms(1, L, K, reoff, rnti, slot, nscid, nid)
            per = [0] * 14
            mask = orc.ptrs_symbols(start, nsym, L, dpos)
            krb = rnti % K if rb_size % K == 0 else rnti % (rb_size % K)
            n_ptrs = len([re for re in range(12 * rb_size) if (re - reoff - krb * 12) % (12 * K) == 0])
            for s in range(start, start + nsym):
                per[s] = (rb_size * ((12 - 6 * cdm) if dtype_ == 0 else (12 - 4 * cdm)) if (dpos >> s) & 1 else rb_size * 12) - (n_ptrs if (mask >> s) & 1 else 0)
            G = sum(per) * Qm
            llr, sh, valid, ph, nre = ref.pdsch_rx_slot_ptrs(P, T, start, nsym, rx, h, G, n_rb_dl=carrier)
            assert [int(v) for v in valid] == per, (valid, per)
            g[f"case{i}"] = np.array(case, np.int32)
            g[f"rx{i}"], g[f"h{i}"], g[f"llr{i}"], g[f"sh{i}"], g[f"phase{i}"], g[f"nre{i}"] = rx, h, llr, np.int32(sh), ph, nre
            i += 1
    g["n"] = np.int32(i)
    # gNB side: nr_generate_pdsch with pduBitmap & 1 (oracle/_ref/libref_pdschtx.so)
    from oracle.bindings import PdschTxParms
    tx_cases = [(512, 25, 1, 9, 3, 11, 8, 1, 1, 6, 1 << 1, 0, 1, 0, 0, 512, 1, 2, 3, 0), (512, 25, 2, 11, 0, 25, 6, 2, 0, 14, (1 << 2) | (1 << 3), 0, 2, 0b0101, 0, 2047, 2, 2, 0, 0),
                (512, 25, 4, 4, 2, 21, 4, 2, 1, 13, 1 << 2, 1, 2, 0b000011, 1, 700, 0, 4, 7, 2)]
    for j, c in enumerate(tx_cases):
        N, carrier, ntx, slot, rb0, nrb, Qm, nl, s0, ns, dpos, dtype_, cdm, ports, scid, amp, L, K, reoff, pm = c
        P = PdschTxParms(N, ntx, slot, rb0, 0, nrb, N - carrier * 6, Qm, nl, s0, ns, dpos, dtype_, cdm, ports, scid, 40 + slot, 501, 0x1234 + slot, amp).set_ptrs(L, K, reoff)
        w = rng.integers(-12000, 12001, size=(4, 4, 2)).astype(np.int16)
        if pm:
            P.set_precoding(pm, w)
        bits = rng.integers(0, 2, size=P.G(), dtype=np.uint8)
        g[f"tx_case{j}"], g[f"tx_w{j}"], g[f"tx_bits{j}"], g[f"tx_out{j}"] = np.array(c, np.int32), w, bits, ref.pdsch_tx_slot(P, bits, carrier)
    g["n_tx"] = np.int32(len(tx_cases))
    np.savez_compressed(OUT, **g)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
