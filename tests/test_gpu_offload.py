"""The "offload" calling convention (libldpc_b200_t2.so): one segment per call, the library de-interleaves, rate-recovers, HARQ-combines and
decodes / encodes, rate-matches and interleaves.  Checked bit for bit against the composition of the oracle functions that are pinned to the
reference's CPU path (nr_deinterleaving_ldpc, nr_rate_matching_ldpc_rx, the decoder-input packing, LDPCdecoder; LDPCencoder,
nr_rate_matching_ldpc, nr_interleaving_ldpc)."""
import numpy as np
import pytest

from openairinterface5g_b200.ldpc import OffloadLdpcLib
from openairinterface5g_b200.synth import random_payloads

pytestmark = pytest.mark.gpu

CASES = [  # BG, Z, F, Qm, E
    (1, 384, 8, 6, 9126), (1, 384, 0, 2, 12000), (1, 208, 40, 4, 6000), (2, 192, 16, 2, 4000), (2, 64, 0, 8, 1600), (1, 384, 424, 4, 16224),
]


def _pack(w, K, F, Z, kc):
    z = np.zeros(kc * Z, np.int16)
    z[2 * Z:K - F] = w[:K - F - 2 * Z]
    z[K - F:K] = 127
    z[K:] = w[K - 2 * Z:(kc - 2) * Z]
    return np.clip(z, -128, 127).astype(np.int8)


def test_offload_encoder_decoder_vs_oracle(ldpc, oracle):
    off = OffloadLdpcLib()
    rng = np.random.default_rng(70)
    for ci, (BG, Z, F, Qm, E) in enumerate(CASES):
        K = (22 if BG == 1 else 10) * Z
        kc = 68 if BG == 1 else 52
        seg = random_payloads(1, K, 70 + ci)[0]
        bits = np.unpackbits(seg)
        bits[K - F:] = 0
        seg = np.packbits(bits)
        # ---- encoder: E rate-matched + interleaved bits of rv 0 and rv 2
        d = oracle.encode(BG, Z, K, seg).copy()
        d[K - F - 2 * Z:K - 2 * Z] = 2
        tx = {}
        for rv in (0, 2):
            rc, e = oracle.rate_matching_tx(0, BG, Z, d, 1, F, K - F - 2 * Z, rv, E)
            assert rc == 0
            f_o = oracle.interleave(E, Qm, e)
            f = off.LDPCencoder(BG, Z, K, F, Qm, rv, E, seg)
            assert np.array_equal(f, f_o), (BG, Z, F, Qm, E, rv)
            tx[rv] = f
        # ---- decoder: first transmission (rv 0, new data), then a retransmission (rv 2) combined with the stored soft buffer
        w = np.zeros((kc - 2) * Z, np.int16)
        for rnd, rv in enumerate((0, 2)):
            noise = rng.normal(0, 6.0, size=E)
            llr = np.clip(np.round((1 - 2 * tx[rv].astype(np.float64)) * 24 + noise), -128, 127).astype(np.int8)
            R = oracle.get_R(rv, E, BG, Z, 0, 0)[0]
            e16 = oracle.deinterleave(E, Qm, llr.astype(np.int16))
            assert oracle.rate_matching_rx(0, BG, Z, w, e16, 1, rv, 1 if rnd == 0 else 0, E, F, K - F - 2 * Z) == 0
            it_o, out_o = oracle.decode(BG, Z, R, 8, _pack(w, K, F, Z, kc))
            it, out = off.LDPCdecoder(BG, Z, R, 8, E, Qm, rv, F, llr, ulsch_id=3, r=ci, setCombIn=0 if rnd == 0 else 1)
            assert it == it_o, (BG, Z, F, Qm, E, rv, it, it_o)
            assert np.array_equal(out, np.asarray(out_o, dtype=np.uint8)[:K // 8]), (BG, Z, F, Qm, E, rv)
        # the combined second round decodes the segment
        assert it <= 8 and np.array_equal(out[:(K - F) // 8], seg[:(K - F) // 8])
