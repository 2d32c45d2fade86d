"""Shared helpers for the test-suite: ldpctest-style case generation (reference ldpctest.c:269-357)."""
import numpy as np
from openairinterface5g_b200.synth import awgn_llr

ALL_Z = [2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 18, 20, 22, 24, 26, 28, 30, 32, 36, 40, 44, 48, 52, 56, 60, 64, 72, 80, 88,
         96, 104, 112, 120, 128, 144, 160, 176, 192, 208, 224, 240, 256, 288, 320, 352, 384]
RATES = {1: (13, 23, 89), 2: (15, 13, 23)}
NCOLS = {(1, 13): 68, (1, 23): 35, (1, 89): 27, (2, 15): 52, (2, 13): 32, (2, 23): 17}


def payloads(BG, Z, n, seed):
    K = (22 if BG == 1 else 10) * Z
    rng = np.random.default_rng(seed)
    P = rng.integers(0, 256, size=(n, (K + 7) // 8), dtype=np.uint8)
    if K % 8:
        P[:, -1] &= (0xFF << (8 - K % 8)) & 0xFF
    return K, P


def decode_all_reference(oracle, reference, llr, BG, Z, R, max_iter, crc=None, threads=None):
    """Every row of llr through the UNMODIFIED reference LDPCdecoder (oracle/_ref/libref_ldpc_dec.so), one blocking call per block on all host
    cores (oracle/cpu_bench.c:orc_decode_all).  crc = (crc_type, crc_len_bits) selects the check_crc stop (nrLDPC_decoder.c:850-862), None the
    parity-check stop of ldpctest.  Returns (iters[n], out[n, ncols*Z/8])."""
    import ctypes as C
    import os
    llr = np.ascontiguousarray(llr, dtype=np.int8)
    n, stride = llr.shape
    ob = NCOLS[(BG, R)] * Z // 8
    out = np.zeros((n, ob), np.uint8)
    its = np.zeros(n, np.int32)
    f = oracle.lib.orc_decode_all
    f.restype = C.c_long
    f.argtypes = [C.c_void_p] * 3 + [C.c_int] * 8 + [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    fn = C.cast(reference.dec.LDPCdecoder, C.c_void_p)
    crc_fn = C.cast(reference.cod.check_crc, C.c_void_p) if crc else None
    f(fn, crc_fn, llr.ctypes.data, n, stride, BG, Z, R, max_iter, crc[0] if crc else 0, crc[1] if crc else 0, out.ctypes.data, ob, its.ctypes.data,
      threads or os.cpu_count() or 8)
    return its, out


def make_case(oracle, BG, Z, R, n, ebn0_db, seed):
    """payload -> oracle encoder -> BPSK/AWGN -> int8 LLRs.  Returns (K, payload bytes, llr[n, ncols*Z])."""
    K, P = payloads(BG, Z, n, seed)
    nc = NCOLS[(BG, R)]
    rate = (22 if BG == 1 else 10) / (nc - 2)
    cw = np.stack([oracle.encode(BG, Z, K, P[i]) for i in range(n)])
    return K, P, awgn_llr(cw, Z, nc, ebn0_db, rate, seed)


# ------------------------------------------------------------------------------------------------ oracle-only PUSCH slot chain
def oracle_pusch_transmit(oracle, P, A, Qm, rb_start, rb_size, nb_rx, slot, rnti, nid, rot, seed, tx_amp=724, h_amp=724.0, snr_db=30.0, dmrs_id=55):
    """CPU restatement chain (oracle functions only) that produces one slot of time-domain samples per rx antenna: TB CRC, segmentation,
    LDPC encoding, rate matching + interleaving, scrambling, QAM mapping, resource mapping (one type-1 DMRS symbol at l = 2 without data),
    flat channel + noise, OFDM modulation.  Returns (payload, frame [nb_rx, 2 * samples_per_frame], genie estimates [nb_rx, 14, N, 2], info)."""
    from openairinterface5g_b200 import transport as T
    rng = np.random.default_rng(seed)
    N = P.N
    payload = rng.integers(0, 256, size=A // 8, dtype=np.uint8)
    crc = oracle.crc(0, payload, A) >> 8
    tb = np.concatenate([payload, np.array([(crc >> 16) & 255, (crc >> 8) & 255, crc & 255], np.uint8)])
    _, C_, K, Z, F, segs = oracle.segmentation(tb, A + 24, 1)
    G = T.nr_get_G(rb_size, 14, 12, 1, 0, Qm, 1)
    E = [T.nr_get_E(G, C_, Qm, 1, r) for r in range(C_)]
    f = []
    for r in range(C_):
        d = oracle.encode(1, Z, K, segs[r]).copy()
        d[K - F - 2 * Z:K - 2 * Z] = 2                                     # NR_NULL filler marks (nr_dlsch_coding.c)
        rc, e = oracle.rate_matching_tx(0, 1, Z, d, C_, F, K - F - 2 * Z, 0, E[r])
        assert rc == 0
        f.append(oracle.interleave(E[r], Qm, e))
    f = np.concatenate(f)
    words = oracle.scramble(f, 0, nid, rnti)
    sym = oracle.modulate(words, G, Qm).reshape(-1, 2).astype(np.int64)
    x = (sym * tx_amp) >> 15
    start_re = (P.first_carrier_offset + rb_start * 12) % N
    sc = (start_re + np.arange(12 * rb_size)) % N
    data_syms = [s for s in range(14) if s != 2]
    ph = rng.uniform(0, 2 * np.pi, nb_rx)
    hi = np.round(np.stack([np.cos(ph), np.sin(ph)], axis=1) * h_amp).astype(np.int64)
    unit = 23170.0 * tx_amp / 32768.0
    est = np.zeros((nb_rx, 14, N, 2), np.int16)
    est[:, 2, :12 * rb_size, :] = np.round(hi * (unit / 1024.0)).astype(np.int16)[:, None, :]
    sigma = unit * (h_amp / 1024.0) * 10.0 ** (-snr_db / 20.0) * 0.70711
    from oracle.bindings import ChestParms
    CP = ChestParms(N, nb_rx, slot, 2, 0, rb_start, 0, rb_size, P.first_carrier_offset, 0, dmrs_id)
    pil = oracle.pusch_dmrs_pilots(CP).reshape(-1, 2).astype(np.float64) * (unit / 32767.0)
    dm_sc = (start_re + 2 * np.arange(6 * rb_size)) % N
    frame = np.zeros((nb_rx, 2 * P.samples_per_frame), np.int16)
    ss = P.slot_timestamp(slot)
    for a in range(nb_rx):
        yr = (hi[a, 0] * x[:, 0] - hi[a, 1] * x[:, 1]) / 1024.0
        yi = (hi[a, 0] * x[:, 1] + hi[a, 1] * x[:, 0]) / 1024.0
        y = np.stack([yr, yi], axis=1) + sigma * rng.standard_normal((x.shape[0], 2))
        grid = np.zeros((14, N, 2), np.int16)
        yq = np.clip(np.round(y), -32768, 32767).astype(np.int16).reshape(len(data_syms), 12 * rb_size, 2)
        for j, s in enumerate(data_syms):
            grid[s, sc] = yq[j]
        hc = (hi[a, 0] + 1j * hi[a, 1]) / 1024.0
        d = hc * (pil[:, 0] - 1j * pil[:, 1])                                # DMRS = conj of the receiver's table, data amplitude
        dn = np.stack([d.real, d.imag], axis=1) + sigma * rng.standard_normal((pil.shape[0], 2))
        grid[2, dm_sc] = np.clip(np.round(dn), -32768, 32767).astype(np.int16)
        t, _ = oracle.ofdm_tx_slot(N, P.mu, P.nb_rb, slot, 14, rot.reshape(-1), grid.reshape(-1))
        frame[a, 2 * ss:2 * ss + t.size] = t
    return payload, frame, est, dict(C=C_, K=K, Z=Z, F=F, E=E, G=G)


def oracle_pusch_receive(oracle, P, info, Qm, rb_start, rb_size, nb_rx, slot, rnti, nid, rot, frame, est=None, max_iter=8, dmrs_id=55, n_layers=1, cdm=2):
    """The receive chain with oracle functions: OFDM demod, channel estimation (every DMRS port), level, inner receiver (one layer: MRC; two layers: the MMSE
    receiver with the estimator's max_ch / nvar, nr_ulsch_demodulation.c:1470-1524, and layer de-mapping :1422-1428), descrambling, de-interleaving, rate
    recovery, decoder-input packing (nr_ulsch_decoding.c:195-210), decoding with CRC24B stop.  Returns (tb bytes, iterations, llr16, log2_maxh)."""
    from oracle.bindings import PuschParms
    from openairinterface5g_b200 import transport as T
    N = P.N
    rxF = np.stack([oracle.ofdm_rx_slot(N, P.mu, P.nb_rb, slot, P.divisor, 0, rot.reshape(-1), frame[a]).reshape(14, N, 2) for a in range(nb_rx)])
    max_ch, nvar = 0, 0
    if est is None:                                                          # estimate from the DMRS symbol (nr_pusch_channel_estimation)
        from oracle.bindings import ChestParms
        ests = []
        for p in range(n_layers):
            e, st = oracle.pusch_channel_estimation(ChestParms(N, nb_rx, slot, 2, p, rb_start, 0, rb_size, P.first_carrier_offset, 0, dmrs_id), rxF)
            ests.append(e); max_ch = max(max_ch, int(st[0])); nvar += int(st[1]) & 0xFFFFFFFF
        est = np.concatenate(ests)                                          # ul_ch_estimates[p * nb_rx + aarx]
        nvar //= 14 * n_layers * nb_rx
    PP = PuschParms(N, nb_rx, rb_start, 0, rb_size, P.first_carrier_offset, Qm, 1 << 2, 0, cdm)
    if n_layers == 1:
        shift, _ = oracle.pusch_log2_maxh(PP, 0, 2, rxF, est)
        llr = np.concatenate([oracle.pusch_inner_rx_symbol(PP, s, 2, shift, rxF, est)[0] for s in range(14) if s != 2])
    else:
        shift, _ = oracle.pusch_log2_maxh_2l(PP, 0, 2, max_ch, rxF, est)
        out = []
        for s in range(14):
            v = oracle.pusch_nb_re(PP, s)
            if v == 0:
                continue
            l2, _ = oracle.pusch_inner_rx_symbol_2l(PP, s, 2, shift, nvar, rxF, est)
            out.append(np.stack([l2[0].reshape(v, Qm), l2[1].reshape(v, Qm)], axis=1).reshape(-1))
        llr = np.concatenate(out)
    llr = oracle.unscramble_llr(llr, 0, nid, rnti)
    C_, K, Z, F, E = info["C"], info["K"], info["Z"], info["F"], info["E"]
    R = T.nr_get_R_ldpc_decoder(0, E[0], 1, Z)[0]
    off, out, its = 0, [], []
    for r in range(C_):
        e = oracle.deinterleave(E[r], Qm, llr[off:off + E[r]]); off += E[r]
        w = np.zeros(66 * Z, np.int16)
        assert oracle.rate_matching_rx(0, 1, Z, w, e, C_, 0, 1, E[r], F, K - F - 2 * Z) == 0
        z = np.zeros(68 * Z, np.int16)
        z[2 * Z:K - F] = w[:K - F - 2 * Z]
        z[K - F:K] = 127
        z[K:] = w[K - 2 * Z:]
        it, hard = oracle.decode(1, Z, R, max_iter, np.clip(z, -128, 127).astype(np.int8), use_crc=1, crc_len_bits=K - F, crc_type=1)
        its.append(it); out.append(np.asarray(hard, dtype=np.uint8)[:(K - F - 24) // 8])
    return np.concatenate(out), np.array(its), llr, shift


def make_tb_llrs(oracle, A, Qm, nl, rb_size, rv, seed, snr_db=8.0, BG=1, nsymb=13):
    """A transport block through the oracle's transmit chain (TB CRC, segmentation, LDPC encoding, rate matching for redundancy version rv, interleaving) and a
    BPSK-per-bit AWGN channel: what nr_ulsch_decoding receives from nr_rx_pusch_tp.  Returns (payload bytes, llr int16[G], dict(C, K, Z, F, E, G))."""
    from openairinterface5g_b200 import transport as T
    rng = np.random.default_rng(seed)
    payload = rng.integers(0, 256, size=A // 8, dtype=np.uint8)
    if A > 3824:
        crc = oracle.crc(0, payload, A) >> 8
        tb = np.concatenate([payload, np.array([(crc >> 16) & 255, (crc >> 8) & 255, crc & 255], np.uint8)])
        B = A + 24
    else:
        crc = oracle.crc(3, payload, A) >> 16
        tb = np.concatenate([payload, np.array([(crc >> 8) & 255, crc & 255], np.uint8)])
        B = A + 16
    _, C_, K, Z, F, segs = oracle.segmentation(tb, B, BG)
    G = T.nr_get_G(rb_size, nsymb, 0, 1, 0, Qm, nl)
    E = [T.nr_get_E(G, C_, Qm, nl, r) for r in range(C_)]
    f = []
    for r in range(C_):
        d = oracle.encode(BG, Z, K, segs[r]).copy()
        d[K - F - 2 * Z:K - 2 * Z] = 2
        rc, e = oracle.rate_matching_tx(0, BG, Z, d, C_, F, K - F - 2 * Z, rv, E[r])
        assert rc == 0
        f.append(oracle.interleave(E[r], Qm, e))
    bits = np.concatenate(f).astype(np.float64)
    sigma = 10 ** (-snr_db / 20.0)
    y = (1.0 - 2.0 * bits) + sigma * rng.standard_normal(bits.size)
    llr = np.clip(np.round(y * 24.0), -32768, 32767).astype(np.int16)
    return payload, llr, dict(C=C_, K=K, Z=Z, F=F, E=E, G=G)


PTRS_CASES = [  # N, nb_rx, rb_start, rb_size, Qm, dmrs_pos, dmrs_type, cdm, carrier, start, nsym, L (log2), K, re_offset, rnti, slot, nscid, nid
    (4096, 2, 0, 273, 6, 1 << 2, 0, 2, 273, 1, 13, 0, 2, 0, 0x1234, 3, 0, 40),
    (4096, 4, 0, 272, 8, 1 << 2, 0, 1, 273, 1, 13, 1, 4, 2, 0x4321, 7, 1, 500),
    (2048, 1, 10, 50, 4, (1 << 2) | (1 << 11), 0, 1, 106, 1, 13, 2, 2, 5, 0x0101, 0, 0, 1007),
    (2048, 2, 30, 76, 2, 1 << 3, 0, 2, 106, 2, 10, 1, 2, 11, 77, 19, 0, 0),
    (1024, 4, 0, 52, 6, 1 << 2, 1, 1, 52, 1, 13, 2, 4, 1, 65535, 5, 1, 3),
    (1024, 2, 20, 31, 4, 1 << 2, 1, 2, 52, 2, 12, 0, 4, 0, 9, 9, 0, 9),
    (512, 4, 3, 11, 8, 1 << 1, 0, 1, 25, 1, 6, 1, 2, 3, 1, 1, 1, 65535),
    (2048, 2, 0, 106, 6, (1 << 2) | (1 << 13), 0, 1, 106, 1, 13, 2, 2, 0, 4242, 2, 0, 123),
    (2048, 2, 0, 105, 6, (1 << 2) | (1 << 7) | (1 << 11), 0, 2, 106, 0, 14, 1, 4, 7, 31, 2, 0, 123),
]
PTRS_SIGNALS = [("random", 2000, 1500), ("random", 32767, 32767), ("coherent", 0, 0.0), ("coherent", 30, 0.05), ("coherent", 0, 0.3)]


def ptrs_inputs(oracle, rng, case, kind, a, b):
    """Inputs of a PDSCH slot with PT-RS for the UE receiver.  kind "random": rx / estimates uniform in +-a / +-b.  kind "coherent": a flat channel per antenna,
    QPSK on every RE, the PT-RS REs carrying the pilots nr_ptrs_cpe_estimation regenerates (Gold sequence of the symbol's PDSCH DMRS), a common phase error of
    b rad per symbol and Gaussian noise of sigma a -- so the estimates are a real phase ramp (b = 0, a = 0: zero imaginary part, the 32768 -> -32768 case)."""
    N, nb_rx, rb_start, rb_size, Qm, dpos, dtype_, cdm, carrier, start, nsym, L, K, reoff, rnti, slot, nscid, nid = case
    if kind == "random":
        return rng.integers(-a, a + 1, size=(nb_rx, 14, N, 2)).astype(np.int16), rng.integers(-b, b + 1, size=(nb_rx, 14, N, 2)).astype(np.int16)
    start_re = (N - carrier * 6 + rb_start * 12) % N
    hh = (rng.normal(size=(nb_rx, 1, 1)) + 1j * rng.normal(size=(nb_rx, 1, 1))) * 600
    tx = (rng.choice([-1, 1], size=(14, N)) + 1j * rng.choice([-1, 1], size=(14, N))) * 700
    pos = oracle.ptrs_symbols(start, nsym, L, dpos)
    krb = rnti % K if rb_size % K == 0 else rnti % (rb_size % K)
    for m in range(14):
        if (pos >> m) & 1:
            g = oracle.gold_words(((((14 * slot + m + 1) * (2 * nid + 1)) << 17) + 2 * nid + nscid) % (1 << 31), 20)
            j = 0
            for re in range(12 * rb_size):
                if (re - reoff - krb * 12) % (K * 12) == 0:
                    b0 = (int(g[(2 * j) >> 5]) >> ((2 * j) & 31)) & 1
                    b1 = (int(g[(2 * j + 1) >> 5]) >> ((2 * j + 1) & 31)) & 1
                    tx[m, (start_re + re) % N] = ((-1 if b0 else 1) + 1j * (-1 if b1 else 1)) * 700
                    j += 1
    rot = np.exp(1j * b * np.arange(14))[None, :, None]
    y = hh * tx[None] * rot / 600 + a * (rng.normal(size=(nb_rx, 14, N)) + 1j * rng.normal(size=(nb_rx, 14, N)))
    rx = np.stack([np.round(y.real), np.round(y.imag)], -1).clip(-32768, 32767).astype(np.int16)
    hf = np.broadcast_to(hh, (nb_rx, 14, N))
    return rx, np.stack([np.round(hf.real), np.round(hf.imag)], -1).astype(np.int16)


RFSIM_CASES = [  # nb_tx, nb_rx, channel_length, channel_offset, path_loss_dB, noise_power_dB, samples, TS, CirSize factor, amplitude
    (1, 1, 1, 0, 0.0, -100.0, 3000, 100000, 4, 3000), (2, 2, 12, 0, -3.0, -20.0, 7680, 123456, 3, 8000), (4, 4, 30, 3, -10.5, -6.0, 5000, 30720 * 7 + 11, 2, 32767),
    (2, 4, 200, 0, 6.0, -40.0, 2048, 999, 2, 1000), (4, 2, 63, -5, -0.1, 0.0, 4097, 2 ** 33 + 5, 2, 20000), (1, 2, 7, 1, 20.0, -3.0, 1000, 50, 1, 32767),
]


def rfsim_inputs(rng, case):
    """Channel taps, circular tx buffer, pre-filled output and noise draws for one rxAddInput call (tests, golden generator)."""
    nb_tx, nb_rx, L, offset, pl, npw, n, TS, cf, amp = case
    cir = cf * (n + L + 16)
    ch = rng.normal(size=(nb_tx * nb_rx, L, 2)) * (0.7 / np.sqrt(L))
    sig = rng.integers(-amp, amp + 1, size=(cir, 2)).astype(np.int16)
    out = rng.integers(-200, 201, size=(n, 2)).astype(np.int16)
    noise = rng.normal(size=(n, 2))
    return cir, ch, sig, out, noise


PRACH_CASES = [  # nb_rx, short_sequence, rootSequenceIndex, NCS, format, mu, sent preamble (-1: noise only), delay in samples of the ZC sequence, amplitude, noise sigma
    (1, 0, 10, 13, 0, 1, 5, 2, 3000, 0), (2, 0, 22, 13, 0, 1, 37, 7, 800, 300), (4, 0, 100, 26, 1, 0, 63, 0, 500, 500), (2, 0, 0, 0, 3, 1, 9, 3, 2000, 100),
    (4, 0, 300, 119, 0, 1, 20, 40, 1200, 400), (1, 1, 5, 12, 8, 1, 11, 1, 4000, 0), (2, 1, 40, 23, 5, 1, 30, 5, 1500, 500), (4, 1, 120, 0, 7, 3, 44, 0, 900, 300),
    (2, 0, 10, 13, 0, 1, -1, 0, 0, 2000), (3, 1, 1, 34, 4, 1, 2, 10, 32767, 0), (4, 0, 837, 46, 2, 1, 50, 12, 20000, 8000),
]


def prach_num_roots(short_sequence, NCS):
    N_ZC = 139 if short_sequence else 839
    return 64 if NCS == 0 else -(-64 // (N_ZC // NCS))


def prach_inputs(rng, case, xu):
    """rxsigF [nb_rx][N_ZC][2] for one PRACH occasion: the sent preamble's root sequence with its cyclic shift and a delay, a random phase per antenna, noise."""
    nb_rx, short, root, NCS, fmt, mu, pre, delay, amp, sigma = case
    N_ZC = 139 if short else 839
    k = np.arange(N_ZC)
    y = np.zeros((nb_rx, N_ZC), np.complex128)
    if pre >= 0:
        per_root = N_ZC // NCS if NCS else 1
        r, v = pre // per_root, pre % per_root
        x = xu[r, :N_ZC, 0].astype(np.float64) + 1j * xu[r, :N_ZC, 1]
        shift = v * NCS - delay * N_ZC / (1024 if not short else 256)
        for a in range(nb_rx):
            y[a] = x / 32768.0 * amp * np.exp(2j * np.pi * k * shift / N_ZC) * np.exp(1j * rng.uniform(0, 6.28))
    y += sigma * (rng.normal(size=y.shape) + 1j * rng.normal(size=y.shape))
    return np.stack([np.round(y.real), np.round(y.imag)], -1).clip(-32768, 32767).astype(np.int16)


def ptrs_fuzz_cases(rng, n, N=512, carrier=25, safe_tail=True):
    """Random valid PDSCH + PT-RS configurations on a small carrier (the tuple layout of PTRS_CASES): allocation, DMRS symbols (1-3, at least one inside the
    allocation, possibly its first symbol), DMRS type / CDM groups, modulation, antennas, PT-RS densities and offsets, RNTI, slot, scrambling."""
    out = []
    while len(out) < n:
        rb_size = int(rng.integers(1, carrier + 1)); rb_start = int(rng.integers(0, carrier - rb_size + 1))
        start = int(rng.integers(0, 4)); nsym = int(rng.integers(3, 15 - start))
        k = int(rng.integers(1, 4))
        syms = rng.choice(np.arange(start, start + nsym), size=min(k, nsym), replace=False)
        dpos = 0
        for s_ in syms:
            dpos |= 1 << int(s_)
        dtype_ = int(rng.integers(0, 2)); cdm = int(rng.integers(1, 3))
        L, K = int(rng.integers(0, 3)), int(rng.choice([2, 4]))
        # The reference's LLR routines round a symbol's RE count up to whole SIMD vectors; the spill lands in the next symbol's LLRs (rewritten right after) except
        # for the slot's LAST symbol, where it runs past the end of nr_rx_pdsch's own malloc'ed layer_llr buffer -- with PT-RS the count is no longer a multiple
        # of 12 and glibc aborts with "double free or corruption" (defect 22).  Keep the last symbol's count a multiple of 16.
        last = start + nsym - 1
        i, l_ref, Ls, mask = 0, start, 1 << L, 0
        while l_ref + i * Ls <= last:                                       # set_ptrs_symb_idx
            hit = [l for l in range(l_ref + i * Ls, max(l_ref + (i - 1) * Ls + 1, l_ref) - 1, -1) if (dpos >> l) & 1]
            if hit:
                l_ref, i = hit[0], 1
                continue
            mask |= 1 << (l_ref + i * Ls)
            i += 1
        v_last = 0
        for l in range(start, last + 1):                                    # the last symbol that carries data
            v = rb_size * ((12 - 6 * cdm) if dtype_ == 0 else (12 - 4 * cdm)) if (dpos >> l) & 1 else 12 * rb_size - (((rb_size + K - 1) // K) if (mask >> l) & 1 else 0)
            if v:
                v_last = v
        if safe_tail and v_last % 16:
            continue
        out.append((N, int(rng.integers(1, 5)), rb_start, rb_size, int(rng.choice([2, 4, 6, 8])), dpos, dtype_, cdm, carrier, start, nsym,
                    L, K, int(rng.integers(0, 12)), int(rng.integers(0, 65536)), int(rng.integers(0, 20)),
                    int(rng.integers(0, 2)), int(rng.integers(0, 65536))))
    return out


def pdsch_tx_fuzz_cases(rng, n, N=512, carrier=25):
    """Random valid PDSCH transmitter configurations: (N, carrier, ntx, slot, rb_start, rb_size, Qm, layers, start, nsym, dmrs_pos, dmrs_type, cdm, ports, scid, amp,
    ptrs or None, precoding index).  DMRS symbols only inside the allocation (the reference derives its length from the whole mask), ports 0 ... layers - 1."""
    out = []
    while len(out) < n:
        rb_size = int(rng.integers(1, carrier + 1)); rb_start = int(rng.integers(0, carrier - rb_size + 1))
        start = int(rng.integers(0, 4)); nsym = int(rng.integers(3, 15 - start))
        syms = rng.choice(np.arange(start, start + nsym), size=min(int(rng.integers(1, 4)), nsym), replace=False)
        dpos = 0
        for s_ in syms:
            dpos |= 1 << int(s_)
        nl = int(rng.integers(1, 5)); dtype_ = int(rng.integers(0, 2))
        cdm = int(rng.integers(2 if nl > 2 else 1, 3))
        ntx = int(rng.integers(nl, 5))
        ptrs = (int(rng.integers(0, 3)), int(rng.choice([2, 4])), int(rng.integers(0, 12))) if rng.integers(0, 2) else None
        pm = int(rng.integers(1, 4)) if (ntx >= 2 and rng.integers(0, 2)) else 0
        out.append((N, carrier, ntx, int(rng.integers(0, 20)), rb_start, rb_size, int(rng.choice([2, 4, 6, 8])), nl, start, nsym, dpos, dtype_, cdm, (1 << nl) - 1,
                    int(rng.integers(0, 2)), int(rng.choice([300, 512, 2047, 30000])), ptrs, pm))
    return out


def chest_fuzz_cases(rng, n, N=512, carrier=25, max_rx=4):
    """Random DMRS type 1 channel-estimation configurations: (N, nb_rx, slot, symbol, port, rb_start, rb_size, carrier, scid, dmrs id, delay of the synthetic channel)."""
    out = []
    for _ in range(n):
        rb_size = int(rng.integers(1, carrier + 1)); rb_start = int(rng.integers(0, carrier - rb_size + 1))
        out.append((N, int(rng.integers(1, max_rx + 1)), int(rng.integers(0, 20)), int(rng.integers(0, 14)), int(rng.integers(0, 4)), rb_start, rb_size, carrier,
                    int(rng.integers(0, 2)), int(rng.integers(0, 65536)), int(rng.integers(-12, 13))))
    return out


def chest_inputs(oracle, rng, P, port, delay, amp_noise=300):
    """rxdataF with the port's DMRS through a channel with a linear phase (a delay the estimator can find) plus noise; every third call full-scale noise only."""
    N, nb_rx, symbol, rb_size = P.fft_size, P.nb_rx, P.symbol, P.rb_size
    if amp_noise is None:
        return rng.integers(-32768, 32768, size=(nb_rx, 14, N, 2)).astype(np.int16)
    rx = rng.integers(-amp_noise, amp_noise + 1, size=(nb_rx, 14, N, 2)).astype(np.int16)
    pil = oracle.pusch_dmrs_pilots(P).reshape(-1, 2).astype(np.float64)
    k0 = (P.rb_start * 12 + P.first_carrier_offset) % N
    idx = (k0 + 2 * np.arange(6 * rb_size) + ((port >> 1) & 1)) % N
    for a in range(nb_rx):
        hh = (900 + 100 * a) * np.exp(1j * (0.3 * a - 2 * np.pi * delay * np.arange(6 * rb_size) * 2 / N))
        y = hh * (pil[:, 0] - 1j * pil[:, 1]) / 32767.0
        rx[a, symbol, idx, 0] += np.round(y.real).astype(np.int16); rx[a, symbol, idx, 1] += np.round(y.imag).astype(np.int16)
    return rx


def prach_fuzz_cases(rng, n):
    """Random PRACH occasions (the tuple layout of PRACH_CASES) with N_CS from the 38.211 tables 6.3.3.1-5 / -7 (unrestricted set)."""
    ncs_long = [0, 13, 15, 18, 22, 26, 32, 38, 46, 59, 76, 93, 119, 167, 279, 419]
    ncs_short = [0, 2, 4, 6, 8, 10, 12, 13, 15, 17, 19, 23, 27, 34, 46, 69]
    out = []
    for _ in range(n):
        short = int(rng.integers(0, 2))
        ncs = int(rng.choice(ncs_short if short else ncs_long))
        fmt = int(rng.integers(4, 13)) if short else int(rng.integers(0, 4))
        pre = int(rng.integers(-1, 64))
        amp = int(rng.choice([400, 1500, 6000, 32767]))
        out.append((int(rng.integers(1, 5)), short, int(rng.integers(0, 137 if short else 837)), ncs, fmt, int(rng.integers(0, 4)), pre, int(rng.integers(0, 8)), amp,
                    int(rng.choice([0, amp // 8, amp // 2]))))
    return out


def rm_fuzz_cases(rng, n):
    """Random rate-matching configurations: (BG, Z, F, E list (1-3 segments, multiples of Qm), rv, Tbslbrm, C, Qm)."""
    zs = [16, 24, 36, 52, 64, 96, 128, 208, 256, 320, 384]
    out = []
    for _ in range(n):
        BG = int(rng.integers(1, 3)); Z = int(rng.choice(zs)); Qm = int(rng.choice([2, 4, 6, 8]))
        K = (22 if BG == 1 else 10) * Z; N = (66 if BG == 1 else 50) * Z
        F = int(rng.integers(0, max(1, min(K - 2 * Z - 8, 6 * Z)) // 8 + 1)) * 8
        ne = int(rng.integers(1, 4))
        E0 = int(rng.integers(max(60, N // 8), 2 * N)) // Qm * Qm
        Es = [E0 + Qm * int(rng.integers(0, 2)) for _ in range(ne)]
        out.append((BG, Z, F, Es, int(rng.integers(0, 4)), 0 if rng.integers(0, 2) else int(rng.integers(3000, 400000)), int(rng.integers(ne, 24)), Qm))
    return out
