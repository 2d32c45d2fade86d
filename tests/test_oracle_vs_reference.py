"""Pins the CPU oracle (oracle/nrb200_oracle.c) bit-exactly against the UNMODIFIED reference compiled by oracle/build_ref.sh.
Runs wherever oracle/_ref exists (this container; the built .so files also travel to the GPU box)."""
import ctypes as C
import numpy as np
import pytest
from common import ALL_Z, RATES, NCOLS, make_case, payloads


def test_lifting_sizes(oracle):
    assert [z for z in range(0, 400) if oracle.ils_of_z(z) >= 0] == ALL_Z


@pytest.mark.parametrize("BG", [1, 2])
def test_encoder_all_z(oracle, reference, BG):
    """ldpctest.c:286-292 cross-check, for every lifting size whose K is a whole number of bytes."""
    for Z in ALL_Z:
        K, P = payloads(BG, Z, 3, Z)
        if K % 8:
            continue
        orig = reference.encode(BG, Z, K, P, orig=True)
        optim = reference.encode(BG, Z, K, P)
        mine = np.stack([oracle.encode(BG, Z, K, P[i]) for i in range(3)])
        assert np.array_equal(mine, orig), (BG, Z)
        if (BG, Z) == (2, 64):
            # reference defect: `case 64: break;` in encode_parity_check_part_optim (ldpc_encode_parity_check.c BG2 switch)
            # while LDPCencoder routes BG2 Zc>=64 there => the default encoder emits all-zero parity for BG2 Z=64.
            assert not np.array_equal(optim, orig) and not optim[:, 8 * Z:].any()
        else:
            assert np.array_equal(optim, orig), (BG, Z)


@pytest.mark.parametrize("BG,R,ebn0", [(1, 13, 1.0), (1, 13, 2.4), (1, 13, 3.0), (1, 23, 4.0), (1, 89, 7.0), (2, 13, 2.5), (2, 23, 5.0)])
def test_decoder_z384(oracle, reference, BG, R, ebn0):
    K, P, llr = make_case(oracle, BG, 384, R, 3, ebn0, seed=R)
    for i in range(3):
        it_r, out_r = reference.decode(BG, 384, R, 8, llr[i])
        it_o, out_o = oracle.decode(BG, 384, R, 8, llr[i])
        assert it_r == it_o and np.array_equal(out_r, out_o)


def test_decoder_all_z_all_rates(oracle, reference):
    for BG in (1, 2):
        for R in RATES[BG]:
            if (BG, R) == (2, 15):
                continue   # AVX2 build has a generator defect there, see test_bg2_r15_avx2_defect
            for Z in ALL_Z[::3] + [384]:
                K, P, llr = make_case(oracle, BG, Z, R, 1, 3.0 if R in (13, 15) else 6.0, seed=Z + R)
                it_r, out_r = reference.decode(BG, Z, R, 8, llr[0])
                it_o, out_o = oracle.decode(BG, Z, R, 8, llr[0])
                assert it_r == it_o and np.array_equal(out_r, out_o), (BG, Z, R)


def test_bg2_r15_avx2_defect(oracle, reference):
    """cnProc_gen_BG2_avx2.c emits `i+=2` for the degree-3 group => the AVX2 build never updates odd 32-byte vectors of those
    check nodes.  The oracle reproduces the AVX2 build bit-exactly with quirk bit 0 and the intended arithmetic without."""
    for Z, e in ((384, 0.5), (64, 3.0), (7, 3.0)):
        K, P, llr = make_case(oracle, 2, Z, 15, 2, e, seed=Z)
        for i in range(2):
            it_r, out_r = reference.decode(2, Z, 15, 8, llr[i])
            oracle.lib.orc_set_quirks(1)
            it_q, out_q = oracle.decode(2, Z, 15, 8, llr[i])
            oracle.lib.orc_set_quirks(0)
            assert it_r == it_q and np.array_equal(out_r, out_q)


def test_bg2_r15_avx512_is_the_intended_arithmetic(oracle, reference512):
    for Z, e in ((384, 0.5), (384, 3.0), (64, 3.0), (7, 3.0)):
        K, P, llr = make_case(oracle, 2, Z, 15, 2, e, seed=Z)
        for i in range(2):
            it_r, out_r = reference512.decode(2, Z, 15, 8, llr[i])
            it_o, out_o = oracle.decode(2, Z, 15, 8, llr[i])
            assert it_r == it_o and np.array_equal(out_r, out_o)


@pytest.mark.parametrize("out_mode", [1, 2])
def test_output_modes(oracle, reference, out_mode):
    """LLRINT8 really yields hard bits in the reference (llr2bit runs in place over p_out, nrLDPC_decoder.c:866-877)."""
    K, P, llr = make_case(oracle, 1, 96, 13, 2, 3.0, seed=3)
    for i in range(2):
        it_r, out_r = reference.decode(1, 96, 13, 8, llr[i], out_mode)
        it_o, out_o = oracle.decode(1, 96, 13, 8, llr[i], out_mode)
        assert it_r == it_o and np.array_equal(out_r, out_o)
        assert set(np.unique(out_r)) <= {0, 1}


@pytest.mark.parametrize("max_iter", [0, 1, 2, 3, 5, 20])
def test_iteration_caps_and_crc_mode(oracle, reference, max_iter):
    BG, Z, R = 1, 128, 13
    K = 22 * Z
    rng = np.random.default_rng(max_iter)
    P = rng.integers(0, 256, size=(4, K // 8), dtype=np.uint8)
    for i in range(4):   # attach a CRC24B so that check_crc can succeed
        crc = oracle.crc(1, P[i], K - 24) >> 8
        P[i, -3:] = [(crc >> 16) & 0xFF, (crc >> 8) & 0xFF, crc & 0xFF]
    from openairinterface5g_b200.synth import awgn_llr
    cw = np.stack([oracle.encode(BG, Z, K, P[i]) for i in range(4)])
    for ebn0 in (1.5, 3.0):
        llr = awgn_llr(cw, Z, 68, ebn0, 1 / 3, max_iter)
        for i in range(4):
            for use_crc in (0, 1):
                it_r, out_r = reference.decode(BG, Z, R, max_iter, llr[i], 0, use_crc, K, 1)
                it_o, out_o = oracle.decode(BG, Z, R, max_iter, llr[i], 0, use_crc, K, 1)
                assert it_r == it_o, (max_iter, ebn0, i, use_crc)
                assert np.array_equal(out_r, out_o), (max_iter, ebn0, i, use_crc)


def test_abort_flag(oracle, reference):
    K, P, llr = make_case(oracle, 1, 64, 13, 1, 3.0, seed=1)
    it_r, out_r = reference.decode(1, 64, 13, 8, llr[0], abort_in=1)
    it_o, out_o = oracle.decode(1, 64, 13, 8, llr[0], abort_in=1)
    assert it_r == it_o == 10 and np.array_equal(out_r, out_o)


def test_crc_all_polys(oracle, reference):
    rng = np.random.default_rng(0)
    for n in (8, 24, 32, 100, 1001, 3840, 8424, 8448):
        d = rng.integers(0, 256, size=(n + 7) // 8 + 4, dtype=np.uint8)
        for poly in range(8):
            assert oracle.crc(poly, d, n) == reference.crc(poly, d, n), (poly, n)
    for crc_type, poly, L in ((0, 0, 24), (1, 1, 24), (2, 3, 16), (3, 6, 8)):
        d = rng.integers(0, 256, size=64, dtype=np.uint8)
        c = oracle.crc(poly, d, 8 * 64 - L) >> (32 - L)
        for b in range(L // 8):
            d[64 - L // 8 + b] = (c >> (8 * (L // 8 - 1 - b))) & 0xFF
        assert oracle.check_crc(d, 512, crc_type) == reference.check_crc(d, 512, crc_type) == 1
        d[3] ^= 4
        assert oracle.check_crc(d, 512, crc_type) == reference.check_crc(d, 512, crc_type) == 0


def test_rate_matching_and_interleaving(oracle, reference):
    rng = np.random.default_rng(5)
    for BG, Z, F, E, rv, Tbslbrm, Cseg in ((1, 384, 0, 9072, 0, 0, 1), (1, 384, 88, 9072, 2, 0, 3), (1, 96, 40, 30000, 1, 0, 2), (2, 128, 16, 5000, 3, 0, 1),
                                           (1, 384, 0, 20000, 0, 200000, 20), (2, 52, 0, 1200, 0, 0, 1), (1, 208, 120, 4100, 3, 90000, 4), (2, 384, 200, 19000, 2, 0, 2)):
        N = (66 if BG == 1 else 50) * Z
        K = (22 if BG == 1 else 10) * Z
        Foffset = K - F - 2 * Z
        w = rng.integers(0, 2, size=N, dtype=np.uint8)
        w[Foffset:Foffset + F] = 2   # NR_NULL
        rc_r, e_r = reference.rate_matching_tx(Tbslbrm, BG, Z, w, Cseg, F, Foffset, rv, E)
        rc_o, e_o = oracle.rate_matching_tx(Tbslbrm, BG, Z, w, Cseg, F, Foffset, rv, E)
        assert rc_r == rc_o == 0 and np.array_equal(e_r, e_o), (BG, Z, F, E, rv)
        for Qm in (2, 4, 6, 8):
            Eq = E - E % Qm
            assert np.array_equal(reference.interleave(Eq, Qm, e_r[:Eq]), oracle.interleave(Eq, Qm, e_r[:Eq]))
            soft = rng.integers(-300, 300, size=Eq, dtype=np.int16)
            assert np.array_equal(reference.deinterleave(Eq, Qm, soft), oracle.deinterleave(Eq, Qm, soft))
        soft = rng.integers(-128, 128, size=E, dtype=np.int16)
        w_r = rng.integers(-50, 50, size=N + 16, dtype=np.int16)
        w_o = w_r.copy()
        for clear in (1, 0):
            assert reference.rate_matching_rx(Tbslbrm, BG, Z, w_r, soft, Cseg, rv, clear, E, F, Foffset) == 0
            assert oracle.rate_matching_rx(Tbslbrm, BG, Z, w_o, soft, Cseg, rv, clear, E, F, Foffset) == 0
            assert np.array_equal(w_r, w_o)
        for rnd in (0, 1):
            assert reference.get_R(rv, E, BG, Z, 0, rnd) == oracle.get_R(rv, E, BG, Z, 0, rnd)


def test_rate_matching_fuzz(oracle, reference):
    """150 random rate-matching configurations (base graph, lifting size, filler bits, E, redundancy version, LBRM on / off, segments, modulation): bit selection,
    interleaving, de-interleaving and rate recovery with soft combining against nr_rate_matching.c."""
    from common import rm_fuzz_cases
    rng = np.random.default_rng(96)
    for BG, Z, F, Es, rv, Tbslbrm, Cseg, Qm in rm_fuzz_cases(rng, 150):
        N = (66 if BG == 1 else 50) * Z
        K = (22 if BG == 1 else 10) * Z
        Foffset = K - F - 2 * Z
        E = Es[0]
        w = rng.integers(0, 2, size=N, dtype=np.uint8)
        w[Foffset:Foffset + F] = 2
        rc_r, e_r = reference.rate_matching_tx(Tbslbrm, BG, Z, w, Cseg, F, Foffset, rv, E)
        rc_o, e_o = oracle.rate_matching_tx(Tbslbrm, BG, Z, w, Cseg, F, Foffset, rv, E)
        assert rc_r == rc_o and (rc_r != 0 or np.array_equal(e_r, e_o)), (BG, Z, F, E, rv, Tbslbrm, Cseg, rc_r, rc_o)
        if rc_r != 0:
            continue
        assert np.array_equal(reference.interleave(E, Qm, e_r[:E]), oracle.interleave(E, Qm, e_r[:E]))
        soft = rng.integers(-300, 300, size=E, dtype=np.int16)
        assert np.array_equal(reference.deinterleave(E, Qm, soft), oracle.deinterleave(E, Qm, soft))
        w_r = rng.integers(-50, 50, size=N + 16, dtype=np.int16)
        w_o = w_r.copy()
        for clear in (1, 0):
            soft = rng.integers(-128, 128, size=E, dtype=np.int16)
            rr = reference.rate_matching_rx(Tbslbrm, BG, Z, w_r, soft, Cseg, rv, clear, E, F, Foffset)
            ro = oracle.rate_matching_rx(Tbslbrm, BG, Z, w_o, soft, Cseg, rv, clear, E, F, Foffset)
            assert rr == ro and np.array_equal(w_r, w_o), (BG, Z, F, E, rv, Tbslbrm, Cseg, clear, rr, ro)


def test_segmentation(oracle, reference):
    rng = np.random.default_rng(9)
    for BG, B in ((1, 8448), (1, 8456), (1, 100000), (1, 424), (2, 3840), (2, 3848), (2, 600), (2, 200), (2, 100), (1, 1277992 // 8 * 8), (2, 40000)):
        data = rng.integers(0, 256, size=B // 8 + 8, dtype=np.uint8)
        r = reference.segmentation(data, B, BG)
        o = oracle.segmentation(data, B, BG)
        assert r[:5] == o[:5], (BG, B, r[:5], o[:5])
        assert np.array_equal(r[5], o[5]), (BG, B)


DFT_SIZES = [64, 128, 256, 512, 768, 1024, 1536, 2048, 3072, 4096, 6144, 8192]


@pytest.mark.parametrize("N", DFT_SIZES)
def test_dft_idft_q15(oracle, reference, N):
    """Bit-exact Q15 DFT/IDFT restatement vs oai_dfts.c for the OFDM sizes, over amplitudes that do and do not saturate."""
    rng = np.random.default_rng(N)
    for inverse in (False, True):
        for amp in (300, 3000, 20000, 32767):
            for scale in (1, 0):
                x = rng.integers(-amp, amp + 1, size=2 * N).astype(np.int16)
                assert np.array_equal(oracle.dft(N, inverse, x, scale), reference.dft(N, inverse, x, scale)), (N, inverse, amp, scale)
        x = rng.choice(np.array([-32768, 32767, 0], dtype=np.int16), size=2 * N)
        assert np.array_equal(oracle.dft(N, inverse, x, 1), reference.dft(N, inverse, x, 1))


@pytest.mark.parametrize("N", [12288, 16384, 18432, 24576, 36864, 49152])
def test_dft_idft_large_q15(oracle, reference, N):
    """The sizes above 8192 that work in the reference (radix-3 / radix-4 levels over 4096, 6144, 8192; 12288 and 18432 hand their scale argument down).
    9216 / 73728 are AssertFatal there, 32768 / 98304 overrun their stack buffers and 65536 reads beyond its twiddle table: those cannot be pinned."""
    rng = np.random.default_rng(N)
    for inverse in (False, True):
        for amp, scale in ((300, 1), (3000, 0), (32767, 1), (3000, 2)):
            x = rng.integers(-amp, amp + 1, size=2 * N).astype(np.int16)
            assert np.array_equal(oracle.dft(N, inverse, x, scale), reference.dft(N, inverse, x, scale)), (N, inverse, amp, scale)


@pytest.mark.parametrize("N,inverse", [(32768, False), (32768, True), (65536, True), (98304, False), (98304, True)])
def test_dft_large_unpinned_sizes_are_dfts(oracle, N, inverse):
    """Sizes whose reference implementation is broken (see above): the restatement applies the same level arithmetic as the working sizes; checked against a
    float transform (gain 1/sqrt(N) with scale 1)."""
    x = np.random.default_rng(N).integers(-300, 301, size=2 * N).astype(np.int16)
    y = oracle.dft(N, inverse, x, 1)
    X = x[0::2] + 1j * x[1::2]
    Y = (np.fft.ifft(X) * N if inverse else np.fft.fft(X)) / np.sqrt(N)
    ya = y[0::2] + 1j * y[1::2]
    # the truncating >> 15 of every level leaves a bias that piles up in the bins around DC, so compare in the rms sense (measured: 1.1-1.2 %)
    assert np.sqrt((np.abs(ya - Y) ** 2).mean()) < 0.02 * np.sqrt((np.abs(Y) ** 2).mean())


FOURWAY_SIZES = [12, 24, 36, 48, 60, 72, 96, 108, 120, 144, 180, 192, 216, 240, 288, 300, 324, 360, 384, 432, 480, 540, 576, 600, 648, 720, 864, 900, 960, 972,
                 1080, 1152, 1200, 1296, 1440, 1500, 1620, 1728, 1800, 1920, 1944, 2160, 2400, 2592, 2700, 2880, 2916, 3000, 3240]


@pytest.mark.parametrize("N", FOURWAY_SIZES + [768])
def test_dft_fourway_q15(oracle, reference, N):
    """The DFT-s-OFDM entry points dft12 ... dft3240 (four interleaved transforms per call) vs the restatement; 768 is the internal dft768p."""
    rng = np.random.default_rng(N)
    name = "dft768p" if N == 768 else None
    for amp in (300, 3000, 20000, 32767):
        for scale in (1, 0):
            x = rng.integers(-amp, amp + 1, size=8 * N).astype(np.int16)
            assert np.array_equal(oracle.dft4(N, x, scale), reference.dft4(N, x, scale, name=name)), (N, amp, scale)
    x = rng.choice(np.array([-32768, 32767, 0], dtype=np.int16), size=8 * N)
    assert np.array_equal(oracle.dft4(N, x, 1), reference.dft4(N, x, 1, name=name))


def test_dft2304_reference_is_not_reproducible(oracle, reference):
    """dft2304 (oai_dfts.c:7288) runs the single-transform dft768 over four-way data and combines stack it never wrote: two calls on the same input differ.
    The restatement returns the 768 x 3 transform instead; it is checked against a float DFT."""
    N = 2304
    x = np.random.default_rng(1).integers(-300, 301, size=8 * N).astype(np.int16)
    b1 = reference.dft4(N, x, 1)
    reference.dft4(3240, np.random.default_rng(2).integers(-30000, 30001, size=8 * 3240).astype(np.int16), 1)     # leaves other bytes on the stack
    b2 = reference.dft4(N, x, 1)
    assert not np.array_equal(b1, b2)
    a = oracle.dft4(N, x, 1)
    X = (x[0::2] + 1j * x[1::2]).reshape(N, 4)
    Y = np.fft.fft(X, axis=0) / np.sqrt(N)
    ya = (a[0::2] + 1j * a[1::2]).reshape(N, 4)
    assert np.abs(ya - Y).max() < 0.06 * np.sqrt((np.abs(Y) ** 2).mean())


@pytest.mark.parametrize("Qm", [2, 4, 6, 8])
def test_pusch_llr(oracle, reference, Qm):
    """nr_ulsch_compute_llr (AVX2 path) vs the restatement, for RE counts that are and are not multiples of 8."""
    rng = np.random.default_rng(Qm)
    for n in (8, 96, 3276, 3272):
        y = rng.integers(-32768, 32768, size=2 * n).astype(np.int16)
        y[:16] = rng.choice(np.array([-32768, 32767, 0, -1], dtype=np.int16), size=16)
        mags = [rng.integers(0, 20000, size=2 * n).astype(np.int16) for _ in range(3)]
        assert np.array_equal(oracle.ulsch_llr(Qm, y, *mags), reference.ulsch_llr(Qm, y, *mags)), (Qm, n)


def test_scrambling_and_modulation(oracle, reference):
    rng = np.random.default_rng(6)
    for size, q, Nid, rnti in ((64, 0, 0, 1), (1000, 0, 123, 0x1234), (9072 * 6, 1, 1007, 65535), (33, 0, 5, 77)):
        bits = rng.integers(0, 2, size=size, dtype=np.uint8)
        sc_o, sc_r = oracle.scramble(bits, q, Nid, rnti), reference.scramble(bits, q, Nid, rnti)
        assert np.array_equal(sc_o, sc_r), (size, q, Nid, rnti)
        for Qm in (2, 4, 6, 8):
            length = (size // (Qm * 8)) * Qm * 8 if Qm != 6 else (size // 24) * 24
            if length < Qm * 8 * 4:
                continue
            assert np.array_equal(oracle.modulate(sc_o, length, Qm), reference.modulate(sc_r, length, Qm)), (size, Qm)


def test_llr_scrambling_modulation_fuzz(oracle, reference):
    """Random sizes and seeds: max-log LLRs (any RE count that is a multiple of 4, negative and full-scale thresholds), scrambling, the QAM mapper and unscrambling."""
    rng = np.random.default_rng(98)
    for n in range(120):
        Qm = int(rng.choice([2, 4, 6, 8]))
        nre = 4 * int(rng.integers(1, 900))
        y = rng.integers(-32768, 32768, size=2 * nre).astype(np.int16)
        mags = [rng.integers(-32768 if n % 4 == 0 else 0, 32768, size=2 * nre).astype(np.int16) for _ in range(3)]
        assert np.array_equal(oracle.ulsch_llr(Qm, y, *mags), reference.ulsch_llr(Qm, y, *mags)), ("llr", Qm, nre)
        size, q, Nid, rnti = int(rng.integers(1, 60000)), int(rng.integers(0, 2)), int(rng.integers(0, 1024)), int(rng.integers(0, 65536))
        bits = rng.integers(0, 2, size=size, dtype=np.uint8)
        sc_o, sc_r = oracle.scramble(bits, q, Nid, rnti), reference.scramble(bits, q, Nid, rnti)
        assert np.array_equal(sc_o, sc_r), ("scramble", size, q, Nid, rnti)
        length = (size // (Qm * 8)) * Qm * 8 if Qm != 6 else (size // 24) * 24
        if length >= Qm * 8 * 4:
            assert np.array_equal(oracle.modulate(sc_o, length, Qm), reference.modulate(sc_r, length, Qm)), ("modulate", size, Qm)
        llr = rng.integers(-32768, 32768, size=size).astype(np.int16)
        assert np.array_equal(oracle.unscramble_llr(llr, q, Nid, rnti), reference.unscramble_llr(llr, q, Nid, rnti)), ("unscramble", size, q, Nid, rnti)


def test_unscrambling(oracle, reference):
    rng = np.random.default_rng(8)
    for size, q, Nid, rnti in ((64, 0, 0, 1), (9072, 1, 1007, 65535), (12 * 273 * 6, 0, 500, 4660)):
        llr = rng.integers(-32768, 32768, size=size).astype(np.int16)
        assert np.array_equal(oracle.unscramble_llr(llr, q, Nid, rnti), reference.unscramble_llr(llr, q, Nid, rnti))


# ------------------------------------------------------------------------------------------ slot-level OFDM front end (a20)
OFDM_CASES = [  # N, mu, nb_rb, slot
    (4096, 1, 273, 0), (4096, 1, 273, 3), (2048, 1, 106, 1), (1024, 0, 52, 2), (1536, 1, 78, 4), (512, 0, 25, 0), (3072, 1, 162, 2), (2048, 2, 66, 5),
]


def test_rotation_tables(oracle, reference):
    for N, mu, nb_rb, _ in OFDM_CASES:
        for div in (8, 4):
            dl, ul, ts = reference.rotation_tables(N, mu, nb_rb, div, 3619200000.0, 3619200000.0 - 1e7)
            n = 2 * (14 << mu)
            assert np.array_equal(oracle.symbol_rotation(mu, 3619200000.0), dl[:n])
            assert np.array_equal(oracle.symbol_rotation(mu, 3619200000.0 - 1e7), ul[:n])
            assert np.array_equal(oracle.timeshift_rotation(N, (N // 128 * 9) // div), ts)


def test_ofdm_tx_slot(oracle, reference):
    rng = np.random.default_rng(20)
    for N, mu, nb_rb, slot in OFDM_CASES:
        rot = oracle.symbol_rotation(mu, 3619200000.0)
        rot224 = np.zeros(448, np.int16); rot224[:rot.size] = rot
        F = np.zeros((14, N, 2), np.int16)
        fco = N - nb_rb * 6
        amp = 32767 if slot == 3 else 4000               # one case drives the saturating / truncating corners of rotate_cpx_vector
        F[:, :nb_rb * 6] = rng.integers(-amp, amp + 1, size=(14, nb_rb * 6, 2))
        F[:, fco:] = rng.integers(-amp, amp + 1, size=(14, nb_rb * 6, 2))
        for use_rot in (True, False):
            y_o, F_o = oracle.ofdm_tx_slot(N, mu, nb_rb, slot, 14, rot if use_rot else None, F)
            y_r, F_r = reference.ofdm_tx_slot(N, mu, nb_rb, slot, 14, rot224 if use_rot else None, F, y_o.size // 2)
            assert np.array_equal(F_o, F_r), (N, mu, nb_rb, slot, "rotated txdataF")
            assert np.array_equal(y_o, y_r), (N, mu, nb_rb, slot, use_rot)


def test_ofdm_rx_slot(oracle, reference):
    rng = np.random.default_rng(21)
    for N, mu, nb_rb, slot in OFDM_CASES:
        _, _, _, frame_len = oracle.ofdm_geometry(N, mu, slot)
        rot = oracle.symbol_rotation(mu, 3609200000.0)
        rot224 = np.zeros(448, np.int16); rot224[:rot.size] = rot
        amp = 32767 if slot == 3 else 3000
        rx = rng.integers(-amp, amp + 1, size=2 * frame_len).astype(np.int16)
        for div, ta in ((8, 0), (8, N // 8), (4, N // 32 + 3), (8, N // 4 + 5)):   # in slot 0 a timing offset wraps around the frame buffer
            for use_rot in (True, False):
                y_o = oracle.ofdm_rx_slot(N, mu, nb_rb, slot, div, ta, rot if use_rot else None, rx)
                y_r = reference.ofdm_rx_slot(N, mu, nb_rb, slot, div, ta, rot224 if use_rot else None, rx)
                assert np.array_equal(y_o, y_r), (N, mu, nb_rb, slot, div, ta, use_rot)


def test_ue_slot_fep(oracle, reference):
    """The UE's OFDM front end nr_slot_fep (synchronised UE) is the gNB's with no timing offset and the DL rotation table: same oracle function."""
    rng = np.random.default_rng(22)
    for N, mu, nb_rb, slot in OFDM_CASES:
        _, _, _, frame_len = oracle.ofdm_geometry(N, mu, slot)
        rot = oracle.symbol_rotation(mu, 3619200000.0)
        rot224 = np.zeros(448, np.int16); rot224[:rot.size] = rot
        rx = rng.integers(-3000, 3001, size=(2, 2 * frame_len)).astype(np.int16)
        for div in (8, 4):
            ts = reference.rotation_tables(N, mu, nb_rb, div, 3619200000.0, 3619200000.0)[2]
            y_r = reference.ue_slot_fep(N, mu, nb_rb, 2, slot, div, rot224, ts, rx)
            for a in range(2):
                assert np.array_equal(oracle.ofdm_rx_slot(N, mu, nb_rb, slot, div, 0, rot, rx[a]), y_r[a]), (N, mu, nb_rb, slot, div, a)


# ------------------------------------------------------------------------------------------ single-layer PUSCH inner receiver (a23 + a25)
def _pusch_case(rng, N, nb_rx, rb_start, rb_size, Qm, dmrs_pos, dmrs_type, cdm, nb_rb_carrier, amp_y=2000, amp_h=1500):
    from oracle.bindings import PuschParms
    P = PuschParms(N, nb_rx, rb_start, 0, rb_size, N - nb_rb_carrier * 6, Qm, dmrs_pos, dmrs_type, cdm)
    rx = rng.integers(-amp_y, amp_y + 1, size=(nb_rx, 14, N, 2)).astype(np.int16)
    h = rng.integers(-amp_h, amp_h + 1, size=(nb_rx, 14, N, 2)).astype(np.int16)
    return P, rx, h


PUSCH_CASES = [  # N, nb_rx, rb_start, rb_size, Qm, dmrs_pos, dmrs_type, cdm_no_data, carrier PRBs
    (4096, 4, 0, 273, 6, 1 << 2, 0, 2, 273), (4096, 2, 0, 273, 8, 1 << 2, 0, 1, 273), (2048, 1, 10, 50, 4, (1 << 2) | (1 << 11), 0, 1, 106),
    (2048, 2, 30, 76, 2, 1 << 3, 0, 1, 106), (1024, 4, 0, 52, 6, 1 << 2, 1, 1, 52), (1024, 2, 20, 32, 4, 1 << 2, 1, 2, 52), (512, 8, 3, 11, 8, 1 << 0, 0, 1, 25),
    (4096, 4, 100, 173, 6, 1 << 2, 1, 1, 273),
]


def test_pusch_inner_rx(oracle, reference):
    rng = np.random.default_rng(40)
    for case in PUSCH_CASES:
        big = case[1] == 8
        P, rx, h = _pusch_case(rng, *case, amp_y=32767 if big else 2000, amp_h=32767 if big else 1500)
        dm = [s for s in range(14) if (P.ul_dmrs_symb_pos >> s) & 1]
        sh_o, avg_o = oracle.pusch_log2_maxh(P, dm[0] if oracle.pusch_nb_re(P, dm[0]) > 0 else dm[0] + 1, dm[0], rx, h)
        meas = dm[0] if oracle.pusch_nb_re(P, dm[0]) > 0 else dm[0] + 1
        sh_r, avg_r = reference.pusch_log2_maxh(P, meas, dm[0], rx, h)
        assert np.array_equal(avg_o, avg_r) and sh_o == sh_r, (case, avg_o, avg_r, sh_o, sh_r)
        for symbol in (dm[0], dm[0] + 1, 13):
            valid = oracle.pusch_nb_re(P, symbol)
            if valid == 0:
                continue
            for shift in {sh_o, 0 if big else max(0, sh_o - 3)}:
                llr_o, comp_o = oracle.pusch_inner_rx_symbol(P, symbol, dm[0], shift, rx, h)
                llr_r, comp_r = reference.pusch_inner_rx_symbol(P, symbol, dm[0], shift, rx, h, valid)
                assert np.array_equal(comp_o, comp_r), (case, symbol, shift, "comp")
                assert np.array_equal(llr_o, llr_r), (case, symbol, shift, "llr")


def test_pusch_inner_rx_fuzz(oracle, reference):
    """150 random single-layer PUSCH allocations / DMRS layouts on a 25-PRB carrier through the reference's inner_rx, three symbols each (a DMRS symbol, the one after it,
    the allocation's last), at the measured shift and a smaller one."""
    from oracle.bindings import PuschParms
    from common import ptrs_fuzz_cases
    rng = np.random.default_rng(87)
    for n, case in enumerate(ptrs_fuzz_cases(rng, 150, safe_tail=False)):
        N, nb_rx, rb_start, rb_size, Qm, dpos, dtype_, cdm, carrier, start, nsym = case[:11]
        ay, ah = ((2000, 1500), (600, 900), (32767, 32767))[n % 3]
        rx = rng.integers(-ay, ay + 1, size=(nb_rx, 14, N, 2)).astype(np.int16)
        h = rng.integers(-ah, ah + 1, size=(nb_rx, 14, N, 2)).astype(np.int16)
        if (dpos & (dpos << 1)) or ((dpos >> 13) & dpos & 1):
            continue                                   # get_nb_re_pusch AssertFatal()s on neighbouring DMRS symbols ("Double DMRS configuration is not yet supported")
        P = PuschParms(N, nb_rx, rb_start, 0, rb_size, N - carrier * 6, Qm, dpos, dtype_, cdm)
        dm = [s for s in range(start, start + nsym) if (dpos >> s) & 1]
        with_data = [s for s in range(start, start + nsym) if oracle.pusch_nb_re(P, s) > 0]
        if not with_data:
            continue
        meas = with_data[0]
        cur = dm[0] if meas < dm[0] else max(s for s in dm if s <= meas)
        sh_o, avg_o = oracle.pusch_log2_maxh(P, meas, cur, rx, h)
        sh_r, avg_r = reference.pusch_log2_maxh(P, meas, cur, rx, h)
        assert np.array_equal(avg_o, avg_r) and sh_o == sh_r, (case[:11], avg_o, avg_r, sh_o, sh_r)
        for symbol in sorted({dm[0], min(dm[0] + 1, start + nsym - 1), with_data[-1]}):
            valid = oracle.pusch_nb_re(P, symbol)
            if valid == 0:
                continue
            chs = dm[0] if symbol < dm[0] else max(s for s in dm if s <= symbol)
            for shift in {sh_o, max(0, sh_o - 3)}:
                llr_o, comp_o = oracle.pusch_inner_rx_symbol(P, symbol, chs, shift, rx, h)
                llr_r, comp_r = reference.pusch_inner_rx_symbol(P, symbol, chs, shift, rx, h, valid)
                assert np.array_equal(comp_o, comp_r) and np.array_equal(llr_o, llr_r), (case[:11], symbol, chs, shift)


# ------------------------------------------------------------------------------------------ PUSCH channel estimation (a22), DMRS type 1
CHEST_CASES = [  # N, nb_rx, slot, symbol, port, rb_start, rb_size, carrier PRBs, scid, dmrs id, delay (samples) of the synthetic channel
    (4096, 4, 1, 2, 0, 0, 273, 273, 0, 77, 0), (4096, 2, 8, 2, 0, 0, 273, 273, 1, 1007, 3), (2048, 2, 3, 3, 0, 10, 50, 106, 0, 5, -2),
    (2048, 1, 19, 11, 1, 30, 76, 106, 0, 65535, 1), (1024, 4, 0, 2, 2, 0, 52, 52, 1, 0, 7), (1024, 2, 5, 0, 3, 20, 32, 52, 0, 300, -30), (512, 8, 2, 2, 0, 3, 11, 25, 0, 9, 0),
]


def test_pusch_channel_estimation(oracle, reference):
    from oracle.bindings import ChestParms
    rng = np.random.default_rng(60)
    for N, nb_rx, slot, symbol, port, rb_start, rb_size, carrier, scid, nid, delay in CHEST_CASES:
        P = ChestParms(N, nb_rx, slot, symbol, port, rb_start, 0, rb_size, N - carrier * 6, scid, nid)
        big = nb_rx == 8
        # a channel with a linear phase (time delay) on the pilots + noise, so that the delay estimator has a peak to find
        pil = oracle.pusch_dmrs_pilots(P).reshape(-1, 2).astype(np.float64)
        rx = rng.integers(-300, 301, size=(nb_rx, 14, N, 2)).astype(np.int16) if not big else rng.integers(-32768, 32768, size=(nb_rx, 14, N, 2)).astype(np.int16)
        if not big:
            k0 = ((rb_start * 12) + P.first_carrier_offset) % N
            delta = (0, 0, 1, 1)[port]
            for a in range(nb_rx):
                h = (900 + 100 * a) * np.exp(1j * (0.3 * a - 2 * np.pi * delay * np.arange(6 * rb_size) * 2 / N))
                tx = (pil[:, 0] - 1j * pil[:, 1]) / 23170.0 / np.sqrt(2)          # transmitted DMRS = conj of the rx table
                y = h * tx
                idx = (k0 + 2 * np.arange(6 * rb_size) + delta) % N
                rx[a, symbol, idx, 0] += np.round(y.real).astype(np.int16); rx[a, symbol, idx, 1] += np.round(y.imag).astype(np.int16)
        est_r, out_r, pil_r = reference.pusch_channel_estimation(P, rx, carrier)
        assert np.array_equal(oracle.pusch_dmrs_pilots(P), pil_r), (N, slot, symbol, port, "pilots")
        est_o, out_o = oracle.pusch_channel_estimation(P, rx)
        assert np.array_equal(out_o, out_r), (N, nb_rx, slot, symbol, port, out_o, out_r)
        assert np.array_equal(est_o[:, symbol], est_r[:, symbol]), (N, nb_rx, slot, symbol, port)


def test_channel_estimation_fuzz(oracle, reference):
    """150 random DMRS type 1 configurations (antennas, slot, symbol, port, allocation, scrambling, channel delay; every third with full-scale noise) through the real
    nr_pusch_channel_estimation and -- up to 4 antennas -- the UE's nr_pdsch_channel_estimation."""
    from oracle.bindings import ChestParms
    from common import chest_fuzz_cases, chest_inputs
    rng = np.random.default_rng(91)
    for n, (N, nb_rx, slot, symbol, port, rb_start, rb_size, carrier, scid, nid, delay) in enumerate(chest_fuzz_cases(rng, 150)):
        P = ChestParms(N, nb_rx, slot, symbol, port, rb_start, 0, rb_size, N - carrier * 6, scid, nid)
        rx = chest_inputs(oracle, rng, P, port, delay, None if n % 3 == 2 else 300)
        est_r, out_r, _ = reference.pusch_channel_estimation(P, rx, carrier)
        est_o, out_o = oracle.pusch_channel_estimation(P, rx)
        assert np.array_equal(out_o, out_r) and np.array_equal(est_o[:, symbol], est_r[:, symbol]), ("gNB", N, nb_rx, slot, symbol, port, rb_start, rb_size, out_o, out_r)
        est_r = reference.pdsch_channel_estimation(P, rx, carrier)
        est_o = oracle.pdsch_channel_estimation(P, rx)
        assert np.array_equal(est_o[:, symbol], est_r[:, symbol]), ("UE", N, nb_rx, slot, symbol, port, rb_start, rb_size)


CHEST_VARIANT_CASES = [  # N, nb_rx, slot, symbol, port, rb_start, rb_size, carrier PRBs, scid, dmrs id, delay
    (4096, 4, 4, 2, 0, 0, 273, 273, 0, 77, 2), (2048, 2, 8, 3, 1, 10, 50, 106, 1, 1007, -3), (1024, 3, 0, 11, 2, 20, 32, 52, 0, 300, 5),
    (1024, 2, 12, 2, 3, 0, 52, 52, 1, 0, 0), (512, 8, 16, 5, 0, 3, 11, 25, 0, 9, 1), (2048, 1, 4, 0, 0, 30, 2, 106, 0, 65535, 0),
]


@pytest.mark.parametrize("dmrs_type,chest_freq", [(1, 0), (0, 1), (1, 1)])
def test_pusch_channel_estimation_variants(oracle, reference, dmrs_type, chest_freq):
    """DMRS type 2 with frequency-domain interpolation, and the one-average-per-PRB estimators (chest_freq = 1) of both DMRS types
    (nr_ul_channel_estimation.c:258-460), against the compiled reference.  Type 2 + chest_freq = 1 reads slot-ring position 0 (a missing
    `soffset`), so those cases use slots that are multiples of 4."""
    from oracle.bindings import ChestParms
    rng = np.random.default_rng(61 + 2 * dmrs_type + chest_freq)
    for N, nb_rx, slot, symbol, port, rb_start, rb_size, carrier, scid, nid, delay in CHEST_VARIANT_CASES:
        P = ChestParms(N, nb_rx, slot, symbol, port, rb_start, 0, rb_size, N - carrier * 6, scid, nid, dmrs_type, chest_freq)
        big = nb_rx == 8
        rx = rng.integers(-300, 301, size=(nb_rx, 14, N, 2)).astype(np.int16) if not big else rng.integers(-32768, 32768, size=(nb_rx, 14, N, 2)).astype(np.int16)
        if not big:          # a delayed flat channel on the pilot REs so that the estimators see something coherent
            pil = oracle.pusch_dmrs_pilots(P).reshape(-1, 2).astype(np.float64)
            k0 = ((rb_start * 12) + P.first_carrier_offset) % N
            npil = pil.shape[0]
            if dmrs_type == 0:
                idx = (k0 + 2 * np.arange(npil)) % N + ((port >> 1) & 1)
            else:
                idx = (k0 + 6 * (np.arange(npil) // 2) + (np.arange(npil) & 1)) % N + ((port >> 1) & 1)
            for a in range(nb_rx):
                h = (2000 + 300 * a) * np.exp(1j * (0.4 * a - 2 * np.pi * delay * (idx - k0) / N))
                y = h * (pil[:, 0] - 1j * pil[:, 1]) / 23170.0 / np.sqrt(2)
                rx[a, symbol, idx, 0] += np.round(y.real).astype(np.int16); rx[a, symbol, idx, 1] += np.round(y.imag).astype(np.int16)
        est_r, out_r, pil_r = reference.pusch_channel_estimation(P, rx, carrier, chest_freq=chest_freq, dmrs_type=dmrs_type)
        npil = (4 if dmrs_type else 6) * rb_size
        assert np.array_equal(oracle.pusch_dmrs_pilots(P)[:2 * npil], pil_r[:2 * npil]), (N, slot, symbol, port, "pilots")
        est_o, out_o = oracle.pusch_channel_estimation(P, rx)
        assert np.array_equal(out_o, out_r), (N, nb_rx, slot, symbol, port, out_o, out_r)
        assert np.array_equal(est_o[:, symbol], est_r[:, symbol]), (N, nb_rx, slot, symbol, port)


def test_chest_time_domain_avg(oracle, reference):
    """nr_chest_time_domain_avg (dmrs_nr.c:343-417): 1-4 DMRS symbols, saturating sums, the three division rules, full-scale inputs."""
    rng = np.random.default_rng(77)
    for N, nb_rx, start, nsym, bitmap, nrb in ((512, 2, 0, 14, 0b00000000000100, 25), (512, 3, 0, 14, 0b00100000000100, 20), (1024, 2, 2, 12, 0b00101000001000, 52),
                                               (1024, 4, 0, 14, 0b00100100100100, 40), (2048, 1, 1, 10, 0b00000010000100, 106), (512, 2, 4, 8, 0b11000000110000, 11)):
        est = rng.integers(-32768, 32768, size=(nb_rx, 14, N, 2)).astype(np.int16)
        est[:, :, ::7] //= 64                                  # a share of small values so that not every sum saturates
        r = reference.chest_time_domain_avg(est, nsym, start, bitmap, nrb)
        o = oracle.chest_time_domain_avg(est, nsym, start, bitmap, nrb)
        assert np.array_equal(o, r), (N, nb_rx, start, nsym, bin(bitmap), nrb)
        assert not np.array_equal(o, est) or bin(bitmap & ((1 << (start + nsym)) - 1)).count("1") == 1


def test_pusch_inner_rx_two_layers_mmse(oracle, reference):
    """nb_layer == 2, Qm >= 6: matched filter per layer + nr_ulsch_mmse_2layers + per-layer LLRs, through the reference's inner_rx."""
    from oracle.bindings import PuschParms
    rng = np.random.default_rng(41)
    for N, nb_rx, rb_start, rb_size, Qm, carrier, nvar in ((4096, 4, 0, 273, 6, 273, 40), (2048, 2, 10, 50, 8, 106, 7), (1024, 4, 20, 32, 6, 52, 0), (2048, 2, 30, 76, 8, 106, 1000),
                                                           (1024, 2, 0, 52, 6, 52, 90000)):
        P = PuschParms(N, nb_rx, rb_start, 0, rb_size, N - carrier * 6, Qm, 1 << 2, 0, 2)
        rx = rng.integers(-2000, 2001, size=(nb_rx, 14, N, 2)).astype(np.int16)
        h = rng.integers(-1500, 1501, size=(2 * nb_rx, 14, N, 2)).astype(np.int16)
        for max_ch in (0, 1500, 30000, 200000):
            sh_o, avg_o = oracle.pusch_log2_maxh_2l(P, 0, 2, max_ch, rx, h)
            sh_r, avg_r = reference.pusch_log2_maxh(P, 0, 2, rx, h, nb_layer=2, max_ch=max_ch)
            assert np.array_equal(avg_o, avg_r) and sh_o == sh_r, (N, nb_rx, max_ch, avg_o, avg_r, sh_o, sh_r)
        for symbol, shift in ((3, 8), (13, 6), (0, 11)):
            valid = oracle.pusch_nb_re(P, symbol)
            if nvar == 0:
                nvar = 1                     # the reference AssertFatal()s on a zero determinant (zero-padded REs) unless noise is added
            llr_o, comp_o = oracle.pusch_inner_rx_symbol_2l(P, symbol, 2, shift, nvar, rx, h)
            llr_r, comp_r = reference.pusch_inner_rx_symbol(P, symbol, 2, shift, rx, h, valid, nb_layer=2, nvar=nvar)
            assert np.array_equal(comp_o, comp_r), (N, nb_rx, Qm, symbol, shift, nvar, "comp")
            assert np.array_equal(llr_o, llr_r), (N, nb_rx, Qm, symbol, shift, nvar, "llr")


def test_pusch_inner_rx_two_layers_ml(oracle, reference):
    """nb_layer == 2, Qm < 6: matched filter per layer, rho and magnitudes, joint max-log ML LLRs (nr_ulsch_qpsk_qpsk / nr_ulsch_qam16_qam16), through inner_rx."""
    from oracle.bindings import PuschParms
    rng = np.random.default_rng(43)
    for N, nb_rx, rb_start, rb_size, Qm, carrier, amp in ((4096, 4, 0, 273, 4, 273, 1500), (2048, 2, 10, 50, 2, 106, 1500), (1024, 4, 20, 32, 4, 52, 6000), (2048, 1, 30, 77, 2, 106, 4000),
                                                          (1024, 2, 0, 51, 4, 52, 300), (1024, 3, 3, 9, 2, 52, 32767)):
        P = PuschParms(N, nb_rx, rb_start, 0, rb_size, N - carrier * 6, Qm, 1 << 2, 0, 2)
        rx = rng.integers(-2000, 2001, size=(nb_rx, 14, N, 2)).astype(np.int16)
        h = rng.integers(-amp, amp + 1, size=(2 * nb_rx, 14, N, 2)).astype(np.int16)
        for max_ch in (0, 1500, 200000):
            sh_o, avg_o = oracle.pusch_log2_maxh_2l(P, 0, 2, max_ch, rx, h)
            sh_r, avg_r = reference.pusch_log2_maxh(P, 0, 2, rx, h, nb_layer=2, max_ch=max_ch)
            assert np.array_equal(avg_o, avg_r) and sh_o == sh_r, (N, nb_rx, max_ch, avg_o, avg_r, sh_o, sh_r)
        for symbol, shift in ((3, 8), (13, 5), (0, 11), (2, 7)):
            valid = oracle.pusch_nb_re(P, symbol)
            if valid == 0:
                continue
            llr_o, comp_o = oracle.pusch_inner_rx_symbol_2l(P, symbol, 2, shift, 0, rx, h)
            llr_r, comp_r = reference.pusch_inner_rx_symbol(P, symbol, 2, shift, rx, h, valid, nb_layer=2, nvar=0)
            assert np.array_equal(comp_o, comp_r), (N, nb_rx, Qm, symbol, shift, "comp")
            assert np.array_equal(llr_o, llr_r), (N, nb_rx, Qm, symbol, shift, "llr", int((llr_o != llr_r).sum()))


def test_pdsch_channel_estimation_ue(oracle, reference):
    """UE-side estimator (nr_pdsch_channel_estimation, DMRS type 1 linear interpolation) on the same cases as the gNB one."""
    from oracle.bindings import ChestParms
    rng = np.random.default_rng(63)
    for N, nb_rx, slot, symbol, port, rb_start, rb_size, carrier, scid, nid, delay in CHEST_CASES:
        if nb_rx > 4:
            nb_rx = 4                      # NB_ANTENNAS_RX of the reference build
        P = ChestParms(N, nb_rx, slot, symbol, port, rb_start, 0, rb_size, N - carrier * 6, scid, nid)
        pil = oracle.pusch_dmrs_pilots(P).reshape(-1, 2).astype(np.float64)
        rx = rng.integers(-300, 301, size=(nb_rx, 14, N, 2)).astype(np.int16) if symbol != 2 or N != 512 else rng.integers(-32768, 32768, size=(nb_rx, 14, N, 2)).astype(np.int16)
        k0 = ((rb_start * 12) + P.first_carrier_offset) % N
        for a in range(nb_rx):
            h = (900 + 100 * a) * np.exp(1j * (0.3 * a - 2 * np.pi * delay * np.arange(6 * rb_size) * 2 / N))
            y = h * (pil[:, 0] - 1j * pil[:, 1]) / 32767.0
            idx = (k0 + 2 * np.arange(6 * rb_size)) % N + ((port >> 1) & 1)
            if N != 512:
                rx[a, symbol].reshape(-1, 2)[idx, 0] += np.round(y.real).astype(np.int16); rx[a, symbol].reshape(-1, 2)[idx, 1] += np.round(y.imag).astype(np.int16)
        est_r = reference.pdsch_channel_estimation(P, rx, carrier)
        est_o = oracle.pdsch_channel_estimation(P, rx)
        assert np.array_equal(est_o[:, symbol], est_r[:, symbol]), (N, nb_rx, slot, symbol, port)


@pytest.mark.parametrize("dmrs_type,chest_freq", [(1, 0), (0, 1), (1, 1)])
def test_pdsch_channel_estimation_ue_variants(oracle, reference, dmrs_type, chest_freq):
    """UE estimator, DMRS type 2 (NFAPI_NR_DMRS_TYPE2_linear_interp) and the per-PRB averages of both types, against the real nr_pdsch_channel_estimation.
    Type 2 ports 0-5 (pointer shift 0 / 2 / 4)."""
    from oracle.bindings import ChestParms
    rng = np.random.default_rng(65 + 2 * dmrs_type + chest_freq)
    cases = [(4096, 2, 4, 2, 0, 0, 273, 273, 0, 77), (2048, 2, 8, 3, 1, 10, 50, 106, 1, 1007), (1024, 3, 0, 11, 2, 20, 32, 52, 0, 300), (1024, 2, 12, 2, 3, 0, 52, 52, 1, 0),
             (512, 4, 16, 5, 0, 3, 11, 25, 0, 9), (2048, 1, 5, 0, 1, 30, 2, 106, 0, 65535)]
    if dmrs_type == 1:
        cases += [(1024, 2, 3, 4, 4, 0, 52, 52, 0, 21), (1024, 2, 7, 6, 5, 8, 30, 52, 1, 22)]
    for N, nb_rx, slot, symbol, port, rb_start, rb_size, carrier, scid, nid in cases:
        P = ChestParms(N, nb_rx, slot, symbol, port, rb_start, 0, rb_size, N - carrier * 6, scid, nid, dmrs_type, chest_freq)
        rx = rng.integers(-32768, 32768, size=(nb_rx, 14, N, 2)).astype(np.int16) if N == 512 else rng.integers(-3000, 3001, size=(nb_rx, 14, N, 2)).astype(np.int16)
        est_r = reference.pdsch_channel_estimation(P, rx, carrier, chest_freq=chest_freq, dmrs_type=dmrs_type)
        est_o = oracle.pdsch_channel_estimation(P, rx)
        assert np.array_equal(est_o[:, symbol], est_r[:, symbol]), (N, nb_rx, slot, symbol, port, dmrs_type, chest_freq)


PDSCH_CASES = [  # N, nb_rx, rb_start, rb_size, Qm, dmrs_pos, dmrs_type, cdm groups, carrier PRBs, start_symbol, nr_symbols
    (4096, 2, 0, 273, 6, 1 << 2, 0, 2, 273, 1, 13), (4096, 4, 0, 273, 8, 1 << 2, 0, 1, 273, 1, 13), (2048, 1, 10, 50, 4, (1 << 2) | (1 << 11), 0, 1, 106, 1, 13),
    (2048, 2, 30, 76, 2, 1 << 3, 0, 2, 106, 2, 10), (1024, 4, 0, 52, 6, 1 << 2, 1, 1, 52, 1, 13), (1024, 2, 20, 32, 4, 1 << 2, 1, 2, 52, 2, 12), (512, 4, 3, 11, 8, 1 << 1, 0, 1, 25, 1, 6),
]


def test_pdsch_rx_slot_ue(oracle, reference):
    """UE-side PDSCH receiver, one layer: the reference's own nr_rx_pdsch symbol loop vs the oracle restatement."""
    from oracle.bindings import PuschParms
    rng = np.random.default_rng(65)
    for N, nb_rx, rb_start, rb_size, Qm, dpos, dtype_, cdm, carrier, start, nsym in PDSCH_CASES:
        big = N == 512
        ay, ah = (32767, 32767) if big else (2000, 1500)
        rx = rng.integers(-ay, ay + 1, size=(nb_rx, 14, N, 2)).astype(np.int16)
        h = rng.integers(-ah, ah + 1, size=(nb_rx, 14, N, 2)).astype(np.int16)
        P = PuschParms(N, nb_rx, rb_start, 0, rb_size, N - carrier * 6, Qm, dpos, dtype_, cdm)
        llr_o, sh_o = oracle.pdsch_rx_slot(P, start, nsym, rx, h)
        llr_r, sh_r, valid = reference.pdsch_rx_slot(P, start, nsym, rx, h, llr_o.size)
        assert sh_o == sh_r, (N, nb_rx, Qm, sh_o, sh_r)
        assert np.array_equal(llr_o, llr_r), (N, nb_rx, rb_start, rb_size, Qm, dpos, dtype_, cdm, np.nonzero(llr_o != llr_r)[0][:5])


PDSCH_NL_CASES = [  # N, nb_rx, rb_start, rb_size, Qm, dmrs_pos, dmrs_type, cdm groups, carrier PRBs, start_symbol, nr_symbols, layers, amplitude (rx, h)
    (4096, 4, 0, 273, 6, 1 << 2, 0, 2, 273, 1, 13, 4, (2000, 1500)), (4096, 4, 0, 273, 8, 1 << 2, 0, 2, 273, 1, 13, 3, (900, 700)),
    (2048, 4, 10, 50, 4, (1 << 2) | (1 << 11), 0, 1, 106, 1, 13, 3, (4000, 6000)), (2048, 3, 30, 76, 2, 1 << 3, 0, 2, 106, 2, 10, 3, (300, 200)),
    (1024, 4, 0, 52, 6, 1 << 2, 1, 1, 52, 1, 13, 4, (2000, 1500)), (1024, 4, 20, 32, 4, 1 << 2, 1, 2, 52, 2, 12, 4, (12000, 9000)),
    (512, 4, 3, 11, 8, 1 << 1, 0, 1, 25, 1, 6, 4, (32767, 32767)), (512, 4, 0, 25, 6, 1 << 2, 0, 1, 25, 0, 14, 3, (60, 40)),
    (2048, 4, 0, 106, 6, (1 << 2) | (1 << 13), 0, 1, 106, 1, 13, 4, (2000, 1500)), (1024, 2, 0, 52, 6, 1 << 2, 0, 2, 52, 1, 13, 3, (2000, 1500)),
]


def test_pdsch_rx_slot_ue_3_4_layers(oracle, reference):
    """UE-side PDSCH receiver with three and four layers: nr_rx_pdsch's generic n_tx code (per-layer MRC, nr_zero_forcing_rx with the recursive nr_determin /
    nr_matrix_inverse in fixed point, determinant thresholds, nr_dlsch_layer_demapping) vs the oracle; incl. fewer rx antennas than layers (singular Gram matrix:
    the arithmetic is still defined) and full-scale inputs."""
    from oracle.bindings import PuschParms
    rng = np.random.default_rng(75)
    for N, nb_rx, rb_start, rb_size, Qm, dpos, dtype_, cdm, carrier, start, nsym, nl, (ay, ah) in PDSCH_NL_CASES:
        rx = rng.integers(-ay, ay + 1, size=(nb_rx, 14, N, 2)).astype(np.int16)
        h = rng.integers(-ah, ah + 1, size=(nl * nb_rx, 14, N, 2)).astype(np.int16)
        P = PuschParms(N, nb_rx, rb_start, 0, rb_size, N - carrier * 6, Qm, dpos, dtype_, cdm)
        llr_o, sh_o = oracle.pdsch_rx_slot(P, start, nsym, rx, h, nl=nl)
        llr_r, sh_r, valid = reference.pdsch_rx_slot(P, start, nsym, rx, h, llr_o.size, nl=nl)
        assert sh_o == sh_r, (N, nb_rx, Qm, nl, sh_o, sh_r)
        assert np.array_equal(llr_o, llr_r), (N, nb_rx, rb_start, rb_size, Qm, nl, dpos, dtype_, cdm, np.nonzero(llr_o != llr_r)[0][:5])


def test_pdsch_rx_slot_ue_ptrs(oracle, reference):
    """PT-RS at the UE: the real nr_rx_pdsch + nr_pdsch_ptrs_processing + ptrs_nr.c (oracle/_ref/libref_pdsch_ptrs.so) vs the oracle restatement -- LLRs of the slot,
    log2_maxh, the per-symbol phase estimates (incl. interpolated ones) and PT-RS RE counts; random full-scale inputs and coherent slots with a phase ramp."""
    from oracle.bindings import PuschParms, PtrsParms
    from common import PTRS_CASES, PTRS_SIGNALS, ptrs_inputs
    rng = np.random.default_rng(71)
    for case in PTRS_CASES:
        N, nb_rx, rb_start, rb_size, Qm, dpos, dtype_, cdm, carrier, start, nsym, L, K, reoff, rnti, slot, nscid, nid = case
        for kind, a, b in PTRS_SIGNALS:
            rx, h = ptrs_inputs(oracle, rng, case, kind, a, b)
            P = PuschParms(N, nb_rx, rb_start, 0, rb_size, N - carrier * 6, Qm, dpos, dtype_, cdm)
            T = PtrsParms(1, L, K, reoff, rnti, slot, nscid, nid)
            llr_o, sh_o, ph_o, nre_o = oracle.pdsch_rx_slot_ptrs(P, T, start, nsym, rx, h)
            llr_r, sh_r, valid, ph_r, nre_r = reference.pdsch_rx_slot_ptrs(P, T, start, nsym, rx, h, llr_o.size, n_rb_dl=carrier)
            assert sh_o == sh_r and np.array_equal(nre_o, nre_r), (case, kind, sh_o, sh_r, nre_o, nre_r)
            assert np.array_equal(ph_o, ph_r), (case, kind, ph_o.tolist(), ph_r.tolist())
            assert np.array_equal(llr_o, llr_r), (case, kind, a, b, np.nonzero(llr_o != llr_r)[0][:5])
            assert int(valid.sum()) * Qm == llr_o.size


def test_pdsch_rx_slot_ue_ptrs_fuzz(oracle, reference):
    """300 random PT-RS configurations on a 25-PRB carrier (allocation, 1-3 DMRS symbols anywhere in it, both DMRS types, densities, offsets, RNTI, slot): every branch
    of set_ptrs_symb_idx / nr_ptrs_process_slot (left and right extrapolation, reused slopes, DMRS as the first symbol) against the real functions."""
    from oracle.bindings import PuschParms, PtrsParms
    from common import ptrs_fuzz_cases, ptrs_inputs
    rng = np.random.default_rng(83)
    for n, case in enumerate(ptrs_fuzz_cases(rng, 300)):
        N, nb_rx, rb_start, rb_size, Qm, dpos, dtype_, cdm, carrier, start, nsym, L, K, reoff, rnti, slot, nscid, nid = case
        kind, a, b = (("random", 2000, 1500), ("coherent", 30, 0.05), ("coherent", 0, 0.2))[n % 3]
        rx, h = ptrs_inputs(oracle, rng, case, kind, a, b)
        P = PuschParms(N, nb_rx, rb_start, 0, rb_size, N - carrier * 6, Qm, dpos, dtype_, cdm)
        T = PtrsParms(1, L, K, reoff, rnti, slot, nscid, nid)
        llr_o, sh_o, ph_o, nre_o = oracle.pdsch_rx_slot_ptrs(P, T, start, nsym, rx, h)
        llr_r, sh_r, valid, ph_r, nre_r = reference.pdsch_rx_slot_ptrs(P, T, start, nsym, rx, h, llr_o.size, n_rb_dl=carrier)
        assert sh_o == sh_r and np.array_equal(nre_o, nre_r) and np.array_equal(ph_o, ph_r), (case, kind, ph_o.tolist(), ph_r.tolist())
        assert np.array_equal(llr_o, llr_r) and int(valid.sum()) * Qm == llr_o.size, (case, kind, np.nonzero(llr_o != llr_r)[0][:5])


def test_pdsch_rx_slot_ue_layers_fuzz(oracle, reference):
    """200 random allocations / DMRS layouts on a 25-PRB carrier through the real nr_rx_pdsch with 1 ... 4 layers (no PT-RS): incl. a last symbol that is a DMRS symbol
    with data whose RE count is not a whole number of PRBs (type 2: 8 or 4 data REs per PRB), where the thresholds exist for the extracted REs only."""
    from oracle.bindings import PuschParms
    from common import ptrs_fuzz_cases
    rng = np.random.default_rng(85)
    for n, case in enumerate(ptrs_fuzz_cases(rng, 200)):
        N, nb_rx, rb_start, rb_size, Qm, dpos, dtype_, cdm, carrier, start, nsym = case[:11]
        nl = 1 + n % 4
        nb_rx = max(nb_rx, 2) if nl > 1 else nb_rx
        ay, ah = ((2000, 1500), (600, 900), (32767, 32767))[n % 3]
        rx = rng.integers(-ay, ay + 1, size=(nb_rx, 14, N, 2)).astype(np.int16)
        h = rng.integers(-ah, ah + 1, size=(nl * nb_rx, 14, N, 2)).astype(np.int16)
        P = PuschParms(N, nb_rx, rb_start, 0, rb_size, N - carrier * 6, Qm, dpos, dtype_, cdm)
        llr_o, sh_o = oracle.pdsch_rx_slot(P, start, nsym, rx, h, nl=nl)
        if llr_o.size == 0:
            continue
        llr_r, sh_r, valid = reference.pdsch_rx_slot(P, start, nsym, rx, h, llr_o.size, nl=nl)
        assert sh_o == sh_r and np.array_equal(llr_o, llr_r), (case[:11], nl, sh_o, sh_r, np.nonzero(llr_o != llr_r)[0][:5])


PDSCH_2L_CASES = [  # N, nb_rx, rb_start, rb_size, Qm, dmrs_pos, dmrs_type, cdm groups, carrier PRBs, start_symbol, nr_symbols, amplitude (rx, h)
    (4096, 2, 0, 273, 6, 1 << 2, 0, 1, 273, 1, 13, (2000, 1500)), (4096, 4, 0, 273, 8, 1 << 2, 0, 2, 273, 1, 13, (900, 700)),
    (2048, 2, 10, 50, 4, (1 << 2) | (1 << 11), 0, 1, 106, 1, 13, (4000, 6000)), (2048, 2, 30, 76, 2, 1 << 3, 0, 2, 106, 2, 10, (300, 200)),
    (1024, 4, 0, 52, 6, 1 << 2, 1, 1, 52, 1, 13, (2000, 1500)), (1024, 2, 20, 32, 4, 1 << 2, 1, 2, 52, 2, 12, (12000, 9000)),
    (512, 4, 3, 11, 8, 1 << 1, 0, 1, 25, 1, 6, (32767, 32767)), (512, 2, 0, 25, 6, 1 << 2, 0, 1, 25, 0, 14, (60, 40)),
]


def test_pdsch_rx_slot_ue_2layers(oracle, reference):
    """UE-side PDSCH receiver, two layers (MRC per layer + nr_zero_forcing_rx + layer de-mapping): nr_rx_pdsch symbol loop vs the oracle."""
    from oracle.bindings import PuschParms
    rng = np.random.default_rng(66)
    for N, nb_rx, rb_start, rb_size, Qm, dpos, dtype_, cdm, carrier, start, nsym, (ay, ah) in PDSCH_2L_CASES:
        rx = rng.integers(-ay, ay + 1, size=(nb_rx, 14, N, 2)).astype(np.int16)
        h = rng.integers(-ah, ah + 1, size=(2 * nb_rx, 14, N, 2)).astype(np.int16)
        P = PuschParms(N, nb_rx, rb_start, 0, rb_size, N - carrier * 6, Qm, dpos, dtype_, cdm)
        llr_o, sh_o = oracle.pdsch_rx_slot(P, start, nsym, rx, h, nl=2)
        llr_r, sh_r, valid = reference.pdsch_rx_slot(P, start, nsym, rx, h, llr_o.size, nl=2)
        assert sh_o == sh_r, (N, nb_rx, Qm, sh_o, sh_r)
        assert np.array_equal(llr_o, llr_r), (N, nb_rx, rb_start, rb_size, Qm, dpos, dtype_, cdm, np.nonzero(llr_o != llr_r)[0][:5])


PDSCH_TX_CASES = [  # N, carrier PRBs, nb_tx, slot, rb_start, rb_size, Qm, layers, start_symbol, nr_symbols, dmrs_pos, dmrs_type, cdm groups, dmrs_ports, scid, amp
    (4096, 273, 2, 1, 0, 273, 6, 2, 1, 13, 1 << 2, 0, 1, 0b0011, 0, 512), (4096, 273, 4, 7, 0, 273, 8, 1, 1, 13, 1 << 2, 0, 2, 0b0001, 0, 512),
    (2048, 106, 2, 3, 10, 50, 4, 2, (1), 13, (1 << 2) | (1 << 11), 0, 2, 0b1100, 1, 700), (2048, 106, 4, 19, 30, 76, 2, 4, 2, 10, 1 << 3, 0, 2, 0b1111, 0, 1000),
    (1024, 52, 2, 5, 0, 52, 6, 2, 1, 13, 1 << 2, 1, 1, 0b000011, 0, 512), (2048, 106, 4, 0, 20, 31, 4, 3, 2, 12, 1 << 2, 1, 2, 0b001101, 1, 300),
    (512, 25, 1, 9, 3, 11, 8, 1, 1, 6, 1 << 1, 0, 1, 0, 0, 512), (512, 25, 2, 11, 0, 25, 6, 2, 0, 14, (1 << 2) | (1 << 3), 0, 2, 0b0101, 0, 2047),
    (1536, 79, 2, 2, 0, 79, 6, 2, 1, 13, (1 << 2) | (1 << 7) | (1 << 11), 1, 3, 0b110000, 0, 512),
]


def test_pdsch_tx_slot(oracle, reference):
    """gNB PDSCH transmitter after the encoder: the reference's nr_generate_pdsch (scrambling ... txdataF) vs the oracle restatement."""
    from oracle.bindings import PdschTxParms
    rng = np.random.default_rng(70)
    for N, carrier, ntx, slot, rb0, nrb, Qm, nl, s0, ns, dpos, dtype_, cdm, ports, scid, amp in PDSCH_TX_CASES:
        P = PdschTxParms(N, ntx, slot, rb0, 0, nrb, N - carrier * 6, Qm, nl, s0, ns, dpos, dtype_, cdm, ports, scid, 40 + slot, 501, 0x1234, amp)
        bits = rng.integers(0, 2, size=P.G(), dtype=np.uint8)
        t_o = oracle.pdsch_tx_slot(P, bits)
        t_r = reference.pdsch_tx_slot(P, bits, carrier)
        assert np.array_equal(t_o, t_r), (N, nrb, Qm, nl, dpos, dtype_, cdm, ports, [tuple(x) for x in np.argwhere(t_o != t_r)[:5]])
        assert np.count_nonzero(t_o) > 0


def test_pdsch_tx_slot_wideband_precoding(oracle, reference):
    """Non-identity precoding (pm_idx > 0, one PRG over the allocation): nr_layer_precoder_simd for RB pairs that end below the symbol's last sub-carrier
    (saturating accumulation over the layers), nr_layer_precoder_cm for the others (wrapping).  Weights near full scale provoke both behaviours; odd and even
    rb_size, allocations that wrap around DC and ones that end exactly at the symbol's last sub-carrier."""
    from oracle.bindings import PdschTxParms
    rng = np.random.default_rng(71)
    cases = [  # N, carrier, ntx, slot, rb0, nrb, Qm, nl
        (4096, 273, 4, 1, 0, 273, 6, 2), (4096, 273, 2, 3, 0, 272, 8, 2), (2048, 106, 4, 5, 10, 51, 4, 1), (1024, 52, 4, 0, 3, 40, 6, 4), (1024, 52, 2, 7, 0, 26, 2, 2),
        (512, 25, 4, 2, 2, 21, 6, 3), (512, 25, 4, 2, 0, 12, 6, 2),
    ]
    for N, carrier, ntx, slot, rb0, nrb, Qm, nl in cases:
        P = PdschTxParms(N, ntx, slot, rb0, 0, nrb, N - carrier * 6, Qm, nl, 1, 13, 1 << 2, 0, 2, (1 << nl) - 1, 0, 40 + slot, 501, 0x1234, 512 if Qm < 8 else 30000)
        w = rng.integers(-32767, 32768, size=(4, 4, 2)).astype(np.int16)
        if Qm < 8:
            w //= 3
        P.set_precoding(1 + (slot % 3), w)
        bits = rng.integers(0, 2, size=P.G(), dtype=np.uint8)
        t_o = oracle.pdsch_tx_slot(P, bits)
        t_r = reference.pdsch_tx_slot(P, bits, carrier)
        assert np.array_equal(t_o, t_r), (N, nrb, Qm, nl, ntx, [tuple(x) for x in np.argwhere(t_o != t_r)[:5]])
        assert np.count_nonzero(t_o[ntx - 1]) > 0                       # every antenna radiates


PDSCH_TX_PTRS = [(0, 2, 0), (1, 4, 2), (2, 2, 5), (1, 2, 11), (2, 4, 1), (0, 4, 0), (1, 2, 3), (2, 2, 0), (1, 4, 7)]     # per PDSCH_TX_CASES entry: L (log2), K, PTRSReOffset


def test_pdsch_tx_slot_ptrs(oracle, reference):
    """PT-RS insertion in nr_generate_pdsch (pduBitmap & 1: nr_dlsch.c:98-111, :287-352): PT-RS symbols take the per-RE mapping branch (truncating scaling),
    every layer carries the pilots, the data skip them and the encoder's length shrinks by unav_res; with and without wideband precoding."""
    from oracle.bindings import PdschTxParms
    rng = np.random.default_rng(73)
    for (N, carrier, ntx, slot, rb0, nrb, Qm, nl, s0, ns, dpos, dtype_, cdm, ports, scid, amp), (L, K, reoff) in zip(PDSCH_TX_CASES, PDSCH_TX_PTRS):
        for pm in (0, 1):
            if pm and ntx < 2:
                continue
            P = PdschTxParms(N, ntx, slot, rb0, 0, nrb, N - carrier * 6, Qm, nl, s0, ns, dpos, dtype_, cdm, ports, scid, 40 + slot, 501, 0x1234 + slot, amp)
            P.set_ptrs(L, K, reoff)
            if pm:
                P.set_precoding(2, rng.integers(-12000, 12001, size=(4, 4, 2)).astype(np.int16))
            bits = rng.integers(0, 2, size=P.G(), dtype=np.uint8)
            t_o = oracle.pdsch_tx_slot(P, bits)
            t_r = reference.pdsch_tx_slot(P, bits, carrier)
            assert np.array_equal(t_o, t_r), (N, nrb, Qm, nl, dpos, L, K, reoff, pm, [tuple(x) for x in np.argwhere(t_o != t_r)[:5]])
            P.ptrs_on = 0
            assert P.G() > bits.size


def test_rfsim_rx_add_input(oracle, reference):
    """rfsimulator channel application: the real rxAddInput (apply_channelmod.c) with the harness's noise draws vs the oracle restatement, every rx
    antenna, with and without noise; 1-4 antennas either side, 1-200 taps, wrap-around of the circular buffer, full-scale samples, a time stamp above 2^32."""
    from common import RFSIM_CASES, rfsim_inputs
    rng = np.random.default_rng(77)
    for case in RFSIM_CASES:
        nb_tx, nb_rx, L, offset, pl, npw, n, TS, cf, amp = case
        cir, ch, sig, out, noise = rfsim_inputs(rng, case)
        for a in range(nb_rx):
            for nz in (noise, None):
                o = oracle.rfsim_rx_add_input(nb_tx, nb_rx, L, offset, pl, npw, ch, sig, out, a, TS, cir, nz)
                r = reference.rfsim_rx_add_input(nb_tx, nb_rx, L, offset, pl, npw, ch, sig, out, a, TS, cir, nz)
                assert np.array_equal(o, r), (case, a, nz is None, np.argwhere(o != r)[:5])
                assert not np.array_equal(o, out)


def test_rfsim_rx_add_input_fuzz(oracle, reference):
    """120 random rxAddInput calls: 1-4 antennas either side, 1-255 taps, offsets of either sign, path loss / noise power, buffer sizes and time stamps (incl. above 2^32)."""
    rng = np.random.default_rng(95)
    for n in range(120):
        nb_tx, nb_rx, L = int(rng.integers(1, 5)), int(rng.integers(1, 5)), int(rng.choice([1, 2, 7, 33, 100, 255]))
        ns, offset = int(rng.integers(1, 1500)), int(rng.integers(-6, 7))
        TS = int(rng.integers(L + 8, 1 << 20)) + (int(rng.integers(1, 4)) << 32 if n % 5 == 0 else 0)
        cir = int(rng.integers(ns + L + 16, 3 * (ns + L + 16)))
        ch = rng.normal(size=(nb_tx * nb_rx, L, 2)) * (0.7 / np.sqrt(L))
        sig = rng.integers(-32768, 32768, size=(cir, 2)).astype(np.int16)
        out = rng.integers(-30000, 30001, size=(ns, 2)).astype(np.int16)
        noise = rng.normal(size=(ns, 2)) if n % 2 else None
        pl, npw, a = float(rng.uniform(-20, 6)), float(rng.uniform(-40, 3)), int(rng.integers(0, nb_rx))
        o = oracle.rfsim_rx_add_input(nb_tx, nb_rx, L, offset, pl, npw, ch, sig, out, a, TS, cir, noise)
        r = reference.rfsim_rx_add_input(nb_tx, nb_rx, L, offset, pl, npw, ch, sig, out, a, TS, cir, noise)
        assert np.array_equal(o, r), (nb_tx, nb_rx, L, ns, offset, TS, cir, a, np.argwhere(o != r)[:5])


def test_db_fixed_times10(oracle, reference):
    """dB_fixed_times10 (TOOLS/dB_routines.c:132-155) on the generated table floor(100 log10 n) (tools/gen_db_table.py) vs the compiled reference."""
    rng = np.random.default_rng(80)
    xs = list(range(0, 70000)) + [int(v) for v in rng.integers(0, 2 ** 32, size=20000, dtype=np.uint64)] + [2 ** 32 - 1, 2 ** 31, 2 ** 24, 2 ** 24 - 1, 2 ** 16, 2 ** 16 - 1]
    for x in xs:
        assert oracle.db_fixed_times10(x) == reference.db_fixed_times10(x), x


def test_rx_nr_prach(oracle, reference):
    """gNB PRACH detector: the real rx_nr_prach on long (839) and short (139) sequences, 1-4 antennas, several NCS / formats / numerologies, with the
    root sequences of the real compute_nr_prach_seq -- detected preamble, energy and timing advance vs the oracle; a sent preamble is found, and noise-only input gives the
    same (arbitrary) answer in both."""
    from common import PRACH_CASES, prach_inputs, prach_num_roots
    rng = np.random.default_rng(81)
    for case in PRACH_CASES:
        nb_rx, short, root, NCS, fmt, mu, pre, delay, amp, sigma = case
        nroots = prach_num_roots(short, NCS)
        xu = reference.prach_seq(short, nroots, root)
        rx = prach_inputs(rng, case, xu)
        got_r = reference.rx_nr_prach(nb_rx, short, root, nroots, NCS, fmt, mu, xu, rx)
        got_o = oracle.rx_nr_prach(nb_rx, short, NCS, fmt, mu, xu, rx)
        assert got_o == got_r, (case, got_o, got_r)
        if pre >= 0 and sigma * 3 < amp <= 20000:            # full-scale input overflows the transform: equality still holds, detection does not
            assert got_r[0] == pre, (case, got_r)


def test_pdsch_tx_slot_fuzz(oracle, reference):
    """250 random PDSCH transmitter configurations (allocation, DMRS layout, 1-4 layers, PT-RS on / off, wideband precoding on / off) through the real nr_generate_pdsch.
    Configurations the oracle declines (a port whose CDM group carries data; the over-mapping case, DESIGN.md defect 10) are skipped."""
    from oracle.bindings import PdschTxParms
    from common import pdsch_tx_fuzz_cases
    rng = np.random.default_rng(89)
    done = 0
    oracle.lib.orc_pdsch_tx_slot.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    for N, carrier, ntx, slot, rb0, nrb, Qm, nl, s0, ns, dpos, dtype_, cdm, ports, scid, amp, ptrs, pm in pdsch_tx_fuzz_cases(rng, 250):
        P = PdschTxParms(N, ntx, slot, rb0, 0, nrb, N - carrier * 6, Qm, nl, s0, ns, dpos, dtype_, cdm, ports, scid, 40 + slot, 501, 0x1234 + slot, amp)
        if ptrs:
            P.set_ptrs(*ptrs)
        if pm:
            P.set_precoding(pm, rng.integers(-12000, 12001, size=(4, 4, 2)).astype(np.int16))
        if P.G() <= 0:
            continue
        bits = rng.integers(0, 2, size=P.G(), dtype=np.uint8)
        out = np.zeros((P.nb_tx, 14, P.fft_size, 2), np.int16)
        rc = oracle.lib.orc_pdsch_tx_slot(C.addressof(P), bits.ctypes.data, out.ctypes.data)
        if rc < 0:
            continue
        assert rc == bits.size
        t_r = reference.pdsch_tx_slot(P, bits, carrier)
        assert np.array_equal(out, t_r), (N, nrb, Qm, nl, dpos, dtype_, cdm, ptrs, pm, [tuple(x) for x in np.argwhere(out != t_r)[:5]])
        done += 1
    assert done > 150


def test_rx_nr_prach_fuzz(oracle, reference):
    """150 random PRACH occasions (antennas, long / short sequences, root index, every N_CS of the unrestricted tables, formats, numerologies, sent preamble or noise only,
    amplitudes up to full scale) through the real rx_nr_prach with the real compute_nr_prach_seq."""
    from common import prach_fuzz_cases, prach_inputs, prach_num_roots
    rng = np.random.default_rng(93)
    for case in prach_fuzz_cases(rng, 150):
        nb_rx, short, root, NCS, fmt, mu, pre, delay, amp, sigma = case
        nroots = prach_num_roots(short, NCS)
        xu = reference.prach_seq(short, nroots, root)
        rx = prach_inputs(rng, case, xu)
        got_r = reference.rx_nr_prach(nb_rx, short, root, nroots, NCS, fmt, mu, xu, rx)
        got_o = oracle.rx_nr_prach(nb_rx, short, NCS, fmt, mu, xu, rx)
        assert got_o == got_r, (case, got_o, got_r)


def test_dft_size_index_enumerators_match_oai_header():
    """The size index dft() / idft() receive is OAI's dft_size_idx_t / idft_size_idx_t enumerator: the library's table (nrb200_dft_size_of_index, no GPU needed) and
    the Python mirror are pinned to get_dft / get_idft compiled from OAI's own tools_defs.h (oracle/ref_harness_dftidx.c)."""
    import ctypes as C
    import os
    from openairinterface5g_b200 import dfts
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    path = os.path.join(ROOT, "oracle", "_ref", "libref_dftidx.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libref_dftidx.so not built (needs /root/reference)")
    ref = C.CDLL(path)
    lib = C.CDLL(os.path.join(ROOT, "openairinterface5g_b200", "libdfts_b200.so"))
    assert ref.refh_dft_count() == len(dfts.DFT_SIZES) and ref.refh_idft_count() == len(dfts.IDFT_SIZES)
    for i, n in enumerate(dfts.DFT_SIZES):
        assert ref.refh_dft_size_at(i) == n and dfts.get_dft(n) == i, (i, n)
        assert lib.nrb200_dft_size_of_index(0, i) == n, (i, n)
    for i, n in enumerate(dfts.IDFT_SIZES):
        assert ref.refh_idft_size_at(i) == n and dfts.get_idft(n) == i, (i, n)
        assert lib.nrb200_dft_size_of_index(1, i) == n, (i, n)
    assert ref.refh_get_dft4096() == dfts.get_dft(4096) and ref.refh_get_idft4096() == dfts.get_idft(4096)


def test_decode_all_driver_matches_single_calls(oracle, reference):
    """oracle/cpu_bench.c:orc_decode_all (the all-cores driver of the large-sample GPU differential tests) returns what one-by-one calls return,
    in both stop modes."""
    from common import decode_all_reference, make_case
    K, P, llr = make_case(oracle, 1, 384, 13, 12, 2.2, 99)
    its, out = decode_all_reference(oracle, reference, llr, 1, 384, 13, 8, threads=4)
    for i in range(12):
        it_r, out_r = reference.decode(1, 384, 13, 8, llr[i])
        assert its[i] == it_r and np.array_equal(out[i], out_r)
    K, P, llr = make_case(oracle, 1, 384, 23, 6, 4.0, 98)
    its, out = decode_all_reference(oracle, reference, llr, 1, 384, 23, 8, crc=(1, K), threads=3)
    for i in range(6):
        it_r, out_r = reference.decode(1, 384, 23, 8, llr[i], use_crc=1, crc_len_bits=K, crc_type=1)
        assert its[i] == it_r and np.array_equal(out[i], out_r)


def test_transform_precoding_sequences_chest_and_inner_rx(oracle, reference):
    """DFT-s-OFDM (transform precoding enabled): every computed low-PAPR type-1 sequence (M_ZC = 30 and >= 36, 3 groups), nr_pusch_channel_estimation with the
    low-PAPR pilots, and inner_rx with nr_freq_equalization + nr_idft (incl. M = 12 with its own scaling and M = 1536 / 3072 through idft())."""
    from oracle.bindings import ChestParms, PuschParms
    n = 0
    for k in range(5, 274):
        for u in (0, 7, 29):
            a = reference.lowpapr_seq(u, 0, 6 * k)
            if a is not None:
                assert np.array_equal(a, oracle.lowpapr_seq(u, 0, 6 * k)), (k, u)
                n += 1
    assert n == 3 * 49
    rng = np.random.default_rng(3)
    carrier = {1024: 52, 2048: 106, 4096: 273}
    for N, nb_rx, rb_start, nb, slot, u, port, cf in ((2048, 2, 10, 25, 3, 5, 0, 0), (4096, 4, 0, 270, 1, 0, 0, 0), (1024, 1, 7, 6, 0, 29, 1, 0), (2048, 2, 0, 5, 2, 11, 3, 0),
                                                     (1024, 2, 3, 2, 4, 3, 0, 0), (2048, 2, 10, 25, 3, 5, 0, 1), (512, 3, 1, 4, 7, 20, 2, 1), (1024, 1, 7, 6, 0, 29, 1, 1)):
        carrier.setdefault(512, 25)
        rx = rng.integers(-3000, 3001, size=(nb_rx, 14, N, 2)).astype(np.int16)
        P = ChestParms(N, nb_rx, slot, 2, port, rb_start, 0, nb, N - 6 * carrier[N], 0, 55, 0, cf)
        reference.chest_set_transform_precoding(1, u, 0)
        oracle.chest_set_lowpapr(reference.lowpapr_seq(u, 0, 6 * nb))
        try:
            e_r, st_r, _ = reference.pusch_channel_estimation(P, rx, carrier[N], chest_freq=cf)
            e_o, st_o = oracle.pusch_channel_estimation(P, rx)
        finally:
            reference.chest_set_transform_precoding(0)
            oracle.chest_set_lowpapr(None)
        assert np.array_equal(e_r, e_o) and np.array_equal(st_r, st_o[:5]), (N, nb, port, cf)
    reference.pusch_set_transform_precoding(1)
    oracle.pusch_set_transform_precoding(1)
    try:
        for N, nb_rx, rb_start, nb, Qm in ((2048, 2, 10, 25, 6), (4096, 4, 0, 270, 4), (1024, 1, 7, 6, 2), (2048, 2, 0, 5, 6), (1024, 2, 3, 1, 4), (4096, 2, 5, 128, 6),
                                          (4096, 1, 0, 256, 2), (1024, 2, 3, 2, 6), (4096, 2, 0, 135, 4)):
            P = PuschParms(N, nb_rx, rb_start, 0, nb, N - 6 * carrier[N], Qm, 1 << 2, 0, 2)
            rx = rng.integers(-2000, 2001, size=(nb_rx, 14, N, 2)).astype(np.int16)
            h = rng.integers(-1500, 1501, size=(nb_rx, 14, N, 2)).astype(np.int16)
            for sym, shift in ((0, 7), (5, 9)):
                l_o, c_o = oracle.pusch_inner_rx_symbol(P, sym, 2, shift, rx, h)
                l_r, c_r = reference.pusch_inner_rx_symbol(P, sym, 2, shift, rx, h, 12 * nb)
                assert np.array_equal(l_o, l_r) and np.array_equal(c_o, c_r[:c_o.size]), (N, nb, Qm, sym)
    finally:
        reference.pusch_set_transform_precoding(0)
        oracle.pusch_set_transform_precoding(0)
