"""GPU parity of the rest of the LDPC boundary: encoder, CRC, CRC-stop decoding, the per-code-block OAI ABI, golden vectors,
concurrency, and size-independent properties at the BASELINE batch size (1024)."""
import os
import threading
import numpy as np
import pytest
from common import ALL_Z, RATES, NCOLS, make_case, payloads

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("BG", [1, 2])
def test_encoder_all_z_vs_oracle(ldpc, oracle, BG):
    for Z in ALL_Z:
        K, P = payloads(BG, Z, 9, Z)
        got = ldpc.encode_batch_host(BG, Z, K, P)
        for i in range(9):
            assert np.array_equal(got[i], oracle.encode(BG, Z, K, P[i])), (BG, Z, i)


def test_encoder_golden_and_abi(ldpc):
    d = np.load(os.path.join(G, "ldpc_encoder.npz"))
    for ci in range(int(d["ncases"][0])):
        BG, Z, K = [int(x) for x in d[f"e{ci}_par"]]
        nout = (66 if BG == 1 else 50) * Z
        want = np.unpackbits(d[f"e{ci}_orig"], axis=1)[:, :nout]
        assert np.array_equal(ldpc.encode_batch_host(BG, Z, K, d[f"e{ci}_in"]), want)
        assert np.array_equal(ldpc.LDPCencoder(BG, Z, K, d[f"e{ci}_in"]), want)        # groups of 8 via macro_num, 9 segments


def test_crc_all_polys(ldpc, oracle):
    d = np.load(os.path.join(G, "coding.npz"))
    for i, n in enumerate(d["crc_lens"]):
        for p in range(8):
            got = ldpc.crc_batch_host(p, d["crc_data"][i:i + 1], int(n))
            assert int(got[0]) == int(d["crc_vals"][i][p]), (p, n)
    rng = np.random.default_rng(1)
    blk = rng.integers(0, 256, size=(64, 1056), dtype=np.uint8)
    got = ldpc.crc_batch_host(1, blk, 8424)
    assert [int(x) for x in got] == [oracle.crc(1, blk[i], 8424) for i in range(64)]


def test_decoder_golden(ldpc):
    d = np.load(os.path.join(G, "ldpc_decoder.npz"))
    for ci in range(int(d["ncases"][0])):
        BG, Z, R, n, mi, om, uc, K = [int(x) for x in d[f"c{ci}_par"]]
        iters, out = ldpc.decode_batch_host(BG, Z, R, mi, d[f"c{ci}_llr"], outMode=om)
        # intended arithmetic = the reference's AVX512 build (identical to its AVX2 build except the BG2 R15 generator defect)
        assert np.array_equal(iters, d[f"c{ci}_iters_avx512"]), ci
        assert np.array_equal(out.view(np.uint8), d[f"c{ci}_out_avx512"].view(np.uint8)), ci


def test_decoder_crc_stop_mode(ldpc, oracle):
    d = np.load(os.path.join(G, "ldpc_decoder.npz"))
    BG, Z, R, n, K = [int(x) for x in d["crc_par"]]
    for mi in (2, 3, 8):
        iters, out = ldpc.decode_batch_host(BG, Z, R, mi, d["crc_llr"], use_crc=1, crc_len_bits=K, crc_type=1)
        assert np.array_equal(iters, d[f"crc{mi}_iters"]), mi
        assert np.array_equal(out, d[f"crc{mi}_out"]), mi
    # more shapes against the oracle, incl. the generic (Z % 4 != 0) kernel
    for BG, Z, R, e in ((1, 384, 13, 2.4), (2, 96, 13, 3.0), (1, 10, 13, 4.0)):
        K = (22 if BG == 1 else 10) * Z
        if K % 8:
            continue
        rng = np.random.default_rng(Z)
        P = rng.integers(0, 256, size=(5, K // 8), dtype=np.uint8)
        for i in range(5):
            crc = oracle.crc(0, P[i], K - 24) >> 8
            P[i, -3:] = [(crc >> 16) & 0xFF, (crc >> 8) & 0xFF, crc & 0xFF]
        from openairinterface5g_b200.synth import awgn_llr
        cw = np.stack([oracle.encode(BG, Z, K, P[i]) for i in range(5)])
        llr = awgn_llr(cw, Z, NCOLS[(BG, R)], e, (22 if BG == 1 else 10) / (NCOLS[(BG, R)] - 2), Z)
        for mi in (1, 4, 8):
            iters, out = ldpc.decode_batch_host(BG, Z, R, mi, llr, use_crc=1, crc_len_bits=K, crc_type=0)
            for i in range(5):
                it_o, out_o = oracle.decode(BG, Z, R, mi, llr[i], 0, 1, K, 0)
                assert iters[i] == it_o and np.array_equal(out[i], out_o), (BG, Z, mi, i)


@pytest.mark.parametrize("out_mode", [1, 2])
def test_decoder_crc_stop_one_bit_per_byte_modes(ldpc, oracle, out_mode):
    """check_crc with outMode BITINT8 / LLRINT8: the reference runs the CRC over the one-bit-per-byte output array as it stands
    (nrLDPC_decoder.c:852-858).  An all-zero code word makes that array's CRC pass (early stop), random payloads never pass (numMaxIter + 1)."""
    from openairinterface5g_b200.synth import awgn_llr
    for BG, Z, R in ((1, 384, 13), (2, 96, 13), (1, 10, 13)):
        K = (22 if BG == 1 else 10) * Z
        if K % 8:
            continue
        rng = np.random.default_rng(Z + out_mode)
        P = rng.integers(0, 256, size=(4, K // 8), dtype=np.uint8)
        P[:2] = 0
        cw = np.stack([oracle.encode(BG, Z, K, P[i]) for i in range(4)])
        llr = awgn_llr(cw, Z, NCOLS[(BG, R)], 3.5, (22 if BG == 1 else 10) / (NCOLS[(BG, R)] - 2), Z)
        iters, out = ldpc.decode_batch_host(BG, Z, R, 6, llr, outMode=out_mode, use_crc=1, crc_len_bits=K, crc_type=1)
        for i in range(4):
            it_o, out_o = oracle.decode(BG, Z, R, 6, llr[i], out_mode, 1, K, 1)
            assert iters[i] == it_o and np.array_equal(out[i].view(np.uint8), np.asarray(out_o).view(np.uint8)), (BG, Z, i, iters[i], it_o)
        assert iters[0] <= 6 and iters[3] == 7


def test_oai_per_block_abi(ldpc, oracle):
    """LDPCdecoder exactly as nr_ulsch_decoding.c:218-222 / ldpctest.c:329-340 call it."""
    from openairinterface5g_b200.ldpc import DecodeAbort, LdpcTimeStats
    K, P, llr = make_case(oracle, 1, 384, 13, 3, 2.6, seed=11)
    prof = LdpcTimeStats()
    for i in range(3):
        ab = DecodeAbort()
        it, out = ldpc.LDPCdecoder(1, 384, 13, 8, llr[i], abort=ab, profiler=prof)
        it_o, out_o = oracle.decode(1, 384, 13, 8, llr[i])
        assert it == it_o and np.array_equal(out, out_o)
        assert bool(ab.failed) == (it > 8)                                   # set_abort on failure (nrLDPC_decoder.c:190-193)
        if it <= 8:
            assert np.array_equal(out[:K // 8], P[i])
    assert prof.total.trials == 3 and prof.total.diff > 0
    # a peer segment already failed: the decoder bails out after the first iteration (nrLDPC_decoder.c:557-560)
    ab = DecodeAbort()
    ab.failed = True
    it, out = ldpc.LDPCdecoder(1, 384, 13, 8, llr[0], abort=ab)
    it_o, out_o = oracle.decode(1, 384, 13, 8, llr[0], abort_in=1)
    assert it == it_o == 10 and np.array_equal(out, out_o)


def test_concurrent_callers(ldpc, oracle):
    """tpool-style concurrency: one blocking call per segment from many threads (SURVEY.md 8b 'Threading')."""
    K, P, llr = make_case(oracle, 1, 384, 13, 16, 3.0, seed=21)
    want = [oracle.decode(1, 384, 13, 8, llr[i]) for i in range(16)]
    got = [None] * 16

    def work(i):
        got[i] = ldpc.LDPCdecoder(1, 384, 13, 8, llr[i])
    ths = [threading.Thread(target=work, args=(i,)) for i in range(16)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    for i in range(16):
        assert got[i][0] == want[i][0] and np.array_equal(got[i][1], want[i][1]), i


def test_packed_and_generic_kernels_agree(oracle):
    """The same batch through both kernels (NRB200_FORCE_GENERIC picks the byte-per-thread anchor kernel)."""
    import subprocess
    import sys
    code = (
        "import sys, numpy as np; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "from openairinterface5g_b200.ldpc import load_LDPClib\n"
        "d = np.load(%r)\n"
        "lib = load_LDPClib()\n"
        "it, out = lib.decode_batch_host(1, 384, 13, 8, d['llr'])\n"
        "np.savez(sys.argv[1], it=it, out=out)\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        K, P, llr = make_case(oracle, 1, 384, 13, 24, 2.5, seed=5)
        np.savez(os.path.join(td, "in.npz"), llr=llr)
        res = []
        for env in ({}, {"NRB200_FORCE_GENERIC": "1"}):
            e = dict(os.environ); e.update(env)
            outp = os.path.join(td, "o%d.npz" % len(res))
            subprocess.check_call([sys.executable, "-c", code % (root, os.path.join(root, "tests"), os.path.join(td, "in.npz")), outp], env=e)
            res.append(np.load(outp))
        assert np.array_equal(res[0]["it"], res[1]["it"]) and np.array_equal(res[0]["out"], res[1]["out"])


def test_avx2_bg2_r15_defect_emulation(oracle):
    """NRB200_EMULATE_AVX2_BG2R15_DEFECT=1 reproduces the reference's AVX2 build bit-exactly for BG2 R=15."""
    import subprocess
    import sys
    import tempfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    d = np.load(os.path.join(G, "ldpc_decoder.npz"))
    ci = [c for c in range(int(d["ncases"][0])) if tuple(d[f"c{c}_par"][[0, 1, 2]]) == (2, 384, 15)][0]
    code = (
        "import sys, numpy as np; sys.path.insert(0, %r)\n"
        "from openairinterface5g_b200.ldpc import load_LDPClib\n"
        "d = np.load(%r)\n"
        "it, out = load_LDPClib().decode_batch_host(2, 384, 15, 8, d['c%d_llr'])\n"
        "assert np.array_equal(it, d['c%d_iters_avx2']) and np.array_equal(out, d['c%d_out_avx2'])\n" % (root, os.path.join(G, "ldpc_decoder.npz"), ci, ci, ci))
    e = dict(os.environ)
    e["NRB200_EMULATE_AVX2_BG2R15_DEFECT"] = "1"
    subprocess.check_call([sys.executable, "-c", code], env=e)


def test_batch1024_properties(ldpc, oracle, reference):
    """BASELINE size (batch 1024, BG1 Z=384): encode -> noisy channel -> decode; size-independent properties AND every block against the compiled reference."""
    import torch
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev); g.manual_seed(3)
    B, Z, K = 1024, 384, 8448
    payload = torch.randint(0, 256, (B, K // 8), dtype=torch.uint8, device=dev, generator=g)
    cw = ldpc.encode_batch_torch(1, Z, K, payload)
    assert cw.shape == (B, 66 * Z) and int(cw.max()) <= 1
    # systematic part reproduces the payload bits (after the 2Z punctured columns)
    bits = ((payload[:, :, None] >> torch.arange(7, -1, -1, device=dev)) & 1).reshape(B, K)
    assert torch.equal(bits[:, 2 * Z:], cw[:, :K - 2 * Z])
    # linearity of the code: enc(a ^ b) == enc(a) ^ enc(b)
    p2 = torch.randint(0, 256, (B, K // 8), dtype=torch.uint8, device=dev, generator=g)
    assert torch.equal(ldpc.encode_batch_torch(1, Z, K, payload ^ p2), cw ^ ldpc.encode_batch_torch(1, Z, K, p2))
    sigma = 1.0 / np.sqrt(2.0 * 10 ** 0.4 / 3.0)   # Eb/N0 4 dB
    y = (1.0 - 2.0 * cw.float()) + sigma * torch.randn(cw.shape, device=dev, generator=g)
    llr = torch.zeros((B, 68 * Z), dtype=torch.int8, device=dev)
    llr[:, 2 * Z:] = torch.clamp(torch.floor(y / (sigma / 16)), -128, 127).to(torch.int8)
    iters, out = ldpc.decode_batch_torch(1, Z, 13, 8, llr)
    torch.cuda.synchronize()
    assert int(iters.max()) <= 8                         # every block converges at 4 dB (ldpctest: BLER 0 well below this SNR)
    assert torch.equal(out[:, :K // 8], payload)
    # every one of the 1024 blocks against the unmodified reference decoder (all host cores), here and at the bench's operating point A (1 dB)
    from common import decode_all_reference
    it_c, out_c = decode_all_reference(oracle, reference, llr.cpu().numpy(), 1, Z, 13, 8)
    assert np.array_equal(iters.cpu().numpy(), it_c) and np.array_equal(out.cpu().numpy(), out_c)
    sigma1 = 1.0 / np.sqrt(2.0 * 10 ** 0.1 / 3.0)
    llr1 = torch.zeros((B, 68 * Z), dtype=torch.int8, device=dev)
    llr1[:, 2 * Z:] = torch.clamp(torch.floor(((1.0 - 2.0 * cw.float()) + sigma1 * torch.randn(cw.shape, device=dev, generator=g)) / (sigma1 / 16)), -128, 127).to(torch.int8)
    it1, out1 = ldpc.decode_batch_torch(1, Z, 13, 8, llr1)
    it1_c, out1_c = decode_all_reference(oracle, reference, llr1.cpu().numpy(), 1, Z, 13, 8)
    assert np.array_equal(it1.cpu().numpy(), it1_c) and np.array_equal(out1.cpu().numpy(), out1_c)
    # idempotence / determinism: same input, same output and iteration counts
    iters2, out2 = ldpc.decode_batch_torch(1, Z, 13, 8, llr)
    torch.cuda.synchronize()
    assert torch.equal(iters, iters2) and torch.equal(out, out2)
    # host-buffer entry point gives the same answer as the device-resident one
    it_h, out_h = ldpc.decode_batch_host(1, Z, 13, 8, llr.cpu().numpy())
    assert np.array_equal(it_h, iters.cpu().numpy()) and np.array_equal(out_h, out.cpu().numpy())
    # empty batch
    it0, out0 = ldpc.decode_batch_host(1, Z, 13, 8, np.zeros((0, 68 * Z), dtype=np.int8))
    assert it0.size == 0


def test_crc_long_messages(ldpc, oracle):
    """Transport-block sized CRCs (nr_postDecode / nr_dlsch_encoding call crc24a / crc16 on the whole TB): chunk-folded kernel vs oracle."""
    rng = np.random.default_rng(12)
    for bitlen in (8456, 8457, 16384, 16385, 24584 + 5, 235648, 1277992):
        data = rng.integers(0, 256, size=(3, (bitlen + 7) // 8 + 3), dtype=np.uint8)
        for p in (0, 1, 3):
            got = ldpc.crc_batch_host(p, data, bitlen)
            assert [int(x) for x in got] == [oracle.crc(p, data[i], bitlen) for i in range(3)], (p, bitlen)
