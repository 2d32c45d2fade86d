"""GPU parity of libdfts_b200.so against the CPU oracle (pinned bit-exactly to oai_dfts.c) and the golden fixtures."""
import os
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SIZES = [64, 128, 256, 512, 768, 1024, 1536, 2048, 3072, 4096, 6144, 8192]


@pytest.fixture(scope="module")
def dfts():
    from openairinterface5g_b200.dfts import load_dftslib
    return load_dftslib()


@pytest.mark.parametrize("N", SIZES)
def test_batch_vs_oracle(dfts, oracle, N):
    rng = np.random.default_rng(N)
    nb = 5 if N >= 2048 else 37                                   # ragged: not a multiple of the transforms-per-CTA packing
    for inverse in (False, True):
        for amp in (300, 3000, 20000, 32767):
            for scale in (1, 0):
                x = rng.integers(-amp, amp + 1, size=(nb, 2 * N)).astype(np.int16)
                x[-1] = rng.choice(np.array([-32768, 32767, 0], dtype=np.int16), size=2 * N)
                got = dfts.batch_host(N, inverse, x, scale)
                for b in (0, nb // 2, nb - 1):
                    assert np.array_equal(got[b], oracle.dft(N, inverse, x[b], scale)), (N, inverse, amp, scale, b)


def test_plugin_abi_calls(dfts, oracle):
    """dft(get_dft(N), in, out, scale) / idft(get_idft(N), ...) exactly as nr_slot_fep / PHY_ofdm_mod call them."""
    from openairinterface5g_b200.dfts import get_dft, get_idft
    rng = np.random.default_rng(1)
    for N in (4096, 2048, 1536, 512):
        x = rng.integers(-2000, 2000, size=2 * N).astype(np.int16)
        assert np.array_equal(dfts.dft(get_dft(N), x, 1), oracle.dft(N, False, x, 1))
        assert np.array_equal(dfts.idft(get_idft(N), x, 1), oracle.dft(N, True, x, 1))


def test_golden(dfts):
    d = np.load(os.path.join(G, "dft.npz"))
    for N in SIZES:
        assert np.array_equal(dfts.batch_host(N, False, d[f"x{N}"], 1), d[f"dft{N}"]), N
        assert np.array_equal(dfts.batch_host(N, True, d[f"x{N}"], 1), d[f"idft{N}"]), N


def test_slot_of_ofdm_symbols_properties(dfts):
    """BASELINE config-3 shape: one 100 MHz slot = 14 symbols x 2 antennas of 4096-point transforms, device resident."""
    import torch
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev); g.manual_seed(4)
    x = torch.randint(-1500, 1500, (28, 2 * 4096), dtype=torch.int16, device=dev, generator=g)
    X = dfts.batch_torch(4096, False, x, 1)
    torch.cuda.synchronize()
    # linearity up to fixed-point rounding and agreement with a float FFT scaled by 1/sqrt(N) (= 1/64)
    xf = x.view(28, 4096, 2).to(torch.float64)
    ref = torch.fft.fft(torch.complex(xf[..., 0], xf[..., 1]), dim=1) / 64.0
    got = X.view(28, 4096, 2).to(torch.float64)
    err = (torch.complex(got[..., 0], got[..., 1]) - ref).abs().max().item()
    assert err < 40.0, err                                        # a few LSBs of accumulated Q15 truncation across 6 stages
    # idft(dft(x)) returns x up to that rounding
    xr = dfts.batch_torch(4096, True, X, 1)
    torch.cuda.synchronize()
    assert (xr.to(torch.int32) - x.to(torch.int32)).abs().max().item() < 80
    # determinism
    assert torch.equal(X, dfts.batch_torch(4096, False, x, 1))


FOURWAY = [12, 24, 36, 48, 60, 72, 96, 108, 120, 144, 180, 192, 216, 240, 288, 300, 324, 360, 384, 432, 480, 540, 576, 600, 648, 720, 864, 900, 960, 972, 1080,
           1152, 1200, 1296, 1440, 1500, 1620, 1728, 1800, 1920, 1944, 2160, 2304, 2400, 2592, 2700, 2880, 2916, 3000, 3240]


@pytest.mark.parametrize("N", FOURWAY)
def test_fourway_vs_oracle(dfts, oracle, N):
    """DFT-s-OFDM family: 3 calls per launch (4 N c16 each), amplitudes that do and do not saturate, scale_flag 1 / 0, and the plug-in call dft(DFT_<N>, ...)."""
    from openairinterface5g_b200.dfts import get_dft
    rng = np.random.default_rng(N)
    for amp in (300, 3000, 32767):
        for scale in (1, 0):
            x = rng.integers(-amp, amp + 1, size=(3, 8 * N)).astype(np.int16)
            got = dfts.batch_host(N, False, x, scale)
            for b in range(3):
                assert np.array_equal(got[b], oracle.dft4(N, x[b], scale)), (N, amp, scale, b)
    x = rng.choice(np.array([-32768, 32767, 0], dtype=np.int16), size=8 * N)
    assert np.array_equal(dfts.dft(get_dft(N), x, 1), oracle.dft4(N, x, 1))


def test_fourway_golden_and_device_batch(dfts):
    d = np.load(os.path.join(G, "dft4.npz"))
    for N in d["sizes"]:
        N = int(N)
        assert np.array_equal(dfts.batch_host(N, False, d[f"x{N}"], 1), d[f"y{N}_s1"]), N
        assert np.array_equal(dfts.batch_host(N, False, d[f"x{N}"], 0), d[f"y{N}_s0"]), N
    # a PUSCH-sized batch on the device: 273 PRB = 3240 + 36 is not one DFT size, 270 PRB = 3240 is; 13 symbols x 4 layers = 52 calls in one launch
    N = 3240
    x = torch.from_numpy(np.tile(d[f"x{N}"], (52, 1))).cuda()
    y = dfts.batch_torch(N, False, x, 1)
    torch.cuda.synchronize()
    assert all(np.array_equal(y[i].cpu().numpy(), d[f"y{N}_s1"]) for i in (0, 25, 51))


@pytest.mark.parametrize("N", [12288, 16384, 18432, 24576, 32768, 36864, 49152, 65536, 98304])
def test_large_sizes_vs_oracle(dfts, oracle, N):
    """Sizes above 8192: gather + batched shared-memory transforms + one global pass per top level; 2 transforms per launch, both directions."""
    rng = np.random.default_rng(N)
    for inverse in (False, True):
        if N == 65536 and not inverse:
            continue                                   # the reference has no dft65536
        for amp, scale in ((300, 1), (3000, 0), (32767, 1), (3000, 2)):
            x = rng.integers(-amp, amp + 1, size=(2, 2 * N)).astype(np.int16)
            got = dfts.batch_host(N, inverse, x, scale)
            for b in range(2):
                assert np.array_equal(got[b], oracle.dft(N, inverse, x[b], scale)), (N, inverse, amp, scale, b)


def test_large_plugin_abi_and_device_batch(dfts, oracle):
    from openairinterface5g_b200.dfts import get_dft, get_idft
    rng = np.random.default_rng(77)
    x = rng.integers(-2000, 2001, size=2 * 24576).astype(np.int16)
    assert np.array_equal(dfts.dft(get_dft(24576), x, 1), oracle.dft(24576, False, x, 1))
    assert np.array_equal(dfts.idft(get_idft(24576), x, 1), oracle.dft(24576, True, x, 1))
    xd = torch.from_numpy(np.tile(x, (5, 1))).cuda()
    y = dfts.batch_torch(24576, True, xd, 1)
    torch.cuda.synchronize()
    want = oracle.dft(24576, True, x, 1)
    assert all(np.array_equal(y[i].cpu().numpy(), want) for i in range(5))
