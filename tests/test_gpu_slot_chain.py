"""End-to-end composition on the GPU: a full-band 100 MHz PUSCH slot (273 PRB, 64QAM, 28 code blocks of K=8448) is synthesised with the
library's transmit kernels, sent through a flat 4-antenna channel into the time domain, and received with the device-resident chain
OFDM demod -> level -> compensation/LLR/descrambling -> rate recovery -> LDPC decode (CRC24B stop) -> TB CRC.  Every kernel on the way is
parity-tested against the oracle elsewhere; this test checks that they compose: the transport block comes back intact."""
import numpy as np
import pytest
import torch

from openairinterface5g_b200.dfts import load_dftslib
from openairinterface5g_b200.slot_chain import PuschSlotChain

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cfg", [dict(), dict(A=18696, N=2048, carrier_rb=106, rb_start=20, rb_size=50, nb_rx=2, Qm=4, slot=3),
                                 dict(A=471272, n_layers=2), dict(A=33640, N=1024, mu=0, carrier_rb=52, rb_size=52, nb_rx=2, Qm=6, slot=1, n_layers=2)])
def test_pusch_slot_roundtrip(ldpc, oracle, cfg):
    dev = torch.device("cuda", 0)
    chain = PuschSlotChain(ldpc, load_dftslib(), dev, **cfg)
    payload, rxdata, est = chain.synthesize(seed=5)
    tb, iters, tbcrc = chain.receive(rxdata)                              # channel estimated from the DMRS symbol
    torch.cuda.synchronize()
    it = iters.cpu().numpy()
    assert (it <= chain.max_iter).all(), it
    got = tb.cpu().numpy().reshape(-1)
    assert np.array_equal(got[:payload.size], payload)
    assert int(tbcrc.cpu()[0]) == 0
    assert int(chain.level.cpu()[8]) > 0
    if cfg:
        # the same slot through the oracle-only chain (one and two layers): LLRs, iteration counts and the transport block must agree bit for bit
        from common import oracle_pusch_receive
        info = dict(C=chain.C, K=chain.K, Z=chain.Z, F=chain.F, E=[int(e) for e in chain.E.cpu()])
        frame = rxdata.cpu().numpy().reshape(chain.nb_rx, -1)
        tb_o, its_o, llr_o, shift_o = oracle_pusch_receive(oracle, chain.P, info, chain.Qm, chain.rb_start, chain.rb_size, chain.nb_rx, chain.slot,
                                                           chain.rnti, chain.nid, chain.rot, frame, None, n_layers=chain.nl, cdm=chain.cdm)
        assert shift_o == int(chain.level.cpu()[8])
        assert np.array_equal(chain.llr16.cpu().numpy(), llr_o)
        assert np.array_equal(it, its_o) and np.array_equal(got[:tb_o.size], tb_o)


def test_pusch_slots_in_flight_match_single_slot(ldpc):
    """PuschSlotPipeline: every stream's slot (CUDA graph replay, device resident and with the samples / transport block crossing PCIe) decodes its own
    payload, and gives the LLRs and iteration counts the same slot gives when run alone."""
    import torch
    from openairinterface5g_b200.dfts import load_dftslib
    from openairinterface5g_b200.slot_chain import PuschSlotChain, PuschSlotPipeline
    dl = load_dftslib()
    dev = torch.device("cuda", 0)
    cfg = dict(A=18696, N=2048, carrier_rb=106, rb_start=20, rb_size=50, nb_rx=2, Qm=4, slot=3)
    pipe = PuschSlotPipeline(ldpc, dl, dev, 3, seed0=300, **cfg)
    pipe.timed_rounds(2)
    assert all(pipe.check())
    pipe.timed_rounds(2, e2e=True)
    assert all(pipe.check(host=True))
    for k in (0, 2):
        alone = PuschSlotChain(ldpc, dl, dev, **cfg)
        payload, rxdata, _ = alone.synthesize(seed=300 + k, snr_db=30.0)
        alone.receive(rxdata)
        torch.cuda.synchronize()
        assert torch.equal(alone.llr16, pipe.chains[k].llr16) and torch.equal(alone.iters, pipe.chains[k].iters) and torch.equal(alone.tb, pipe.chains[k].tb)


def test_two_layer_receiver_reads_estimator_state_on_device(ldpc):
    """nrb200_pusch_rx_t.d_est_state: max_ch / nvar taken from the estimator's device state give the LLRs, level and transport block of the path that
    carries them through the host side of the descriptor; and the two-layer slot is capturable (PuschSlotPipeline with n_layers = 2)."""
    from openairinterface5g_b200.slot_chain import PuschSlotPipeline
    dl = load_dftslib()
    dev = torch.device("cuda", 0)
    cfg = dict(A=33640, N=1024, mu=0, carrier_rb=52, rb_size=52, nb_rx=2, Qm=6, slot=1, n_layers=2)
    a, b = PuschSlotChain(ldpc, dl, dev, **cfg), PuschSlotChain(ldpc, dl, dev, **cfg)
    payload, rxdata, _ = a.synthesize(seed=9)
    b.host_scalars = True
    a.receive(rxdata); b.receive(rxdata)
    torch.cuda.synchronize()
    assert torch.equal(a.level, b.level) and torch.equal(a.llr16, b.llr16) and torch.equal(a.iters, b.iters) and torch.equal(a.tb, b.tb)
    assert b.desc.noise_var > 0 and np.array_equal(a.tb.cpu().numpy().reshape(-1)[:payload.size], payload)
    pipe = PuschSlotPipeline(ldpc, dl, dev, 2, seed0=400, **cfg)
    pipe.timed_rounds(2, e2e=True)
    assert all(pipe.check(host=True))


@pytest.mark.parametrize("cfg", [dict(A=18696, N=2048, carrier_rb=106, rb_start=20, rb_size=50, nb_rx=2, Qm=4, slot=3), dict(A=471272, n_layers=2)])
def test_slot_entry_point_equals_staged_calls(ldpc, cfg):
    """nrb200_sch_slot_rx_dev (the whole PUSCH slot in one library call, include/nrb200_slot.h) against the same stages issued one entry point at a
    time: every intermediate and the result identical."""
    dev = torch.device("cuda", 0)
    dl = load_dftslib()
    a, b = PuschSlotChain(ldpc, dl, dev, **cfg), PuschSlotChain(ldpc, dl, dev, **cfg)
    payload, rxdata, _ = a.synthesize(seed=9)
    a.receive(rxdata)
    b.receive(rxdata, staged=True)
    torch.cuda.synchronize()
    for name in ("rxF", "est", "level", "llr16", "llr8", "iters", "tb", "tbcrc"):
        assert torch.equal(getattr(a, name), getattr(b, name)), name
    nb = (a.K - a.F) // 8                                      # the decoder writes ncols(R) * Z / 8 bytes per segment; the rows are wider
    assert torch.equal(a.hard[:, :nb], b.hard[:, :nb])
    assert np.array_equal(a.tb.cpu().numpy().reshape(-1)[:payload.size], payload)


@pytest.mark.parametrize("cfg", [dict(A=19464, N=1024, mu=0, carrier_rb=52, rb_start=2, rb_size=50, nb_rx=2, Qm=4, slot=1, transform_precoding=(7, 0)),
                                 dict(A=9992, N=2048, carrier_rb=106, rb_start=10, rb_size=25, nb_rx=4, Qm=4, slot=3, transform_precoding=(29, 0)),
                                 dict(A=3752, N=1024, mu=0, carrier_rb=52, rb_start=0, rb_size=20, nb_rx=2, Qm=2, slot=6, transform_precoding=(0, 0))])
def test_pusch_slot_with_transform_precoding_roundtrip(ldpc, cfg):
    """DFT-s-OFDM uplink slot: the synthesiser spreads every symbol's modulation symbols with an M-point DFT and sends the low-PAPR type-1 DMRS; the receive
    chain (one library call: estimation with the low-PAPR pilots, equalisation + nr_idft + LLRs, rate recovery, decode) must return the payload.
    QPSK and 16QAM: the reference's 64QAM thresholds after nr_freq_equalization do not fit its own constellation scale (DESIGN.md defect 20,
    tests/test_golden_oracle.py::test_transform_precoding_64qam_cannot_be_demapped), so a 64QAM DFT-s-OFDM block does not decode there either."""
    dev = torch.device("cuda", 0)
    chain = PuschSlotChain(ldpc, load_dftslib(), dev, **cfg)
    payload, rxdata, _ = chain.synthesize(seed=11)
    tb, iters, tbcrc = chain.receive(rxdata)
    torch.cuda.synchronize()
    assert (iters.cpu().numpy() <= chain.max_iter).all(), iters
    assert np.array_equal(tb.cpu().numpy().reshape(-1)[:payload.size], payload) and int(tbcrc.cpu()[0]) == 0
    # the same slot without the flag does not decode: the receiver really takes the other path
    plain = PuschSlotChain(ldpc, load_dftslib(), dev, **{k: v for k, v in cfg.items() if k != "transform_precoding"})
    plain.receive(rxdata)
    torch.cuda.synchronize()
    assert int(plain.tbcrc.cpu()[0]) != 0
