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


@pytest.mark.parametrize("cfg", [dict(), dict(A=18696, N=2048, carrier_rb=106, rb_start=20, rb_size=50, nb_rx=2, Qm=4, slot=3)])
def test_pusch_slot_roundtrip(ldpc, cfg):
    dev = torch.device("cuda", 0)
    chain = PuschSlotChain(ldpc, load_dftslib(), dev, **cfg)
    payload, rxdata, est = chain.synthesize(seed=5)
    tb, iters, tbcrc = chain.receive(rxdata, est)
    torch.cuda.synchronize()
    it = iters.cpu().numpy()
    assert (it <= chain.max_iter).all(), it
    got = tb.cpu().numpy().reshape(-1)
    assert np.array_equal(got[:payload.size], payload)
    assert int(tbcrc.cpu()[0]) == 0
    assert int(chain.level.cpu()[8]) > 0
