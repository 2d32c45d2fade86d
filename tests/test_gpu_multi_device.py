"""One process driving several GPUs (SURVEY 8e / BASELINE config 5: OAI is one process, its code blocks shard across the GPUs of the node with no
data-path collective).  Needs >= 2 visible devices: `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi_device.py -m gpu`; skipped on one GPU."""
import os
import subprocess
import sys
import threading
import numpy as np
import pytest
from common import make_case

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ndev(ldpc):
    n = ldpc.device_count()
    if n < 2:
        pytest.skip("one visible GPU")
    return n


def test_batch_spread_over_devices_equals_one_device(ldpc, oracle, ndev):
    K, P, llr = make_case(oracle, 1, 384, 13, 301, 2.3, seed=3)
    it1, out1 = ldpc.decode_batch_host(1, 384, 13, 8, llr)
    for nd in range(2, ndev + 1):
        itm, outm = ldpc.decode_batch_host_multi(1, 384, 13, 8, llr, nd)
        assert np.array_equal(itm, it1) and np.array_equal(outm, out1), nd
    for i in (0, 150, 300):
        it_o, out_o = oracle.decode(1, 384, 13, 8, llr[i])
        assert it1[i] == it_o and np.array_equal(out1[i], np.asarray(out_o).view(np.uint8))


def test_thread_selects_its_device(ldpc, oracle, ndev):
    """nrb200_set_device is per host thread: two threads on two GPUs at once, device-resident buffers of their own device."""
    import torch
    K, P, llr = make_case(oracle, 1, 384, 13, 40, 2.4, seed=8)
    want = ldpc.decode_batch_host(1, 384, 13, 8, llr)
    got = {}

    def work(d):
        ldpc.set_device(d)
        with torch.cuda.device(d):
            t = torch.from_numpy(llr).to(f"cuda:{d}")
            it, out = ldpc.decode_batch_torch(1, 384, 13, 8, t)
            torch.cuda.synchronize(d)
            got[d] = (it.cpu().numpy(), out.cpu().numpy())
    ths = [threading.Thread(target=work, args=(d,)) for d in range(ndev)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    ldpc.set_device(0)
    for d in range(ndev):
        assert np.array_equal(got[d][0], want[0]) and np.array_equal(got[d][1], want[1]), d


def test_oai_entry_points_spread_calls_with_nrb200_devices(oracle, ndev):
    """NRB200_DEVICES=all: LDPCdecoder calls go round the GPUs, the offload convention pins (ulsch_id, r) with the sticky rule and HARQ combining still
    works across rounds (the soft buffer is found again on the same device).  Own process: the variable is read once."""
    code = r'''
import sys, numpy as np
sys.path.insert(0, "tests")
from common import make_case
from oracle.bindings import Oracle
from openairinterface5g_b200.ldpc import load_LDPClib, OffloadLdpcLib
orc = Oracle(); lib = load_LDPClib()
K, P, llr = make_case(orc, 1, 384, 13, 12, 2.4, 5)
ok = True
for i in range(12):
    it, out = lib.LDPCdecoder(1, 384, 13, 8, llr[i])
    a = orc.decode(1, 384, 13, 8, llr[i])
    ok = ok and it == a[0] and np.array_equal(out, np.asarray(a[1]).view(np.uint8))
off = OffloadLdpcLib()
rng = np.random.default_rng(1)
Z, E, Qm = 384, 9072, 6
for u in range(4):
    for r in range(3):
        pay = rng.integers(0, 256, K // 8, dtype=np.uint8)
        tx = off.LDPCencoder(1, Z, K, 0, Qm, 0, E, pay)
        l8 = np.where(tx == 0, 20, -20).astype(np.int8)
        it, hard = off.LDPCdecoder(1, Z, 23, 8, E, Qm, 0, 0, l8, ulsch_id=u, r=r, setCombIn=0)
        ok = ok and np.array_equal(hard[:K // 8], pay)
        it2, hard2 = off.LDPCdecoder(1, Z, 23, 8, E, Qm, 0, 0, (l8 // 4).astype(np.int8), ulsch_id=u, r=r, setCombIn=1)
        ok = ok and np.array_equal(hard2[:K // 8], pay)
print("MULTI_OK" if ok else "MULTI_BAD", lib.device_count())
'''
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=dict(os.environ, NRB200_DEVICES="all"), capture_output=True, text=True, timeout=600)
    assert "MULTI_OK" in r.stdout, r.stdout[-800:] + r.stderr[-2000:]
