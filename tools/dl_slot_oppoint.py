import sys, os, json, numpy as np, torch
sys.path.insert(0, "/root/repo")
from openairinterface5g_b200.ldpc import load_LDPClib
from openairinterface5g_b200.dfts import load_dftslib
from openairinterface5g_b200.dl_slot_chain import PdschSlotChain
dev = torch.device("cuda", 0)
lib, dl = load_LDPClib(), load_dftslib()
ch = PdschSlotChain(lib, dl, dev)
for coupling in (0.0, 0.15, 0.35):
    for gain in (2.0, 3.0, 4.0, 5.0):
        nfail, slots_ok, its, rails = 0, 0, [], []
        for pseed in range(200, 212):
            p = torch.from_numpy(np.random.default_rng(pseed).integers(0, 256, size=ch.A // 8, dtype=np.uint8)).to(dev)
            rx = ch.channel(ch.transmit(p), seed=pseed, snr_db=40.0, gain=gain, coupling=coupling)
            tb, it, crc = ch.receive(rx)
            torch.cuda.synchronize()
            f = int((it > ch.max_iter).sum()); nfail += f; slots_ok += int(f == 0 and int(crc[0]) == 0); its.append(float(it.float().mean()))
            rails.append(round(float((ch.llr16.abs() >= 127).float().mean()), 2))
        print(json.dumps({"coupling": coupling, "gain": gain, "failed_cb_of_624": nfail, "slots_ok_of_12": slots_ok, "mean_it": round(float(np.mean(its)), 2), "rails": rails}), flush=True)
