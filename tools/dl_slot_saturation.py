import sys, os, json, time, numpy as np, torch
sys.path.insert(0, "/root/repo")
from openairinterface5g_b200.ldpc import load_LDPClib
from openairinterface5g_b200.dfts import load_dftslib
from openairinterface5g_b200.dl_slot_chain import PdschSlotChain, PdschSlotPipeline
dev = torch.device("cuda", 0)
lib, dl = load_LDPClib(), load_dftslib()
K = 16
pipe = PdschSlotPipeline(lib, dl, dev, K)
for _ in range(3): pipe.round()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(60): pipe.round()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(json.dumps({"issue_us_per_slot": 1e6 * (t1 - t0) / (60 * K), "total_us_per_slot": 1e6 * (t2 - t0) / (60 * K)}), flush=True)
REP = 8
st = {
    "tb_crc+segmentation": lambda ch, p, rx: ch.lib.tb_segment_torch(1, ch.A, p, ch.segs, ch.crc1),
    "ldpc_encode": lambda ch, p, rx: lib.encode_batch_torch(1, ch.Z, ch.K, ch.segs, out=ch.cw),
    "rm_tx": lambda ch, p, rx: lib.rm_tx_torch(1, ch.Z, ch.Qm, 0, ch.C, 0, ch.F, ch.cw, ch.E, ch.Eoff, ch.f),
    "pdsch_tx": lambda ch, p, rx: lib.pdsch_tx_slot_torch(ch.txd, ch.f, ch.txF),
    "ofdm_mod": lambda ch, p, rx: dl.ofdm_mod_slot_torch(ch.dtx, ch.txF, ch.txdata),
    "ofdm_demod": lambda ch, p, rx: dl.ofdm_demod_slot_torch(ch.drx, rx, ch.ts, ch.rxF),
    "channel_estimation": lambda ch, p, rx: lib.pusch_chest_torch(ch.cdesc, ch.rxF, ch.est, ch.chest_scratch, ch.chest_state),
    "level+zf_rx": lambda ch, p, rx: lib.pusch_inner_rx_torch(ch.rxd, ch.rxF, ch.est, ch.llr16, level=ch.level),
    "rm_rx": lambda ch, p, rx: lib.rm_rx_torch(1, ch.Z, ch.Qm, 0, ch.C, 0, ch.F, ch.llr16, ch.E, ch.Eoff, ch.harq, ch.llr8, clear=1),
    "ldpc_decode": lambda ch, p, rx: lib.decode_batch_torch(1, ch.Z, ch.R, ch.max_iter, ch.llr8, use_crc=1, crc_len_bits=ch.K - ch.F, crc_type=1, out=ch.hard, iters=ch.iters),
    "tb copy + crc": lambda ch, p, rx: (ch.tb.view(-1).copy_(ch.hard[:, :ch.nbytes].reshape(-1)), lib.crc_batch_torch(0, ch.tb, ch.A + 24, out=ch.tbcrc)),
}
out = {}
cur = torch.cuda.current_stream(dev)
for name, f in st.items():
    graphs = []
    for k in range(K):
        s = pipe.streams[k]
        with torch.cuda.stream(s):
            f(pipe.chains[k], pipe.payload[k], pipe.rx[k]); s.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                for _ in range(REP):
                    f(pipe.chains[k], pipe.payload[k], pipe.rx[k])
        graphs.append(g)
    def rnd():
        for k in range(K):
            with torch.cuda.stream(pipe.streams[k]):
                graphs[k].replay()
    for _ in range(3): rnd()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(cur); pipe.fork(e0)
    n = 20
    for _ in range(n): rnd()
    pipe.join(cur); e1.record(cur)
    torch.cuda.synchronize()
    out[name] = round(1e3 * e0.elapsed_time(e1) / (n * K * REP), 2)
out["sum"] = round(sum(out.values()), 2)
print(json.dumps(out), flush=True)
