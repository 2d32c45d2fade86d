#!/usr/bin/env python3
"""BASELINE config 5 shape: one gNB slot period's worth of PUSCH for 16 UEs (100 MHz, 4 rx, 2 layers each: channel estimation, MMSE receiver, rate recovery,
LDPC decode with CRC stop, TB CRC) sharded across the GPUs of one node.  A UE's transport block -- hence every one of its code blocks and its HARQ soft buffers --
stays on one GPU (shard.sticky_gpu(ue, 0, world): the sticky rule of SURVEY.md 8e at UE granularity); the data path needs no collective.  NCCL carries the
init-time broadcast of the base-graph tables (as in bench.py) and the max-over-ranks reduction of the device time.  Launch:
  python tools/bench_multi_ue.py                                                             (1 GPU)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_multi_ue.py --gpus N
Prints one JSON line on rank 0: UE-slots/s over all GPUs, device resident and with every slot's samples / transport block crossing PCIe."""
import argparse
import json
import os
import sys
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from openairinterface5g_b200.ldpc import load_LDPClib          # noqa: E402
from openairinterface5g_b200.dfts import load_dftslib           # noqa: E402
from openairinterface5g_b200.shard import sticky_gpu            # noqa: E402
from openairinterface5g_b200.slot_chain import PuschSlotPipeline   # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--ues", type=int, default=16)
    ap.add_argument("--rounds", type=int, default=24)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        import ctypes
        saved = os.dup(1)                     # NCCL's version banner goes to stderr (see bench.py)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            ctypes.CDLL(None).fflush(None)
            os.dup2(saved, 1)
            os.close(saved)
    lib, dl = load_LDPClib(), load_dftslib()
    mine = [ue for ue in range(args.ues) if sticky_gpu(ue, 0, world) == rank]
    pipe = PuschSlotPipeline(lib, dl, dev, len(mine), seed0=500 + 100 * rank, A=471272, n_layers=2) if mine else None
    res = {}
    for mode, e2e in (("device_resident", False), ("e2e", True)):
        if dist is not None:
            dist.barrier()
        ms = pipe.timed_rounds(args.rounds, e2e=e2e) if pipe else 0.0
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        ok = torch.tensor([sum(pipe.check(host=e2e)) if pipe else 0], dtype=torch.int64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.all_reduce(ok, op=dist.ReduceOp.SUM)
        res[mode] = {"ue_slots_per_s": args.ues * args.rounds / (float(t.item()) / 1e3), "decoded_ok": f"{int(ok.item())}/{args.ues}"}
    if rank == 0:
        print(json.dumps({"workload": f"{args.ues} UEs x PUSCH slot 100MHz 273PRB 64QAM 4rx 2 layers, 56 CB K=8448 each, sharded by UE over {world} GPU(s)",
                          "n_gpus": world, "ues_per_gpu": [sum(1 for ue in range(args.ues) if sticky_gpu(ue, 0, world) == r) for r in range(world)],
                          "rounds": args.rounds, **res,
                          "gnb_slot_periods_per_s": res["device_resident"]["ue_slots_per_s"] / args.ues,
                          "realtime_factor_vs_2000_slots_per_s": res["device_resident"]["ue_slots_per_s"] / args.ues / 2000.0}), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
