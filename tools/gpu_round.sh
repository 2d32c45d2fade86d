#!/usr/bin/env bash
# One gpurun call: GPU parity tests, the bench line, the ncu launch list and one full capture of the decode kernel.
# Usage (from the CPU box): gpurun --timeout 1800 -- 'bash tools/gpu_round.sh [tag] [quick]'
set -u
TAG=${1:-r01}
MODE=${2:-full}
mkdir -p gpurun_out
make -C oracle liboracle.so >/dev/null 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu_${TAG}.txt
nproc >> gpurun_out/gpu_${TAG}.txt; lscpu | grep -E "Model name|^CPU\(s\)" | cut -c1-200 >> gpurun_out/gpu_${TAG}.txt
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_${TAG}.txt
if [ "$MODE" = "variants" ]; then
echo "== variants"
for t in 768 672 576 480 384; do NRB200_PACKED_THREADS=$t timeout 120 python tools/kernel_time.py 1.0 2>&1 | tail -1; done | tee gpurun_out/variants_${TAG}.txt
timeout 120 python tools/kernel_time.py 3.0 2>&1 | tail -1 | tee -a gpurun_out/variants_${TAG}.txt
fi
echo "== micro-benchmark (ALU pipe / opcode-blend / code-footprint ceilings)"; timeout 120 tools/ubench/_bin/alu_ceiling | tee gpurun_out/alu_ceiling_${TAG}.json
echo "== extras"; timeout 400 python tools/bench_extras.py 2>&1 | tail -34 | tee gpurun_out/extras_${TAG}.jsonl
echo "== extras, OFDM front end with 592 antenna-slots per launch"; NRB200_OFDM_ANTENNA_SLOTS=592 timeout 300 python tools/bench_extras.py 2>&1 | grep ofdm_ | tee gpurun_out/extras_ofdm592_${TAG}.jsonl | cut -c1-260
echo "== ncu full (decode kernel): first, so that the bench line of this very run carries roofline.traffic / on_chip of the kernel it times"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ldpc_decode -s 3 -c 1 -f -o gpurun_out/prof_decode_${TAG} \
  python bench.py --steps 3 --warmup 3 --no-cpu --no-check --no-slot --no-ubench --nbuf 2 > gpurun_out/ncu_full_${TAG}.log 2>&1
ncu -i gpurun_out/prof_decode_${TAG}.ncu-rep --page raw --csv > gpurun_out/prof_decode_${TAG}_raw.csv 2>/dev/null \
  && python tools/ncu_traffic_json.py gpurun_out/prof_decode_${TAG}_raw.csv gpurun_out/ncu_decode_traffic_${TAG}.json 1024 "gpurun_out/prof_decode_${TAG}.ncu-rep" | cut -c1-300 \
  && cp gpurun_out/ncu_decode_traffic_${TAG}.json profiles/ncu_decode_traffic.json   # the bench below reads it (stamped with the kernel sources' SHA-1)
echo "== bench"; timeout 600 python bench.py --steps 50 --warmup 5 2>gpurun_out/bench_${TAG}.err | tee gpurun_out/bench_${TAG}.json
if [ "$MODE" = "full" ]; then
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench point B (Eb/N0 3 dB, early stop)"; timeout 300 python bench.py --steps 50 --warmup 5 --ebn0 3.0 --no-cpu 2>>gpurun_out/bench_${TAG}.err | tee gpurun_out/bench_${TAG}_pointB.json
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>>gpurun_out/bench_${TAG}.err | tee gpurun_out/bench_${TAG}_reference.json
echo "== slot chain"; timeout 200 python tools/bench_slot.py 2>&1 | tail -10 | tee gpurun_out/slot_${TAG}.jsonl
echo "== DL slot chain"; timeout 200 python tools/bench_dl_slot.py 2>&1 | tail -12 | tee gpurun_out/dlslot_${TAG}.jsonl
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_dlslot_${TAG}.csv \
  python tools/bench_dl_slot.py once > gpurun_out/ncu_launches_dlslot_${TAG}.log 2>&1
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_${TAG}.csv \
  python bench.py --steps 5 --warmup 3 --no-cpu --no-check --no-slot --no-ubench --nbuf 2 > gpurun_out/ncu_launches_${TAG}.log 2>&1
fi
echo "== per-call loader ABI (tools/abi_bench.c: this library and the compiled reference, same harness)"
timeout 300 python tools/bench_abi.py 2.0 1.0 2>&1 | tail -16 | tee gpurun_out/abi_${TAG}.jsonl
echo "== nr_ulsch_decoding through the OAI-side caller: reference function + CPU decoder vs interposer"
timeout 400 python tools/bench_ulsch_tb.py 2.0 2>&1 | tee gpurun_out/ulsch_tb_${TAG}.jsonl | cut -c1-420
echo "== cluster decoder: time of one small launch, phase marks"
for c in 8 4 2 0; do NRB200_CLUSTER=$c timeout 120 python tools/cluster_time.py 1.0 1 8 2>&1 | tail -2; done | tee gpurun_out/cluster_time_${TAG}.txt
timeout 120 python tools/cluster_time.py 1.0 2>&1 | tail -6 | tee -a gpurun_out/cluster_time_${TAG}.txt
timeout 120 python tools/cluster_phases.py 1.0 1 2>&1 | tail -9 | tee gpurun_out/cluster_phases_${TAG}.txt
echo "== DFT"
for inv in 1 0; do timeout 120 python tools/dft_time.py 4096 $inv 8880; NRB200_DFT_TMA=0 timeout 120 python tools/dft_time.py 4096 $inv 8880; done 2>&1 | tee gpurun_out/dft_${TAG}.txt
if [ "$MODE" = "full" ]; then
echo "== ncu full (cluster decoder, one block on 8 CTAs)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cluster -s 6 -c 1 -f -o gpurun_out/prof_cluster_${TAG} python tools/cluster_time.py 1.0 1 > gpurun_out/ncu_cluster_${TAG}.log 2>&1
echo "== ncu full (4096-point TMA kernel)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dft4096 -s 3 -c 1 -f -o gpurun_out/prof_dft4096_${TAG} python tools/dft_time.py 4096 1 8880 > gpurun_out/ncu_dft_${TAG}.log 2>&1
fi
ls -la gpurun_out | tail -8
if [ "$MODE" = "full" ]; then
echo "== rfsimulator channel kernel: time and one full ncu capture"
timeout 60 python tools/rfsim_time.py 2 40 614400 | tee gpurun_out/rfsim_${TAG}.txt
timeout 60 python tools/rfsim_time.py 4 40 614400 | tee -a gpurun_out/rfsim_${TAG}.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rfsim -s 5 -c 1 -f -o gpurun_out/prof_rfsim_${TAG} python tools/rfsim_time.py 2 40 614400 > gpurun_out/ncu_rfsim_${TAG}.log 2>&1
fi
