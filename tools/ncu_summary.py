#!/usr/bin/env python3
"""Summarise an ncu report (raw + source CSV pages exported with `ncu -i X --page raw|source --csv`) for profiles/."""
import csv, collections, sys
raw=list(csv.reader(open(sys.argv[1]))); h=raw[0]; d=dict(zip(h,raw[2]))
units=dict(zip(h,raw[1]))
keys=['gpu__time_duration.sum','launch__grid_size','launch__block_size','launch__registers_per_thread','launch__shared_mem_per_block_dynamic','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__warps_eligible.avg.per_cycle_active']
print("| metric | value |\n|---|---|")
for k in keys: print(f"| {k} | {d.get(k)} {units.get(k,'')} |")
print("\n| stall reason (per issue-active) | ratio |\n|---|---|")
for k in h:
    if 'smsp__average_warps_issue_stalled' in k and k.endswith('per_issue_active.ratio'):
        v=float(d[k]); 
        if v>0.05: print(f"| {k.split('stalled_')[-1].replace('_per_issue_active.ratio','')} | {v:.2f} |")
rows=list(csv.reader(open(sys.argv[2]))); hdr=rows[1]
ix_src=hdr.index('Source'); ie=hdr.index('Instructions Executed'); ist=hdr.index('Warp Stall Sampling (All Samples)')
tot=0; by=collections.Counter(); st=collections.Counter(); tots=0
for r in rows[2:]:
    try: n=int(r[ie]); s=int(r[ist])
    except: continue
    t=r[ix_src].strip().split()
    if not t: continue
    op=t[1] if t[0].startswith('@') else t[0]
    by[op.split('.')[0]]+=n; tot+=n; st[op.split('.')[0]]+=s; tots+=s
ew=float(sys.argv[3]) if len(sys.argv)>3 else 1024*9*316*96/32
print(f"\ntotal warp instructions {tot}\n\n| opcode | share | stall-sample share | per edge-word |\n|---|---|---|---|")
for k,v in by.most_common(16): print(f"| {k} | {100*v/tot:.1f}% | {100*st[k]/tots:.1f}% | {v/ew:.2f} |")
