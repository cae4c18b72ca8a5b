#!/usr/bin/env python
"""Summarise an .ncu-rep: headline metrics per kernel + per-CUDA-line instruction / stall-sample shares.
usage: tools/ncu_summary.py report.ncu-rep [top_lines]"""
import csv, subprocess, sys, io

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__grid_size', 'launch__block_size', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'lts__t_bytes.sum', 'sm__cycles_elapsed.max']
for r in rows[2:]:
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(f"{w} = {r[i]} {units[i]}")
    for i, h in enumerate(hdr):
        if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio'):
            v = float(r[i])
            if v >= 0.15:
                print(f"  stall {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]}: {v:.2f}")
    print('---')
mix = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
fn = 0; agg = {}; h2 = None
for r in csv.reader(io.StringIO(mix)):
    if r and r[0] == "Function Name":
        fn += 1
        if fn > 1: break
        continue
    if r and r[0] == "Line No": h2 = r; continue
    if h2 is None or len(r) < 9: continue
    if r[2] == '-' and r[0].isdigit():
        try: s = int(r[6]); ins = int(r[7])
        except ValueError: continue
        a = agg.setdefault(int(r[0]), [0, 0, r[1].strip()]); a[0] += ins; a[1] += s
ti = sum(a[0] for a in agg.values()) or 1; ts = sum(a[1] for a in agg.values()) or 1
print("total warp-inst", ti, "samples", ts)
for ln, (ins, s, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{ins/ti*100:5.1f}% inst {s/ts*100:5.1f}% samp  L{ln}: {src[:105]}")
