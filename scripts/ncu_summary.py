#!/usr/bin/env python
"""Summarise an .ncu-rep (first kernel): key raw metrics + top stalled SASS lines.
usage: scripts/ncu_summary.py file.ncu-rep [n_top]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
d = dict(zip(rows[0], rows[2])); u = dict(zip(rows[0], rows[1]))
keys = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'smsp__inst_executed.sum',
        # tensor pipe (no tensor counter appeared in any profile)
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor_subpipe_imma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum', 'sm__inst_executed_pipe_uniform.sum',
        'l1tex__data_pipe_tc_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum', 'lts__t_sectors_srcunit_tex_lookup_hit.sum',
        'lts__t_sectors_srcunit_tex_lookup_miss.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__sass_thread_inst_executed_op_ffma_pred_on.sum', 'sm__cycles_elapsed.max', 'sm__cycles_active.avg']
for k in keys:
    if k in d: print(f"{k} = {d[k]} {u.get(k,'')}")
for k in d:  # 'realtime' tensor counters carry a section prefix in the raw page
    if 'pipe_tensor' in k and 'realtime' in k: print(f"{k} = {d[k]} {u.get(k,'')}")
try:
    c = float(d['l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']); w = float(d['l1tex__data_pipe_lsu_wavefronts_mem_shared.sum'])
    print(f"shared-memory bank-conflict ratio (LSU): {c:.0f} conflict wavefronts of {w:.0f} = {c / w:.3f}")
except Exception:
    pass
for k in sorted(d):
    if 'issue_stalled' in k and k.endswith('per_issue_active.ratio') and float(d[k] or 0) > 0.05:
        print(f"  {k.split('issue_stalled_')[1].split('_per_issue')[0]:>22s} {float(d[k]):.3f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}; data = rows[2:]
S = lambda r: int(r[ix['# Samples']]); IE = lambda r: int(r[ix['Instructions Executed']])
print('total samples', sum(S(r) for r in data), 'warp instructions', sum(IE(r) for r in data))
ffma = [r for r in data if 'FFMA' in r[ix['Source']]]
lds = [r for r in data if 'LDS' in r[ix['Source']]]
print('FFMA: samples', sum(S(r) for r in ffma), 'inst', sum(IE(r) for r in ffma), '| LDS: samples', sum(S(r) for r in lds), 'inst', sum(IE(r) for r in lds))
cols = ['stall_long_sb', 'stall_short_sb', 'stall_barrier', 'stall_mio', 'stall_wait', 'stall_math', 'stall_not_selected', 'stall_dispatch']
tot = {c: sum(int(r[ix[c]]) for r in data) for c in cols}
print('stall totals', tot)
for r in sorted(data, key=lambda r: -S(r))[:ntop]:
    print(str(S(r)).rjust(5), str(IE(r)).rjust(9), ' '.join(f"{c[6:9]}{r[ix[c]]:>4s}" for c in cols), r[ix['Source']][:70])
