#!/usr/bin/env python3
"""Summarise an .ncu-rep: curated raw metrics per kernel + hottest CUDA source lines.  usage: ncu_summary.py file.ncu-rep [kernel-regex]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; kre = sys.argv[2] if len(sys.argv) > 2 else None
K = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
     'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum', 'smsp__thread_inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
     'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
     'launch__grid_size', 'launch__block_size', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
     'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
     'lts__t_bytes.sum', 'l1tex__t_bytes.sum', 'smsp__average_warp_latency_per_inst_issued.ratio', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
     'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
     'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
     'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
     'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
     'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
     'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio',
     'smsp__average_warps_issue_stalled_selected_per_issue_active.ratio', 'smsp__thread_inst_executed_per_inst_executed.ratio']
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'] + (['-k', 'regex:' + kre] if kre else []), capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
seen = set()
for r in rows[2:]:
    name = r[hdr.index('Kernel Name')].split('(')[0]
    if name in seen: continue
    seen.add(name)
    print('=== ' + name)
    for k in K:
        if k in hdr: print('  %-88s %s' % (k, r[hdr.index(k)]))
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'] + (['-k', 'regex:' + kre] if kre else []), capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if 'Instructions Executed' in r]
if hi:
    h = rows[hi[0]]
    ie = h.index('Instructions Executed')
    ws = h.index('Warp Stall Sampling (All Samples)') if 'Warp Stall Sampling (All Samples)' in h else None
    acc = []
    for r in rows[hi[0] + 1:]:
        if len(r) > ie and r[0].isdigit():
            try: acc.append((float(r[ie]), float(r[ws]) if ws is not None and r[ws] not in ('', '-') else 0.0, r[0], r[1][:110]))
            except Exception: pass
    tot = sum(a[0] for a in acc) or 1; tots = sum(a[1] for a in acc) or 1
    print('--- hottest CUDA lines (inst%% / stall-sample%%), total warp inst %.4g' % tot)
    for v, s_, ln, text in sorted(acc, reverse=True)[:int(sys.argv[3]) if len(sys.argv) > 3 else 30]:
        print('  %5.1f%% %5.1f%%  L%s: %s' % (100 * v / tot, 100 * s_ / tots, ln, text.strip()))
