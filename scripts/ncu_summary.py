#!/usr/bin/env python
"""Compact summary of one `ncu --set full` capture: usage ncu_summary.py report.ncu-rep [--json out.json --model-steps N --source TEXT]
With --json also writes the numbers bench.py's roofline object reads (profiles/roofline_ncu.json)."""
import csv, json, subprocess, sys
rep = sys.argv[1]
def _opt(name, default=None):
    return sys.argv[sys.argv.index(name) + 1] if name in sys.argv else default
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
def g(k):
    return float(m[k][1].replace(',', '')) if k in m else float('nan')
keep = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sass__inst_executed_local_loads', 'sass__inst_executed_local_stores',
        'sm__cycles_elapsed.max', 'sm__cycles_elapsed.max.per_second']
for k in keep:
    if k in m:
        print('%-85s %12s %s' % (k, m[k][1], m[k][0]))
for k in sorted(m):
    if k.startswith('smsp__average_warps_issue_stalled') and k.endswith('per_issue_active.ratio'):
        print('%-85s %12s' % (k, m[k][1]))
cyc = g('sm__cycles_elapsed.max')
tot = 0
for op, w in (('dfma', 2), ('dmul', 1), ('dadd', 1)):
    k = 'smsp__sass_thread_inst_executed_op_%s_pred_on.sum.per_cycle_elapsed' % op
    if k in m:
        n = g(k) * cyc
        tot += w * n
        print('executed %s thread-instructions: %.4e' % (op.upper(), n))
t = g('gpu__time_duration.sum')
unit = m['gpu__time_duration.sum'][0]
sec = t * {'ms': 1e-3, 'us': 1e-6, 's': 1, 'ns': 1e-9}.get(unit, 1e-3)
print('executed FP64 flop (2*DFMA + DMUL + DADD): %.4e in %.4f s -> %.2f TFLOP/s executed' % (tot, sec, tot / sec / 1e12))

if _opt('--json'):
    dram = (g('dram__bytes_read.sum') * {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1}.get(m['dram__bytes_read.sum'][0], 1) +
            g('dram__bytes_write.sum') * {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1}.get(m['dram__bytes_write.sum'][0], 1))
    out = dict(dram_bytes_per_launch=int(dram), fp64_flop_per_launch=tot, model_steps_per_launch=int(float(_opt('--model-steps', '0'))),
               fp64_pipe_active_pct=g('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed'),
               issue_active_pct=g('smsp__issue_active.avg.pct_of_peak_sustained_active'),
               warps_active_pct=g('sm__warps_active.avg.pct_of_peak_sustained_active'), kernel_ms_under_ncu=sec * 1e3,
               source=_opt('--source', rep) + ': one fb200_evolve_kernel launch under ncu --set full; fp64_flop = 2 x DFMA + DMUL + DADD executed '
                      'thread-instructions; fp64_pipe_active = sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed')
    json.dump(out, open(_opt('--json'), 'w'), indent=1)
