"""Summarise an .ncu-rep (raw page) into the handful of counters that matter here."""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum', 'lts__t_sector_hit_rate.pct',
        'smsp__warps_eligible.avg.per_cycle_active', 'sm__cycles_elapsed.max',
        'sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active',
        'TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed']
for r in rows[2:]:
    print('-----', r[hdr.index('Kernel Name')], 'grid', r[hdr.index('Grid Size')] if 'Grid Size' in hdr else '')
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(f'  {w:72s} {r[i]:>18s} {units[i]}')
    st = []
    for i, h in enumerate(hdr):
        if h.startswith('smsp__pcsamp_warps_issue_stalled_') and not h.endswith('_not_issued'):
            try:
                st.append((float(r[i]), h.replace('smsp__pcsamp_warps_issue_stalled_', '')))
            except ValueError:
                pass
    tot = sum(v for v, _ in st) or 1
    print('  stalls: ' + ', '.join(f'{h} {100*v/tot:.0f}%' for v, h in sorted(st, reverse=True)[:8]))
