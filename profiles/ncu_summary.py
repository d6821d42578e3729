"""Print the handful of ncu metrics the round summaries quote, per kernel, from a .ncu-rep (read on the CPU box).
    python profiles/ncu_summary.py gpurun_out/x.ncu-rep"""
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'inst_executed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'smsp__sass_inst_executed_op_local_ld.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed']
raw = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print('-----', r[hdr.index('Kernel Name')][:90])
    for k in KEYS:
        if k in hdr:
            print(f'  {k} [{units[hdr.index(k)]}] {r[hdr.index(k)]}')
    st = []
    for i, h in enumerate(hdr):
        if 'smsp__average_warps_issue_stalled' in h and 'per_issue_active' in h and 'not_issued' not in h:
            try:
                st.append((float(r[i]), h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')))
            except ValueError:
                pass
    print('  stalls/issue:', ', '.join(f'{n} {v:.2f}' for v, n in sorted(st, reverse=True)[:6]))
    for k in ['l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed_op_shared_ld.sum',
              'smsp__inst_executed_op_shared_st.sum', 'smsp__inst_executed_op_ldgsts.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_lgds.sum',
              'l1tex__data_pipe_lsu_wavefronts.sum', 'smsp__inst_executed_op_global_st.sum']:
        if k in hdr:
            print(f'  {k} {r[hdr.index(k)]}')
