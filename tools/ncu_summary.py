"""Developer tool: ncu report -> the markdown summary kept under profiles/.
usage: ncu_summary.py report.ncu-rep "<command line that was profiled>" > profiles/rNN_ncu_<kernel>.md"""
import csv, io, subprocess, sys
rep, cmd = sys.argv[1], sys.argv[2]
idx = int(sys.argv[3]) if len(sys.argv) > 3 else 0   # which captured launch
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h, units, r = rows[0], rows[1], rows[2 + idx]
g = lambda k: r[h.index(k)] if k in h else None
u = lambda k: units[h.index(k)] if k in h else ''
print('# ncu --set full: `%s`\n' % g('Kernel Name'))
print('`%s`\n' % cmd)
print('grid %s block %s\n' % (g('Grid Size'), g('Block Size')))
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active']
print('| metric | value | unit |\n|---|---|---|')
for k in want:
    if g(k) is not None: print('| %s | %s | %s |' % (k, g(k), u(k)))
st = [(float(r[i]), k) for i, k in enumerate(h) if k.startswith('smsp__average_warps_issue_stalled_') and k.endswith('_per_issue_active.ratio') and r[i]]
st.sort(reverse=True)
print('\nTop stall reasons (warps per issue-active cycle): ' + ', '.join('%s=%.2f' % (k.split('stalled_')[1].split('_per_issue')[0], v) for v, k in st[:6]))
