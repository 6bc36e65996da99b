"""Key numbers of one kernel from an .ncu-rep (ncu --set full): duration, instructions, issue / pipe utilisation, DRAM bytes,
occupancy, registers and the warp-stall breakdown.   python scripts/ncu_summary.py <file.ncu-rep> [title]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, units, v = rows[0], rows[1], rows[2]
d = {k: (v[i], units[i]) for i, k in enumerate(h)}


def g(k):
    return d.get(k, ('n/a', ''))


if len(sys.argv) > 2:
    print(sys.argv[2])
print('kernel                                  %s' % g('Kernel Name')[0])
print('grid x block (cluster)                  %s x %s (%s)' % (g('launch__grid_size')[0], g('launch__block_size')[0], g('launch__cluster_dim_x')[0]))
for k, label in [('gpu__time_duration.sum', 'gpu__time_duration.sum'),
                 ('smsp__inst_executed.sum', 'smsp__inst_executed.sum (warp instructions)'),
                 ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue slots busy (% of active cycles)'),
                 ('sm__warps_active.avg.pct_of_peak_sustained_active', 'achieved occupancy (% of max warps)'),
                 ('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'FMA pipe active (%)'),
                 ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor pipe active (%)'),
                 ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM throughput (% of peak)'),
                 ('dram__bytes_read.sum', 'dram__bytes_read.sum'), ('dram__bytes_write.sum', 'dram__bytes_write.sum'),
                 ('dram__throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM throughput (% of peak)'),
                 ('launch__registers_per_thread', 'registers / thread'),
                 ('launch__shared_mem_per_block_dynamic', 'dynamic smem / CTA'),
                 ('sm__cycles_elapsed.avg', 'sm__cycles_elapsed.avg')]:
    val, un = g(k)
    print('%-40s%s %s' % (label, val, un))
st = []
for i, k in enumerate(h):
    if 'issue_stalled' in k and k.endswith('_per_issue_active.ratio'):
        try:
            st.append((float(v[i]), k.split('issue_stalled_')[1].replace('_per_issue_active.ratio', '')))
        except ValueError:
            pass
print('warp-stall reasons (average stalled warps per issued instruction):')
for x, k in sorted(st, reverse=True)[:9]:
    print('    %-22s %.3f' % (k, x))
