"""Turn the round's ncu captures under gpurun_out/ into the tracked tables of profiles/ (bring-up tool).
usage: python tools/make_profiles.py   (expects gpurun_out/r02_final.ncu-rep and gpurun_out/r02_launches_final.csv)"""
import collections
import csv
import json
import os
import subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, 'gpurun_out'), os.path.join(ROOT, 'profiles')
raw = subprocess.run(['ncu', '-i', os.path.join(G, 'r02_final.ncu-rep'), '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = rows[0]
col = lambda n: h.index(n)
names = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
         'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
         'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
         'launch__grid_size', 'launch__block_size', 'smsp__inst_executed.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
         'smsp__sass_inst_executed_op_global_ld.sum', 'smsp__sass_inst_executed_op_global_st.sum', 'sm__cycles_elapsed.max',
         'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
         'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
         'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
         'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'lts__t_sector_hit_rate.pct']
out = collections.OrderedDict()
for r in rows[2:]:
    out[r[col('Kernel Name')].split('(')[0]] = {n: (r[col(n)], rows[1][col(n)]) for n in names if n in h}
md = ['| metric | ' + ' | '.join(out) + ' |', '|---|' + '---|' * len(out)]
for n in names:
    md.append('| `' + n + '` | ' + ' | '.join(f'{float(v[n][0]):,.2f} {v[n][1]}' for v in out.values()) + ' |')
g = lambda k, n: float(out[k][n][0]) if k in out else 0.0
fwd = sum(g(k, 'dram__bytes_read.sum') + g(k, 'dram__bytes_write.sum') for k in ('sample_knn_kernel', 'render_fwd_kernel'))
bwd = sum(g(k, 'dram__bytes_read.sum') + g(k, 'dram__bytes_write.sum') for k in ('trunk_bwd_umma_kernel', 'geo_bwd_umma_kernel', 'render_bwd_kernel'))
json.dump({'source': 'ncu --set full --clock-control none, one launch per kernel (profiles/r02_ncu_summary.md)',
           'unit': 'bytes per launch of the C-ABI call (dram__bytes_read.sum + dram__bytes_write.sum of its main kernels)',
           'fwd_bytes': fwd * 1e6, 'bwd_bytes': bwd * 1e6,
           'per_kernel_MB': {k: {'read': g(k, 'dram__bytes_read.sum'), 'write': g(k, 'dram__bytes_write.sum')} for k in out}},
          open(os.path.join(P, 'r02_ncu_traffic.json'), 'w'), indent=1)
open(os.path.join(G, 'r02_table.md'), 'w').write('\n'.join(md))
# launch list: one period of the literal iteration
rows = [r for r in csv.reader(open(os.path.join(G, 'r02_launches_final.csv'))) if len(r) > 10]
h = rows[0]
ki, vi, ii = h.index('Kernel Name'), h.index('Metric Value'), h.index('ID')
with open(os.path.join(P, 'r02_launches_final.csv'), 'w', newline='') as f:
    w = csv.writer(f)
    w.writerow(['ID', 'Kernel Name', 'gpu__time_duration.sum (ns)'])
    for r in rows[1:]:
        w.writerow([r[ii], r[ki][:110], r[vi]])
seq = [(r[ki].split('(')[0], float(r[vi].replace(',', ''))) for r in rows[1:]]
fw = [i for i, (n, v) in enumerate(seq) if n.endswith('render_fwd_kernel')]
period = fw[1] - fw[0]
a, b = len(seq) - period, len(seq)
tot = sum(v for n, v in seq[a:b])
agg = collections.OrderedDict()
for n, v in seq[a:b]:
    key = n if n.startswith('lsr::') else 'torch / ATen kernels of the caller-side ops (randint, cat, median, index_put, loss, zero-fill, ...)'
    agg.setdefault(key, [0, 0.0])
    agg[key][0] += 1
    agg[key][1] += v
tab = '\n'.join(f'| `{k}` | {c} | {v / 1e3:.1f} | {100 * v / tot:.1f} % |' for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]))
open(os.path.join(G, 'step_table.md'), 'w').write(
    f'{b - a} launches = one literal mapping iteration (one period of the launch list), {tot / 1e3:.1f} us of kernel time (cold-cache, '
    f'serialised under ncu)\n\n| kernel | launches | us | share |\n|---|---|---|---|\n' + tab)
print(open(os.path.join(G, 'r02_table.md')).read())
print(open(os.path.join(G, 'step_table.md')).read())
print(json.load(open(os.path.join(P, 'r02_ncu_traffic.json'))))
