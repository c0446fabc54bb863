#!/usr/bin/env python
"""Table of an `ncu --metrics ... --csv` launch list: one line per kernel launch with the metrics abbreviated."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
i0 = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr = rows[i0]; data = rows[i0 + 1:]
ki, mi, vi, idi = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('ID')
ks = collections.OrderedDict()
for r in data:
    if len(r) <= vi: continue
    k = ks.setdefault(r[idi], {'name': r[ki].replace('<unnamed>::', '').replace('void ', '')[:36]})
    k[r[mi]] = r[vi]
short = {'gpu__time_duration.sum': 'us', 'dram__bytes_read.sum': 'MBrd', 'dram__throughput.avg.pct_of_peak_sustained_elapsed': 'dram%', 'launch__grid_size': 'grid',
         'sm__cycles_active.min': 'cmin', 'sm__cycles_active.max': 'cmax', 'sm__cycles_active.avg': 'cavg', 'sm__cycles_elapsed.max': 'cel',
         'smsp__issue_active.avg.pct_of_peak_sustained_active': 'issue%', 'sm__warps_active.avg.pct_of_peak_sustained_active': 'warps%',
         'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed': 'smem%', 'sm__inst_executed_pipe_tensor.sum.pct_of_peak_sustained_active': 'tensor%',
         'lts__t_sector_hit_rate.pct': 'l2hit'}
def f(x):
    try: return float(x.replace(',', ''))
    except Exception: return x
for id_, k in ks.items():
    out = [f"{k['name']:36s}"]
    for m, s in short.items():
        if m not in k: continue
        v = f(k[m])
        if s == 'us': v = v / 1e3
        if s == 'MBrd': v = v / 1e6
        out.append(f"{s}={v:.1f}" if isinstance(v, float) else f"{s}={v}")
    if 'gpu__time_duration.sum' in k and 'dram__bytes_read.sum' in k:
        out.append(f"GB/s={f(k['dram__bytes_read.sum']) / f(k['gpu__time_duration.sum']):.0f}")
    print(' '.join(out))
