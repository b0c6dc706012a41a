"""Per-kernel totals of an ncu launch list (``ncu --csv --metrics gpu__time_duration.sum[,dram__bytes_read.sum,...]``).

    python scripts/launch_summary.py profiles/r02_launches_frame_c2.csv [--json out.json]
"""
import csv
import json
import re
import sys
from collections import OrderedDict, defaultdict

path = sys.argv[1]
rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
h = rows[0]
ik, im, iv, iu, iid = (h.index(n) for n in ('Kernel Name', 'Metric Name', 'Metric Value', 'Metric Unit', 'ID'))
per = defaultdict(lambda: defaultdict(float))
count = defaultdict(set)
tot = defaultdict(float)
scale = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
for r in rows[1:]:
    name = re.sub(r'\(.*', '', r[ik])
    name = re.sub(r'^void |dboa::|\(anonymous namespace\)::', '', name)
    v = float(r[iv].replace(',', '')) * scale.get(r[iu], 1.0)
    per[name][r[im]] += v
    tot[r[im]] += v
    count[name].add(r[iid])
n_launch = sum(len(s) for s in count.values())
t_all = tot['gpu__time_duration.sum']
print(f'{path}: {n_launch} launches, sum of launch durations {t_all:.1f} us (cold-cache, serialised: shares, not absolutes)')
for m in ('dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum'):
    if m in tot:
        print(f'  {m:24s} {tot[m] / 1e6:10.2f} MB')
print(f'{"kernel":58s} {"n":>4s} {"us":>9s} {"share":>6s}' + ('   dram rd MB  wr MB' if 'dram__bytes_read.sum' in tot else ''))
for name, d in sorted(per.items(), key=lambda kv: -kv[1]['gpu__time_duration.sum']):
    t = d['gpu__time_duration.sum']
    line = f'{name[:58]:58s} {len(count[name]):4d} {t:9.1f} {t / t_all:6.1%}'
    if 'dram__bytes_read.sum' in d:
        line += f'   {d["dram__bytes_read.sum"] / 1e6:9.2f} {d["dram__bytes_write.sum"] / 1e6:6.2f}'
    print(line)
if '--json' in sys.argv:
    out = OrderedDict(source=path, launches=n_launch, sum_launch_us=t_all)
    for m in ('dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum'):
        if m in tot:
            out[m.split('.')[0].replace('__', '_')] = int(tot[m])
    json.dump(out, open(sys.argv[sys.argv.index('--json') + 1], 'w'), indent=1)
