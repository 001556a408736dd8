#!/usr/bin/env python
"""Sums DRAM bytes / time over the launches of ONE lip_ggn_vp call from an
`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` log of bench.py (mlp workload) and writes
profiles/r02_traffic_ggn_vp.json (read by bench.py for roofline.traffic; r01_traffic_ggn_vp.json is round 1's capture).  usage: traffic_summary.py traffic_step.csv out.json"""
import collections, csv, json, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
rows = list(csv.DictReader(lines))
launch = collections.OrderedDict()
for r in rows:
    d = launch.setdefault(r['ID'], {'name': r['Kernel Name']})
    v = float(r['Metric Value'].replace(',', '')); u = r['Metric Unit']; n = r['Metric Name']
    if n.startswith('dram__bytes'):
        v *= {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
    else:
        v *= {'ns': 1e-3, 'us': 1, 'ms': 1e3, 'usecond': 1, 'msecond': 1e3, 'nsecond': 1e-3}.get(u, 1)
    d[n] = v
L = list(launch.values())
idx = [i for i, l in enumerate(L) if 'dot_partial' in l['name']]
grp = [l for l in L[idx[0] + 1:idx[1]] if ('lip::' in l['name'] or 'unnamed>::' in l['name']) and 'reduce_partials' not in l['name']
       and 'at::' not in l['name']]
rd = sum(l.get('dram__bytes_read.sum', 0) for l in grp); wr = sum(l.get('dram__bytes_write.sum', 0) for l in grp)
t = sum(l.get('gpu__time_duration.sum', 0) for l in grp)
by = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for l in grp:
    n = l['name'].split('(')[0][-60:]
    by[n][0] += 1; by[n][1] += l.get('gpu__time_duration.sum', 0); by[n][2] += l.get('dram__bytes_read.sum', 0); by[n][3] += l.get('dram__bytes_write.sum', 0)
out = {"what": "DRAM traffic of ONE lip_ggn_vp call (C3b, 256 Rademacher probes): sum over its launches, ncu --metrics "
               "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none (bench.py --no-cpu --no-slq "
               "--no-e2e --steps 1 --warmup 1, second call)",
       "launches": len(grp), "dram_bytes_read": rd, "dram_bytes_write": wr, "traffic_bytes": rd + wr, "serialized_us": t,
       "by_kernel": {k: {"launches": v[0], "us": v[1], "read_bytes": v[2], "write_bytes": v[3]} for k, v in by.items()}}
json.dump(out, open(sys.argv[2], 'w'), indent=1)
print(f"{len(grp)} launches, read {rd/1e9:.2f} GB, write {wr/1e9:.2f} GB, serialized {t:.0f} us")
for k, v in sorted(by.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1]:9.1f} us {v[0]:3d}x  R {v[2]/1e9:6.2f} GB  W {v[3]/1e9:6.2f} GB  {k}")
