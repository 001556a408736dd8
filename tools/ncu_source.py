#!/usr/bin/env python
"""Print the SASS of one kernel from `ncu --page source --csv` output with per-instruction stall samples.
usage: ncu_source.py src.csv [pattern] [context]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
pat = sys.argv[2] if len(sys.argv) > 2 else 'UTCHMMA'
ctx = int(sys.argv[3]) if len(sys.argv) > 3 else 40
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
hdr = rows[hi]
body = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[0] != 'Address']
ia = hdr.index('Source'); isamp = hdr.index('# Samples'); iex = hdr.index('Instructions Executed')
stalls = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
def I(x):
    try: return int(x)
    except Exception: return 0
tot = sum(I(r[isamp]) for r in body)
print('total samples', tot, 'instructions', len(body))
idx = [i for i, r in enumerate(body) if pat in r[ia]]
print(pat, 'count', len(idx))
if pat == 'TOP':
    order = sorted(range(len(body)), key=lambda i: -I(body[i][isamp]))[:ctx]
    for i in sorted(order):
        r = body[i]; st = {hdr[j][6:]: I(r[j]) for j in stalls if I(r[j]) > 0}
        print(f"{i:5d} {I(r[isamp]):6d} {r[iex]:>9s}  {r[ia][:90]:90s} {st}")
    sys.exit()
lo = max(0, idx[0] - ctx); hi2 = min(len(body), idx[-1] + ctx)
for i in range(lo, hi2):
    r = body[i]; s = I(r[isamp])
    st = {hdr[j][6:]: I(r[j]) for j in stalls if I(r[j]) > 0}
    print(f"{i:5d} {s:6d} {r[iex]:>9s}  {r[ia][:100]:100s} {st if s > 100 else ''}")
