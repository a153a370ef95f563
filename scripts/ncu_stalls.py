"""Aggregate the warp-stall samples of one kernel from `ncu -i X.ncu-rep --page source --csv` (SASS view)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h = rows[1]
data = [r for r in rows[2:] if len(r) == len(h) and r[0].startswith('0x')]
ix = {n: i for i, n in enumerate(h)}
stalls = [n for n in h if n.startswith('stall_') and 'Not Issued' not in n]
tot = sum(int(r[ix['# Samples']]) for r in data)
print('total samples', tot, 'instructions', len(data))
agg = {s: sum(int(r[ix[s]]) for r in data) for s in stalls}
for s, v in sorted(agg.items(), key=lambda kv: -kv[1])[:12]:
    print(f'{s:28s} {v:8d} {100 * v / tot:5.1f}%')
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
for r in sorted(data, key=lambda r: -int(r[ix['# Samples']]))[:n]:
    st = sorted(((int(r[ix[s]]), s[6:]) for s in stalls), reverse=True)[:2]
    print(r[ix['Address']][-5:], r[ix['# Samples']].rjust(6), r[ix['Instructions Executed']].rjust(9), r[ix['Source']].strip()[:64].ljust(64), st)
