"""Print the hottest SASS instructions (by warp-stall samples) of an `ncu --page source --csv` export."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.006
hdr = rows[1]
isrc = hdr.index('Source')
ist = hdr.index('Warp Stall Sampling (All Samples)')
iex = hdr.index('Instructions Executed')
data = [r for r in rows[2:] if len(r) > ist and r[ist].isdigit()]
tot = sum(int(r[ist]) for r in data)
print('total samples', tot, 'instructions', len(data), 'executed', sum(int(r[iex]) for r in data if r[iex].isdigit()))
for n, r in enumerate(data):
    s = int(r[ist])
    if s > tot * frac:
        print(f"{n:5d} {100.0 * s / tot:5.1f}%  exec={r[iex]:>8s}  {r[isrc].strip()[:90]}")
