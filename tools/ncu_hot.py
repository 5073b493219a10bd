"""Top stall locations of one kernel from an ncu --page source --csv export:  python tools/ncu_hot.py file.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr, data = rows[1], rows[2:]
ia, isrc, iex = hdr.index('# Samples'), hdr.index('Source'), hdr.index('Instructions Executed')
stall = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[ia]) for r in data if r[ia].isdigit())
totex = sum(int(r[iex]) for r in data if r[iex].isdigit())
print('total samples', tot, 'total warp-instructions', totex)
agg = {}
for r in data:
    for i in stall:
        if r[i].isdigit(): agg[hdr[i][6:]] = agg.get(hdr[i][6:], 0) + int(r[i])
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
for r in sorted([r for r in data if r[ia].isdigit()], key=lambda r: -int(r[ia]))[:n]:
    st = {hdr[i][6:]: int(r[i]) for i in stall if r[i].isdigit() and int(r[i]) > 0}
    st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{int(r[ia]):6d} {100*int(r[ia])/tot:5.1f}% ex={r[iex]:>8s} {r[isrc].strip()[:72]:72s} {st}")
