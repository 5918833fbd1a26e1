"""Summarise `ncu -i X.ncu-rep --page source --csv`: stall-reason totals and the hottest SASS instructions.
usage: python tools/ncu_src_top.py src.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
data = []
for r in rows[2:]:  # a multi-kernel export repeats the two header rows: keep the first kernel only
    if len(r) < len(hdr):
        break
    data.append(r)
ci = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[ci['# Samples']]) for r in data)
print(rows[0][1], "total samples", tot, "instructions", len(data))
agg = {s: sum(int(r[ci[s]]) for r in data) for s in stalls}
for s, v in sorted(agg.items(), key=lambda kv: -kv[1])[:10]:
    print("  %-26s %7d %5.1f%%" % (s, v, 100 * v / tot))
for i, r in enumerate(data):
    r.append(i)
top = sorted(data, key=lambda r: -int(r[ci['# Samples']]))[:n]
for r in top:
    st = sorted(((int(r[ci[s]]), s[6:]) for s in stalls), reverse=True)[:3]
    print(r[ci['# Samples']].rjust(6), "#%-5d" % r[-1], r[1].strip()[:64].ljust(64), r[ci['Instructions Executed']].rjust(8), st)
