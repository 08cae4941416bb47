"""Per-instruction view of an .ncu-rep source page: executed-count segments, stall mix, top stalled SASS lines."""
import csv, subprocess, sys
from collections import Counter
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]
first = [r for r in rows[2:] if len(r) > 10 and r[5].isdigit()]
iS, iE, iSm = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
tot = sum(int(r[iE]) for r in first); tots = sum(int(r[iSm]) for r in first)
print(len(first), "sass lines,", tot, "warp instructions,", tots, "samples")
segs = []; cur = None
for i, r in enumerate(first):
    e = int(r[iE])
    if cur and abs(e - cur[2]) <= 0.02 * max(cur[2], 1): cur[1] = i; cur[3] += e; cur[4] += int(r[iSm])
    else:
        cur = [i, i, e, e, int(r[iSm])]; segs.append(cur)
segs.sort(key=lambda s: -s[4])
for s in segs[:16]:
    print(f'lines {s[0]:6d}-{s[1]:6d} n={s[1]-s[0]+1:4d} count~{s[2]:8d} inst={100*s[3]/tot:5.1f}% samples={100*s[4]/tots:5.1f}%  first: {first[s[0]][iS][:50]}')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
agg = Counter()
for r in first:
    for i in stall_cols:
        if r[i].isdigit(): agg[hdr[i]] += int(r[i])
print({k: round(100 * v / tots, 1) for k, v in agg.most_common(10)})
idx = sorted(range(len(first)), key=lambda i: -int(first[i][iSm]))[:int(sys.argv[2]) if len(sys.argv) > 2 else 20]
for i in idx:
    r = first[i]
    st = sorted(((int(r[c]), hdr[c]) for c in stall_cols if r[c].isdigit()), reverse=True)[:2]
    print(str(i).rjust(6), r[iSm].rjust(5), r[iE].rjust(8), r[iS][:62].ljust(62), st)
