"""Per-region view of an `ncu --page source --csv` export (SASS level): instructions per warp per unit and sample share.
usage: python tools/ncu_sass_regions.py src.csv <units (e.g. warps x tiles)> [chunk] [kernel-name substring]"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
units = float(sys.argv[2]); chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 50
hdr = None; blocks = []
for r in rows:
    if not r: continue
    if r[0] == "Kernel Name": blocks.append((r[1], [])); continue
    if r[0] == "Address": hdr = r; continue
    blocks[-1][1].append(r)
want = sys.argv[4] if len(sys.argv) > 4 else None
k, rs = next(b for b in blocks if b[1] and (want is None or want in b[0]))
iS = hdr.index("# Samples"); iI = hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[iS]) for r in rs)
print(k, "SASS lines", len(rs), "samples", tot, "inst/unit", sum(int(r[iI]) for r in rs) / units)
for c in range(0, len(rs), chunk):
    seg = rs[c:c + chunk]
    n = sum(int(r[iI]) for r in seg) / units
    s = sum(int(r[iS]) for r in seg) / tot * 100
    if n < 0.05 and s < 0.05: continue
    ops = collections.Counter(r[1].split()[0] if not r[1].strip().startswith('@') else r[1].split()[1] for r in seg)
    st = collections.Counter()
    for r in seg:
        for i in stall_cols: st[hdr[i][6:]] += int(r[i])
    top = ", ".join(f"{a}:{100 * b / max(1, sum(st.values())):.0f}%" for a, b in st.most_common(3))
    print(f"{c:5d} inst/unit {n:7.1f} samp {s:5.1f}%  [{top}]  {dict(ops.most_common(5))}")
