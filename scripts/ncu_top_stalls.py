"""Print the hottest SASS lines (by warp-stall samples) of one kernel from an ncu --page source --csv dump."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
si = hdr.index("Warp Stall Sampling (All Samples)"); src = hdr.index("Source")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_")]
data = []
tot = 0
for r in rows[2:]:
    try: s = int(r[si])
    except: continue
    tot += s
    data.append((s, r))
data.sort(key=lambda x: -x[0])
print("total samples", tot)
for s, r in data[: int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    top = sorted(((int(r[i] or 0), hdr[i]) for i in stall_cols), reverse=True)[:2]
    print(f"{s:6d} {100*s/tot:5.1f}%  {r[src][:90]:90s} {top}")
