"""Summarise an ncu launch-list CSV (gpu__time_duration.sum per launch)."""
import csv, sys
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ui = hdr.index('Metric Unit')
tot = 0
for r in rows[1:]:
    v = float(r[vi].replace(',', ''))
    if r[ui] == 'ns': v /= 1000.0
    tot += v
    print(f"{r[ki].split('(')[0][:34]:34s} {v:9.1f} us")
print(f"{'TOTAL':34s} {tot:9.1f} us")
