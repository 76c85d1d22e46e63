"""Turn gpurun_out/prof_<tag>.ncu-rep (+ launches_<tag>.csv, bench_<tag>.json) into the tracked evidence
under profiles/: a per-kernel markdown table, traffic.json (dram bytes per launch, per stage) and copies of
the launch list / bench line.   Usage: python scripts/make_profile_summary.py <tag>"""
import csv, io, json, os, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
rep = os.path.join(ROOT, "gpurun_out", f"prof_{tag}.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
cols = {
    "Kernel Name": "kernel", "gpu__time_duration.sum": "us (under ncu)", "launch__grid_size": "grid",
    "launch__registers_per_thread": "regs", "dram__bytes_read.sum": "dram rd", "dram__bytes_write.sum": "dram wr",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram %", "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm %",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue %", "sm__warps_active.avg.pct_of_peak_sustained_active": "warps %",
    "lts__t_sector_hit_rate.pct": "L2 hit %", "smsp__inst_executed.sum": "warp inst",
}
idx = {k: hdr.index(k) for k in cols if k in hdr}
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
stage_of = lambda k: ("keygen" if "keygen" in k else "project" if "project" in k else "raster" if "raster" in k else
                      "bin" if ("bin_emit" in k or "tile_ranges" in k) else "sort")
lines = [f"# ncu --set full summary, tag {tag} (one frame of bench.py C3: 6M f16, 1080p, scale 0.02)", "",
         "Times under ncu are cold-cache and serialised: compare SHARES; bench.py's stage times are the live ones.", "",
         "| " + " | ".join(cols[k] + (f" [{units[idx[k]]}]" if units[idx[k]] else "") for k in idx) + " |",
         "|" + "---|" * len(idx)]
traffic = {}
warp_inst = {}
order = []
for r in rows[2:]:
    vals = []
    for k in idx:
        v = r[idx[k]]
        if k == "Kernel Name":
            v = v.split("(")[0].replace("bgs::", "").replace("void ", "")
        else:
            try: v = f"{float(v.replace(',', '')):.1f}"
            except ValueError: pass
        vals.append(v)
    lines.append("| " + " | ".join(vals) + " |")
    name = r[idx["Kernel Name"]]
    def to_bytes(col):
        v = float(r[idx[col]].replace(",", "")); u = units[idx[col]].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
    order.append((name, to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")))
    if "smsp__inst_executed.sum" in idx:
        st0 = stage_of(name)
        if st0 in ("keygen", "project", "raster"):
            warp_inst[st0] = warp_inst.get(st0, 0) + int(float(r[idx["smsp__inst_executed.sum"]].replace(",", "")))
# the pair sort's onesweep passes come after bin_emit in launch order: attribute them to "bin"
seen_bin = False
for name, b in order:
    st = stage_of(name)
    if "bin_emit" in name: seen_bin = True
    if st == "sort" and seen_bin: st = "bin"
    traffic[st] = traffic.get(st, 0) + b
lines += ["", "dram bytes per frame by stage (read + write): " + ", ".join(f"{k} {v/1e6:.1f} MB" for k, v in traffic.items())]
open(os.path.join(ROOT, "profiles", f"{tag}_ncu_summary.md"), "w").write("\n".join(lines) + "\n")
out = {k: int(v) for k, v in traffic.items()}
out["warp_inst"] = warp_inst          # executed warp-instructions per launch (issue roofline of the compute-bound kernels)
json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
for f in (f"launches_{tag}.csv", f"bench_{tag}.json", f"pytest_gpu_{tag}.log"):
    p = os.path.join(ROOT, "gpurun_out", f)
    if os.path.exists(p): shutil.copy(p, os.path.join(ROOT, "profiles", f))
print("\n".join(lines))
