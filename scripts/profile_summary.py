"""Summarise an ncu report (--set full) and a per-launch CSV into profiles/<tag>_*.{csv,md}.
usage: python scripts/profile_summary.py <tag> <prof.ncu-rep> <launches.csv> [bench.json]"""
import collections, csv, io, json, subprocess, sys

tag, rep, launches = sys.argv[1:4]
bench = sys.argv[4] if len(sys.argv) > 4 else None
data_by_rep = []
for one in rep.split(","):                      # several reports (one per kernel filter) may be given, comma-separated
    raw = subprocess.run(["ncu", "-i", one, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    data_by_rep.append((rows[0], rows[1], rows[2:]))
cols = collections.OrderedDict([
    ("kernel", "Kernel Name"), ("grid", "launch__grid_size"), ("block", "launch__block_size"),
    ("regs", "launch__registers_per_thread"), ("time_us", "gpu__time_duration.sum"),
    ("dram_read_MB", "dram__bytes_read.sum"), ("dram_write_MB", "dram__bytes_write.sum"),
    ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("l1_hit_pct", "l1tex__t_sector_hit_rate.pct"), ("l2_hit_pct", "lts__t_sector_hit_rate.pct"),
    ("lsu_wavefront_pct", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed"),
    ("issue_active_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("warp_inst", "smsp__inst_executed.sum"),
])
def conv(v, u, want):
    try: x = float(v)
    except ValueError: return v
    scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
    if want.endswith("_MB") and u in scale: x *= scale[u]
    if want == "time_us": x *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1.0)
    return round(x, 3)
out = []
for hdr, units, data in data_by_rep:
    for r in data:
        rec = collections.OrderedDict()
        for k, name in cols.items():
            if name in hdr:
                i = hdr.index(name)
                rec[k] = r[i][:60] if k == "kernel" else conv(r[i], units[i], k)
        out.append(rec)
with open("profiles/%s_ncu_full.csv" % tag, "w", newline="") as f:
    w = csv.DictWriter(f, fieldnames=list(out[0].keys())); w.writeheader(); w.writerows(out)
# launch list aggregate
agg = collections.OrderedDict()
for row in csv.DictReader(l for l in open(launches) if not l.startswith("==")):
    k = row["Kernel Name"][:70]; v = float(row["Metric Value"]); u = row["Metric Unit"]
    v = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)
    agg.setdefault(k, []).append(v)
with open("profiles/%s_launches.csv" % tag, "w", newline="") as f:
    w = csv.writer(f); w.writerow(["kernel", "launches", "avg_us", "min_us", "max_us", "total_us"])
    for k, v in agg.items():
        w.writerow([k, len(v), round(sum(v) / len(v), 2), round(min(v), 2), round(max(v), 2), round(sum(v), 1)])
with open("profiles/%s_summary.md" % tag, "w") as f:
    f.write("# ncu summary %s\n\n" % tag)
    f.write("Source: `%s` (ncu --set full --clock-control none, cold-cache, serialised) and `%s` "
            "(gpu__time_duration.sum per launch, eager run of `bench.py --no-graph`).\n\n" % (rep, launches))
    if bench:
        b = json.load(open(bench))
        f.write("Bench line of the same build (`%s`): value %.0f %s, %.3f ms/step, e2e %.0f, roofline %s frac %.3f "
                "(achieved %.0f GB/s of %.0f), path_frac %.3f, clocks %s\n\n" % (
                    bench, b["value"], b["unit"], b["ms_per_step"], b["e2e"]["value"], b["roofline"]["kernel"],
                    b["roofline"]["frac"], b["roofline"]["achieved"], b["roofline"]["peak"],
                    b["roofline"].get("path_frac", 0), json.dumps(b.get("clocks"))))
    f.write("## ncu --set full (one launch per kernel)\n\n| " + " | ".join(out[0].keys()) + " |\n|" + "---|" * len(out[0]) + "\n")
    seen = set()
    for rec in out:
        key = (rec["kernel"], rec.get("grid"))
        if key in seen: continue
        seen.add(key)
        f.write("| " + " | ".join(str(v) for v in rec.values()) + " |\n")
    f.write("\n## launch list (b2d kernels; share of the step = total_us / sum)\n\n| kernel | launches | avg_us | total_us |\n|---|---|---|---|\n")
    tot = sum(sum(v) for k, v in agg.items() if "b2d::" in k and "roi_align_any" not in k)
    for k, v in agg.items():
        if "b2d::" in k:
            f.write("| %s | %d | %.1f | %.1f (%.1f%%) |\n" % (k, len(v), sum(v) / len(v), sum(v), 100 * sum(v) / tot))
# per-launch DRAM traffic of each stage's dominant kernel (bench.py roofline.traffic)
stage_kernel = {"roi_align": "k_roi_align", "proposals": "k_nms_mask", "rpn_targets": "k_assign_label", "roi_targets": "k_assign_label"}
traffic = {}
for st, pat in stage_kernel.items():
    for rec in out:
        if pat in rec["kernel"] and "dram_read_MB" in rec:
            traffic[st] = int((rec["dram_read_MB"] + rec["dram_write_MB"]) * 1e6)
            break
if traffic:                                      # a report without the step's dominant kernel (e.g. a K6-only capture) leaves it alone
    traffic["source"] = "profiles/%s_ncu_full.csv" % tag
    json.dump(traffic, open("profiles/traffic.json", "w"), indent=1)
print("wrote profiles/%s_*" % tag)
