"""Per-CUDA-line hot spots from an ncu report (needs -lineinfo and --import-source on).
usage: python scripts/src_lines.py <rep> [min_pct] [kernel-regex]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; minp = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]
if len(sys.argv) > 3: cmd += ["--kernel-name", "regex:" + sys.argv[3]]
raw = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
cur = None; hdr = None; recs = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and r[0].isdigit() and len(r) > 8 and r[2] == "-":
        recs.append((cur, int(r[0]), r[1], int(r[hdr.index("# Samples")]), int(r[hdr.index("Instructions Executed")])))
ts = sum(x[3] for x in recs); tn = sum(x[4] for x in recs)
print("samples", ts, "warp-inst", tn)
for f, ln, src, sm, n in recs:
    if sm > ts * minp / 100 or n > tn * minp / 100:
        print("%-18s %4d  inst %5.1f%%  smp %5.1f%%  %s" % (f, ln, 100 * n / max(tn, 1), 100 * sm / max(ts, 1), src.strip()[:100]))
