"""Hot spots of one kernel from an ncu report's source page.
usage: python scripts/src_hot.py <rep> <kernel-regex> [min_pct]"""
import collections, csv, io, subprocess, sys
rep, pat = sys.argv[1:3]
minp = float(sys.argv[3]) if len(sys.argv) > 3 else 2.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
# first kernel block only
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
end = starts[1] if len(starts) > 1 else len(rows)
print(rows[starts[0]][1])
hdr, data = rows[starts[0] + 1], rows[starts[0] + 2:end]
iS, iN, iSm = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
ts = sum(int(r[iSm]) for r in data); tn = sum(int(r[iN]) for r in data)
print('samples', ts, 'warp-inst', tn, 'sass rows', len(data))
tot = collections.Counter()
for r in data:
    for j, h in enumerate(hdr):
        if h.startswith('stall_') and 'Not' not in h and r[j].isdigit(): tot[h] += int(r[j])
print([(k, round(100 * v / max(ts, 1), 1)) for k, v in tot.most_common(6)])
step = max(len(data) // 16, 1)
for a in range(0, len(data), step):
    seg = data[a:a + step]
    print(a, "inst %.1f%% samples %.1f%%" % (100 * sum(int(r[iN]) for r in seg) / tn, 100 * sum(int(r[iSm]) for r in seg) / ts))
for k, r in enumerate(data):
    sm = int(r[iSm])
    if sm > ts * minp / 100:
        st = {h: int(r[j]) for j, h in enumerate(hdr) if h.startswith('stall_') and 'Not' not in h and r[j].isdigit() and int(r[j]) > 0}
        print(k, r[iS].strip()[:55].ljust(55), r[iN], "%.2f%%" % (100 * sm / ts), sorted(st.items(), key=lambda x: -x[1])[:2])
