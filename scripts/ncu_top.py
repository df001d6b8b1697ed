"""Summarise an ncu report's SASS source page: total stall breakdown and the hottest instructions."""
import csv, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
lines = out.splitlines()
# first kernel only
start = [i for i, l in enumerate(lines) if l.startswith('"Kernel Name"')]
seg = lines[start[0] + 1:(start[1] if len(start) > 1 else len(lines))]
rows = list(csv.reader(seg))
hdr = rows[0]; rows = [r for r in rows[1:] if len(r) == len(hdr)]
ci = {h: i for i, h in enumerate(hdr)}
S = ci["# Samples"]
tot = sum(int(r[S]) for r in rows)
print("instructions", len(rows), "samples", tot, "inst executed", sum(int(r[ci["Instructions Executed"]]) for r in rows))
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(int(r[ci[h]]) for r in rows) for h in stalls}
for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:10]:
    print("  %-28s %6.2f%%" % (h, 100.0 * v / max(1, tot)))
print("hottest instructions:")
order = sorted(range(len(rows)), key=lambda i: -int(rows[i][S]))[:topn]
for i in sorted(order):
    r = rows[i]
    top = sorted(stalls, key=lambda h: -int(r[ci[h]]))[:2]
    print("%5d %5.2f%% exec=%-8s %-60s %s" % (i, 100.0 * int(r[S]) / tot, r[ci["Instructions Executed"]], r[ci["Source"]].strip()[:60],
                                       ",".join("%s=%s" % (h[6:], r[ci[h]]) for h in top)))
