"""Join an ncu report's per-SASS-instruction counters with nvdisasm line info -> per-source-line instruction and
stall-sample totals.  usage: ncu_lines.py <report.ncu-rep> <cubin> <mangled-kernel-substring> [topN]"""
import csv, re, subprocess, sys
rep, cubin, kname = sys.argv[1:4]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
# find function section
lines_for_inst = []
cur_line = None
infunc = False
for l in dis:
    if l.startswith(".text.") or re.match(r"^\s*\.section\s+\.text\.", l):
        infunc = kname in l
        continue
    if not infunc:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur_line = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"^\s+/\*[0-9a-f]{4}\*/", l):
        lines_for_inst.append(cur_line)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout.splitlines()
start = [i for i, l in enumerate(out) if l.startswith('"Kernel Name"')]
rows = list(csv.reader(out[start[0] + 1:(start[1] if len(start) > 1 else len(out))]))
hdr = rows[0]; rows = [r for r in rows[1:] if len(r) == len(hdr)]
ci = {h: i for i, h in enumerate(hdr)}
print("sass instrs: ncu %d, nvdisasm %d" % (len(rows), len(lines_for_inst)))
agg = {}
tot_i = tot_s = 0
for i, r in enumerate(rows):
    key = lines_for_inst[i] if i < len(lines_for_inst) else None
    ie = int(r[ci["Instructions Executed"]]); sm = int(r[ci["# Samples"]])
    a = agg.setdefault(key, [0, 0]); a[0] += ie; a[1] += sm
    tot_i += ie; tot_s += sm
src_cache = {}
def src(key):
    if not key: return ""
    f, n = key
    import glob
    if f not in src_cache:
        c = glob.glob("/root/repo/research_new_hnsw_b200/csrc/" + f)
        src_cache[f] = open(c[0]).read().splitlines() if c else []
    L = src_cache[f]
    return L[n - 1].strip()[:90] if 0 < n <= len(L) else ""
for key, (ie, sm) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    print("%5.1f%% inst %5.1f%% smp  %-22s %s" % (100.0 * ie / tot_i, 100.0 * sm / tot_s, "%s:%d" % key if key else "?", src(key)))
