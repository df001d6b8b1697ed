"""ncu report -> the text summary committed under profiles/: selected raw metrics per captured kernel, warp-stall
breakdown and hottest SASS instructions of each.  usage: ncu_summary.py <report.ncu-rep> [header line ...]"""
import csv, subprocess, sys
rep = sys.argv[1]
print("# " + " ".join(sys.argv[2:]) if len(sys.argv) > 2 else "# " + rep)
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "l1tex__t_sector_hit_rate.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__warps_active.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "launch__waves_per_multiprocessor",
        "sm__cycles_elapsed.avg.per_second", "smsp__cycles_active.avg"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
ci = {h: i for i, h in enumerate(hdr)}
names = [r[ci["Kernel Name"]] for r in data]
print("kernels captured:")
for i, n in enumerate(names):
    print("  [%d] %s" % (i, n))
for k in KEYS + sorted(h for h in hdr if "warp_issue_stalled" in h and h.endswith("_per_warp_active.pct")):
    if k in ci:
        print("%-88s %-10s | %s" % (k, units[ci[k]], " | ".join(r[ci[k]] for r in data)))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout.splitlines()
starts = [i for i, l in enumerate(src) if l.startswith('"Kernel Name"')]
seen = set()
for si, st in enumerate(starts):
    kname = src[st].split('","')[1].split("(")[0] if '","' in src[st] else "?"
    seg = src[st + 1:(starts[si + 1] if si + 1 < len(starts) else len(src))]
    key = (kname, len(seg))
    if key in seen:  # the source page lists every kernel once per view
        continue
    seen.add(key)
    r = list(csv.reader(seg))
    if not r:
        continue
    h = r[0]; body = [x for x in r[1:] if len(x) == len(h)]
    c = {x: i for i, x in enumerate(h)}
    if "# Samples" not in c:
        continue
    tot = sum(int(x[c["# Samples"]]) for x in body) or 1
    print("\n== SASS of %s: %d instructions, %d samples, %d warp instructions executed" %
          (kname, len(body), tot, sum(int(x[c["Instructions Executed"]]) for x in body)))
    stalls = [x for x in h if x.startswith("stall_") and "Not Issued" not in x]
    agg = {x: sum(int(b[c[x]]) for b in body) for x in stalls}
    print("stall samples: " + ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / tot) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    ops = {}
    for b in body:
        t = b[c["Source"]].strip().split()
        op = (t[1] if t and t[0].startswith("@") and len(t) > 1 else (t[0] if t else "?")).split(".")[0]
        ops[op] = ops.get(op, 0) + int(b[c["Instructions Executed"]])
    print("executed by opcode: " + ", ".join("%s %d" % kv for kv in sorted(ops.items(), key=lambda kv: -kv[1])[:16]))
    print("hottest instructions (share of samples, executed, SASS, top stall reasons):")
    for i in sorted(sorted(range(len(body)), key=lambda i: -int(body[i][c["# Samples"]]))[:25]):
        b = body[i]
        top = sorted(stalls, key=lambda x: -int(b[c[x]]))[:2]
        print("%5d %5.2f%% exec=%-9s %-64s %s" % (i, 100.0 * int(b[c["# Samples"]]) / tot, b[c["Instructions Executed"]],
                                                  b[c["Source"]].strip()[:64], ",".join("%s=%s" % (x[6:], b[c[x]]) for x in top)))
