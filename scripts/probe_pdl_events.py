"""Does a cudaEventRecord (and a side stream waiting on it) between two search launches take away their programmatic
overlap?  (No: the FIRST measurement of a process is ~10 % faster than any later one whatever the mode -- the baseline is
therefore repeated between the modes.)  One GPU, C2-shaped 200 k-point graph built on the GPU, 10 k queries, ef=16."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import research_new_hnsw_b200 as pkg
from research_new_hnsw_b200.synth import lowrank_data
n, d, nq, k, ef = int(os.environ.get("N", 200000)), 128, 10000, 10, int(os.environ.get("EF", 16))
X = lowrank_data(n, d, seed=1); Q = [torch.from_numpy(lowrank_data(nq, d, seed=2 + i)).cuda() for i in range(2)]
g = pkg.HierarchicalNSW(pkg.L2Space(d), n, 32, 200); g.addPoints(X); g.flush()
ol = torch.empty((nq, k), dtype=torch.int64, device="cuda"); od = torch.empty((nq, k), dtype=torch.float32, device="cuda")
main = torch.cuda.current_stream(); side = torch.cuda.Stream(priority=-1)
junk = torch.zeros(1 << 20, device="cuda")
def run(mode, steps=60):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(steps):
        g.searchKnnDevice(Q[s % 2].data_ptr(), nq, k, ef, ol.data_ptr(), od.data_ptr(), 0, 0, main.cuda_stream)
        if mode >= 1:
            ev = torch.cuda.Event(); ev.record(main)
        if mode >= 2:
            side.wait_event(ev)
        if mode >= 3:
            with torch.cuda.stream(side):
                junk.add_(1.0)          # a small kernel on the high-priority side stream
        if mode >= 4:
            with torch.cuda.stream(side):
                ev2 = torch.cuda.Event(); ev2.record(side)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps
NAMES = ["back to back", "+ event record", "+ side stream waits on it", "+ small side kernel", "+ side event record"]
order = [int(x) for x in os.environ.get("MODES", "0,1,0,2,0,3,0,4,0").split(",")]
for mode in order:   # mode 0 is repeated at the end: the numbers must not depend on the position in this list
    run(mode, 10)
    print("%-30s %.4f ms/step" % (NAMES[mode], min(run(mode) for _ in range(3))), flush=True)
