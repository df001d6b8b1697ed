"""One-GPU emulation of what the sharded pipeline adds around the search kernel at N = 8: per step, on a high-priority side
stream ordered behind the search by an event, (a) 7 device-to-device copies of one packed block (stand-ins for the
peer pushes), (b) the merge kernel over 8 blocks.  Which part costs the search stream its time?  The back-to-back baseline
is repeated between the modes (the first measurement of a process is faster than any later one)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import research_new_hnsw_b200 as pkg
from research_new_hnsw_b200 import capi
from research_new_hnsw_b200.synth import lowrank_data
n, d, nq, k, ef, W = int(os.environ.get("N", 1000000)), 128, 10000, 10, int(os.environ.get("EF", 16)), 8
X = lowrank_data(n, d, seed=1); Q = [torch.from_numpy(lowrank_data(nq, d, seed=2 + i)).cuda() for i in range(2)]
g = pkg.HierarchicalNSW(pkg.L2Space(d), n, 32, 200, storage=1 if os.environ.get("BF16") else 0); g.addPoints(X); g.flush()
block = nq * k * 12
area = torch.zeros(W * block, dtype=torch.uint8, device="cuda")
peers = torch.zeros(W * block, dtype=torch.uint8, device="cuda")
ol = torch.empty((nq, k), dtype=torch.int64, device="cuda"); od = torch.empty((nq, k), dtype=torch.float32, device="cuda")
main = torch.cuda.current_stream(); side = torch.cuda.Stream(priority=-1)
def run(mode, steps=60):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(steps):
        pl = area.data_ptr()
        g.searchKnnDevice(Q[s % 2].data_ptr(), nq, k, ef, pl, pl + nq * k * 8, 0, 0, main.cuda_stream)
        if mode == 0:
            continue
        ev = torch.cuda.Event(); ev.record(main); side.wait_event(ev)
        with torch.cuda.stream(side):
            if mode in (1, 3):
                for r in range(1, W):
                    peers[r * block:(r + 1) * block].copy_(area[:block], non_blocking=True)
            if mode in (2, 3):
                capi.merge_topk_packed_device(area.data_ptr(), block, W, nq, k, ol.data_ptr(), od.data_ptr(), side.cuda_stream)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps
# blocks 1..7 of the area: plausible rows so the merge does real work
r0 = g.searchKnnBatch(lowrank_data(nq, d, seed=9), k, ef=ef)
blk = np.concatenate([r0["labels"].astype(np.uint64).view(np.uint8).ravel(), r0["dists"].view(np.uint8).ravel()])
for r in range(W):
    area[r * block:(r + 1) * block] = torch.from_numpy(blk).cuda()
NAMES = ["back to back", "+ 7 block copies on the side stream", "+ merge of 8 blocks on the side stream", "+ copies + merge"]
for mode in [int(x) for x in os.environ.get("MODES", "0,1,0,2,0,3,0").split(",")]:
    run(mode, 10)
    print("%-40s %.4f ms/step" % (NAMES[mode], min(run(mode) for _ in range(3))), flush=True)
