"""Ad-hoc probe for C4: BruteforceSearch exact k=100 on N x 768 inner product."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import research_new_hnsw_b200 as pkg
from research_new_hnsw_b200.synth import lowrank_data
from oracle import bind
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
d, k, nq = 768, 100, 10000
os.environ["B200HNSW_BF_STATS"] = "1"
X = lowrank_data(n, d, seed=1, latent=64, noise=0.1, normalize=True)
Q = lowrank_data(nq, d, seed=2, latent=64, noise=0.1, normalize=True)
g = pkg.BruteforceSearch(pkg.InnerProductSpace(d), n)
t = time.time(); g.addPoints(X); print("upload %.2fs" % (time.time() - t), flush=True)
for path, qs in (("tensor", nq), ("tensor", nq), ("scan", 2000), ("tensor", 128), ("tensor", 1)):
    os.environ["B200HNSW_BF_PATH"] = path
    t = time.time(); r = g.searchKnnBatch(Q[:qs], k); wall = time.time() - t
    st = g.stats()
    print("%-6s nq=%5d: kernels %.2f ms wall %.3fs -> %.0f QPS, %.1f TFLOP/s (2*nq*N*d/t), path=%d cand/query=%.0f" % (
        path, qs, st["last_kernel_ms"], wall, qs / (st["last_kernel_ms"] / 1e3), 2.0 * qs * n * d / (st["last_kernel_ms"] / 1e3) / 1e12,
        st["hops_base"], st["hops_upper"] / qs), flush=True)
    if path == "tensor" and qs == nq: keep = r
    if path == "scan": print("   scan == tensor on %d queries: ids %s dists %s" % (qs, np.array_equal(r["labels"], keep["labels"][:qs]), np.array_equal(r["dists"], keep["dists"][:qs])))
ref = bind.Ref(bind.best_ref_level()); b = ref.bf_new(bind.IP, d, n); b.add(X)
rr = b.search(Q[:64], k, threads=os.cpu_count())
print("reference (%s, %d threads): %.1f QPS; ids equal %s dists equal(sse only) %s" % (bind.best_ref_level(), os.cpu_count(), 64 / rr["seconds"],
      np.array_equal(rr["labels"], keep["labels"][:64]), np.array_equal(rr["dists"], keep["dists"][:64])))
if bind.best_ref_level() != "sse":
    ref = bind.Ref("sse"); b = ref.bf_new(bind.IP, d, n); b.add(X); rr = b.search(Q[:32], k, threads=os.cpu_count())
    print("reference (sse as shipped): %.1f QPS; ids equal %s dists bit-equal %s" % (32 / rr["seconds"], np.array_equal(rr["labels"], keep["labels"][:32]), np.array_equal(rr["dists"], keep["dists"][:32])))
