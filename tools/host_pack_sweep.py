"""e2e encode/decode wall time of config 2 vs. number of host packing threads (dev tool)"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import grok_b200 as G
import oracle_pipeline as P
sys.path.insert(0, ROOT)
import bench
numa = bench.bind_to_gpu_numa_node(0)[0]
print("numa", numa, "cpus", len(os.sched_getaffinity(0)))
W = H = 8192
cp = G.make_coding(W, H, 3, 12, numres=6, tile=(1024, 1024))
base = P.synthetic_image(1024, 1024, 3, 12, seed=1)
pinned = os.environ.get("PINNED", "1") == "1"
alloc = (lambda: G.pinned_empty((H, W), np.int32)) if pinned else (lambda: np.empty((H, W), np.int32))
planes = [alloc() for _ in range(3)]
out = [alloc() for _ in range(3)]
for c in range(3):
    planes[c][:] = np.tile(base[c], (8, 8))
eng = G.Engine(0)
for nt in [int(x) for x in (sys.argv[1:] or ["0", "4", "8", "12", "16", "24", "32", "48"])]:
    G.set_host_threads(nt)
    best = [1e9, 1e9]
    for it in range(6):
        t0 = time.perf_counter()
        res = eng.encode(cp, planes)
        t1 = time.perf_counter()
        eng.decode(cp, res.blocks, res.bytes, out)
        t2 = time.perf_counter()
        res.free()
        if it >= 2:
            best = [min(best[0], t1 - t0), min(best[1], t2 - t1)]
    ok = all(np.array_equal(a, b) for a, b in zip(out, planes))
    print("threads %2d  encode %.2f ms  decode %.2f ms  total %.2f ms  lossless %s" % (nt, best[0] * 1e3, best[1] * 1e3, sum(best) * 1e3, ok), flush=True)
