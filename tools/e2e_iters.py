"""per-iteration wall times of b2k_encode + b2k_decode on config 2 (dev tool): THREADS=-1 (auto policy) | n"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import grok_b200 as G
import oracle_pipeline as P
import bench
bench.bind_to_gpu_numa_node(0)
W = H = 8192
cp = G.make_coding(W, H, 3, 12, numres=6, tile=(1024, 1024))
base = P.synthetic_image(1024, 1024, 3, 12, seed=1)
planes = [G.pinned_empty((H, W), np.int32) for _ in range(3)]
out = [G.pinned_empty((H, W), np.int32) for _ in range(3)]
for c in range(3):
    planes[c][:] = np.tile(base[c], (8, 8))
eng = G.Engine(0)
G.set_host_threads(int(os.environ.get("THREADS", "-1")))
ts = []
for it in range(int(os.environ.get("ITERS", "26"))):
    t0 = time.perf_counter()
    res = eng.encode(cp, planes)
    t1 = time.perf_counter()
    eng.decode(cp, res.blocks, res.bytes, out)
    t2 = time.perf_counter()
    res.free()
    ts.append(((t1 - t0) * 1e3, (t2 - t1) * 1e3, G.host_pack_last()))
print(" ".join("%.1f+%.1f%s" % (a, b, "" if m == (1, 1) else str(m)) for a, b, m in ts))
print("mean of last 20: %.2f ms" % np.mean([a + b for a, b, _ in ts[-20:]]))
