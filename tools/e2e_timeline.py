"""host-side timeline of one b2k_encode + b2k_decode of config 2 (B2K_DEBUG_TIMING=1), and wall times vs B2K_CHUNKS (dev tool)"""
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
G.set_host_threads(int(os.environ.get("THREADS", "24")))
best = [1e9, 1e9]
for it in range(7):
    sys.stderr.write("---- iteration %d ----\n" % it)
    t0 = time.perf_counter()
    res = eng.encode(cp, planes)
    t1 = time.perf_counter()
    eng.decode(cp, res.blocks, res.bytes, out)
    t2 = time.perf_counter()
    res.free()
    if it >= 2:
        best = [min(best[0], t1 - t0), min(best[1], t2 - t1)]
print("chunks %s threads %s: encode %.2f ms decode %.2f ms total %.2f ms lossless %s" % (
    os.environ.get("B2K_CHUNKS", "8"), os.environ.get("THREADS", "24"), best[0] * 1e3, best[1] * 1e3, sum(best) * 1e3,
    all(np.array_equal(a, b) for a, b in zip(out, planes))))
