"""print the device-event breakdown of b2k_encode / b2k_decode on config 2 (dev tool)"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import grok_b200 as G
import oracle_pipeline as P
W = H = 8192
cp = G.make_coding(W, H, 3, 12, numres=6, tile=(1024, 1024))
base = P.synthetic_image(1024, 1024, 3, 12, seed=1)
planes = [G.pinned_empty((H, W), np.int32) for _ in range(3)]
out = [G.pinned_empty((H, W), np.int32) for _ in range(3)]
for c in range(3):
    planes[c][:] = np.tile(base[c], (8, 8))
eng = G.Engine(0)
p16 = [G.pinned_empty((H, W), np.uint16) for _ in range(3)]
o16 = [G.pinned_empty((H, W), np.uint16) for _ in range(3)]
for c in range(3):
    p16[c][:] = planes[c]
for it in range(5):
    t0 = time.perf_counter()
    res = eng.encode(cp, planes)
    t1 = time.perf_counter()
    ms = eng.decode(cp, res.blocks, res.bytes, out)
    t2 = time.perf_counter()
    print("iter", it, "encode wall %.2f ms" % ((t1 - t0) * 1e3), {k: round(v, 2) for k, v in res.timings.items()},
          "| decode wall %.2f ms (device %.2f)" % ((t2 - t1) * 1e3, ms), "bytes", res.num_bytes)
    res.free()

for it in range(6):
    t0 = time.perf_counter()
    res = eng.encode(cp, p16)
    t1 = time.perf_counter()
    ms = eng.decode(cp, res.blocks, res.bytes, o16)
    t2 = time.perf_counter()
    print("u16 iter", it, "encode wall %.2f ms" % ((t1 - t0) * 1e3), {k: round(v, 2) for k, v in res.timings.items()},
          "| decode wall %.2f ms" % ((t2 - t1) * 1e3), "ok", all(np.array_equal(a, b) for a, b in zip(o16, p16)))
    res.free()
