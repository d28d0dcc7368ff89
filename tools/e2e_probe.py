"""End-to-end (host buffers in, host buffers out) encode / decode times of config 2 through b2k_encode / b2k_decode, for the
pipeline settings in the environment (B2K_CHUNKS = tile chunks per call): python tools/e2e_probe.py [reps]
Prints the median of `reps` calls for pinned 16-bit planes and pinned int32 planes."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import grok_b200 as G
import oracle_pipeline as P

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 9
W = H = 8192
cp = G.make_coding(W, H, 3, 12, numres=6, tile=(1024, 1024))
img = P.synthetic_image(W, H, 3, 12, 20260924)
eng = G.Engine(0)
for name, dt in (("u16", np.uint16), ("i32", np.int32)):
    src = [G.pinned_empty((H, W), dt) for _ in range(3)]
    dst = [G.pinned_empty((H, W), dt) for _ in range(3)]
    for a, b in zip(src, img):
        a[:] = b
    te, td = [], []
    for i in range(reps + 3):
        t0 = time.perf_counter(); r = eng.encode(cp, src); t1 = time.perf_counter()
        blocks, data = r.blocks, r.bytes
        t2 = time.perf_counter(); eng.decode(cp, blocks, data, dst); t3 = time.perf_counter()
        r.free()
        if i >= 3:
            te.append((t1 - t0) * 1e3); td.append((t3 - t2) * 1e3)
    ok = all(np.array_equal(a, b) for a, b in zip(dst, src))
    print("B2K_CHUNKS=%s %s: encode %.2f ms  decode %.2f ms  (min %.2f / %.2f)  lossless %s" % (
        os.environ.get("B2K_CHUNKS", "default"), name, np.median(te), np.median(td), min(te), min(td), ok), flush=True)
