"""Streaming throughput probe (GPU): encode-only, decode-only and chained streams at several depths, config 2, 16-bit planes.
usage: python tools/stream_probe.py [frames]"""
import os, sys, threading, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import grok_b200 as G
import oracle_pipeline as P

W = H = 8192
N = int(sys.argv[1]) if len(sys.argv) > 1 else 12
cp = G.make_coding(W, H, 3, 12, numres=6, tile=(1024, 1024))
img = P.synthetic_image(W, H, 3, 12, 20260924)
p16 = [G.pinned_empty((H, W), np.uint16) for _ in range(3)]
for a, b in zip(p16, img):
    a[:] = b
eng = G.Engine(0)
res0 = eng.encode(cp, p16)
blocks0, bytes0 = res0.blocks.copy(), G.pinned_empty((res0.num_bytes,), np.uint8)
bytes0[:] = res0.bytes
res0.free()
for _ in range(3):
    t0 = time.perf_counter(); r = eng.encode(cp, p16); t1 = time.perf_counter()
    o = [G.pinned_empty((H, W), np.uint16) for _ in range(3)]
    t2 = time.perf_counter(); eng.decode(cp, r.blocks, r.bytes, o); t3 = time.perf_counter(); r.free()
print("sync: encode %.2f ms, decode %.2f ms" % ((t1 - t0) * 1e3, (t3 - t2) * 1e3), flush=True)

for depth in (1, 2, 3, 4):
    done = threading.Semaphore(0)
    def on_enc(tag, res, status):
        if res is not None: res.free()
        done.release()
    for rep in range(2):
        enc = G.EncodeStream(cp, depth=depth, sample_bytes=2, on_encoded=on_enc)
        t0 = time.perf_counter()
        for i in range(N): enc.submit(p16, i)
        for _ in range(N): done.acquire()
        dt = (time.perf_counter() - t0) / N
        enc.end()
    print("encode stream depth %d: %.2f ms/frame" % (depth, dt * 1e3), flush=True)
    outs = [[G.pinned_empty((H, W), np.uint16) for _ in range(3)] for _ in range(depth + 1)]
    def on_dec(tag, status): done.release()
    for rep in range(2):
        dec = G.DecodeStream(depth=depth, sample_bytes=2, on_decoded=on_dec)
        t0 = time.perf_counter()
        for i in range(N): dec.submit(cp, blocks0, bytes0, outs[i % len(outs)], i)
        for _ in range(N): done.acquire()
        dt = (time.perf_counter() - t0) / N
        dec.end()
    print("decode stream depth %d: %.2f ms/frame" % (depth, dt * 1e3), flush=True)
