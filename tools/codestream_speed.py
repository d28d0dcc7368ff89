"""host T2 speed on config 2's block table (49,728 blocks, ~150 MB of coded bytes): b2k_codestream_write / _parse.
Block contents do not matter to T2, so the arena is random bytes with realistic lengths (CPU only, no GPU needed)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import grok_b200 as G
cp = G.make_coding(8192, 8192, 3, 12, numres=6, tile=(1024, 1024))
tab = G.enumerate_blocks(cp)
rng = np.random.default_rng(1)
area = (tab["x1"] - tab["x0"]).astype(np.int64) * (tab["y1"] - tab["y0"])
length = np.maximum(3, (area * 0.74).astype(np.int64) + rng.integers(-40, 40, len(tab)))
tab["length"] = length
tab["offset"] = np.concatenate([[0], np.cumsum(length)[:-1]])
tab["numbps"] = 1
tab["numpasses"] = 1
data = rng.integers(0, 255, int(length.sum()), dtype=np.uint8)
for flags in (0, G.CS_TLM | G.CS_PLT):
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        cs = G.codestream_write(cp, tab, data, flags)
        best = min(best, time.perf_counter() - t0)
    print("write flags=%d: %.1f ms for %d blocks, %.1f MB -> %.2f GB/s, %.1f Mblocks/s" % (flags, best * 1e3, len(tab), len(cs) / 1e6, len(cs) / best / 1e9, len(tab) / best / 1e6))
best = 1e9
for _ in range(3):
    t0 = time.perf_counter()
    cp2, blocks = G.codestream_parse(cs)
    best = min(best, time.perf_counter() - t0)
print("parse: %.1f ms" % (best * 1e3), "lengths ok", np.array_equal(blocks["length"], tab["length"]))
