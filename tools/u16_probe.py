import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import grok_b200 as G
import oracle_pipeline as P
W = H = 8192
cp = G.make_coding(W, H, 3, 12, numres=6, tile=(1024, 1024))
base = P.synthetic_image(1024, 1024, 3, 12, seed=1)
p16 = [G.pinned_empty((H, W), np.uint16) for _ in range(3)]
o16 = [G.pinned_empty((H, W), np.uint16) for _ in range(3)]
for c in range(3):
    p16[c][:] = np.tile(base[c], (8, 8))
eng = G.Engine(0)
for it in range(3):
    res = eng.encode(cp, p16)
    eng.decode(cp, res.blocks, res.bytes, o16)
    res.free()
