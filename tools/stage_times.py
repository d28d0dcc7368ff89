"""Device-resident stage times of config 2 (forward, HT encode, HT decode, inverse) for the library in B2K_LIB (or the product):
python tools/stage_times.py [steps]   -- prints one line; used for A/B runs of kernel variants (tools/build_variant.py)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import grok_b200 as G
import oracle_pipeline as P
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
W = H = 8192
cp = G.make_coding(W, H, 3, 12, numres=6, tile=(1024, 1024))
img = P.synthetic_image(W, H, 3, 12, 20260924)
eng = G.Engine(0)
job = eng.job(cp)
job.upload(img)
for _ in range(3):
    job.roundtrip()
ms, stage, l1, nbytes = job.roundtrip_n(steps)
out = [np.zeros_like(p) for p in img]
job.download(out)
ok = all(np.array_equal(a, b) for a, b in zip(out, img))
print("%-40s step %.3f ms | fwd %.3f enc %.3f dec %.3f inv %.3f | bytes %d lossless %s" % (
    os.environ.get("B2K_LIB", "product").split("/")[-2] if os.environ.get("B2K_LIB") else "product",
    ms / steps, stage[0] / steps, stage[1] / steps, stage[2] / steps, stage[3] / steps, nbytes, ok))

# the block-coder stage pipelined over block ranges (b2k_job_roundtrip_pipelined_n): sweep with PIPE="chunks:streams,..."
for spec in [x for x in os.environ.get("PIPE", "").split(",") if x]:
    ch, ns = (int(v) for v in spec.split(":"))
    for _ in range(2):
        job.roundtrip_pipelined_n(1, ch, ns)
    ms, stage, l1, nb2 = job.roundtrip_pipelined_n(steps, ch, ns)
    job.download(out)
    ok = all(np.array_equal(a, b) for a, b in zip(out, img))
    print("  pipelined chunks %2d streams %d: step %.3f ms | fwd %.3f coder %.3f inv %.3f | l1 %.3f | bytes %d (same %s) lossless %s" % (
        ch, ns, ms / steps, stage[0] / steps, stage[1] / steps, stage[2] / steps, l1 / steps, nb2, nb2 == nbytes, ok))
