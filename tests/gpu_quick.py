"""quick GPU bring-up script (not a pytest): python tests/gpu_quick.py"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import grok_b200 as G
import oracle_pipeline as P
import oracle_lib as O

def check(cp, planes, name):
    eng = check.eng
    job = eng.job(cp)
    job.upload(planes)
    ms = job.forward()
    got = [np.zeros_like(p) for p in planes]
    job.download_coeffs(got)
    ref = P.forward(cp, planes)
    ok = all(np.array_equal(g, r) for g, r in zip(got, ref))
    nbad = sum(int((g != r).sum()) for g, r in zip(got, ref))
    print("%s: forward %s (%.3f ms) mismatches=%d" % (name, "OK" if ok else "FAIL", ms, nbad))
    if not ok:
        for c, (g, r) in enumerate(zip(got, ref)):
            ys, xs = np.nonzero(g != r)
            if len(ys):
                print("  comp", c, "first bad at", ys[0], xs[0], "got", g[ys[0], xs[0]], "want", r[ys[0], xs[0]], "count", len(ys),
                      "rows", ys.min(), ys.max(), "cols", xs.min(), xs.max())
    # T1 encode
    job.upload_coeffs(ref)
    ms, total = job.t1_encode()
    res = job.fetch_result()
    blks = P.enumerate_all(cp)
    assert len(blks) == res.num_blocks, (len(blks), res.num_blocks)
    rects = P.tile_rects(cp)
    bad = 0
    for i, (t, c, b) in enumerate(blks):
        gb = res.blocks[i]
        assert (gb["x0"], gb["y0"], gb["x1"], gb["y1"], gb["buf_x"], gb["buf_y"]) == (b.x0, b.y0, b.x1, b.y1, b.buf_x, b.buf_y), (i, gb, (b.x0,b.y0,b.x1,b.y1,b.buf_x,b.buf_y))
        if b.x1 == b.x0 or b.y1 == b.y0:
            continue
        want = P.encode_block(cp, ref, rects[t], c, b)
        have = res.block_bytes(i)
        if not np.array_equal(want, have):
            bad += 1
            if bad <= 3:
                n = min(len(want), len(have))
                d = np.nonzero(want[:n] != have[:n])[0]
                print("  block", i, "res", b.resno, "orient", b.orient, "size", b.x1-b.x0, b.y1-b.y0, "len", len(want), len(have), "first diff", d[:5])
    print("%s: t1 encode %s (%.3f ms, %d bytes, %d blocks) bad=%d" % (name, "OK" if bad == 0 else "FAIL", ms, total, len(blks), bad))
    # decode path
    ms = job.t1_decode()
    got = [np.zeros_like(p) for p in planes]
    job.download_coeffs(got)
    okd = all(np.array_equal(g, r) for g, r in zip(got, ref))
    print("%s: t1 decode %s (%.3f ms)" % (name, "OK" if okd else "FAIL", ms))
    ms = job.inverse()
    rec = [np.zeros_like(p) for p in planes]
    job.download(rec)
    okr = all(np.array_equal(g, r) for g, r in zip(rec, planes))
    print("%s: inverse roundtrip %s (%.3f ms)" % (name, "OK" if okr else "FAIL", ms))
    if not okr:
        for c, (g, r) in enumerate(zip(rec, planes)):
            ys, xs = np.nonzero(g != r)
            if len(ys):
                print("  comp", c, "first bad at", ys[0], xs[0], "got", g[ys[0], xs[0]], "want", r[ys[0], xs[0]], "count", len(ys))
    res.free()
    job.close()
    return ok and bad == 0 and okd and okr

if __name__ == "__main__":
    check.eng = G.Engine(0)
    allok = True
    cp = G.make_coding(512, 512, 1, 8, numres=6)
    allok &= check(cp, P.synthetic_image(512, 512, 1, 8, 1234), "cfg1 512x512 grey8")
    cp = G.make_coding(2048, 1024, 3, 12, numres=6, tile=(1024, 1024))
    allok &= check(cp, P.synthetic_image(2048, 1024, 3, 12, 2026), "2 tiles rgb12")
    cp = G.make_coding(333, 217, 3, 12, numres=4, origin=(3, 5))
    allok &= check(cp, P.synthetic_image(333, 217, 3, 12, 7, origin=(3, 5)), "odd origin rgb12")
    cp = G.make_coding(100, 75, 4, 16, numres=3, tile=(61, 40), cblk=(32, 32))
    allok &= check(cp, P.synthetic_image(100, 75, 4, 16, 9), "4comp 16bit odd tiles")
    print("ALL OK" if allok else "SOME FAILED")
