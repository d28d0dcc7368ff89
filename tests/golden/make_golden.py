#!/usr/bin/env python3
"""Regenerates tests/golden/*.npz from the REFERENCE's own kernels (oracle/_ref/libgrok_ref.so,
compiled by oracle/Makefile from /root/reference).  Run in the build container only; the
fixtures are committed so the GPU box (no reference tree) can check against them.

  ht_blocks.npz  : sign-magnitude code blocks + the bytes ojph_encode_codeblock{32,_avx2,_avx512}
                   produce for them (all variants agree; asserted here) + what
                   ojph_decode_codeblock32 returns for those bytes.
  dwt_cases.npz  : tile components + grk::dwt53 / grk::dwt97 multi-level forward outputs.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as O  # noqa: E402


def main():
    R = O.ref()
    assert R is not None, "build oracle/_ref first (make -C oracle ref)"
    rng = np.random.default_rng(20260924)
    blocks = {}
    n = 0
    shapes = [(64, 64), (32, 32), (64, 17), (5, 3), (1, 1), (2, 7), (33, 64), (128, 32), (4, 64), (63, 63), (16, 16),
              (64, 1), (1, 64), (3, 2)]
    for (w, h) in shapes:
        for kind in ("noise", "sparse", "zero", "max", "smooth"):
            kmax = int(rng.integers(2, 19))
            lim = (1 << kmax) - 1
            if kind == "noise":
                c = rng.integers(-lim, lim + 1, (h, w))
            elif kind == "sparse":
                c = rng.integers(-lim, lim + 1, (h, w)) * (rng.random((h, w)) < 0.07)
            elif kind == "zero":
                c = np.zeros((h, w), np.int64)
            elif kind == "max":
                c = np.where((np.add.outer(np.arange(h), np.arange(w)) & 1) == 0, lim, -lim)
            else:
                c = (rng.standard_normal((h, w)) * min(lim, 20)).astype(np.int64)
                c = np.clip(c, -lim, lim)
            sm = O.to_sgnmag(c, kmax)
            outs = [O.ref_ht_encode(sm, kmax, v) for v in (0, 1, 2)]
            outs = [o for o in outs if o is not None]
            assert all(np.array_equal(outs[0], o) for o in outs), "reference encoder variants disagree"
            rc, dec = O.ref_ht_decode(outs[0], kmax, w, h, 0)
            assert rc == 0
            blocks["in%03d" % n] = sm
            blocks["kmax%03d" % n] = np.int32(kmax)
            blocks["out%03d" % n] = outs[0]
            blocks["dec%03d" % n] = dec
            n += 1
    blocks["count"] = np.int32(n)
    np.savez_compressed(os.path.join(HERE, "ht_blocks.npz"), **blocks)

    cases = {}
    geoms = [(0, 0, 64, 64, 4), (3, 5, 61, 47, 3), (1, 0, 17, 33, 6), (0, 1, 128, 9, 5), (7, 7, 1, 20, 3), (2, 3, 40, 1, 2),
             (0, 0, 5, 5, 6)]
    for i, (x0, y0, w, h, numres) in enumerate(geoms):
        stride = ((w + 15) // 16) * 16 + 16
        a = O.aligned_zeros((h + 2, stride), np.int32)
        src = rng.integers(-2048, 2048, (h, w)).astype(np.int32)
        a[:h, :w] = src
        R.ref_dwt53_fwd_2d(a, stride, x0, y0, x0 + w, y0 + h, numres, 0)
        f = O.aligned_zeros((h + 2, stride), np.float32)
        f[:h, :w] = src.astype(np.float32)
        R.ref_dwt97_fwd_2d(f, stride, x0, y0, x0 + w, y0 + h, numres, 0.0, 0)
        cases["geom%d" % i] = np.array([x0, y0, w, h, numres], np.int32)
        cases["src%d" % i] = src
        cases["dwt53_%d" % i] = a[:h, :w].copy()
        cases["dwt97_%d" % i] = f[:h, :w].copy()
    cases["count"] = np.int32(len(geoms))
    np.savez_compressed(os.path.join(HERE, "dwt_cases.npz"), **cases)
    print("wrote", n, "HT blocks and", len(geoms), "DWT cases")


if __name__ == "__main__":
    main()
