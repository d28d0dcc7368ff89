#!/usr/bin/env python3
"""Regenerates tests/golden/*.npz from the REFERENCE's own kernels (oracle/_ref/libgrok_ref.so,
compiled by oracle/Makefile from /root/reference).  Run in the build container only; the
fixtures are committed so the GPU box (no reference tree) can check against them.

  ht_blocks.npz  : sign-magnitude code blocks + the bytes ojph_encode_codeblock{32,_avx2,_avx512}
                   produce for them (all variants agree; asserted here) + what
                   ojph_decode_codeblock32 returns for those bytes.
  dwt_cases.npz  : tile components + grk::dwt53 / grk::dwt97 multi-level forward outputs.
  ht_refine.npz  : code blocks with 2 and 3 coding passes: the reference's encoder never writes SigProp /
                   MagRef, so the streams come from the oracle's test-only refinement encoder on top of the
                   reference-identical cleanup bytes; the fixture holds what ojph_decode_codeblock32 /
                   _ssse3 / _avx2 (all agree; asserted here) make of them, plain and stripe-causal.
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

    ref = {}
    rng2 = np.random.default_rng(20260925)
    n = 0
    for (w, h) in shapes:
        for dens in (0.03, 0.3, 1.0):
            M = int(rng2.integers(8, 28))         # missing MSBs of the cleanup pass; its LSB plane is p = 30 - M
            p = 30 - M
            nb = int(rng2.integers(1, p + 3))
            mag = rng2.integers(0, 1 << nb, (h, w)).astype(np.uint64) * (rng2.random((h, w)) < dens)
            mag = np.minimum(mag, (1 << (31 - (p - 1))) - 1)
            v = (mag << np.uint64(p - 1)).astype(np.uint32)
            sm = np.where(v != 0, v | (rng2.integers(0, 2, (h, w)).astype(np.uint32) << 31), 0).astype(np.uint32)
            cup = O.ref_ht_encode(sm, M, 0)
            for causal in (False, True):
                for npass in (2, 3):
                    seg = O.ht_encode_refine(sm, M, npass, causal)
                    data = np.concatenate([cup, seg])
                    outs = [O.ref_ht_decode(data, M, w, h, variant=vv, num_passes=npass, len2=len(seg), causal=causal) for vv in (0, 1, 2)]
                    assert all(rc == 0 for rc, _ in outs)
                    assert all(np.array_equal(outs[0][1], o) for _, o in outs), "reference decoder variants disagree"
                    ref["data%03d" % n] = data
                    ref["meta%03d" % n] = np.array([w, h, M, npass, len(seg), int(causal)], np.int32)
                    ref["dec%03d" % n] = outs[0][1]
                    n += 1
    ref["count"] = np.int32(n)
    np.savez_compressed(os.path.join(HERE, "ht_refine.npz"), **ref)
    nref = n
    n = int(blocks["count"])

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
    print("wrote", n, "HT blocks,", nref, "refinement blocks and", len(geoms), "DWT cases")


if __name__ == "__main__":
    main()
