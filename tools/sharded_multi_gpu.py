#!/usr/bin/env python3
"""BASELINE config 4 style run: one image, tiles sharded over N GPUs (tile t -> rank t % N), every
rank block-codes its own tiles with no data-path collective, then the coded segments are gathered
to rank 0 (the codestream writer) with NCCL -- an all_gather of segment sizes followed by a
gather of the variable-length byte arenas over NVLink -- and rank 0 decodes ALL tiles from the
gathered segments and checks the image is bit-exact, then assembles ONE tiled HTJ2K codestream from them
(b2k_codestream_write: main header + TLM, per tile SOT + PLT + packets in tile-index order) and decodes
that file again through b2k_codestream_parse + b2k_decode.

  torchrun --nproc-per-node N tools/sharded_multi_gpu.py [--size 16384] [--comps 4] [--prec 16]
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    import grok_b200 as G
    import oracle_pipeline as P

    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=16384)
    ap.add_argument("--comps", type=int, default=4)
    ap.add_argument("--prec", type=int, default=16)
    ap.add_argument("--tile", type=int, default=1024)
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    W = H = a.size
    cp = G.make_coding(W, H, a.comps, a.prec, numres=6, tile=(a.tile, a.tile), mct=1 if a.comps >= 3 else 0)
    # every rank builds the same synthetic image (tile-repeated generator: cheap at 16K)
    base = P.synthetic_image(a.tile, a.tile, a.comps, a.prec, seed=20260926)
    reps = W // a.tile
    planes = [G.pinned_empty((H, W), np.int32) for _ in range(a.comps)]
    for c in range(a.comps):
        planes[c][:] = np.tile(base[c], (reps, reps))
    eng = G.Engine(local)
    res = eng.encode(cp, planes, tile_mod=world, tile_rem=rank)   # warm-up (allocations)
    res.free()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = eng.encode(cp, planes, tile_mod=world, tile_rem=rank)
    torch.cuda.synchronize()
    t_enc = time.perf_counter() - t0
    nbytes = res.num_bytes
    # ---- gather the coded segments on rank 0 over NCCL ----
    t1 = time.perf_counter()
    if world > 1:
        sizes = [torch.zeros(2, dtype=torch.int64, device="cuda") for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([nbytes, res.num_blocks], dtype=torch.int64, device="cuda"))
        sizes = [(int(s[0]), int(s[1])) for s in sizes]
        seg = torch.from_numpy(res.bytes.copy()).cuda()
        tab = torch.from_numpy(res.blocks.view(np.uint8).reshape(-1).copy()).cuda()
        if rank == 0:
            segs = [torch.empty(n, dtype=torch.uint8, device="cuda") for n, _ in sizes]
            tabs = [torch.empty(k * G.BLOCK_DTYPE.itemsize, dtype=torch.uint8, device="cuda") for _, k in sizes]
            dist.gather(seg, segs, dst=0)
            dist.gather(tab, tabs, dst=0)
        else:
            dist.gather(seg, None, dst=0)
            dist.gather(tab, None, dst=0)
        torch.cuda.synchronize()
    t_gather = time.perf_counter() - t1
    tmax = torch.tensor([t_enc, t_gather], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ok = True
    if rank == 0:
        # writer side: merge the per-rank tables into the full enumeration order and decode everything
        if world > 1:
            merged = G.merge_shards(cp, [(np.frombuffer(tabs[r].cpu().numpy().tobytes(), dtype=G.BLOCK_DTYPE), segs[r].cpu().numpy())
                                         for r in range(world)])          # b2k_result_merge
            full, data = merged.blocks.copy(), merged.bytes.copy()
            merged.free()
        else:
            full, data = res.blocks.copy(), res.bytes.copy()
        out = [np.zeros((H, W), np.int32) for _ in range(a.comps)]
        eng2 = eng
        eng2.decode(cp, full, data, out)
        ok = all(np.array_equal(x, y) for x, y in zip(out, planes))
        tw = time.perf_counter()
        cs = G.codestream_write(cp, full, data)                     # the T2 / writer step, on the host
        t_write = time.perf_counter() - tw
        cp2, out2 = eng.decode_codestream(cs)
        ok = ok and all(np.array_equal(x, y) for x, y in zip(out2, planes))
        pix = W * H
        print({"config": "%dx%dx%d %d-bit lossless, %d tiles sharded over %d GPU(s)" % (W, H, a.comps, a.prec, reps * reps, world),
               "encode_ms_max_over_ranks": float(tmax[0]) * 1e3, "nccl_gather_ms": float(tmax[1]) * 1e3,
               "encode_Mpix_s": pix / float(tmax[0]) / 1e6, "coded_bytes_total": int(len(data)), "codestream_bytes": int(len(cs)),
               "codestream_write_ms": t_write * 1e3,
               "round_trip_bit_exact": bool(ok)})
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
