#!/usr/bin/env python3
"""BASELINE config 5: random-window decode of a large TLM / PLT indexed code stream, one window per GPU at a time.

Rank 0 builds a SIZE x SIZE x 3 12-bit image from 1024x1024 tiles (seeded, per-tile offset), encodes it on its GPU into ONE
tiled HTJ2K code stream (TLM + PLT) and broadcasts the bytes (NCCL).  Then every rank decodes its share of the seeded
2048x2048 windows with Engine.decode_window (b2k_codestream_parse_window: only the touched tiles' packets are parsed and
decoded) and checks the pixels against the generator.  Prints one JSON line: per-window latency, windows/s, Mpixels/s.

  python tools/config5_roi_bench.py [--size 32768] [--rois 8]
  torchrun --nproc-per-node 8 tools/config5_roi_bench.py --size 32768 --rois 64
"""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]


def main():
    import torch
    import grok_b200 as G
    import oracle_pipeline as P
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=32768)
    ap.add_argument("--rois", type=int, default=8)
    ap.add_argument("--roi", type=int, default=2048)
    ap.add_argument("--reduce", type=int, default=0)
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    S, T = a.size, 1024
    reps = S // T
    base = P.synthetic_image(T, T, 3, 12, seed=20260927)

    def tile_pixels(c, t):
        return ((base[c] + 37 * t) & 0xFFF).astype(np.uint16)

    eng = G.Engine(local)
    t_enc = 0.0
    if rank == 0:
        cp = G.make_coding(S, S, 3, 12, numres=6, tile=(T, T))
        planes = [np.empty((S, S), np.uint16) for _ in range(3)]
        for t in range(reps * reps):
            ty, tx = divmod(t, reps)
            for c in range(3):
                planes[c][ty * T:(ty + 1) * T, tx * T:(tx + 1) * T] = tile_pixels(c, t)
        t0 = time.perf_counter()
        cs = eng.encode_codestream(cp, planes, flags=G.CS_TLM | G.CS_PLT)
        t_enc = time.perf_counter() - t0
        del planes
        n = torch.tensor([len(cs)], dtype=torch.int64, device="cuda")
    else:
        n = torch.zeros(1, dtype=torch.int64, device="cuda")
    if dist is not None:
        dist.broadcast(n, 0)
        buf = torch.from_numpy(cs).cuda() if rank == 0 else torch.empty(int(n[0]), dtype=torch.uint8, device="cuda")
        dist.broadcast(buf, 0)
        cs = buf.cpu().numpy()
        del buf
    rng = np.random.default_rng(20260927)
    wins = [(int(rng.integers(0, S - a.roi)), int(rng.integers(0, S - a.roi))) for _ in range(a.rois)]
    mine = wins[rank::world]
    eng.decode_window(cs, (mine[0][0], mine[0][1], mine[0][0] + a.roi, mine[0][1] + a.roi), a.reduce)     # warm-up
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    lat, ok = [], True
    t0 = time.perf_counter()
    for (x0, y0) in mine:
        t1 = time.perf_counter()
        _, got = eng.decode_window(cs, (x0, y0, x0 + a.roi, y0 + a.roi), a.reduce)
        lat.append(time.perf_counter() - t1)
        if a.reduce == 0:
            for c in range(3):       # check against the generator, tile by tile
                for ty in range(y0 // T, (y0 + a.roi - 1) // T + 1):
                    for tx in range(x0 // T, (x0 + a.roi - 1) // T + 1):
                        ya, yb, xa, xb = max(y0, ty * T), min(y0 + a.roi, (ty + 1) * T), max(x0, tx * T), min(x0 + a.roi, (tx + 1) * T)
                        want = tile_pixels(c, ty * reps + tx)[ya - ty * T:yb - ty * T, xa - tx * T:xb - tx * T]
                        ok &= bool(np.array_equal(got[c][ya - y0:yb - y0, xa - x0:xb - x0], want))
    dt = time.perf_counter() - t0
    tt = torch.tensor([dt, float(not ok)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    if rank == 0:
        dt = float(tt[0])
        print(json.dumps({"config": "config5: %dx%dx3 12-bit, %d tiles, TLM+PLT code stream of %d bytes; %d windows of %dx%d, reduce %d, over %d GPU(s)"
                                    % (S, S, reps * reps, len(cs), a.rois, a.roi, a.roi, a.reduce, world),
                          "encode_s_rank0": t_enc, "window_latency_ms_rank0": [round(x * 1e3, 2) for x in lat],
                          "windows_per_s": a.rois / dt, "Mpixels_per_s": a.rois * (a.roi >> a.reduce) ** 2 / dt / 1e6,
                          "pixels_match_generator": bool(tt[1] == 0)}))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
