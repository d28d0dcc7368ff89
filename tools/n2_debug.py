import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, torch.distributed as dist
import grok_b200 as G, oracle_pipeline as P
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
use_nccl = world > 1 and "--nonccl" not in sys.argv
if world > 1:
    dist.init_process_group("nccl" if use_nccl else "gloo", **({"device_id": torch.device("cuda", local)} if use_nccl else {}))
W = H = 8192
cp = G.make_coding(W, H, 3, 12, numres=6, tile=(1024, 1024))
base = P.synthetic_image(1024, 1024, 3, 12, seed=1)
planes = [np.tile(base[c], (8, 8)) for c in range(3)]
eng = G.Engine(local); job = eng.job(cp); job.upload(planes)
for it in range(3):
    job.forward(); job.t1_encode(); job.t1_decode(); job.inverse()
if world > 1: dist.barrier()
torch.cuda.synchronize()
acc = np.zeros(4); t0 = time.perf_counter()
for it in range(10):
    acc += np.array([job.forward(), job.t1_encode()[0], job.t1_decode(), job.inverse()])
torch.cuda.synchronize()
print("rank", rank, "stages ms", (acc / 10).round(3), "wall/step %.3f" % ((time.perf_counter() - t0) * 100), flush=True)
if world > 1: dist.barrier()
