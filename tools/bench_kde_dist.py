"""BASELINE config 3 (kde! LOOCV of 100k x 4-D points) with the rows of every nLOO_LL step sharded over the ranks
of a torchrun launch (NCCL all-reduce of two scalars per golden-section step).  One JSON line on rank 0.
  python -m torch.distributed.run --nproc-per-node N tools/bench_kde_dist.py [--n 100000]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import kde_b200 as K
from kde_b200 import dist as kd
from tests.util import mixture

n = int(sys.argv[sys.argv.index("--n") + 1]) if "--n" in sys.argv else 100_000
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
os.environ.setdefault("NCCL_DEBUG", "WARN")
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # keep stdout for the JSON line
torch.cuda.set_device(local)
K.init(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pts = mixture(np.random.default_rng(20261017), 4, n)
kd.kde_sharded(pts[:, :2000])  # warm-up (module load, NCCL communicator)
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
p = kd.kde_sharded(pts)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
if world > 1:
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
if rank == 0:
    one = None
    if world == 1:
        t0 = time.perf_counter(); q = K.kde(pts); one = time.perf_counter() - t0
    print(json.dumps({"workload": "C3: kde! LOOCV, %d points, 4-D, f64" % n, "n_gpus": world, "wall_s": dt,
                      "native_single_call_s": one, "bandwidth": K.getBW(p)[:, 0].tolist(),
                      "scaling": "strong (rows of each nLOO_LL step sharded, all-reduce per step)"}))
if world > 1:
    dist.destroy_process_group()
