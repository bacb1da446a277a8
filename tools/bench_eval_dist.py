"""BASELINE config 5 sharded over the ranks of a torchrun launch (query points block-partitioned, the
1M-component tree replicated, NCCL all-gather of the densities).  Prints one JSON line on rank 0.
  python -m torch.distributed.run --nproc-per-node N tools/bench_eval_dist.py [--n 1000000] [--f32]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import kde_b200 as K
from kde_b200 import _lib, dist as kd
from tests.util import mixture, silverman

n = int(sys.argv[sys.argv.index("--n") + 1]) if "--n" in sys.argv else 1_000_000
prec = K.F32 if "--f32" in sys.argv else K.F64
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
os.environ.setdefault("NCCL_DEBUG", "WARN")
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # keep stdout for the JSON line
torch.cuda.set_device(local)
K.init(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rng = np.random.default_rng(20261017)
pts, pos = mixture(rng, 3, n), mixture(rng, 3, n)
p = K.kde(pts, silverman(pts))
a, b = kd.shard_range(n, rank, world)
dev = torch.device("cuda", local)
d_pos = torch.from_numpy(np.ascontiguousarray(pos[:, a:b].T)).to(dev)
d_out = torch.empty(b - a, dtype=torch.float64, device=dev)
g_out = torch.empty(n, dtype=torch.float64, device=dev) if world > 1 else d_out
L = _lib.lib()
def step():
    _lib.check(L.kdeb200_eval_device(p._dev(), d_pos.data_ptr(), b - a, 0, prec, d_out.data_ptr(), torch.cuda.current_stream().cuda_stream))
    if world > 1:
        if n % world == 0:
            dist.all_gather_into_tensor(g_out, d_out)
        else:
            kd.all_gather_blocks(d_out, n)
for _ in range(2):
    step()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); step(); step(); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 2
t = torch.tensor([ms], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"workload": "C5: %d components x %d queries, 3-D, %s" % (n, n, "f32" if prec else "f64"), "n_gpus": world,
                      "ms_per_eval_call": float(t.item()), "evals_per_s": float(n) * n / (float(t.item()) * 1e-3), "scaling": "strong (queries sharded)"}))
if world > 1:
    dist.destroy_process_group()
