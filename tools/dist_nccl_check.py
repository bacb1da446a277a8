"""NCCL check of the sharded host API on >= 2 GPUs (torchrun): product, evaluation, LOO likelihood equal the
single-GPU results.  python -m torch.distributed.run --nproc-per-node 2 tools/dist_nccl_check.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import kde_b200 as K
from kde_b200 import dist as kd

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
os.environ.setdefault("NCCL_DEBUG", "WARN")
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # keep stdout for the JSON line
torch.cuda.set_device(local)
K.init(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rng = np.random.default_rng(5)
trees = [K.kde(rng.standard_normal((3, 500)) + 0.3 * j, [0.3]) for j in range(4)]
Np, T, seed = 10_001, 5, 77
gp, gi = kd.prod_sharded(trees, Np, Niter=T, seed=seed)
fp, fi = K.prodAppxMSGibbsS(None, trees, None, None, Niter=T, Np=Np, seed=seed)
assert np.array_equal(gp, fp) and np.array_equal(gi, fi)
pos = rng.standard_normal((3, 12_345))
assert np.array_equal(kd.eval_sharded(trees[0], pos), K.evaluateDualTree(trees[0], pos))
H = kd.loo_entropy_sharded(trees[1], trees[1].bandwidthMin[:3])
assert abs(H - K.entropy(trees[1])) < 1e-12 * abs(H)
dist.barrier()
if rank == 0:
    print("dist_nccl_check ok on %d GPUs" % world)
dist.destroy_process_group()
