"""-m gpu: the sharded paths with the CUDA worker.  Two ranks share cuda:0 (the test box has one
GPU), exchange over gloo; on the 8-GPU box bench.py runs the same code over NCCL."""
import os

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.test_dist_gloo import free_port

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import kde_b200 as K
        from kde_b200 import dist as kd
        K.init(0)
        rng = np.random.default_rng(7)
        trees = [K.kde(rng.standard_normal((3, 300)) + 0.3 * j, [0.3]) for j in range(4)]
        Np, T, seed = 1001, 5, 99
        gp, gi = kd.prod_sharded(trees, Np, Niter=T, seed=seed)
        fp, fi = K.prodAppxMSGibbsS(None, trees, None, None, Niter=T, Np=Np, seed=seed)
        assert np.array_equal(gp, fp) and np.array_equal(gi, fi)
        pos = rng.standard_normal((3, 777))
        assert np.array_equal(kd.eval_sharded(trees[0], pos), K.evaluateDualTree(trees[0], pos))
        H = kd.loo_entropy_sharded(trees[1], trees[1].bandwidthMin[:3])
        assert abs(H - K.entropy(trees[1])) < 1e-12 * abs(H)
        pts = np.concatenate([rng.standard_normal((2, 700)) * 0.5 - 1.0, rng.standard_normal((2, 500)) * 0.3 + 1.5], axis=1)
        ks, k1 = kd.kde_sharded(pts), K.kde(pts)          # rows of every nLOO_LL step split over the two ranks
        assert np.max(np.abs(K.getBW(ks)[:, 0] / K.getBW(k1)[:, 0] - 1.0)) < 1e-9
        if rank == 0:
            out.put("ok")
    finally:
        dist.destroy_process_group()


def test_sharded_cuda_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    mp.spawn(_worker, args=(2, free_port(), q), nprocs=2, join=True)
    assert q.get(timeout=5) == "ok"
