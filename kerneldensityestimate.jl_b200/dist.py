"""Multi-GPU layer: one process per GPU, trees replicated, independent units sharded.

The reference has no distributed layer at all; the hot path shards trivially because every unit
is independent and addressed by its global index:

  path                 unit            partition                 exchange (torch.distributed)
  Gibbs product (S1)   sample s        contiguous block / rank   all-gather of points and labels
  evaluation (S2)      query point     contiguous block / rank   all-gather of p
  LOO likelihood (S3)  leaf row j      contiguous block / rank   all-reduce(sum) + all-reduce(max flag)

Random variates are addressed by the GLOBAL sample index (injected arrays are sliced per sample,
Philox is keyed by (seed, sample, draw)), so the result is independent of the number of ranks.
Backend: NCCL over NVLink on the GPU box (tensors stay on the device), gloo on CPU for the
world_size-2 tests of this host logic.
"""
import ctypes as C

import numpy as np

from . import _lib
from . import api


def shard_range(n, rank, world):
    """Contiguous block partition of range(n): the first n % world ranks get one extra unit."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _dist():
    import torch.distributed as dist
    return dist


def _world(group=None):
    dist = _dist()
    if not dist.is_available() or not dist.is_initialized():
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def _exchange_device(device=None, group=None):
    """Where exchange tensors live: the caller's choice, else CUDA under NCCL and host under gloo."""
    import torch
    if device is not None:
        return device
    dist = _dist()
    if dist.is_initialized() and dist.get_backend(group) == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def all_gather_blocks(local, n_total, group=None):
    """Assemble per-rank row blocks (shard_range order) into the full array on every rank.
    `local` is a torch tensor [n_local, ...] on the backend's device; uneven blocks are padded."""
    import torch
    dist = _dist()
    rank, world = _world(group)
    if world == 1:
        return local
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    nmax = max(b - a for a, b in sizes)
    pad = torch.zeros((nmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world * nmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    if all(b - a == nmax for a, b in sizes):
        return out
    return torch.cat([out[r * nmax: r * nmax + (b - a)] for r, (a, b) in enumerate(sizes)], dim=0)


# ----------------------------------------------------------------------------- Gibbs ---------
def prod_sharded(trees, Np, Niter=5, seed=0, addEntropy=True, partialDimMask=None, randU=None, randN=None,
                 group=None, compute=None, device=None):
    """prodAppxMSGibbsS over all ranks of `group`: rank r draws samples shard_range(Np, r, world),
    then the blocks are all-gathered.  Returns (points d x Np, indices M x Np) as numpy on every rank.

    compute(s0, s1) -> (points d x n, indices M x n) overrides the per-shard worker (the gloo tests
    inject the oracle here to exercise the sharding logic without a GPU)."""
    import torch
    rank, world = _world(group)
    s0, s1 = shard_range(Np, rank, world)
    if compute is None:
        def compute(a, b):
            return api.prodAppxMSGibbsS(None, trees, None, None, Niter=Niter, addEntropy=addEntropy, Np=Np,
                                        randU=randU, randN=randN, partialDimMask=partialDimMask, seed=seed, s0=a, s1=b)
    pts, idx = compute(s0, s1)
    if world == 1:
        return pts, idx
    dev = _exchange_device(device, group)
    tp = torch.from_numpy(np.ascontiguousarray(pts.T)).to(dev)   # [n, d]
    ti = torch.from_numpy(np.ascontiguousarray(idx.T)).to(dev)   # [n, M]
    gp = all_gather_blocks(tp, Np, group).cpu().numpy()
    gi = all_gather_blocks(ti, Np, group).cpu().numpy()
    return gp.T, gi.T


def prod_sharded_device(handles, ndens, dims, Np_total, Niter, seed, d_points, d_indices, g_points=None,
                        g_indices=None, group=None, stream=None):
    """Device-resident variant used by bench.py: writes this rank's block into the torch tensors
    d_points [n, d] / d_indices [n, M] and (world > 1) all-gathers into g_points / g_indices."""
    import torch
    dist = _dist()
    rank, world = _world(group)
    s0, s1 = shard_range(Np_total, rank, world)
    st = torch.cuda.current_stream().cuda_stream if stream is None else stream
    _lib.check(_lib.lib().kdeb200_gibbs_device(handles, ndens, Np_total, Niter, 1, None, None, 0, None, 0, seed, s0,
                                               s1, d_points.data_ptr(), d_indices.data_ptr(), None, st))
    if world > 1:
        if Np_total % world != 0:  # all_gather_into_tensor needs equal blocks; the host variant pads instead
            raise api.KDEError("prod_sharded_device: Np_total (%d) must be a multiple of the world size (%d); "
                               "use prod_sharded for ragged splits" % (Np_total, world))
        dist.all_gather_into_tensor(g_points, d_points[: s1 - s0], group=group)
        dist.all_gather_into_tensor(g_indices, d_indices[: s1 - s0], group=group)


# ----------------------------------------------------------------------------- evaluation ----
def eval_sharded(bd, pos, group=None, compute=None, device=None):
    """evaluateDualTree(bd, pos) with the query points block-partitioned over the ranks."""
    import torch
    rank, world = _world(group)
    pos = np.asarray(pos, dtype=np.float64)
    M = pos.shape[1]
    a, b = shard_range(M, rank, world)
    if compute is None:
        def compute(lo, hi):
            return api.evaluateDualTree(bd, pos[:, lo:hi]) if hi > lo else np.zeros(0)
    p = np.asarray(compute(a, b), dtype=np.float64)
    if world == 1:
        return p
    dev = _exchange_device(device, group)
    return all_gather_blocks(torch.from_numpy(p).to(dev), M, group).cpu().numpy()


def loo_entropy_sharded(bd, bw_var=None, group=None, compute=None, device=None, n_rows=None):
    """entropy(bd) with the leaf rows block-partitioned: all-reduce of (sum_j W_j log L_j, zero flag)."""
    import torch
    dist = _dist()
    rank, world = _world(group)
    N = bd.bt.num_points if n_rows is None else n_rows
    a, b = shard_range(N, rank, world)
    if compute is None:
        def compute(lo, hi):
            s, f = C.c_double(0.0), C.c_int(0)
            bw = None if bw_var is None else np.ascontiguousarray(bw_var, dtype=np.float64)
            _lib.check(_lib.lib().kdeb200_loo_partial(bd._dev(stale_bw_ok=bw is not None), _lib.fptr(bw), lo, hi,
                                                      C.byref(s), C.byref(f)))
            return s.value, f.value
    s, f = compute(a, b)
    if world > 1:
        dev = _exchange_device(device, group)
        ts = torch.tensor([s], dtype=torch.float64, device=dev)
        tf = torch.tensor([int(f)], dtype=torch.int32, device=dev)
        dist.all_reduce(ts, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(tf, op=dist.ReduceOp.MAX, group=group)
        s, f = float(ts.item()), int(tf.item())
    return float("inf") if f else -s


def kde_sharded(points, group=None, device=None, loo_factory=None):
    """kde!(points) (src/KDE01.jl:3-27) with the leaf rows of every nLOO_LL evaluation block-partitioned over the
    ranks: one all-reduce of (sum, zero flag) per golden-section step.  Every rank runs the same
    golden loop on the same all-reduced likelihoods and returns the same density.  The loop is native
    (kdeb200_kde_lcv_sharded, the exchange is a callback into torch.distributed); loo_factory(bd) -> compute(lo, hi)
    replaces the CUDA worker and selects the step-by-step mirror instead (gloo tests)."""
    import torch
    pts = api._as_matrix(points)
    d, N = pts.shape
    rank, world = _world(group)
    if loo_factory is None:
        dist = _dist()
        dev = _exchange_device(device, group) if world > 1 else None
        failure = []

        def exchange(psum, pflag, count, _user):
            # the d golden-section searches run in lock-step: ONE all-reduce per step carries the `count` partial
            # likelihoods and zero flags of that step (flags ride along as doubles; sum > 0 <=> some rank flagged)
            try:
                if world > 1:
                    buf = [psum[i] for i in range(count)] + [float(pflag[i]) for i in range(count)]
                    t = torch.tensor(buf, dtype=torch.float64, device=dev)
                    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
                    v = t.tolist()
                    for i in range(count):
                        psum[i], pflag[i] = v[i], int(v[count + i] > 0.0)
                return 0
            except Exception as e:  # exceptions must not unwind through the C frames
                failure.append(e)
                return 1
        cb = _lib.allreduce_v_fn(exchange)
        a, b = shard_range(N, rank, world)
        flat = np.ascontiguousarray(pts.T).ravel()
        bw = np.zeros(d)
        rc = _lib.lib().kdeb200_kde_lcv_sharded_v(d, N, _lib.fptr(flat), a, b, cb, None, _lib.fptr(bw), None)
        if failure:
            raise failure[0]
        _lib.check(rc)
        return api.kde(pts, bw)
    p = api.kde(pts, [1.0])
    bwds = np.zeros(d)

    def ent(bd):
        return loo_entropy_sharded(bd, bw_var=bd.bandwidthMin[:bd.bt.dims], group=group, device=device,
                                   compute=loo_factory(bd))
    for i in range(d):
        bwds[i] = api.getBW(api.ksize(api.marginal(p, [i + 1]), _entropy=ent))[0, 0]
    return api.kde(pts, bwds)
