#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 hot path (contract: see the task statement).

Workload (BASELINE.json configs[3], the one `metric` is quoted on; it fits one GPU):
    prodAppxMSGibbsS of 8 densities x 4096 components, 3-D, Niter=5, 1,000,000 product samples
    per GPU per step, free-running Philox streams, synthetic Gaussian-mixture data.
A "step" is one such call.  `value` = product samples/s with the trees resident in HBM
(kdeb200_gibbs_device on torch's current stream, CUDA events, L2 flushed between steps);
`e2e` = the same through the host API (prodAppxMSGibbsS mirror -> kdeb200_gibbs): tree
flatten + H2D and the D2H of points and labels inside the timed region.
N > 1 (torchrun, one rank per GPU): every rank draws its own 1M-sample slice of an N x 1M-sample
run (chains are addressed by global sample index, so results do not depend on N) and the slices
are assembled with an NCCL all-gather inside the timed step -> "scaling": "weak".

`--impl reference` times the CPU arm: the literal C restatement of the reference (oracle/,
kind "port"; the reference itself is Julia and there is no julia binary on the box) with all
host threads over independent chains, on a bounded sample of the same workload.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 20261017
NDENS, NCOMP, DIM, NITER = 8, 4096, 3, 5
SAMPLES_PER_GPU = 1_000_000
# algorithmic FP64-pipe slots per sample for this shape: 393216 leaf-level
# evaluations x 22 + 196512 internal-level evaluations x 47
ALG_SLOTS_PER_SAMPLE = 393216 * 22 + 196512 * 47
ISSUED_FP64_PER_SAMPLE = 31035518826 * 32 / 75776  # executed DFMA+DMUL+DADD warp-instructions x 32 lanes / samples (ncu)


def synth_points(j):
    """Density j: K=4 isotropic Gaussians (sigma 0.6) on corners of {+-2}^3, shifted 0.25*j along dim 1."""
    rng = np.random.default_rng(SEED + j)
    corners = np.array([[-2, -2, -2], [-2, -2, 2], [-2, 2, -2], [-2, 2, 2]], dtype=np.float64)
    comp = rng.integers(0, 4, size=NCOMP)
    pts = corners[comp].T + 0.6 * rng.standard_normal((DIM, NCOMP))
    pts[0, :] += 0.25 * j
    return pts


def silverman(pts):
    d, N = pts.shape
    return pts.std(axis=1, ddof=1) * (4.0 / ((d + 2.0) * N)) ** (1.0 / (d + 4.0))


def host_cores():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return os.cpu_count() or 1


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, line in self.rows:
            if t < t0 or t > t1:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except Exception:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def run_reference(args, rank, world):
    """CPU arm: oracle port, all host threads, bounded sample of the same workload."""
    if rank != 0:
        return
    from oracle import oracle as O
    trees = [O.OKDE.kde_bw(synth_points(j), silverman(synth_points(j))) for j in range(NDENS)]
    cores = host_cores()  # torchrun exports OMP_NUM_THREADS=1: ask the OS instead
    n = args.ref_samples if args.ref_samples > 0 else 16 * cores
    nU, nN = O.prod_sizes(trees, n, NITER)
    rng = np.random.default_rng(SEED)
    U, G = rng.random(nU), rng.standard_normal(nN)
    for _ in range(args.warmup if args.warmup < 1 else 1):
        O.gibbs(trees, min(n, cores), NITER, U, G, nthreads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.gibbs(trees, n, NITER, U, G, nthreads=cores)
    dt = (time.perf_counter() - t0) / args.steps
    val = n / dt
    sample = "%d of %d samples per step (C4 shape, injected numpy streams), %d OpenMP threads over chains" % (
        n, SAMPLES_PER_GPU, cores)
    print(json.dumps({
        "impl": "reference", "metric": "product samples/sec (Gibbs, Niter=5)", "value": val, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(),
        "cpu_baseline": {"value": val, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def workload_config():
    return {"workload": "C4: prodAppxMSGibbsS, 8 densities x 4096 components, 3-D, Niter=5, 1M samples per GPU per step",
            "ndens": NDENS, "components": NCOMP, "dims": DIM, "niter": NITER, "samples_per_gpu": SAMPLES_PER_GPU,
            "rng": "Philox4x32-10 (seed %d), free-running" % SEED,
            "bandwidth": "Silverman", "l2": "flushed between timed steps (256 MiB write); trees (5.3 MB) are L2-resident by design",
            "multi_gpu": "weak: rank r draws samples [r*1M,(r+1)*1M) of one N*1M-sample run, NCCL all-gather of points+labels in the timed step"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--samples", type=int, default=SAMPLES_PER_GPU, help="samples per GPU per step (default = the named config)")
    ap.add_argument("--ref-samples", type=int, default=0)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.warmup < 3:
        args.warmup = 3  # timing rule: W >= 3
    # stdout carries exactly ONE JSON line: NCCL writes its banner / logs to stdout by default (seen at 4 and 8
    # GPUs even at WARN level), so send them to stderr whatever level the caller asked for
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"

    import torch
    import torch.distributed as dist
    import kde_b200 as K
    from kde_b200 import _lib, api as _api, dist as _dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    K.init(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = _lib.lib()
    n_per = args.samples
    Np_total = n_per * world
    s0, s1 = rank * n_per, (rank + 1) * n_per

    pts = [synth_points(j) for j in range(NDENS)]
    trees = [K.kde(p, silverman(p)) for p in pts]
    handles = _api._handles(trees)
    nlev, perU, perN, evals = K.gibbs_sizes(trees, NITER)
    tree_bytes = 0
    for t in trees:
        b = C.c_int64(0)
        L.kdeb200_tree_info(t._dev(), None, None, None, C.byref(b))
        tree_bytes += b.value

    dev = torch.device("cuda", local)
    d_pts = torch.empty((n_per, DIM), dtype=torch.float64, device=dev)
    d_idx = torch.empty((n_per, NDENS), dtype=torch.int64, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    g_pts = g_idx = None
    if world > 1:
        g_pts = torch.empty((Np_total, DIM), dtype=torch.float64, device=dev)
        g_idx = torch.empty((Np_total, NDENS), dtype=torch.int64, device=dev)

    def step():  # the sharded product: this rank's block + NCCL all-gather (kde_b200.dist)
        _dist.prod_sharded_device(handles, NDENS, DIM, Np_total, NITER, SEED, d_pts, d_idx, g_pts, g_idx)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    time.sleep(0.3)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_wall0 = time.time()
    for a, b in evs:
        flush.fill_(1)  # L2 flush, outside the event bracket
        a.record()
        step()
        b.record()
    barrier()
    t_wall1 = time.time()
    ms = [a.elapsed_time(b) for a, b in evs]
    total_ms = float(sum(ms))
    if world > 1:
        tt = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms = float(tt.item())
    clocks = sampler.stop(t_wall0, t_wall1)
    value = Np_total * args.steps / (total_ms * 1e-3)

    # kernel-only duration for the roofline: same launch, no collective, events on the launch stream
    kms = []
    for _ in range(max(2, min(args.steps, 3))):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(L.kdeb200_gibbs_device(handles, NDENS, Np_total, NITER, 1, None, None, 0, None, 0, SEED, s0, s1,
                                          d_pts.data_ptr(), d_idx.data_ptr(), None, st))
        b.record()
        torch.cuda.synchronize()
        kms.append(a.elapsed_time(b))
    k_ms = float(np.mean(kms))

    # e2e through the host API with host buffers (tree flatten + H2D, D2H of points and labels)
    e2e_n = n_per
    e2e_t = []
    for i in range(1 + max(1, min(args.steps, 2))):
        for t in trees:
            t._invalidate()
        flush.fill_(1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        p_host, i_host = K.prodAppxMSGibbsS(None, trees, None, None, Niter=NITER, Np=Np_total, seed=SEED, s0=s0,
                                            s1=s0 + e2e_n)
        float(p_host[0, 0])
        dt = time.perf_counter() - t0
        if i > 0:
            e2e_t.append(dt)
    e2e_dt = float(np.mean(e2e_t))
    if world > 1:
        tt = torch.tensor([e2e_dt], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_dt = float(tt.item())
    e2e_val = e2e_n * world / e2e_dt

    out = None
    if rank == 0:
        # roofline denominators measured live: dependency-free DFMA stream
        dfma, _ = K.pipe_peak(0, 200000)
        ach = ALG_SLOTS_PER_SAMPLE * n_per / (k_ms * 1e-3)  # algorithmic FP64-pipe slots / s
        peaks_file = {}
        try:
            peaks_file = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        roof = {"bound": "fp64_fma_pipe", "achieved": ach * 2 / 1e12, "peak": dfma * 2 / 1e12, "unit": "TFLOP/s",
                "frac": ach / dfma, "traffic": None,
                "traffic_note": "ncu --set full on a 75,776-sample launch of the same kernel (profiles/r01_gibbs_v9_ncu.txt): "
                                "dram read 0.19 GB + write 1.88 GB (local-memory checkpoints leaving L2), i.e. ~27 KB/sample "
                                "or <0.5% of HBM bandwidth; a full 1M-sample launch does not finish under ncu replay",
                "kernel": "gibbs_kernel<3,false>", "kernel_ms": k_ms,
                # honest "issued" view next to the algorithmic one (SURVEY.md 8d asks for both): the kernel issues
                # 13.1e6 FP64 instructions per sample (ncu, profiles/r01_gibbs_v9_ncu.txt) -- fewer than the model's
                # 17.9e6 slots because exp costs 7 instructions instead of 14 -- so frac can exceed 1
                "issued_fp64_instr_per_sample": ISSUED_FP64_PER_SAMPLE,
                "issued_frac": ISSUED_FP64_PER_SAMPLE * n_per / (k_ms * 1e-3) / dfma,
                "algorithmic_fp64_slots_per_sample": ALG_SLOTS_PER_SAMPLE, "kernel_evals_per_sample": evals,
                "peak_source": "DFMA microbenchmark (kdeb200_pipe_peak) measured in this run; nominal 64/clk/SM x 148 x 1.965 GHz = 37.2 TFLOP/s",
                # the HBM view of the same launch, for the record: compulsory bytes = trees in + points and labels out
                "hbm_view": {"algorithmic_bytes": tree_bytes + n_per * (DIM * 8 + NDENS * 8),
                             "achieved_GBs": (tree_bytes + n_per * (DIM * 8 + NDENS * 8)) / (k_ms * 1e-3) / 1e9,
                             "peak_GBs": peaks_file.get("hbm_gbs"),
                             "frac": ((tree_bytes + n_per * (DIM * 8 + NDENS * 8)) / (k_ms * 1e-3) / 1e9 / peaks_file["hbm_gbs"])
                             if peaks_file.get("hbm_gbs") else None},
                "note": "path is FP64-FMA-pipe bound, not HBM/tensor (SURVEY.md 8d): the HBM view above is ~1e-5 of the measured copy bandwidth"}
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            from oracle import oracle as O
            otrees = [O.OKDE.kde_bw(p, silverman(p)) for p in pts]
            cores = host_cores()
            rng = np.random.default_rng(SEED)
            n0 = 2 * cores
            nU, nN = O.prod_sizes(otrees, n0, NITER)
            U, G = rng.random(nU), rng.standard_normal(nN)
            t0 = time.perf_counter()
            O.gibbs(otrees, n0, NITER, U, G, nthreads=cores)
            pilot = time.perf_counter() - t0
            n = int(max(n0, min(65536, n0 * args.cpu_seconds / max(pilot, 1e-3))))
            n = (n // cores) * cores
            nU, nN = O.prod_sizes(otrees, n, NITER)
            U, G = rng.random(nU), rng.standard_normal(nN)
            t0 = time.perf_counter()
            O.gibbs(otrees, n, NITER, U, G, nthreads=cores)
            dt = time.perf_counter() - t0
            t0 = time.perf_counter()
            O.gibbs(otrees, 16, NITER, U, G, nthreads=1)
            one = 16 / (time.perf_counter() - t0)
            cpu = {"value": n / dt, "unit": "samples/s", "cores": cores, "kind": "port", "one_thread_value": one,
                   "sample": "%d of %d samples of the same workload in %.1f s (literal C restatement of the reference, OpenMP over chains; the reference itself is single-threaded Julia, not installed)" % (n, n_per, dt)}
        out = {
            "metric": "product samples/sec (Gibbs, Niter=5)", "value": value, "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(), "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": "samples/s", "h2d_bytes_per_step": tree_bytes,
                    "d2h_bytes_per_step": e2e_n * (DIM * 8 + NDENS * 8), "ms_per_step": e2e_dt * 1e3,
                    "api": "kde_b200.prodAppxMSGibbsS -> kdeb200_tree_create x8 + kdeb200_gibbs (host buffers)"},
            "gpu_launches": args.steps, "roofline": roof, "cpu_baseline": cpu,
            "kernel_evals_per_s": evals * value,
        }
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
