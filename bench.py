#!/usr/bin/env python
"""bench.py -- logistic-GLMM observations/sec for one fused ELBO + gradient + sparse-Hessian
evaluation (BASELINE.json metric) on N B200s, plus the end-to-end number through the public
Objective API, the roofline of the dominant kernel and the CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c3]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

A step = one pass of the hot path over one batch of synthetic observations already resident in
HBM: lrvb_glmm_eval(order 2) [+ NCCL all-reduce of the packed (KL, global gradient, global
Hessian block) when N > 1] + device CSR assembly of the arrowhead Hessian.  N = 1 workload is
BASELINE.json configs[1] (N=1M, K=20, G=10k, Q=8); with N > 1 every rank holds a shard of that
size (weak scaling; groups are rank-private, globals all-reduced).  Prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # per-GPU shard: observations, fixed effects, groups, Gauss-Hermite points
    "c2": dict(N=1_000_000, K=20, G=10_000, Q=8,
               name="logistic GLMM N=1M K=20 G=10k Q=8 per GPU (BASELINE configs[1])"),
    "c3": dict(N=10_000_000, K=50, G=100_000, Q=8,
               name="logistic GLMM N=10M K=50 G=100k Q=8 per GPU (BASELINE configs[2] shard)"),
    "c1": dict(N=5_000, K=5, G=100, Q=4, name="logistic GLMM N=5k K=5 G=100 Q=4 (BASELINE configs[0])"),
}
METRIC = "GLMM obs/sec for ELBO+grad+sparse Hessian"
UNIT = "obs/s"
CPU_SAMPLE = dict(N=500_000, G=5_000)   # bounded sample of the workload for the CPU arm


# ------------------------------------------------------------------------------------------
class ClockSampler(object):
    """SM clock / throttle reasons sampled DURING the timed region, in-process through NVML
    (a thread polling every few ms; spawning nvidia-smi stalls the GPU for milliseconds and its
    100 ms period misses a short timed region).  Falls back to one nvidia-smi query."""
    REASONS = (("hw_slowdown", 0x8), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40),
               ("hw_power_brake_slowdown", 0x80), ("sw_power_cap", 0x4))

    def __init__(self, index, period_s=0.002):
        self.index, self.period = index, period_s
        self.samples, self.reasons, self.power = [], set(), []
        self.handle, self.nvml, self.max_mhz = None, None, None
        self._stop = threading.Event()
        self.thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if vis:
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if index < len(ids) and ids[index].isdigit():
                    phys = int(ids[index])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception as exc:       # noqa: BLE001
            self.error = "NVML unavailable: %s" % exc

    def _poll(self):
        nv = self.nvml
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)))
                mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                for name, bit in self.REASONS:
                    if mask & bit:
                        self.reasons.add(name)
                if len(self.samples) % 32 == 1:    # the power query is the slow one: sparse
                    self.power.append(nv.nvmlDeviceGetPowerUsage(self.handle) / 1000.0)
            except Exception:          # noqa: BLE001
                pass
            self._stop.wait(self.period)

    def start(self):
        if self.nvml is not None:
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()

    def stop(self):
        if self.thread is not None:
            self._stop.set()
            self.thread.join(timeout=2)
        if self.samples:
            return dict(sm_mhz=float(np.median(self.samples)), sm_max_mhz=self.max_mhz,
                        reasons=sorted(self.reasons), samples=len(self.samples),
                        power_w_max=max(self.power) if self.power else None, source="nvml")
        try:
            out = subprocess.run(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=clocks.sm,clocks.max.sm",
                 "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=20).stdout
            a, b2 = [float(v) for v in out.strip().split(",")[:2]]
            return dict(sm_mhz=a, sm_max_mhz=b2, reasons=[], samples=1,
                        source="nvidia-smi after the timed region (NVML polling failed)")
        except Exception:              # noqa: BLE001
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["clock sampling unavailable"], samples=0)


def synth_shard(torch, N, K, G, seed, device):
    """Seeded synthetic shard on the device: X ~ N(0,1), balanced group-sorted ids,
    y ~ Bernoulli(sigmoid(X beta + u_g))."""
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    X = torch.randn(N, K, dtype=torch.float64, device=device, generator=gen)
    base, rem = divmod(N, G)
    counts = torch.full((G,), base, dtype=torch.int64, device=device)
    counts[:rem] += 1
    g = torch.repeat_interleave(torch.arange(G, device=device), counts)
    beta = 0.5 * torch.randn(K, dtype=torch.float64, device=device, generator=gen)
    u = 0.3 + 0.5 * torch.randn(G, dtype=torch.float64, device=device, generator=gen)
    p = torch.sigmoid(X @ beta + u[g])
    y = (torch.rand(N, dtype=torch.float64, device=device, generator=gen) < p).double()
    return X, y, g


def measure_dgemm_tflops(torch, n=8192, reps=5):
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    for _ in range(2):
        a @ b
    best = float("inf")
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        a @ b
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a, b
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12


def cpu_oracle_step(sample, K, Q, seed=123):
    """One CPU pass (oracle port) over a bounded sample: KL + gradient + arrowhead blocks + CSR."""
    from oracle import glmm_oracle as go
    Ns, Gs = sample["N"], sample["G"]
    X, y, g = go.make_glmm_data(Ns, K, Gs, seed)
    gh_x, gh_w = np.polynomial.hermite.hermgauss(Q)
    o = go.GLMMOracle(X, y, g, gh_x, gh_w, G=Gs)
    x = go.make_free(o.lay.D, seed)

    def step():
        t = time.perf_counter()
        H = o.kl_hessian_csr(x)     # kl_blocks (value, gradient, Hessian blocks) + CSR emission
        return time.perf_counter() - t, H.nnz
    return step


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        n = [p.get("num_threads", 1) for p in threadpool_info() if p.get("user_api") == "blas"]
        return max(n) if n else 1
    except Exception:
        return os.cpu_count() or 1


# ------------------------------------------------------------------------------------------
def run_reference(args, wl):
    """--impl reference: the reference's CPU path for this metric (oracle port -- the reference
    itself needs autograd<1.4, not installable here), rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    step = cpu_oracle_step(CPU_SAMPLE, wl["K"], wl["Q"])
    for _ in range(max(1, min(args.warmup, 2))):
        step()
    times = [step()[0] for _ in range(args.steps)]
    ms = 1e3 * float(np.mean(times))
    value = CPU_SAMPLE["N"] / (ms * 1e-3)
    cores = blas_threads()
    sample = "N=%d obs, K=%d, G=%d, Q=%d (1/%d of the per-GPU workload) per step" % (
        CPU_SAMPLE["N"], wl["K"], CPU_SAMPLE["G"], wl["Q"], wl["N"] // CPU_SAMPLE["N"])
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["name"], "cpu_sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": sample,
                         "note": "numpy analytic restatement of the reference path (oracle/); "
                                 "BLAS parts use %d threads, elementwise parts 1" % cores},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_ours(args, wl):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    # stdout must carry the one JSON line only, but native libraries write there too (NCCL prints its
    # version banner with printf): point fd 1 at stderr for the whole run and keep the real stdout
    # for the result line
    sys.stdout.flush()
    result_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    if args.gpus != world and rank == 0:
        print("warning: --gpus %d but WORLD_SIZE %d" % (args.gpus, world), file=sys.stderr)

    import lrvb_b200 as vb
    from lrvb_b200 import _native as nat
    lib = nat.load()

    N, K, G, Q = wl["N"], wl["K"], wl["G"], wl["Q"]
    X, y, g = synth_shard(torch, N, K, G, seed=1000 * 2 + rank, device=device)
    if world > 1:
        from lrvb_b200 import distributed as vbd
        model = vbd.ShardedLogisticGLMM.from_local_shard(X, y, g, num_local_groups=G,
                                                         num_gh_points=Q)
    else:
        model = vb.LogisticGLMM(X, y, g, num_gh_points=Q, num_groups=G)
    obj = vb.Objective(model.glmm_par, model)
    D = model.D
    gen = torch.Generator(device=device)
    gen.manual_seed(7)
    x_dev = 0.1 * torch.randn(D, dtype=torch.float64, device=device, generator=gen)
    if world > 1:
        dist.broadcast(x_dev, 0)
    nsteps_total = args.warmup + args.steps
    x_host = [x_dev.cpu().numpy() + 1e-3 * np.random.default_rng(s).standard_normal(D)
              for s in range(min(nsteps_total, 8))]
    if world > 1:   # every rank must evaluate at the same points
        for xh in x_host:
            t = torch.from_numpy(xh).to(device)
            dist.broadcast(t, 0)
            xh[:] = t.cpu().numpy()

    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=device)
    local = model.local if world > 1 else model

    def device_step():
        model.evaluate(x_dev, 2, force=True)
        return model.hessian_csr()

    # ---------------- device-resident timing ----------------
    csr = None
    for _ in range(args.warmup):
        csr = device_step()                 # keep the previous result alive as the timed loop does:
    torch.cuda.synchronize()                # the allocator then owns both ping-pong result blocks
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if rank == 0:
        sampler.start()
    for _ in range(3):                      # the sampler thread's first NVML calls happen untimed
        flush.zero_()
        csr = device_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        flush.zero_()
        csr = device_step()                 # the first collective after a barrier pays a re-sync: untimed
        torch.cuda.synchronize()
    launches0 = lib.lrvb_launch_count()
    step_ms, gram_ms, obs_ms, eval_ms = [], [], [], []
    import ctypes
    ms3 = (ctypes.c_float * 3)()
    torch.cuda.synchronize()
    import gc
    gc.collect()
    gc.disable()                            # no collector pauses inside the timed steps
    for _ in range(args.steps):
        flush.zero_()                       # L2 flush (126 MB L2 < 256 MiB), outside the timing
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        csr = device_step()
        e1.record()
        e1.synchronize()
        step_ms.append(e0.elapsed_time(e1))
    torch.cuda.synchronize()
    gc.enable()
    launches = (lib.lrvb_launch_count() - launches0) // max(1, args.steps)
    if world > 1:
        dist.barrier()
    clocks = sampler.stop() if rank == 0 else None
    total_ms = torch.tensor([float(np.sum(step_ms))], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(total_ms.item()) / args.steps
    value = world * N / (ms_per_step * 1e-3)
    if rank == 0:
        sm = np.sort(np.asarray(step_ms))
        worst = np.argsort(step_ms)[-3:][::-1]
        print("step ms: min %.3f median %.3f p90 %.3f max %.3f; slowest steps %s" % (
            sm[0], sm[len(sm) // 2], sm[int(0.9 * (len(sm) - 1))], sm[-1],
            ", ".join("#%d %.3f" % (i, step_ms[i]) for i in worst)), file=sys.stderr)
    nnz = csr.nnz

    # ---------------- per-kernel durations for the roofline (separate, untimed-for-the-metric loop:
    # the CUDA events the library records between its kernels cut the chain of programmatic
    # dependent launches, so the metric above is measured without them) ----------------
    nat.check(lib.lrvb_glmm_set_timing(local._h, 1))
    for i in range(min(args.steps, 30) + 3):
        flush.zero_()
        csr = device_step()
        torch.cuda.synchronize()
        if i >= 3:
            nat.check(lib.lrvb_glmm_last_timing(local._h, ms3))
            eval_ms.append(ms3[0]); obs_ms.append(ms3[1]); gram_ms.append(ms3[2])
    nat.check(lib.lrvb_glmm_set_timing(local._h, 0))

    # ---------------- LRVB covariance of the global parameters (second half of the metric) -------
    # (H^-1)[:Dg,:Dg] by the Schur complement of the local blocks on the cached Hessian: DMMA Gram
    # over the border matrix [+ all-reduce of S when sharded] + SPD inverse; device timed, max over ranks
    for _ in range(3):
        cov = model.global_covariance()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ncov = 20
    c0.record()
    for _ in range(ncov):
        if world > 1:
            model._sinv = None              # the sharded model caches the result: recompute
        cov = model.global_covariance()
    c1.record()
    c1.synchronize()
    cov_ms = torch.tensor([c0.elapsed_time(c1) / ncov], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(cov_ms, op=dist.ReduceOp.MAX)
    cov_ms = float(cov_ms.item())

    # ---------------- end to end through the public API (host buffers) ----------------

    def e2e_step(i):
        xh = x_host[i % len(x_host)]
        if world > 1:
            # sharded: every rank brings ITS part of the distributed Hessian / gradient to its host
            model.evaluate(xh, 2)
            H = model.hessian_csr().to_scipy()
            gr = model.grad_local_layout().cpu().numpy()
            kl = float(model.kl_tensor().item())
            return H, gr, kl
        H = obj.fun_free_hessian(xh)     # scipy CSR on the host (order-2 evaluation, cached)
        gr = obj.fun_free_grad(xh)       # numpy (D,)
        kl = obj.fun_free(xh)            # float
        return H, gr, kl
    for i in range(args.warmup):
        e2e_step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        H, gr, kl = e2e_step(args.warmup + i)
    torch.cuda.synchronize()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = world * N / (float(e2e_s.item()) / args.steps)
    h2d = 8 * D
    d2h = 8 + 8 * D + 12 * int(H.nnz) + 4 * (D + 1)

    if rank == 0:
        # rooflines of the two kernels that carry the step: the fused per-observation pass (HBM
        # roofline per SURVEY.md 8(d): 8K + 12 algorithmic bytes per observation) and the packed
        # Gram kernel (FP64 tensor roofline: 4K^2 + 2K flops per observation).  `roofline` is the
        # one with the larger measured duration; the other is reported beside it.
        peak_tf = measure_dgemm_tflops(torch)
        flops_per_obs = 4 * K * K + 2 * K
        bytes_per_obs = 8 * K + 12
        gms, oms = float(np.mean(gram_ms)), float(np.mean(obs_ms))
        traffic = {}
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get(args.workload, {})
            except Exception:
                traffic = {}
        gram_kernel = ("lrvb::k_gram_small (DMMA.8x8x4, packed [x|s] triangle per warp)" if K < 16 else
                       "lrvb::k_gram_mid (DMMA.8x8x4, packed [x|s] triangle per warp / warp team)" if K <= 104 else
                       "lrvb::k_gram_big (DMMA.8x8x4, packed [x|s] rectangles)")
        roof_gram = {"bound": "tensor", "kernel": gram_kernel,
                     "achieved": N * flops_per_obs / (gms * 1e-3) / 1e12, "peak": peak_tf,
                     "unit": "TFLOP/s", "traffic": traffic.get("k_gram_dram_bytes"),
                     "peak_source": "cuBLAS DGEMM 8192^3 measured in this run (FP64 is not in "
                                    "MEASURED_PEAKS.json); DMMA.8x8x4 microbenchmark: 37.1 TF",
                     "kernel_ms": gms, "flops_per_obs": flops_per_obs}
        roof_gram["frac"] = roof_gram["achieved"] / peak_tf
        hbm = _hbm_peak()
        roof_obs = {"bound": "hbm", "kernel": "lrvb::k_obs_fused<2> (quadrature + per-group sums, TMA ring per warp)"
                    if K <= 62 else "lrvb::k_obs<2>",
                    "achieved": N * bytes_per_obs / (oms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                    "traffic": traffic.get("k_obs_dram_bytes"),
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy)" if os.path.exists(
                        os.path.join(ROOT, "MEASURED_PEAKS.json")) else "B200_PROFILING.md fallback",
                    "kernel_ms": oms, "bytes_per_obs": bytes_per_obs,
                    "note": "the kernel is bound by fp64 instruction throughput (exp / log1p chains of the "
                            "quadrature: ncu fp64 pipe 49%, issue 56%, 12 warps / SM at 161 registers), not by HBM; see profiles/"}
        roof_obs["frac"] = roof_obs["achieved"] / hbm
        roofline = dict(roof_obs if oms >= gms else roof_gram)
        roofline["other_kernel"] = roof_gram if oms >= gms else roof_obs
        roofline["eval_ms"] = float(np.mean(eval_ms))
        roofline["hbm_frac_of_step"] = (N * bytes_per_obs / (ms_per_step * 1e-3) / 1e9) / hbm
        cpu = None
        if world == 1 or True:
            stepf = cpu_oracle_step(CPU_SAMPLE, K, Q)
            stepf()
            ts = [stepf()[0] for _ in range(3)]
            cores = blas_threads()
            cpu = {"value": CPU_SAMPLE["N"] / float(np.median(ts)), "unit": UNIT, "cores": cores,
                   "kind": "port",
                   "sample": "N=%d obs, K=%d, G=%d, Q=%d (1/%d of the per-GPU workload), median of 3"
                             % (CPU_SAMPLE["N"], K, CPU_SAMPLE["G"], Q, N // CPU_SAMPLE["N"])}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["name"], "obs_per_gpu": N, "K": K, "groups_per_gpu": G,
                       "gh_points": Q, "free_dim": D, "hessian_nnz": int(nnz),
                       "parallelism": "obs-sharded x%d, NCCL all-reduce of (KL, grad_g, H_gg)" % world
                       if world > 1 else "single GPU",
                       "l2": "flushed between timed steps (256 MiB write); X alone is %d MB"
                             % (N * K * 8 // 10 ** 6)},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h,
                    "api": "Objective.fun_free_hessian/fun_free_grad/fun_free with numpy x; "
                           "scipy CSR + numpy gradient + float returned to the host"},
            "lrvb_covariance": _cov_entry(cov_ms, K, G, world, peak_tf),
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
            "clocks": clocks,
        }
        sys.stdout.flush()
        os.write(result_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.stdout.flush()
    os.dup2(result_fd, 1)
    os.close(result_fd)


def _cov_entry(cov_ms, K, G, world, peak_tf):
    """Second half of the metric: wall time of the global-parameter LRVB covariance and its fraction
    of the FP64 tensor roofline.  Algorithmic flops: per group B_g^T L_g^-1 B_g on the upper
    triangle (2 Dg^2 + 8 Dg), plus Dg^3 for the SPD inverse; at Dg = 44 the work is ~40 MFLOP per
    GPU, i.e. the number is launch / latency bound, not tensor bound."""
    Dg = 4 + 2 * K
    flops = world * G * (2 * Dg * Dg + 8 * Dg) + Dg ** 3
    tf = flops / (cov_ms * 1e-3) / 1e12
    return {"ms": cov_ms, "what": "(H^-1)[:Dg,:Dg] of the cached Hessian: Schur complement (DMMA) "
            "[+ sum over ranks] + SPD inverse, device time, max over ranks", "Dg": Dg,
            "algorithmic_flops": flops, "achieved_tflops": tf, "peak_tflops": peak_tf,
            "frac": tf / peak_tf, "bound": "latency (work far below one wave of the tensor pipe)"}


def _hbm_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0   # B200_PROFILING.md fallback


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3   # timing rules: W >= 3
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()
