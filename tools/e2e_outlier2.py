"""Which operation of the 7th end-to-end call is slow: wall clock of every stage of hessian_scipy for the first
calls, with the device / pinned-host allocator counters."""
import os, sys, time, gc
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import lrvb_b200 as vb
import bench

wl = bench.WORKLOADS["c2"]
dev = torch.device("cuda", 0)
X, y, g = bench.synth_shard(torch, wl["N"], wl["K"], wl["G"], 2000, dev)
model = vb.LogisticGLMM(X, y, g, num_gh_points=wl["Q"], num_groups=wl["G"])
rng = np.random.default_rng(0)
xs = [0.1 * rng.standard_normal(model.D) for _ in range(8)]
gc.collect(); gc.disable()
keep = None
def stats():
    d = torch.cuda.memory_stats()
    try:
        h = torch.cuda.host_memory_stats()
        hs = (h.get("num_host_alloc", -1), h.get("num_host_free", -1), h.get("segment.allocated", -1))
    except Exception as e:
        hs = (-1, -1, -1)
    return (d["num_device_alloc"], d["num_device_free"]) + hs
for i in range(16):
    x = xs[i % 8]
    T = [time.perf_counter()]
    model.evaluate(x, 2); T.append(time.perf_counter())
    csr = model.hessian_csr(); T.append(time.perf_counter())
    model._enqueue_host_copies(); T.append(time.perf_counter())
    pat = csr._resolve(); nnz = pat.nnz; T.append(time.perf_counter())
    hv = torch.empty(nnz, dtype=torch.float64, pin_memory=True); T.append(time.perf_counter())
    hv.copy_(csr._val[:nnz], non_blocking=True); T.append(time.perf_counter())
    torch.cuda.current_stream().synchronize(); T.append(time.perf_counter())
    a = hv.numpy(); T.append(time.perf_counter())
    keep = (a, csr); T.append(time.perf_counter())
    d = 1e3 * np.diff(T)
    print("call %2d: eval %.3f csr %.3f enq %.3f resolve %.3f pin %.3f copy %.3f sync %.3f numpy %.3f rebind %.3f | total %.3f | %s" % (
        (i + 1,) + tuple(d) + (d.sum(), stats())))
