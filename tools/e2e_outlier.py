"""Per-call wall-clock of the end-to-end step (numpy x in; scipy CSR + numpy gradient + float out) over many
steps: median of every sub-call and every step that took more than 1.5x the median, with the sub-call that
was slow.  Finds the outliers that inflate the MEAN e2e step of bench.py."""
import os, sys, time, gc
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import lrvb_b200 as vb
import bench

wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
dev = torch.device("cuda", 0)
X, y, g = bench.synth_shard(torch, wl["N"], wl["K"], wl["G"], 2000, dev)
model = vb.LogisticGLMM(X, y, g, num_gh_points=wl["Q"], num_groups=wl["G"])
obj = vb.Objective(model.glmm_par, model)
rng = np.random.default_rng(0)
xs = [0.1 * rng.standard_normal(model.D) for _ in range(8)]
for i in range(5):
    obj.fun_free_hessian(xs[i]); obj.fun_free_grad(xs[i]); obj.fun_free(xs[i])
gc.collect(); gc.disable()
torch.cuda.synchronize()
rows = []
keep = None
for i in range(steps):
    x = xs[(5 + i) % 8]
    t0 = time.perf_counter()
    H = obj.fun_free_hessian(x)
    t1 = time.perf_counter()
    gr = obj.fun_free_grad(x)
    t2 = time.perf_counter()
    kl = obj.fun_free(x)
    t3 = time.perf_counter()
    keep = (H, gr, kl)          # as in bench.py: the previous result dies when the new one is bound
    t4 = time.perf_counter()
    rows.append((t1 - t0, t2 - t1, t3 - t2, t4 - t3))
r = 1e3 * np.asarray(rows)
tot = r.sum(1)
med = np.median(tot)
print("steps %d: total median %.3f mean %.3f min %.3f max %.3f ms" % (steps, med, tot.mean(), tot.min(), tot.max()))
print("median per call: hessian %.3f grad %.3f fun %.3f rebind %.3f" % tuple(np.median(r, 0)))
for i in np.nonzero(tot > 1.5 * med)[0]:
    print("  step %3d: total %.3f = hessian %.3f grad %.3f fun %.3f rebind %.3f" % ((i, tot[i]) + tuple(r[i])))
