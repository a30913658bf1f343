"""Per-call breakdown of bench.py's end-to-end step at C2 (which part of the 1-3 ms is host code)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import lrvb_b200 as vb
import bench


def main():
    wl = bench.WORKLOADS["c2"]
    X, y, g = bench.synth_shard(torch, wl["N"], wl["K"], wl["G"], 2000, torch.device("cuda", 0))
    model = vb.LogisticGLMM(X, y, g, num_gh_points=8, num_groups=wl["G"])
    obj = vb.Objective(model.glmm_par, model)
    rng = np.random.default_rng(0)
    xs = [rng.normal(size=model.D) * 0.1 for _ in range(8)]
    acc = {}

    def lap(name, t0):
        t1 = time.perf_counter()
        acc[name] = acc.get(name, 0.0) + (t1 - t0)
        return t1
    for rep in range(25):
        if rep == 5:
            acc.clear()
        x = xs[rep % 8]
        t = time.perf_counter()
        model.evaluate(x, 2); t = lap("evaluate (launch)", t)
        obj._set_par(x, "free"); t = lap("_set_par", t)
        csr = model.hessian_csr(); t = lap("hessian_csr (launch)", t)
        H = csr.to_scipy(); t = lap("to_scipy (sync + D2H)", t)
        gr = obj.fun_free_grad(x); t = lap("fun_free_grad", t)
        kl = obj.fun_free(x); t = lap("fun_free", t)
    for k, v in acc.items():
        print("%-28s %.3f ms" % (k, v / 20 * 1e3))
    print("total %.3f ms" % (sum(acc.values()) / 20 * 1e3))
    import cProfile, pstats
    pr = cProfile.Profile()
    pr.enable()
    for rep in range(20):
        x = xs[rep % 8]
        H = obj.fun_free_hessian(x); gr = obj.fun_free_grad(x); kl = obj.fun_free(x)
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(18)


if __name__ == "__main__":
    main()
