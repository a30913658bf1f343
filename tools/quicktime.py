"""Scratch timing of the evaluation pipeline at BASELINE configs (not the bench)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import sys, time
import numpy as np, torch
import lrvb_b200 as vb
from oracle import glmm_oracle as go

def run(N, K, G, Q, reps=10):
    rng = np.random.default_rng(0)
    X = torch.randn(N, K, dtype=torch.float64, device="cuda")
    base, rem = divmod(N, G)
    counts = torch.full((G,), base, dtype=torch.int64); counts[:rem] += 1
    g = torch.repeat_interleave(torch.arange(G), counts).cuda()
    y = (torch.rand(N, device="cuda") < 0.5).double()
    model = vb.LogisticGLMM(X, y, g, num_gh_points=Q, num_groups=G)
    x = torch.randn(model.D, dtype=torch.float64, device="cuda") * 0.1
    for order in (0, 1, 2):
        for _ in range(3):
            model.evaluate(x, order, force=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            model.evaluate(x, order, force=True)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print("N=%d K=%d G=%d Q=%d order=%d: %.3f ms  %.3f Gobs/s" % (N, K, G, Q, order, ms, N / ms / 1e6), flush=True)
    model.evaluate(x, 2, force=True)
    v = torch.randn(model.D, dtype=torch.float64, device="cuda")
    for name, fn in (("hvp", lambda: model.hvp(v)), ("csr", lambda: model.hessian_csr()),
                     ("schur+inv", lambda: model.global_covariance())):
        for _ in range(2): fn()
        torch.cuda.synchronize(); t = time.time()
        for _ in range(5): fn()
        torch.cuda.synchronize()
        print("   %s: %.3f ms" % (name, (time.time() - t) / 5 * 1e3), flush=True)
    t = time.time(); xs, info, iters = model.cg(v, None, precond=1, rtol=1e-8, maxiter=2000); torch.cuda.synchronize()
    print("   cg(block_jacobi): info=%d iters=%d %.1f ms" % (info, iters, (time.time() - t) * 1e3), flush=True)

if __name__ == "__main__":
    run(5000, 5, 100, 4)
    run(1000000, 20, 10000, 8)
    run(10000000, 50, 100000, 8, reps=3)
    a = torch.randn(8192, 8192, dtype=torch.float64, device="cuda"); b = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
    for _ in range(2): a @ b
    torch.cuda.synchronize(); t = time.time(); a @ b; torch.cuda.synchronize()
    print("DGEMM 8192^3: %.1f TFLOP/s" % (2 * 8192**3 / (time.time() - t) / 1e12))
