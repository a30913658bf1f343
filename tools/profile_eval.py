"""Tiny driver for ncu: a few order-2 evaluations (+ CSR, HVP) at a named config."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import sys
import torch
import lrvb_b200 as vb

CFG = {"c1": (5000, 5, 100, 4), "c2": (1000000, 20, 10000, 8), "c3": (10000000, 50, 100000, 8),
       "c3s": (2000000, 50, 20000, 8), "c4s": (500000, 200, 5000, 8)}

def main(name, reps=2, extras=True):
    N, K, G, Q = CFG[name]
    torch.manual_seed(0)
    X = torch.randn(N, K, dtype=torch.float64, device="cuda")
    base, rem = divmod(N, G)
    counts = torch.full((G,), base, dtype=torch.int64); counts[:rem] += 1
    g = torch.repeat_interleave(torch.arange(G), counts).cuda()
    y = (torch.rand(N, device="cuda") < 0.5).double()
    model = vb.LogisticGLMM(X, y, g, num_gh_points=Q, num_groups=G)
    x = torch.randn(model.D, dtype=torch.float64, device="cuda") * 0.1
    for _ in range(reps):
        model.evaluate(x, 2, force=True)
        if extras:
            model.hessian_csr()
            model.hvp(x)
    torch.cuda.synchronize()
    print("ok", float(model.kl_tensor()))

if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 2)
