"""Gram-kernel time (events inside lrvb_glmm_eval) for several K at fixed N; prints achieved DMMA rates."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import ctypes, sys
import torch
import lrvb_b200 as vb
from lrvb_b200 import _native as nat

def run(N, K, G=1000, Q=int(os.environ.get("GT_Q", "4"))):
    X = torch.randn(N, K, dtype=torch.float64, device="cuda")
    g = torch.repeat_interleave(torch.arange(G), N // G).cuda()
    y = (torch.rand(N, device="cuda") < 0.5).double()
    model = vb.LogisticGLMM(X, y, g, num_gh_points=Q, num_groups=G)
    lib = nat.load()
    nat.check(lib.lrvb_glmm_set_timing(model._h, 1))
    x = torch.randn(model.D, dtype=torch.float64, device="cuda") * 0.1
    ms3 = (ctypes.c_float * 3)()
    ts = []
    for i in range(6):
        model.evaluate(x, 2, force=True)
        torch.cuda.synchronize()
        nat.check(lib.lrvb_glmm_last_timing(model._h, ms3))
        if i >= 2: ts.append(ms3[2])
    t = sorted(ts)[len(ts) // 2]
    T2 = (2 * K + 7) // 8
    tiles = T2 * (T2 + 1) // 2 + (1 if K % 8 else 0)
    alg = N * (4 * K * K + 2 * K) / (t * 1e-3) / 1e12
    exe = N / 4 * tiles * 512 / (t * 1e-3) / 1e12
    print("N=%d K=%d: gram %.3f ms  obs %.3f ms  algorithmic %.1f TF  executed %.1f TF (%.0f%% of 37.1)" % (
        N, K, t, ms3[1], alg, exe, 100 * exe / 37.1), flush=True)

if __name__ == "__main__":
    Ks = [int(a) for a in sys.argv[1:]] or [24, 32, 50, 64, 100, 200]
    for K in Ks:
        run(2_000_000 if K <= 64 else 1_000_000, K)
