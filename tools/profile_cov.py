"""Tiny driver for ncu: the LRVB global covariance (Schur + SPD inverse) and one direct solve at a named config."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lrvb_b200 as vb
CFG = {"c2": (1000000, 20, 10000, 8), "c3s": (2000000, 50, 20000, 8), "c3g": (1000000, 50, 100000, 8)}
N, K, G, Q = CFG[sys.argv[1]]
torch.manual_seed(0)
X = torch.randn(N, K, dtype=torch.float64, device="cuda")
g = torch.repeat_interleave(torch.arange(G), N // G).cuda()
y = (torch.rand(N, device="cuda") < 0.5).double()
model = vb.LogisticGLMM(X, y, g, num_gh_points=Q, num_groups=G)
x = torch.randn(model.D, dtype=torch.float64, device="cuda") * 0.1
model.evaluate(x, 2, force=True)
for _ in range(2):
    cov = model.global_covariance()
    sol = model.solve(x)
torch.cuda.synchronize()
print("ok", float(cov[0, 0]))
