"""Timing of the batched exponential-family terms at BASELINE configs[4] (1M local factors)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
import lrvb_b200 as vb
ef = vb.ExponentialFamilies
M = 1_000_000
g = torch.Generator(device="cuda"); g.manual_seed(5)
shape = torch.exp(0.5 * torch.randn(M, dtype=torch.float64, device="cuda", generator=g))
rate = torch.exp(0.5 * torch.randn(M, dtype=torch.float64, device="cuda", generator=g))
alpha = 10 * torch.rand(5, M, dtype=torch.float64, device="cuda", generator=g) + 0.1
tau = 5 * torch.rand(M, 2, dtype=torch.float64, device="cuda", generator=g) + 0.1
A = torch.randn(M, 2, 2, dtype=torch.float64, device="cuda", generator=g)
v = A @ A.transpose(1, 2) + torch.eye(2, dtype=torch.float64, device="cuda")
df = 3 + 5 * torch.rand(M, dtype=torch.float64, device="cuda", generator=g)
info = torch.exp(0.5 * torch.randn(M, dtype=torch.float64, device="cuda", generator=g))
cases = [
    ("gamma_entropy_batched", lambda: ef.gamma_entropy_batched(shape, rate), 24),
    ("get_e_log_gamma", lambda: ef.get_e_log_gamma(shape, rate), 24),
    ("dirichlet_entropy (d=5)", lambda: ef.dirichlet_entropy(alpha), 48),
    ("get_e_log_dirichlet (d=5)", lambda: ef.get_e_log_dirichlet(alpha), 80),
    ("beta_entropy_batched", lambda: ef.beta_entropy_batched(tau), 24),
    ("univariate_normal_entropy_batched", lambda: ef.univariate_normal_entropy_batched(info), 16),
    ("wishart_entropy (k=2, batched)", lambda: ef.wishart_entropy(df, v), 48),
]
for name, fn, bytes_per in cases:
    try:
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print("%-36s %.3f ms  %.1f G factors/s  %.0f GB/s" % (name, ms, M / ms / 1e6, M * bytes_per / ms / 1e6), flush=True)
    except Exception as exc:
        print(name, "failed:", exc)
