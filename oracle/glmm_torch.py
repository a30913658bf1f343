"""torch-CPU-fp64 autodiff restatement of the GLMM KL.  TEST INFRASTRUCTURE ONLY.

Stands in for the reference's ``autograd`` differentiation of the composed objective
(SparseObjectives.py:102-113: grad / hessian / hessian_vector_product of ``fun_free``).
The forward code below is a line-for-line torch spelling of oracle/glmm_oracle.py's
``kl`` (itself pinned against the reference's forward code by the golden vectors), so
``torch.func.grad / hessian / jvp(grad)`` give derivatives that are independent of the
hand-derived analytic formulas.  It is also the multi-threaded CPU baseline bench.py times
as "the reference's autograd path" (BASELINE.md section 3).
"""
import math

import numpy as np
import torch


_SHIFT = 50


def digamma_acc(a):
    """digamma by upward recurrence to a+50: torch's CPU trigamma (the derivative autodiff
    uses) is only ~5e-10 accurate below x~20 but full precision at x>=50."""
    j = torch.arange(_SHIFT, dtype=a.dtype)
    return torch.digamma(a + _SHIFT) - torch.sum(1.0 / (a[..., None] + j), dim=-1)


def lgamma_acc(a):
    j = torch.arange(_SHIFT, dtype=a.dtype)
    return torch.lgamma(a + _SHIFT) - torch.sum(torch.log(a[..., None] + j), dim=-1)


class GLMMTorch:
    def __init__(self, oracle):
        o = oracle
        self.o = o
        self.X = torch.from_numpy(o.X)
        self.S = self.X * self.X
        self.y = torch.from_numpy(o.y)
        self.g = torch.from_numpy(o.g)
        self.w = None if o.w is None else torch.from_numpy(o.w)
        self.c = torch.from_numpy(np.sqrt(2) * o.gh_x)
        self.what = torch.from_numpy(o.gh_w / np.sqrt(np.pi))
        lb = o.lower_bounds()
        self.con = torch.from_numpy(np.isfinite(lb))
        self.lb = torch.from_numpy(np.where(np.isfinite(lb), lb, 0.0))

    def kl(self, free):
        o, lay, pr = self.o, self.o.lay, self.o.prior
        K, G = lay.K, lay.G
        vec = torch.where(self.con, torch.exp(free) + self.lb, free)   # Parameters.py:47-61
        mu_m, mu_i, a, b = vec[0], vec[1], vec[2], vec[3]
        beta_m, beta_i = vec[4:4 + K], vec[4 + K:4 + 2 * K]
        u_m, u_i = vec[lay.u_mean:lay.u_mean + G], vec[lay.u_info:lay.u_info + G]
        z_m = u_m[self.g] + self.X @ beta_m
        z_v = (1.0 / u_i)[self.g] + self.S @ (1.0 / beta_i)
        z_s = torch.sqrt(z_v)
        t = z_m[:, None] + z_s[:, None] * self.c[None, :]
        A = torch.sum(self.what * torch.log1p(torch.exp(t)), dim=1)    # Modeling.py:48
        ell = self.y * z_m - A
        if self.w is not None:
            ell = self.w * ell
        e_tau = a / b
        e_log_tau = digamma_acc(a) - torch.log(b)
        re = torch.sum(-0.5 * e_tau * ((mu_m - u_m) ** 2 + 1 / mu_i + 1 / u_i)) + 0.5 * G * e_log_tau
        loglik = torch.sum(ell) + re
        l2pi = math.log(2 * math.pi)
        ent = (0.5 * (-torch.log(mu_i) + 1 + l2pi)
               + 0.5 * torch.sum(-torch.log(beta_i) + 1 + l2pi)
               + 0.5 * torch.sum(-torch.log(u_i) + 1 + l2pi)
               + a - torch.log(b) + lgamma_acc(a) + (1 - a) * digamma_acc(a))
        prior = (-0.5 * pr.mu_info * ((mu_m - pr.mu_mean) ** 2 + 1 / mu_i)
                 + torch.sum(-0.5 * pr.beta_info * ((beta_m - pr.beta_mean) ** 2 + 1 / beta_i))
                 + (pr.tau_shape - 1) * e_log_tau - pr.tau_rate * e_tau)
        return -(loglik + ent + prior)

    def value(self, free):
        return float(self.kl(torch.as_tensor(free, dtype=torch.float64)))

    def grad(self, free):
        f = torch.as_tensor(free, dtype=torch.float64)
        return torch.func.grad(self.kl)(f).numpy()

    def hessian(self, free):
        f = torch.as_tensor(free, dtype=torch.float64)
        return torch.func.hessian(self.kl)(f).numpy()

    def hvp(self, free, v):
        f = torch.as_tensor(free, dtype=torch.float64)
        v = torch.as_tensor(v, dtype=torch.float64)
        return torch.func.jvp(torch.func.grad(self.kl), (f,), (v,))[1].numpy()
