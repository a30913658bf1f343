"""Oracle-backed stand-in for GLMM.LogisticGLMM (torch CPU tensors) so that the sharded algebra
of lrvb_b200.distributed can run over gloo without a GPU.  TEST INFRASTRUCTURE ONLY."""
import numpy as np
import scipy.sparse
import torch

from oracle import glmm_oracle as go


class _CSR(object):
    def __init__(self, m):
        self.m = m

    def to_scipy(self):
        return self.m


class OracleLocal(object):
    def __init__(self, X, y, g, G, w, include_global, gh_x=None, gh_w=None, num_gh_points=4, **_):
        if gh_x is None:
            gh_x, gh_w = np.polynomial.hermite.hermgauss(num_gh_points)
        self.o = go.GLMMOracle(X, y, g, gh_x, gh_w, weights=w, G=G)
        # the terms that depend on no observation and no group (priors / entropies of mu, beta, tau)
        self.o_empty = go.GLMMOracle(np.zeros((0, X.shape[1])), np.zeros(0), np.zeros(0, np.int64),
                                     gh_x, gh_w, G=0)
        self.include_global = include_global
        self.N, self.K, self.G = self.o.N, self.o.K, G
        self.Dg, self.D = self.o.lay.Dg, self.o.lay.D
        self.device = torch.device("cpu")
        self._out_global = torch.zeros(1 + self.Dg + self.Dg ** 2, dtype=torch.float64)
        self._grad_local = torch.zeros(2 * G, dtype=torch.float64)

    def evaluate(self, x, order, coords="free", force=False):
        assert coords == "free"
        x = x.numpy()
        kl, grad, blk = self.o.kl_blocks(x)
        if not self.include_global:
            xe = x[:self.Dg]
            kle, grade, blke = self.o_empty.kl_blocks(xe)
            kl -= kle
            grad = grad.copy()
            grad[:self.Dg] -= grade
            blk = dict(A=blk["A"] - blke["A"], B=blk["B"], L=blk["L"])
        self._blk = {k: v.copy() for k, v in blk.items() if k != "A"}
        Dg = self.Dg
        self._out_global[0] = kl
        self._out_global[1:1 + Dg] = torch.from_numpy(grad[:Dg])
        self._out_global[1 + Dg:] = torch.from_numpy(blk["A"].reshape(-1))
        self._grad_local[:] = torch.from_numpy(grad[Dg:])

    @property
    def blk(self):
        """B, L of the last evaluation and the LIVE global block: like the device model, A is the
        tail of ``_out_global``, so an in-place all-reduce of that buffer updates it."""
        d = dict(self._blk)
        d["A"] = self._out_global[1 + self.Dg:].numpy().reshape(self.Dg, self.Dg).copy()
        return d

    def blocks(self):
        return (torch.from_numpy(self.blk["A"]), torch.from_numpy(self.blk["B"]),
                torch.from_numpy(self.blk["L"]))

    def hvp_cached(self, v, out=None, include_A=True):
        blk = dict(self.blk)
        if not include_A:
            blk["A"] = np.zeros_like(blk["A"])
        return torch.from_numpy(self.o.hvp_blocks(blk, v.numpy()))

    def hessian_csr(self):
        H = scipy.sparse.csr_matrix(go.GLMMOracle.blocks_to_dense(self.o.lay, self.blk))
        H.sort_indices()
        return _CSR(H)

    def _linv(self):
        L = self.blk["L"]
        det = L[:, 0] * L[:, 2] - L[:, 1] ** 2
        return L[:, 2] / det, -L[:, 1] / det, L[:, 0] / det

    def schur_cached(self, include_A=True):
        i00, i01, i11 = self._linv()
        B0, B1 = self.blk["B"][:, 0, :], self.blk["B"][:, 1, :]
        M = (B0.T @ (i00[:, None] * B0) + B0.T @ (i01[:, None] * B1)
             + B1.T @ (i01[:, None] * B0) + B1.T @ (i11[:, None] * B1))
        return torch.from_numpy((self.blk["A"] if include_A else 0.0) - M)

    @staticmethod
    def spd_inverse_(S):
        S.copy_(torch.linalg.inv(S))
        return S

    def solve_reduce_rhs(self, b, include_bg=True):
        i00, i01, i11 = self._linv()
        Dg, G = self.Dg, self.G
        B0, B1 = self.blk["B"][:, 0, :], self.blk["B"][:, 1, :]
        out = []
        for row in b.numpy().reshape(-1, self.D):
            bm, bi = row[Dg:Dg + G], row[Dg + G:]
            t0, t1 = i00 * bm + i01 * bi, i01 * bm + i11 * bi
            out.append((row[:Dg] if include_bg else 0.0) - (B0.T @ t0 + B1.T @ t1))
        return torch.from_numpy(np.stack(out))

    def solve_finish(self, Sinv, rhs, b):
        i00, i01, i11 = self._linv()
        Dg, G = self.Dg, self.G
        B0, B1 = self.blk["B"][:, 0, :], self.blk["B"][:, 1, :]
        out = []
        for row, r in zip(b.numpy().reshape(-1, self.D), rhs.numpy()):
            xg = Sinv.numpy() @ r
            r0, r1 = row[Dg:Dg + G] - B0 @ xg, row[Dg + G:] - B1 @ xg
            out.append(np.concatenate([xg, i00 * r0 + i01 * r1, i01 * r0 + i11 * r1]))
        return torch.from_numpy(np.stack(out))
