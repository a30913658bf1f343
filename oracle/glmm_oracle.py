"""numpy restatement of the logistic-GLMM LRVB hot path.  TEST INFRASTRUCTURE ONLY.

See oracle/__init__.py for who may import this and for the parity status.
All citations are relative to /root/reference/LinearResponseVariationalBayes/.

The model is composed from reference primitives:

  parameters, pushed in this order into one ModelParamsDict
  (ParameterDictionary.py:39-46 -> flat layout):
      mu   = UVNParam        (NormalParams.py:26-30)  mean, info>=lb
      tau  = GammaParam      (GammaParams.py:4-8)     shape>=lb, rate>=lb
      beta = UVNParamVector  (NormalParams.py:51-56)  mean[K], info[K]
      u    = UVNParamVector                            mean[G], info[G]

  z_mean_n = E[u]_{g[n]} + x_n . E[beta]               (.e()   NormalParams.py:58-59)
  z_var_n  = Var[u]_{g[n]} + x_n^2 . Var[beta]         (.var() NormalParams.py:62-63)
  l_n      = y_n z_mean_n - GH(z_mean_n, sqrt(z_var_n)) (Modeling.py:35-52)
  loglik   = sum_n w_n l_n
           + sum_g [ -1/2 E[tau]((E[mu]-E[u_g])^2 + Var[mu] + Var[u_g]) + 1/2 E[log tau] ]
                                                       (GammaParams.py:9-13)
  entropy  = uvn_entropy(mu) + uvn_entropy(beta) + uvn_entropy(u) + gamma_entropy(tau)
                                                       (ExponentialFamilies.py:23-25, 33-35)
  prior    = uvn_prior(mu) + sum_k uvn_prior(beta_k) + gamma_prior(tau)
                                                       (ExponentialFamilies.py:191-195)
  KL(free) = -(loglik + entropy + prior)
"""
import math
from dataclasses import dataclass

import numpy as np
import scipy.sparse
import scipy.sparse.linalg
import scipy.special


# --------------------------------------------------------------------------
# Forward primitives (restated)
# --------------------------------------------------------------------------

def gh_logistic_term(z_mean, z_sd, gh_x, gh_w, aggregate_all=True):
    """Modeling.py:35-52 get_e_logistic_term_guass_hermite, restated verbatim."""
    z_mean = np.asarray(z_mean, dtype=np.float64)
    z_sd = np.asarray(z_sd, dtype=np.float64)
    assert z_mean.shape == z_sd.shape  # Modeling.py:38
    z_vals = np.sqrt(2) * z_sd[..., None] * gh_x + z_mean[..., None]
    logit_term = gh_w * np.log1p(np.exp(z_vals)) / np.sqrt(np.pi)
    if aggregate_all:
        return np.sum(logit_term)
    return np.sum(logit_term, axis=-1)


def univariate_normal_entropy(info_obs):
    """ExponentialFamilies.py:23-25."""
    return 0.5 * np.sum(-1 * np.log(info_obs) + 1 + np.log(2 * math.pi))


def gamma_entropy(shape, rate):
    """ExponentialFamilies.py:33-35."""
    return np.sum(shape - np.log(rate) + scipy.special.gammaln(shape)
                  + (1 - shape) * scipy.special.digamma(shape))


def get_e_log_gamma(shape, rate):
    """ExponentialFamilies.py:111-112."""
    return scipy.special.digamma(shape) - np.log(rate)


def uvn_prior(prior_mean, prior_info, e_obs, var_obs):
    """ExponentialFamilies.py:191-192."""
    return -0.5 * (prior_info * ((e_obs - prior_mean) ** 2 + var_obs))


def gamma_prior(prior_shape, prior_rate, e_obs, e_log_obs):
    """ExponentialFamilies.py:194-195."""
    return (prior_shape - 1) * e_log_obs - prior_rate * e_obs


def constrain(free_vec, lb, ub=float("inf")):
    """Parameters.py:47-61."""
    if ub <= lb:
        raise ValueError("Upper bound must be greater than lower bound")
    if ub == float("inf"):
        if lb == -float("inf"):
            return np.array(free_vec, dtype=np.float64, copy=True)
        return np.exp(free_vec) + lb
    if lb == -float("inf"):
        return ub - np.exp(-1 * free_vec)
    exp_vec = np.exp(free_vec)
    return (ub - lb) * exp_vec / (1 + exp_vec) + lb


def unconstrain(vec, lb, ub=float("inf")):
    """Parameters.py:31-44."""
    if ub <= lb:
        raise ValueError("Upper bound must be greater than lower bound")
    if ub == float("inf"):
        if lb == -float("inf"):
            return np.array(vec, dtype=np.float64, copy=True)
        return np.log(vec - lb)
    if lb == -float("inf"):
        return -1 * np.log(ub - vec)
    return np.log(vec - lb) - np.log(ub - vec)


# --------------------------------------------------------------------------
# Sparse emission (restated)
# --------------------------------------------------------------------------

def get_sparse_sub_matrix(sub_matrix, row_indices, col_indices, row_dim, col_dim):
    """SparseObjectives.py:604-619, restated: python double loop, exact zeros dropped,
    COO triplets -> csr_matrix (duplicates summed, columns sorted, int32)."""
    vals, rows, cols = [], [], []
    for row in range(sub_matrix.shape[0]):
        for col in range(sub_matrix.shape[1]):
            if sub_matrix[row, col] != 0:
                vals.append(sub_matrix[row, col])
                rows.append(int(row_indices[row]))
                cols.append(int(col_indices[col]))
    return scipy.sparse.csr_matrix((vals, (rows, cols)), (row_dim, col_dim))


def get_sparse_sub_matrix_fast(sub_matrix, row_indices, col_indices, row_dim, col_dim):
    """Vectorised equivalent of get_sparse_sub_matrix (same triplet order)."""
    sub_matrix = np.asarray(sub_matrix)
    r, c = np.nonzero(sub_matrix)  # C-order == the reference's loop order
    rows = np.asarray(row_indices)[r].astype(np.int64)
    cols = np.asarray(col_indices)[c].astype(np.int64)
    return scipy.sparse.csr_matrix((sub_matrix[r, c], (rows, cols)), (row_dim, col_dim))


def get_sparse_sub_hessian(sub_hessian, full_indices, full_hess_dim):
    """SparseObjectives.py:591-597."""
    return get_sparse_sub_matrix(sub_hessian, full_indices, full_indices,
                                 full_hess_dim, full_hess_dim)


# --------------------------------------------------------------------------
# The GLMM
# --------------------------------------------------------------------------

@dataclass
class GLMMPrior:
    mu_mean: float = 0.0
    mu_info: float = 0.01
    beta_mean: float = 0.0
    beta_info: float = 0.01
    tau_shape: float = 3.0
    tau_rate: float = 3.0


@dataclass
class GLMMBounds:
    """Lower bounds of the constrained sub-parameters (min_info / min_shape / min_rate)."""
    mu_info: float = 0.0
    tau_shape: float = 0.0
    tau_rate: float = 0.0
    beta_info: float = 0.0
    u_info: float = 0.0


class Layout:
    """Flat free/vector layout (ParameterDictionary.py:39-46, 64-65)."""

    def __init__(self, K, G):
        self.K, self.G = int(K), int(G)
        self.mu_mean, self.mu_info, self.tau_shape, self.tau_rate = 0, 1, 2, 3
        self.beta_mean = 4
        self.beta_info = 4 + K
        self.Dg = 4 + 2 * K
        self.u_mean = self.Dg
        self.u_info = self.Dg + G
        self.D = self.Dg + 2 * G


def _sigmoid(t):
    return scipy.special.expit(t)


class GLMMOracle:
    """Analytic fp64 oracle: KL, gradient, dense/sparse Hessian, HVP, CG, LRVB covariance."""

    def __init__(self, X, y, g, gh_x, gh_w, weights=None, prior=None, bounds=None, G=None):
        self.X = np.ascontiguousarray(X, dtype=np.float64)
        self.y = np.asarray(y, dtype=np.float64)
        self.g = np.asarray(g, dtype=np.int64)
        self.N, self.K = self.X.shape
        self.G = int(G) if G is not None else int(self.g.max()) + 1
        self.w = None if weights is None else np.asarray(weights, dtype=np.float64)
        self.gh_x = np.asarray(gh_x, dtype=np.float64)
        self.gh_w = np.asarray(gh_w, dtype=np.float64)
        self.prior = prior or GLMMPrior()
        self.bounds = bounds or GLMMBounds()
        self.lay = Layout(self.K, self.G)

    # ---- packing ---------------------------------------------------------
    def lower_bounds(self):
        """Per-free-coordinate lower bound (-inf for unconstrained means)."""
        lay, b = self.lay, self.bounds
        lb = np.full(lay.D, -np.inf)
        lb[lay.mu_info] = b.mu_info
        lb[lay.tau_shape] = b.tau_shape
        lb[lay.tau_rate] = b.tau_rate
        lb[lay.beta_info:lay.beta_info + lay.K] = b.beta_info
        lb[lay.u_info:lay.u_info + lay.G] = b.u_info
        return lb

    def free_to_vector(self, free):
        free = np.asarray(free, dtype=np.float64)
        if free.size != self.lay.D:
            raise ValueError("Wrong size for parameter glmm_par.  Expected {}, got {}".format(
                self.lay.D, free.size))
        lb = self.lower_bounds()
        con = np.isfinite(lb)
        vec = free.copy()
        vec[con] = np.exp(free[con]) + lb[con]  # Parameters.py:55
        return vec

    def vector_to_free(self, vec):
        lb = self.lower_bounds()
        con = np.isfinite(lb)
        free = np.array(vec, dtype=np.float64, copy=True)
        free[con] = np.log(vec[con] - lb[con])  # Parameters.py:40
        return free

    def unpack(self, vec):
        lay = self.lay
        K, G = lay.K, lay.G
        return dict(
            mu_m=vec[0], mu_i=vec[1], a=vec[2], b=vec[3],
            beta_m=vec[4:4 + K], beta_i=vec[4 + K:4 + 2 * K],
            u_m=vec[lay.u_mean:lay.u_mean + G], u_i=vec[lay.u_info:lay.u_info + G])

    # ---- per-observation quadrature ---------------------------------------
    def obs_terms(self, p, order=2):
        """z, l_n and its derivatives in (z_mean, z_var) coordinates."""
        S = self.X * self.X
        z_m = p["u_m"][self.g] + self.X @ p["beta_m"]
        z_v = (1.0 / p["u_i"])[self.g] + S @ (1.0 / p["beta_i"])
        z_s = np.sqrt(z_v)
        what = self.gh_w / np.sqrt(np.pi)
        c = np.sqrt(2) * self.gh_x
        t = z_m[:, None] + z_s[:, None] * c[None, :]
        A = np.sum(what * np.log1p(np.exp(t)), axis=1)  # Modeling.py:48
        w = np.ones(self.N) if self.w is None else self.w
        out = dict(z_m=z_m, z_v=z_v, S=S, ell=w * (self.y * z_m - A))
        if order >= 1:
            sg = _sigmoid(t)
            A_m = np.sum(what * sg, axis=1)
            A_s = np.sum(what * sg * c, axis=1)
            out["l_m"] = w * (self.y - A_m)
            out["l_v"] = w * (-A_s / (2 * z_s))
        if order >= 2:
            sp = sg * (1 - sg)
            A_mm = np.sum(what * sp, axis=1)
            A_ms = np.sum(what * sp * c, axis=1)
            A_ss = np.sum(what * sp * c * c, axis=1)
            out["l_mm"] = w * (-A_mm)
            out["l_mv"] = w * (-A_ms / (2 * z_s))
            out["l_vv"] = w * (-(A_ss / (4 * z_v) - A_s / (4 * z_s ** 3)))
        return out

    # ---- value -----------------------------------------------------------
    def kl_vector(self, vec):
        p = self.unpack(np.asarray(vec, dtype=np.float64))
        pr = self.prior
        o = self.obs_terms(p, order=0)
        e_tau = p["a"] / p["b"]
        e_log_tau = get_e_log_gamma(p["a"], p["b"])
        re = np.sum(-0.5 * e_tau * ((p["mu_m"] - p["u_m"]) ** 2 + 1 / p["mu_i"] + 1 / p["u_i"])
                    + 0.5 * e_log_tau)
        loglik = np.sum(o["ell"]) + re
        entropy = (univariate_normal_entropy(p["mu_i"]) + univariate_normal_entropy(p["beta_i"])
                   + univariate_normal_entropy(p["u_i"]) + gamma_entropy(p["a"], p["b"]))
        prior = (uvn_prior(pr.mu_mean, pr.mu_info, p["mu_m"], 1 / p["mu_i"])
                 + np.sum(uvn_prior(pr.beta_mean, pr.beta_info, p["beta_m"], 1 / p["beta_i"]))
                 + gamma_prior(pr.tau_shape, pr.tau_rate, e_tau, e_log_tau))
        return float(-(loglik + entropy + prior))

    def kl(self, free):
        return self.kl_vector(self.free_to_vector(free))

    # ---- derivatives in vector coordinates --------------------------------
    def _group_sum(self, v):
        return np.bincount(self.g, weights=v, minlength=self.G)

    def vector_derivs(self, vec, hessian=True):
        """Returns (kl, grad_vec, blocks) with blocks = dict(A (Dg x Dg), B (G,2,Dg), L (G,3))
        for the Hessian of KL in *vector* coordinates: A global block, B[g,0,:]/B[g,1,:] the
        rows (u_mean_g, :)/(u_info_g, :) of the border, L[g] = (mm, mi, ii) of the local 2x2."""
        lay = self.lay
        K, G, Dg = lay.K, lay.G, lay.Dg
        vec = np.asarray(vec, dtype=np.float64)
        p = self.unpack(vec)
        pr = self.prior
        X = self.X
        o = self.obs_terms(p, order=2 if hessian else 1)
        S = o["S"]
        mu_m, mu_i, a, b = p["mu_m"], p["mu_i"], p["a"], p["b"]
        bi, ui, um = p["beta_i"], p["u_i"], p["u_m"]
        psi1 = scipy.special.polygamma(1, a)
        psi2 = scipy.special.polygamma(2, a)
        E = a / b
        dm = mu_m - um                      # (G,)
        Sg = dm ** 2 + 1 / mu_i + 1 / ui    # (G,)
        Ssum, dsum = np.sum(Sg), np.sum(dm)

        # ---- gradient of (loglik + entropy + prior) =: F ; KL = -F
        gF = np.zeros(lay.D)
        # data term, through phi = (beta_m, v=1/beta_i, u_m, r=1/u_i)
        g_bm = X.T @ o["l_m"]
        g_v = S.T @ o["l_v"]
        g_um = self._group_sum(o["l_m"])
        g_r = self._group_sum(o["l_v"])
        gF[lay.beta_mean:lay.beta_mean + K] += g_bm
        gF[lay.beta_info:lay.beta_info + K] += g_v * (-1 / bi ** 2)
        gF[lay.u_mean:lay.u_mean + G] += g_um
        gF[lay.u_info:lay.u_info + G] += g_r * (-1 / ui ** 2)
        # random-effect term
        gF[lay.mu_mean] += -E * dsum
        gF[lay.mu_info] += 0.5 * E * G / mu_i ** 2
        gF[lay.tau_shape] += -0.5 * Ssum / b + 0.5 * G * psi1
        gF[lay.tau_rate] += 0.5 * a * Ssum / b ** 2 - 0.5 * G / b
        gF[lay.u_mean:lay.u_mean + G] += E * dm
        gF[lay.u_info:lay.u_info + G] += 0.5 * E / ui ** 2
        # entropies
        gF[lay.mu_info] += -0.5 / mu_i
        gF[lay.beta_info:lay.beta_info + K] += -0.5 / bi
        gF[lay.u_info:lay.u_info + G] += -0.5 / ui
        gF[lay.tau_shape] += 1 + (1 - a) * psi1
        gF[lay.tau_rate] += -1 / b
        # priors
        gF[lay.mu_mean] += -pr.mu_info * (mu_m - pr.mu_mean)
        gF[lay.mu_info] += 0.5 * pr.mu_info / mu_i ** 2
        gF[lay.beta_mean:lay.beta_mean + K] += -pr.beta_info * (p["beta_m"] - pr.beta_mean)
        gF[lay.beta_info:lay.beta_info + K] += 0.5 * pr.beta_info / bi ** 2
        gF[lay.tau_shape] += (pr.tau_shape - 1) * psi1 - pr.tau_rate / b
        gF[lay.tau_rate] += -(pr.tau_shape - 1) / b + pr.tau_rate * a / b ** 2

        kl = self.kl_vector(vec)
        if not hessian:
            return kl, -gF, None

        # ---- Hessian of F
        A = np.zeros((Dg, Dg))
        B = np.zeros((G, 2, Dg))
        L = np.zeros((G, 3))
        bm0, bi0 = lay.beta_mean, lay.beta_info
        dv = -1 / bi ** 2          # dv/d(beta_info)
        dr = -1 / ui ** 2          # dr/d(u_info)
        la, lb_, lc = o["l_mm"], o["l_mv"], o["l_vv"]
        M1 = X.T @ (la[:, None] * X)
        M2 = X.T @ (lb_[:, None] * S)
        M3 = S.T @ (lc[:, None] * S)
        A[bm0:bm0 + K, bm0:bm0 + K] += M1
        A[bm0:bm0 + K, bi0:bi0 + K] += M2 * dv[None, :]
        A[bi0:bi0 + K, bm0:bm0 + K] += (M2 * dv[None, :]).T
        A[bi0:bi0 + K, bi0:bi0 + K] += M3 * dv[:, None] * dv[None, :] + np.diag(g_v * 2 / bi ** 3)
        # borders (data): per-group sums
        def gsum_rows(wv, M):
            out = np.zeros((G, K))
            if self.N and np.all(self.g[1:] >= self.g[:-1]):
                # group-sorted: segmented sums (same result as np.add.at, much faster)
                starts = np.flatnonzero(np.r_[True, self.g[1:] != self.g[:-1]])
                out[self.g[starts]] = np.add.reduceat(wv[:, None] * M, starts, axis=0)
            else:
                np.add.at(out, self.g, wv[:, None] * M)
            return out
        B[:, 0, bm0:bm0 + K] += gsum_rows(la, X)                       # (u_m, beta_m)
        B[:, 0, bi0:bi0 + K] += gsum_rows(lb_, S) * dv[None, :]        # (u_m, beta_i)
        B[:, 1, bm0:bm0 + K] += gsum_rows(lb_, X) * dr[:, None]        # (u_i, beta_m)
        B[:, 1, bi0:bi0 + K] += gsum_rows(lc, S) * dv[None, :] * dr[:, None]
        L[:, 0] += self._group_sum(la)
        L[:, 1] += self._group_sum(lb_) * dr
        L[:, 2] += self._group_sum(lc) * dr * dr + g_r * 2 / ui ** 3
        # random-effect term
        A[0, 0] += -E * G
        A[0, 2] += -dsum / b
        A[2, 0] += -dsum / b
        A[0, 3] += a * dsum / b ** 2
        A[3, 0] += a * dsum / b ** 2
        A[1, 1] += -E * G / mu_i ** 3
        A[1, 2] += 0.5 * G / (b * mu_i ** 2)
        A[2, 1] += 0.5 * G / (b * mu_i ** 2)
        A[1, 3] += -0.5 * a * G / (b ** 2 * mu_i ** 2)
        A[3, 1] += -0.5 * a * G / (b ** 2 * mu_i ** 2)
        A[2, 2] += 0.5 * G * psi2
        A[2, 3] += 0.5 * Ssum / b ** 2
        A[3, 2] += 0.5 * Ssum / b ** 2
        A[3, 3] += -a * Ssum / b ** 3 + 0.5 * G / b ** 2
        B[:, 0, 0] += E
        B[:, 0, 2] += dm / b
        B[:, 0, 3] += -a * dm / b ** 2
        B[:, 1, 2] += 0.5 / (b * ui ** 2)
        B[:, 1, 3] += -0.5 * a / (b ** 2 * ui ** 2)
        L[:, 0] += -E
        L[:, 2] += -E / ui ** 3
        # entropies
        A[1, 1] += 0.5 / mu_i ** 2
        A[bi0:bi0 + K, bi0:bi0 + K] += np.diag(0.5 / bi ** 2)
        L[:, 2] += 0.5 / ui ** 2
        A[2, 2] += -psi1 + (1 - a) * psi2
        A[3, 3] += 1 / b ** 2
        # priors
        A[0, 0] += -pr.mu_info
        A[1, 1] += -pr.mu_info / mu_i ** 3
        A[bm0:bm0 + K, bm0:bm0 + K] += np.diag(np.full(K, -pr.beta_info))
        A[bi0:bi0 + K, bi0:bi0 + K] += np.diag(-pr.beta_info / bi ** 3)
        A[2, 2] += (pr.tau_shape - 1) * psi2
        A[2, 3] += pr.tau_rate / b ** 2
        A[3, 2] += pr.tau_rate / b ** 2
        A[3, 3] += (pr.tau_shape - 1) / b ** 2 - 2 * pr.tau_rate * a / b ** 3
        return kl, -gF, dict(A=-A, B=-B, L=-L)

    # ---- free coordinates ---------------------------------------------------
    def _free_jac(self, vec):
        """d vec / d free (diagonal) and d2 vec / d free2 (Parameters.py:53-55)."""
        lb = self.lower_bounds()
        con = np.isfinite(lb)
        j1 = np.ones(self.lay.D)
        j2 = np.zeros(self.lay.D)
        j1[con] = vec[con] - lb[con]
        j2[con] = vec[con] - lb[con]
        return j1, j2

    def kl_grad(self, free):
        vec = self.free_to_vector(free)
        _, gv, _ = self.vector_derivs(vec, hessian=False)
        j1, _ = self._free_jac(vec)
        return gv * j1

    def kl_blocks(self, free):
        """(kl, grad_free, blocks_free): convert_vector_to_free_hessian (Parameters.py:397-424)
        specialised to diagonal transforms: H_free = J H_vec J + diag(g_vec * d2vec/dfree2)."""
        lay = self.lay
        Dg, G = lay.Dg, lay.G
        vec = self.free_to_vector(free)
        kl, gv, blk = self.vector_derivs(vec, hessian=True)
        j1, j2 = self._free_jac(vec)
        jg = j1[:Dg]
        jm = j1[lay.u_mean:lay.u_mean + G]
        ji = j1[lay.u_info:lay.u_info + G]
        A = blk["A"] * jg[:, None] * jg[None, :] + np.diag(gv[:Dg] * j2[:Dg])
        B = blk["B"] * jg[None, None, :]
        B[:, 0, :] *= jm[:, None]
        B[:, 1, :] *= ji[:, None]
        L = blk["L"].copy()
        L[:, 0] = L[:, 0] * jm * jm + gv[lay.u_mean:lay.u_mean + G] * j2[lay.u_mean:lay.u_mean + G]
        L[:, 1] = L[:, 1] * jm * ji
        L[:, 2] = L[:, 2] * ji * ji + gv[lay.u_info:lay.u_info + G] * j2[lay.u_info:lay.u_info + G]
        return kl, gv * j1, dict(A=A, B=B, L=L)

    @staticmethod
    def blocks_to_dense(lay, blk):
        D, Dg, G = lay.D, lay.Dg, lay.G
        H = np.zeros((D, D))
        H[:Dg, :Dg] = blk["A"]
        um = np.arange(lay.u_mean, lay.u_mean + G)
        ui = np.arange(lay.u_info, lay.u_info + G)
        H[um, :Dg] = blk["B"][:, 0, :]
        H[:Dg, um] = blk["B"][:, 0, :].T
        H[ui, :Dg] = blk["B"][:, 1, :]
        H[:Dg, ui] = blk["B"][:, 1, :].T
        H[um, um] = blk["L"][:, 0]
        H[um, ui] = blk["L"][:, 1]
        H[ui, um] = blk["L"][:, 1]
        H[ui, ui] = blk["L"][:, 2]
        return H

    def kl_hessian_dense(self, free):
        _, _, blk = self.kl_blocks(free)
        return self.blocks_to_dense(self.lay, blk)

    def kl_hessian_csr(self, free, slow=False):
        """Sparse Hessian by the reference recipe: the dense global block and
        one dense block per group are emitted with get_sparse_sub_hessian
        (SparseObjectives.py:591-619) and the csr pieces are summed.  Each structural entry
        is emitted exactly once (global block from the global piece; border + local entries
        from the owning group's piece)."""
        lay = self.lay
        D, Dg, G = lay.D, lay.Dg, lay.G
        _, _, blk = self.kl_blocks(free)
        if slow:
            H = get_sparse_sub_hessian(blk["A"], np.arange(Dg), D)
            for g in range(G):
                idx = np.concatenate([np.arange(Dg), [lay.u_mean + g, lay.u_info + g]])
                sub = np.zeros((Dg + 2, Dg + 2))
                sub[Dg, :Dg] = blk["B"][g, 0]
                sub[:Dg, Dg] = blk["B"][g, 0]
                sub[Dg + 1, :Dg] = blk["B"][g, 1]
                sub[:Dg, Dg + 1] = blk["B"][g, 1]
                sub[Dg, Dg] = blk["L"][g, 0]
                sub[Dg, Dg + 1] = sub[Dg + 1, Dg] = blk["L"][g, 1]
                sub[Dg + 1, Dg + 1] = blk["L"][g, 2]
                H = H + get_sparse_sub_hessian(sub, idx, D)
            H = scipy.sparse.csr_matrix(H)
            H.sort_indices()
            return H
        # vectorised: identical triplets, one csr construction
        rows, cols, vals = [], [], []
        r, c = np.nonzero(blk["A"])
        rows.append(r); cols.append(c); vals.append(blk["A"][r, c])
        for s, off in ((0, lay.u_mean), (1, lay.u_info)):
            gg, cc = np.nonzero(blk["B"][:, s, :])
            v = blk["B"][gg, s, cc]
            rows += [gg + off, cc]; cols += [cc, gg + off]; vals += [v, v]
        gi = np.arange(G)
        for (col, ro, co) in ((0, lay.u_mean, lay.u_mean), (1, lay.u_mean, lay.u_info),
                              (1, lay.u_info, lay.u_mean), (2, lay.u_info, lay.u_info)):
            nz = blk["L"][:, col] != 0
            rows.append(gi[nz] + ro); cols.append(gi[nz] + co); vals.append(blk["L"][nz, col])
        H = scipy.sparse.csr_matrix(
            (np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), (D, D))
        H.sort_indices()
        return H

    def hvp_blocks(self, blk, v):
        """Arrowhead H v from blocks."""
        lay = self.lay
        Dg, G = lay.Dg, lay.G
        vg, vm, vi = v[:Dg], v[lay.u_mean:lay.u_mean + G], v[lay.u_info:lay.u_info + G]
        out = np.empty(lay.D)
        B0, B1 = blk["B"][:, 0, :], blk["B"][:, 1, :]
        out[:Dg] = blk["A"] @ vg + B0.T @ vm + B1.T @ vi
        out[lay.u_mean:lay.u_mean + G] = B0 @ vg + blk["L"][:, 0] * vm + blk["L"][:, 1] * vi
        out[lay.u_info:lay.u_info + G] = B1 @ vg + blk["L"][:, 1] * vm + blk["L"][:, 2] * vi
        return out

    def kl_hvp(self, free, v):
        _, _, blk = self.kl_blocks(free)
        return self.hvp_blocks(blk, np.asarray(v, dtype=np.float64))

    # ---- solves ---------------------------------------------------------------
    def cg_solve(self, free, b, x0=None, M=None, rtol=1e-8, maxiter=None):
        """ConjugateGradient.py:63-85 with the scipy>=1.14 spelling of tol (rtol, atol=0)."""
        _, _, blk = self.kl_blocks(free)
        D = self.lay.D
        op = scipy.sparse.linalg.LinearOperator((D, D), matvec=lambda v: self.hvp_blocks(blk, v))
        return scipy.sparse.linalg.cg(op, b, x0=x0, rtol=rtol, atol=0.0, M=M, maxiter=maxiter)

    def schur_global_cov(self, free):
        """(H^-1)_gg = (A - sum_g B_g^T L_g^-1 B_g)^-1  (SURVEY A.4)."""
        _, _, blk = self.kl_blocks(free)
        A, B, L = blk["A"], blk["B"], blk["L"]
        det = L[:, 0] * L[:, 2] - L[:, 1] ** 2
        i00, i01, i11 = L[:, 2] / det, -L[:, 1] / det, L[:, 0] / det
        B0, B1 = B[:, 0, :], B[:, 1, :]
        S = A - (B0.T @ (i00[:, None] * B0) + B0.T @ (i01[:, None] * B1)
                 + B1.T @ (i01[:, None] * B0) + B1.T @ (i11[:, None] * B1))
        return np.linalg.inv(S), S


# --------------------------------------------------------------------------
# Synthetic data
# --------------------------------------------------------------------------

def make_glmm_data(N, K, G, seed, intercept=False):
    """Seeded synthetic logistic-GLMM data, group-sorted."""
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((N, K))
    if intercept:
        X[:, 0] = 1.0
    base, rem = divmod(N, G)
    counts = np.full(G, base, dtype=np.int64)
    counts[:rem] += 1
    g = np.repeat(np.arange(G, dtype=np.int64), counts)
    beta = rng.normal(0.0, 0.5, K)
    u = rng.normal(0.3, 0.5, G)
    p = scipy.special.expit(X @ beta + u[g])
    y = (rng.random(N) < p).astype(np.float64)
    return X, y, g


def make_free(D, seed, scale=0.1):
    return np.random.default_rng(seed + 7).normal(0.0, scale, D)
