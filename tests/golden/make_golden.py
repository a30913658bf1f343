"""Generates tests/golden/*.npz by running the UNMODIFIED reference (/root/reference, imported
under oracle/ref_shim.py) on seeded inputs.  Run in the build container only:

    python tests/golden/make_golden.py

What is pinned
  * forward values of Modeling.get_e_logistic_term_guass_hermite, ExponentialFamilies.* and the
    Parameters / ModelParamsDict packing -- straight from the reference's code;
  * the composed GLMM KL evaluated with the reference's own parameter bundles
    (UVNParam, GammaParam, UVNParamVector .e()/.var()/.e_log()/.entropy()), Modeling and
    ExponentialFamilies functions;
  * its gradient by the complex-step method applied to that same reference code (exact to
    rounding: the reference's forward code is complex-analytic once scipy's real-only gammaln
    is continued by loggamma), and Hessian / Hessian-vector products by central differences of
    the complex-step gradient with Richardson extrapolation (~1e-9 relative).
The reference differentiates with autograd (not installable here); these derivative vectors are
the strongest pin of "what autograd would return" available from the reference's own code.
"""
import os
import sys

import numpy as np
import scipy.special

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.ref_shim import import_reference  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def complex_safe_gammaln():
    real_gammaln = scipy.special.gammaln

    def gammaln(z):
        if np.iscomplexobj(z):
            return scipy.special.loggamma(z)
        return real_gammaln(z)
    scipy.special.gammaln = gammaln


def build_par(vb, K, G, lb):
    par = vb.ModelParamsDict("glmm_par")
    par.push_param(vb.UVNParam("mu", min_info=lb))
    par.push_param(vb.GammaParam("tau", min_shape=lb, min_rate=lb))
    par.push_param(vb.UVNParamVector("beta", K, min_info=lb))
    par.push_param(vb.UVNParamVector("u", G, min_info=lb))
    return par


def ref_kl(mods, par, data, prior, free):
    vb, M, ef = mods
    X, y, g, w, gh_x, gh_w = data
    par.set_free(free)
    mu, tau, beta, u = par["mu"], par["tau"], par["beta"], par["u"]
    z_mean = u.e()[g] + np.matmul(X, beta.e())
    z_var = u.var()[g] + np.matmul(X ** 2, beta.var())
    z_sd = np.sqrt(z_var)
    A = M.get_e_logistic_term_guass_hermite(z_mean, z_sd, gh_x, gh_w, aggregate_all=False)
    loglik = np.sum(w * (y * z_mean - A))
    e_tau, e_log_tau = tau.e(), tau.e_log()
    loglik = loglik + np.sum(
        -0.5 * e_tau * ((mu.e() - u.e()) ** 2 + mu.var() + u.var()) + 0.5 * e_log_tau)
    entropy = mu.entropy() + beta.entropy() + u.entropy() + tau.entropy()
    pm, pi, bm, bi, ps, pr = prior
    log_prior = (ef.uvn_prior(pm, pi, mu.e(), mu.var())
                 + np.sum(ef.uvn_prior(bm, bi, beta.e(), beta.var()))
                 + ef.gamma_prior(ps, pr, e_tau, e_log_tau))
    return np.squeeze(-(loglik + entropy + log_prior))


def cs_grad(f, x, h=1e-30):
    g = np.zeros(x.size)
    for i in range(x.size):
        xc = x.astype(complex)
        xc[i] += 1j * h
        g[i] = np.imag(f(xc)) / h
    return g


def fd_hvp(f, x, v, h=2e-3):
    def d(hh):
        return (cs_grad(f, x + hh * v) - cs_grad(f, x - hh * v)) / (2 * hh)
    return (4 * d(h / 2) - d(h)) / 3   # Richardson: O(h^4)


def glmm_case(mods, name, N, K, G, Q, seed, lb, weights, empty, full_hessian):
    vb = mods[0]
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((N, K))
    pool = np.setdiff1d(np.arange(G), empty)
    g = np.sort(rng.choice(pool, size=N))
    beta = rng.normal(0, 0.5, K)
    uu = rng.normal(0.3, 0.5, G)
    y = (rng.random(N) < scipy.special.expit(X @ beta + uu[g])).astype(np.float64)
    w = rng.uniform(0.5, 1.5, N) if weights else np.ones(N)
    gh_x, gh_w = np.polynomial.hermite.hermgauss(Q)
    prior = (0.1, 0.01, -0.2, 0.02, 3.0, 2.5)
    par = build_par(vb, K, G, lb)
    D = par.free_size()
    free = rng.normal(0, 0.3, D)
    data = (X, y, g, w, gh_x, gh_w)

    def f(x):
        return ref_kl(mods, par, data, prior, x)
    kl = float(np.real(f(free)))
    grad = cs_grad(f, free)
    par.set_free(free)
    vec = par.get_vector()
    dirs = rng.standard_normal((3, D))
    hvps = np.stack([fd_hvp(f, free, v) for v in dirs])
    out = dict(X=X, y=y, g=g, w=w, gh_x=gh_x, gh_w=gh_w, prior=np.array(prior), lb=lb, G=G,
               free=free, vector=vec, kl=kl, grad=grad, dirs=dirs, hvps=hvps,
               names=np.array([str(s) for s in par.names()]))
    if full_hessian:
        out["hessian"] = np.stack([fd_hvp(f, free, e) for e in np.eye(D)])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "D =", D, "KL =", kl)


def forward_cases(mods):
    vb, M, ef = mods
    rng = np.random.default_rng(20261018)
    out = {}
    for Q in (4, 8, 20):
        gh_x, gh_w = np.polynomial.hermite.hermgauss(Q)
        zm = rng.normal(0, 3, 257)
        zs = np.exp(rng.normal(0, 0.7, 257))
        zm[:4] = [-40.0, 40.0, 0.0, 300.0]
        out["gh%d_zm" % Q], out["gh%d_zs" % Q] = zm, zs
        out["gh%d_each" % Q] = M.get_e_logistic_term_guass_hermite(zm, zs, gh_x, gh_w, False)
        out["gh%d_all" % Q] = M.get_e_logistic_term_guass_hermite(zm, zs, gh_x, gh_w, True)
    shape, rate = np.exp(rng.normal(0, 0.5, 101)), np.exp(rng.normal(0, 0.5, 101))
    shape[:3] = [0.05, 3.0, 150.0]
    out.update(gam_shape=shape, gam_rate=rate, gamma_entropy=ef.gamma_entropy(shape, rate),
               e_log_gamma=ef.get_e_log_gamma(shape, rate))
    info = np.exp(rng.normal(0, 1, 77))
    out.update(uvn_info=info, uvn_entropy=ef.univariate_normal_entropy(info))
    alpha = 10 * rng.random((5, 63)) + 0.1
    out.update(dir_alpha=alpha, dirichlet_entropy=ef.dirichlet_entropy(alpha),
               e_log_dirichlet=ef.get_e_log_dirichlet(alpha), e_dirichlet=ef.get_e_dirichlet(alpha))
    alpha3 = 10 * rng.random((4, 6, 3)) + 0.1
    out.update(dir_alpha3=alpha3, dirichlet_entropy3=ef.dirichlet_entropy(alpha3))
    tau = 5 * rng.random((41, 2)) + 0.2
    out.update(beta_tau=tau, beta_entropy=ef.beta_entropy(tau))
    p = rng.random((19, 6))
    p /= p.sum(1, keepdims=True)
    out.update(mn_p=p, multinoulli_entropy=ef.multinoulli_entropy(p))
    for k in (2, 3, 5):
        Ms = 7
        v = np.stack([(lambda a: a @ a.T + np.eye(k))(rng.standard_normal((k, k))) for _ in range(Ms)])
        df = k + 1 + 5 * rng.random(Ms)
        out["wis%d_v" % k], out["wis%d_df" % k] = v, df
        out["wis%d_entropy" % k] = np.array([ef.wishart_entropy(df[i], v[i]) for i in range(Ms)])
        out["wis%d_e_log_det" % k] = np.array([ef.e_log_det_wishart(df[i], v[i]) for i in range(Ms)])
        out["wis%d_e_log_inv_diag" % k] = np.stack(
            [ef.e_log_inv_wishart_diag(df[i], v[i]) for i in range(Ms)])
        out["mvn%d_entropy" % k] = np.array([ef.multivariate_normal_entropy(v[i]) for i in range(Ms)])
    out["mv_digamma"] = np.array([ef.multivariate_digamma(3.7, 4), ef.multivariate_digamma(2.2, 2)])
    out["mv_gammaln"] = np.array([ef.multivariate_gammaln(3.7, 4), ef.multivariate_gammaln(2.2, 2)])
    # packing: constrain / unconstrain over a bounds grid, ModelParamsDict layout
    fv = rng.normal(0, 1, 11)
    for i, (lb, ub) in enumerate([(-np.inf, np.inf), (0.0, np.inf), (-1.5, np.inf), (-np.inf, 2.0),
                                  (-1.0, 3.0)]):
        c = vb.Parameters.constrain(fv, lb, ub)
        out["con%d" % i] = c
        out["unc%d" % i] = vb.Parameters.unconstrain(c, lb, ub)
    out["con_free"] = fv
    par = build_par(vb, 3, 4, 0.25)
    x = rng.normal(0, 0.4, par.free_size())
    par.set_free(x)
    out.update(pack_free=x, pack_vector=par.get_vector(), pack_free_back=par.get_free(),
               pack_names=np.array([str(s) for s in par.names()]),
               pack_free_index=np.array([[r.start, r.stop] for r in par.free_indices_dict.values()]),
               pack_e_log_tau=par["tau"].e_log(), pack_tau_entropy=par["tau"].entropy(),
               pack_u_entropy=par["u"].entropy(), pack_u_var=par["u"].var())
    np.savez_compressed(os.path.join(OUT, "forward.npz"), **out)
    print("forward.npz:", len(out), "arrays")


def _reference_packing_modules():
    """The reference's MatrixParameters / SimplexParams call two functions that only exist in
    autograd's numpy wrapper (``np.make_diagonal``) and in scipy < 1.0 (``sp.misc.logsumexp``).  They
    are supplied through module-local namespace proxies -- the reference source is not modified and
    numpy / scipy themselves are not patched."""
    import types
    import scipy
    import scipy.special
    import LinearResponseVariationalBayes.MatrixParameters as mp
    import LinearResponseVariationalBayes.SimplexParams as sx

    def make_diagonal(d, offset=0, axis1=-1, axis2=-2):
        d = np.asarray(d)
        out = np.zeros(d.shape + (d.shape[-1],), dtype=d.dtype)
        i = np.arange(d.shape[-1])
        out[..., i, i] = d
        return out
    np_proxy = types.ModuleType("numpy_with_make_diagonal")
    np_proxy.__dict__.update({k: v for k, v in np.__dict__.items() if not k.startswith("__")})
    np_proxy.make_diagonal = make_diagonal
    mp.np = np_proxy
    sp_proxy = types.ModuleType("scipy_with_misc_logsumexp")
    sp_proxy.__dict__.update({k: v for k, v in scipy.__dict__.items() if not k.startswith("__")})
    sp_proxy.misc = types.SimpleNamespace(logsumexp=scipy.special.logsumexp)
    sx.sp = sp_proxy
    return mp, sx


def _richardson_jac(f, x, h=1e-2, levels=4):
    """Central differences of f at x, Richardson-extrapolated over h, h/2, ...: (len f, len x)."""
    def cd(hh):
        cols = []
        for i in range(x.size):
            e = np.zeros_like(x)
            e[i] = hh
            cols.append((f(x + e) - f(x - e)) / (2 * hh))
        return np.stack(cols, axis=-1)
    T = [cd(h / 2 ** l) for l in range(levels)]
    for m in range(1, levels):
        T = [(4 ** m * T[l + 1] - T[l]) / (4 ** m - 1) for l in range(len(T) - 1)]
    return T[0]


def packing_cases():
    """Outputs of the reference's own packing functions (value maps; simplex Jacobian / Hessian from
    constrain_*_from_moment and from SimplexParam's sparse assembly) and extrapolated differences of
    its pos_def_matrix_free_to_vector (the reference differentiates that with autograd)."""
    mp, sx = _reference_packing_modules()
    rng = np.random.default_rng(77)
    out = {}
    for k in (1, 2, 3, 4):
        v = k * (k + 1) // 2
        for lb in (0.0, 0.3):
            tag = "pd_k%d_lb%s" % (k, "0" if lb == 0.0 else "p3")
            free = rng.normal(scale=0.7, size=(5, v))
            mats = np.array([mp.unpack_posdef_matrix(f, diag_lb=lb) for f in free])
            out[tag + "_free"] = free
            out[tag + "_mat"] = mats
            out[tag + "_pack"] = np.array([mp.pack_posdef_matrix(m, diag_lb=lb) for m in mats])
            out[tag + "_vec"] = np.array([mp.pos_def_matrix_free_to_vector(f, diag_lb=lb) for f in free])
            out[tag + "_sym"] = np.array([mp.unvectorize_symmetric_matrix(mp.vectorize_ld_matrix(m))
                                          for m in mats])
            jac = np.array([_richardson_jac(lambda t: mp.pos_def_matrix_free_to_vector(t, diag_lb=lb), f)
                            for f in free])
            out[tag + "_jac_fd"] = jac
            if k <= 3:
                hess = []
                for f in free:
                    # Hessian of every vector entry = Jacobian of the extrapolated Jacobian
                    Hm = _richardson_jac(
                        lambda t: _richardson_jac(
                            lambda u: mp.pos_def_matrix_free_to_vector(u, diag_lb=lb), t).reshape(-1),
                        f).reshape(v, v, v)
                    hess.append(Hm)
                out[tag + "_hess_fd"] = np.array(hess)
    # PosDefMatrixParamVector round trip through the reference class
    pv = mp.PosDefMatrixParamVector(length=4, matrix_size=3, diag_lb=0.2)
    fr = rng.normal(scale=0.5, size=pv.free_size())
    pv.set_free(fr)
    out["pdvec_free"] = fr
    out["pdvec_val"] = np.array(pv.get())
    out["pdvec_vector"] = pv.get_vector()
    out["pdvec_free_back"] = pv.get_free()
    for d in (2, 3, 5):
        tag = "sx_d%d" % d
        free = rng.normal(scale=1.5, size=(6, d - 1))
        free[0] *= 20.0           # a nearly degenerate simplex
        z = sx.constrain_simplex_matrix(free)
        out[tag + "_free"] = free
        out[tag + "_z"] = z
        out[tag + "_unc"] = sx.unconstrain_simplex_matrix(z)
        out[tag + "_jac"] = np.array([sx.constrain_grad_from_moment(zr) for zr in z])
        out[tag + "_hess"] = np.array([sx.constrain_hess_from_moment(zr) for zr in z])
        par = sx.SimplexParam(shape=(6, d))
        # np.product (removed from numpy 2) is only used by free_size / vector_size
        par.free_size = lambda d=d: 6 * (d - 1)
        par.vector_size = lambda d=d: 6 * d
        out[tag + "_jac_sparse"] = par.free_to_vector_jac(free.flatten()).toarray()
        hs = par.free_to_vector_hess(free.flatten())
        out[tag + "_hess_sparse"] = np.array([h.toarray() for h in hs])
    np.savez_compressed(os.path.join(OUT, "packing.npz"), **out)
    print("packing.npz:", len(out), "arrays")


def sparse_pattern_cases(vb):
    """The reference's own sparse emission (SparseObjectives.py:591-619, get_sparse_sub_hessian /
    get_sparse_sub_matrix + scipy's COO -> CSR summation) applied to seeded dense blocks laid out like the
    arrowhead Hessian: a global block and one (Dg + 2) x (Dg + 2) block per group, with exact zeros, an
    all-zero group and float-typed index vectors (make_index_param returns floats, :581-584).  Pins
    oracle.get_sparse_sub_matrix / get_sparse_sub_hessian and the pattern the oracle's kl_hessian_csr emits."""
    import scipy.sparse
    so = vb.SparseObjectives if hasattr(vb, "SparseObjectives") else __import__(
        "LinearResponseVariationalBayes.SparseObjectives", fromlist=["x"])
    rng = np.random.default_rng(77)
    K, G = 3, 5
    Dg, D = 4 + 2 * K, 4 + 2 * K + 2 * G
    out = {"pat_K": K, "pat_G": G}
    A = rng.standard_normal((Dg, Dg)); A = A + A.T
    A[0, 1] = A[1, 0] = 0.0
    A[:2, 4:] = 0.0; A[4:, :2] = 0.0
    H = so.get_sparse_sub_hessian(A, np.arange(Dg, dtype=float), D)
    subs = []
    for g in range(G):
        sub = np.zeros((Dg + 2, Dg + 2))
        if g != 3:                                   # group 3: no observations, all-zero border
            b = rng.standard_normal((2, Dg))
            b[1, 0] = 0.0                            # structural zero (mu.mean x u.info)
            if g == 1:
                b[0, 5] = 0.0                        # an accidental exact zero
            sub[Dg:, :Dg] = b
            sub[:Dg, Dg:] = b.T
        l = rng.standard_normal(3)
        sub[Dg, Dg], sub[Dg, Dg + 1], sub[Dg + 1, Dg], sub[Dg + 1, Dg + 1] = l[0], l[1], l[1], l[2]
        idx = np.concatenate([np.arange(Dg), [Dg + g, Dg + G + g]]).astype(float)
        H = H + so.get_sparse_sub_hessian(sub, idx, D)
        subs.append(sub)
    H = scipy.sparse.csr_matrix(H)
    H.sort_indices()
    out.update(pat_A=A, pat_subs=np.stack(subs), pat_indptr=H.indptr, pat_indices=H.indices, pat_data=H.data)
    # rectangular sub-matrix with distinct row / column index vectors
    M = rng.standard_normal((3, 4)); M[1, 2] = 0.0
    R = so.get_sparse_sub_matrix(M, np.array([7.0, 2.0, 2.0]), np.array([0.0, 5.0, 1.0, 5.0]), 9, 6)
    R = scipy.sparse.csr_matrix(R); R.sum_duplicates(); R.sort_indices()
    out.update(pat_M=M, pat_R_indptr=R.indptr, pat_R_indices=R.indices, pat_R_data=R.data)
    np.savez_compressed(os.path.join(OUT, "sparse_pattern.npz"), **out)
    print("sparse_pattern.npz: nnz", H.nnz, "of", D * D)


def main():
    vb = import_reference()
    sparse_pattern_cases(vb)
    import LinearResponseVariationalBayes.Modeling as M
    import LinearResponseVariationalBayes.ExponentialFamilies as ef
    mods = (vb, M, ef)
    forward_cases(mods)          # before gammaln is continued: pure reference behaviour
    packing_cases()
    complex_safe_gammaln()
    glmm_case(mods, "glmm_small", N=400, K=3, G=12, Q=4, seed=11, lb=0.0, weights=True,
              empty=[5], full_hessian=True)
    glmm_case(mods, "glmm_bounds", N=1500, K=5, G=30, Q=8, seed=12, lb=0.05, weights=False,
              empty=[], full_hessian=False)
    glmm_case(mods, "glmm_k9", N=900, K=9, G=20, Q=6, seed=13, lb=0.0, weights=True,
              empty=[0, 19], full_hessian=False)


if __name__ == "__main__":
    main()
