"""GPU: the CUDA path against golden vectors produced by the UNMODIFIED reference code
(tests/golden/make_golden.py) -- no oracle in between."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLD, name + ".npz"), allow_pickle=False)


@pytest.mark.parametrize("name", ["glmm_small", "glmm_bounds", "glmm_k9"])
def test_glmm_against_reference_golden(vb, name):
    d = load(name)
    pm, pi, bm, bi, ps, pr = d["prior"]
    lb = float(d["lb"])
    model = vb.LogisticGLMM(d["X"], d["y"], d["g"], gh_x=d["gh_x"], gh_w=d["gh_w"], weights=d["w"],
                            prior=vb.GLMMPrior(pm, pi, bm, bi, ps, pr), num_groups=int(d["G"]),
                            min_info=lb, min_shape=lb, min_rate=lb)
    obj = vb.Objective(model.glmm_par, model)
    x = d["free"]
    assert list(model.glmm_par.names()) == [str(s) for s in d["names"]]
    kl = obj.fun_free(x)
    assert abs(kl - float(d["kl"])) <= 1e-12 + 1e-9 * abs(float(d["kl"]))
    np.testing.assert_allclose(model.glmm_par.get_vector(), d["vector"], rtol=1e-14)
    g = obj.fun_free_grad(x)
    np.testing.assert_allclose(g, d["grad"], rtol=1e-9, atol=1e-9 * np.abs(d["grad"]).max())
    for v, hv in zip(d["dirs"], d["hvps"]):
        # golden HVPs are finite differences of the reference's exact gradient: ~1e-8
        np.testing.assert_allclose(obj.fun_free_hvp(x, v), hv, rtol=1e-7,
                                   atol=1e-7 * np.abs(hv).max())
    if "hessian" in d.files:
        H = obj.fun_free_hessian(x).toarray()
        np.testing.assert_allclose(H, d["hessian"], rtol=1e-7, atol=1e-7 * np.abs(H).max())


def test_forward_functions_against_reference_golden(vb):
    f = load("forward")
    ef, M = vb.ExponentialFamilies, vb.Modeling
    tol = dict(rtol=1e-11, atol=1e-12)
    for Q in (4, 8, 20):
        gh_x, gh_w = np.polynomial.hermite.hermgauss(Q)
        zm, zs = f["gh%d_zm" % Q], f["gh%d_zs" % Q]
        np.testing.assert_allclose(M.get_e_logistic_term_guass_hermite(zm, zs, gh_x, gh_w, False),
                                   f["gh%d_each" % Q], **tol)
        np.testing.assert_allclose(M.get_e_logistic_term_guass_hermite(zm, zs, gh_x, gh_w, True),
                                   f["gh%d_all" % Q], **tol)
    np.testing.assert_allclose(ef.gamma_entropy(f["gam_shape"], f["gam_rate"]), f["gamma_entropy"], **tol)
    np.testing.assert_allclose(ef.get_e_log_gamma(f["gam_shape"], f["gam_rate"]), f["e_log_gamma"], **tol)
    np.testing.assert_allclose(ef.univariate_normal_entropy(f["uvn_info"]), f["uvn_entropy"], **tol)
    np.testing.assert_allclose(ef.dirichlet_entropy(f["dir_alpha"]), f["dirichlet_entropy"], **tol)
    np.testing.assert_allclose(ef.dirichlet_entropy(f["dir_alpha3"]), f["dirichlet_entropy3"], **tol)
    np.testing.assert_allclose(ef.get_e_log_dirichlet(f["dir_alpha"]), f["e_log_dirichlet"], **tol)
    np.testing.assert_allclose(ef.get_e_dirichlet(f["dir_alpha"]), f["e_dirichlet"], **tol)
    np.testing.assert_allclose(ef.beta_entropy(f["beta_tau"]), f["beta_entropy"], **tol)
    np.testing.assert_allclose(ef.multinoulli_entropy(f["mn_p"]), f["multinoulli_entropy"], **tol)
    for k in (2, 3, 5):
        v, df = f["wis%d_v" % k], f["wis%d_df" % k]
        np.testing.assert_allclose(ef.wishart_entropy(df, v), f["wis%d_entropy" % k], **tol)
        np.testing.assert_allclose(ef.e_log_det_wishart(df, v), f["wis%d_e_log_det" % k], **tol)
        np.testing.assert_allclose(ef.e_log_inv_wishart_diag(df, v), f["wis%d_e_log_inv_diag" % k], **tol)
        # the reference's unbatched call: a single (k,k) matrix with scalar df
        np.testing.assert_allclose(ef.wishart_entropy(df[0], v[0]), f["wis%d_entropy" % k][0], **tol)
        np.testing.assert_allclose(ef.multivariate_normal_entropy(v[1]), f["mvn%d_entropy" % k][1],
                                   rtol=1e-10, atol=1e-11)
    np.testing.assert_allclose([ef.multivariate_digamma(3.7, 4), ef.multivariate_digamma(2.2, 2)],
                               f["mv_digamma"], **tol)
    np.testing.assert_allclose([ef.multivariate_gammaln(3.7, 4), ef.multivariate_gammaln(2.2, 2)],
                               f["mv_gammaln"], **tol)


def test_device_csr_export_against_the_references_sparse_emission(vb):
    """The device CSR export (csrc/csr.cu: full export, refill, conditional re-export) on the very blocks of
    tests/golden/sparse_pattern.npz -- emitted there by the reference's own get_sparse_sub_hessian +
    scipy summation: indptr / indices / data bit-exact, exact zeros dropped, empty group included."""
    import torch
    from oracle import glmm_oracle as go
    d = load("sparse_pattern")
    K, G = int(d["pat_K"]), int(d["pat_G"])
    Dg = 4 + 2 * K
    X, y, g = go.make_glmm_data(200, K, G, seed=5)
    model = vb.LogisticGLMM(X, y, g, num_gh_points=4, num_groups=G)
    model._csr_refill_min = 0
    model.evaluate(go.make_free(model.D, 5), 2)
    A, B, L = model.blocks()
    subs = d["pat_subs"]

    def load_blocks(scale=1.0):
        A.copy_(torch.from_numpy(d["pat_A"] * scale))
        B.copy_(torch.from_numpy(np.stack([subs[:, Dg, :Dg], subs[:, Dg + 1, :Dg]], axis=1) * scale))
        L.copy_(torch.from_numpy(np.stack([subs[:, Dg, Dg], subs[:, Dg, Dg + 1], subs[:, Dg + 1, Dg + 1]], axis=1) * scale))

    load_blocks()
    for rep in range(3):            # refill of the model's data pattern -> mismatch -> conditional export; then refills
        H = model.hessian_csr().to_scipy()
        assert H.indptr.dtype == np.int32 and H.indices.dtype == np.int32
        np.testing.assert_array_equal(H.indptr, d["pat_indptr"])
        np.testing.assert_array_equal(H.indices, d["pat_indices"])
        np.testing.assert_array_equal(H.data, d["pat_data"] * (1.0 if rep == 0 else 2.0 ** (rep)))
        load_blocks(2.0 ** (rep + 1))
