"""CPU: the oracle restatement (oracle/) against golden vectors produced by the UNMODIFIED
reference code (tests/golden/make_golden.py), plus internal cross-checks of the oracle
(analytic formulas vs torch-fp64 autodiff; slow reference-style CSR emission vs vectorised)."""
import os

import numpy as np
import pytest

from oracle import glmm_oracle as go

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLD, name + ".npz"), allow_pickle=False)


def oracle_from_golden(d):
    pm, pi, bm, bi, ps, pr = d["prior"]
    lb = float(d["lb"])
    return go.GLMMOracle(d["X"], d["y"], d["g"], d["gh_x"], d["gh_w"], weights=d["w"],
                         prior=go.GLMMPrior(pm, pi, bm, bi, ps, pr),
                         bounds=go.GLMMBounds(lb, lb, lb, lb, lb), G=int(d["G"]))


def test_gh_logistic_forward_matches_reference():
    f = load("forward")
    for Q in (4, 8, 20):
        gh_x, gh_w = np.polynomial.hermite.hermgauss(Q)
        zm, zs = f["gh%d_zm" % Q], f["gh%d_zs" % Q]
        np.testing.assert_allclose(go.gh_logistic_term(zm, zs, gh_x, gh_w, False),
                                   f["gh%d_each" % Q], rtol=1e-14, atol=0)
        np.testing.assert_allclose(go.gh_logistic_term(zm, zs, gh_x, gh_w, True),
                                   f["gh%d_all" % Q], rtol=1e-14, atol=0)


def test_ef_forward_matches_reference():
    f = load("forward")
    np.testing.assert_allclose(go.gamma_entropy(f["gam_shape"], f["gam_rate"]), f["gamma_entropy"],
                               rtol=1e-14)
    np.testing.assert_allclose(go.get_e_log_gamma(f["gam_shape"], f["gam_rate"]), f["e_log_gamma"],
                               rtol=1e-14)
    np.testing.assert_allclose(go.univariate_normal_entropy(f["uvn_info"]), f["uvn_entropy"],
                               rtol=1e-14)


def test_transforms_match_reference():
    f = load("forward")
    fv = f["con_free"]
    grid = [(-np.inf, np.inf), (0.0, np.inf), (-1.5, np.inf), (-np.inf, 2.0), (-1.0, 3.0)]
    for i, (lb, ub) in enumerate(grid):
        c = go.constrain(fv, lb, ub)
        np.testing.assert_allclose(c, f["con%d" % i], rtol=1e-15)
        np.testing.assert_allclose(go.unconstrain(c, lb, ub), f["unc%d" % i], rtol=1e-13, atol=1e-15)


@pytest.mark.parametrize("name", ["glmm_small", "glmm_bounds", "glmm_k9"])
def test_glmm_oracle_matches_reference(name):
    d = load(name)
    o = oracle_from_golden(d)
    x = d["free"]
    np.testing.assert_allclose(o.free_to_vector(x), d["vector"], rtol=1e-15)
    np.testing.assert_allclose(o.kl(x), float(d["kl"]), rtol=1e-13)
    g = o.kl_grad(x)
    # complex-step gradient of the reference's forward code is exact to rounding
    np.testing.assert_allclose(g, d["grad"], rtol=1e-10, atol=1e-10 * np.abs(d["grad"]).max())
    for v, hv in zip(d["dirs"], d["hvps"]):
        np.testing.assert_allclose(o.kl_hvp(x, v), hv, rtol=1e-7, atol=1e-7 * np.abs(hv).max())
    if "hessian" in d.files:
        H = o.kl_hessian_dense(x)
        np.testing.assert_allclose(H, d["hessian"], rtol=1e-7, atol=1e-7 * np.abs(H).max())
        # structural zeros of the arrowhead pattern are exact zeros of the reference's Hessian
        assert np.abs(d["hessian"][H == 0]).max() < 1e-6 * np.abs(H).max()


def test_analytic_oracle_vs_torch_autodiff():
    from oracle.glmm_torch import GLMMTorch
    X, y, g = go.make_glmm_data(600, 4, 15, seed=3)
    gh_x, gh_w = np.polynomial.hermite.hermgauss(6)
    w = np.random.default_rng(3).uniform(0.5, 1.5, 600)
    o = go.GLMMOracle(X, y, g, gh_x, gh_w, weights=w, bounds=go.GLMMBounds(0.1, 0.2, 0.3, 0.05, 0.02))
    t = GLMMTorch(o)
    x = go.make_free(o.lay.D, 3, scale=0.3)
    np.testing.assert_allclose(o.kl(x), t.value(x), rtol=1e-13)
    np.testing.assert_allclose(o.kl_grad(x), t.grad(x), rtol=1e-10, atol=1e-11)
    H = o.kl_hessian_dense(x)
    np.testing.assert_allclose(H, t.hessian(x), rtol=1e-9, atol=1e-10 * np.abs(H).max())
    v = np.random.default_rng(4).standard_normal(o.lay.D)
    np.testing.assert_allclose(o.kl_hvp(x, v), t.hvp(x, v), rtol=1e-9, atol=1e-10 * np.abs(H).max())


def test_csr_emission_slow_equals_fast_and_pattern_formula():
    X, y, g = go.make_glmm_data(300, 3, 10, seed=5)
    gh_x, gh_w = np.polynomial.hermite.hermgauss(4)
    o = go.GLMMOracle(X, y, g, gh_x, gh_w)
    x = go.make_free(o.lay.D, 5)
    slow, fast = o.kl_hessian_csr(x, slow=True), o.kl_hessian_csr(x)
    assert slow.indptr.dtype == np.int32 and fast.indices.dtype == np.int32
    np.testing.assert_array_equal(slow.indptr, fast.indptr)
    np.testing.assert_array_equal(slow.indices, fast.indices)
    np.testing.assert_allclose(slow.data, fast.data, rtol=1e-15)
    K, G = 3, 10
    assert fast.nnz == 4 * K * K + 14 + G * (8 * K + 14)
    np.testing.assert_allclose(fast.toarray(), o.kl_hessian_dense(x), rtol=1e-15)
    sub = np.array([[1.0, 0.0], [2.0, 3.0]])
    a = go.get_sparse_sub_matrix(sub, [4, 1], [0, 2], 5, 5)
    b = go.get_sparse_sub_matrix_fast(sub, [4, 1], [0, 2], 5, 5)
    assert (a != b).nnz == 0 and a.nnz == 3


def test_oracle_cg_and_schur_agree_with_dense_solve():
    X, y, g = go.make_glmm_data(2000, 3, 8, seed=6)
    gh_x, gh_w = np.polynomial.hermite.hermgauss(4)
    o = go.GLMMOracle(X, y, g, gh_x, gh_w)
    x = go.make_free(o.lay.D, 6)
    H = o.kl_hessian_dense(x)
    assert np.linalg.eigvalsh(H).min() > 0
    b = np.random.default_rng(7).standard_normal(o.lay.D)
    xs, info = o.cg_solve(x, b, rtol=1e-10)
    assert info == 0
    assert np.max(np.abs(xs - np.linalg.solve(H, b))) < 1e-8   # test_objectives.py:552-554
    cov, _ = o.schur_global_cov(x)
    np.testing.assert_allclose(cov, np.linalg.inv(H)[:o.lay.Dg, :o.lay.Dg], rtol=1e-9, atol=1e-12)


def test_sparse_emission_against_the_references_own_helpers():
    """oracle.get_sparse_sub_matrix / get_sparse_sub_hessian (and the dense -> CSR route the parity tests use)
    against tests/golden/sparse_pattern.npz, which make_golden.py produced with the reference's OWN
    SparseObjectives.get_sparse_sub_hessian / get_sparse_sub_matrix (:591-619): indptr / indices bit-exact,
    data bit-exact (no arithmetic is involved beyond scipy's duplicate summation)."""
    import scipy.sparse
    from oracle import glmm_oracle as go
    d = np.load(os.path.join(GOLD, "sparse_pattern.npz"))
    K, G = int(d["pat_K"]), int(d["pat_G"])
    Dg, D = 4 + 2 * K, 4 + 2 * K + 2 * G
    H = go.get_sparse_sub_hessian(d["pat_A"], np.arange(Dg, dtype=float), D)
    for g in range(G):
        idx = np.concatenate([np.arange(Dg), [Dg + g, Dg + G + g]]).astype(float)
        H = H + go.get_sparse_sub_hessian(d["pat_subs"][g], idx, D)
    H = scipy.sparse.csr_matrix(H)
    H.sort_indices()
    assert H.indptr.dtype == np.int32 and H.indices.dtype == np.int32
    np.testing.assert_array_equal(H.indptr, d["pat_indptr"])
    np.testing.assert_array_equal(H.indices, d["pat_indices"])
    np.testing.assert_array_equal(H.data, d["pat_data"])
    # the same matrix through the arrowhead blocks (what GLMMOracle.kl_hessian_csr and the GPU export emit)
    subs = d["pat_subs"]
    blk = dict(A=d["pat_A"], B=np.stack([subs[:, Dg, :Dg], subs[:, Dg + 1, :Dg]], axis=1),
               L=np.stack([subs[:, Dg, Dg], subs[:, Dg, Dg + 1], subs[:, Dg + 1, Dg + 1]], axis=1))
    Hd = scipy.sparse.csr_matrix(go.GLMMOracle.blocks_to_dense(go.Layout(K, G), blk))
    Hd.sort_indices()
    np.testing.assert_array_equal(Hd.indptr, d["pat_indptr"])
    np.testing.assert_array_equal(Hd.indices, d["pat_indices"])
    np.testing.assert_array_equal(Hd.data, d["pat_data"])
    R = go.get_sparse_sub_matrix(d["pat_M"], np.array([7.0, 2.0, 2.0]), np.array([0.0, 5.0, 1.0, 5.0]), 9, 6)
    R = scipy.sparse.csr_matrix(R); R.sum_duplicates(); R.sort_indices()
    np.testing.assert_array_equal(R.indptr, d["pat_R_indptr"])
    np.testing.assert_array_equal(R.indices, d["pat_R_indices"])
    np.testing.assert_allclose(R.data, d["pat_R_data"], rtol=0, atol=0)
