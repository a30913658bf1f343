"""GPU parity at the sizes bench.py actually runs: BASELINE.json configs[1] at
its full size (C2: N=1M, K=20, G=10k), a slice of configs[2] wide enough for the three-phase row-pointer
scan of the CSR export (D > 256k), and a K=200 case (configs[3]'s width: k_gram_wide, the wide-model
observation kernel + k_group, blocked multi-CTA SPD inverse).  Same tolerance as everywhere: 1e-9 relative /
1e-12 absolute, sparsity pattern bit-exact."""
import numpy as np
import pytest

from helpers import assert_close, make_case, make_model, make_oracle

pytestmark = pytest.mark.gpu


def _check(vb, case, hvp=True):
    oracle = make_oracle(case)
    model = make_model(vb, case)
    obj = vb.Objective(model.glmm_par, model)
    x = case["free"]
    klo, ge, _ = oracle.kl_blocks(x)
    assert_close(obj.fun_free(x), klo, what="KL")
    assert_close(obj.fun_free_grad(x), ge, scale=np.abs(ge).max() * 1e-3, what="grad")
    H, He = obj.fun_free_hessian(x), oracle.kl_hessian_csr(x)
    assert H.indptr.dtype == np.int32 and H.indices.dtype == np.int32
    np.testing.assert_array_equal(H.indptr, He.indptr)
    np.testing.assert_array_equal(H.indices, He.indices)
    assert_close(H.data, He.data, scale=np.abs(He.data).max() * 1e-6, what="hessian data")
    if hvp:
        v = np.random.default_rng(5).standard_normal(x.size)
        hve = He @ v
        assert_close(obj.fun_free_hvp(x, v), hve, scale=np.abs(hve).max() * 1e-3, what="hvp")
    return oracle, model, obj, H


@pytest.mark.timeout(900)
def test_c2_full_size(vb):
    case = make_case(N=1_000_000, K=20, G=10_000, Q=8, seed=2000)
    _, model, _, H = _check(vb, case)
    assert H.nnz == 4 * 20 * 20 + 14 + 10_000 * (8 * 20 + 14)
    # second evaluation at another point through the same handle (the static-pattern CSR path)
    oracle = make_oracle(case)
    x2 = case["free"] + 0.05 * np.random.default_rng(1).standard_normal(case["free"].size)
    obj = vb.Objective(model.glmm_par, model)
    H2, He2 = obj.fun_free_hessian(x2), oracle.kl_hessian_csr(x2)
    np.testing.assert_array_equal(H2.indices, He2.indices)
    assert_close(H2.data, He2.data, scale=np.abs(He2.data).max() * 1e-6, what="hessian data (2nd point)")


@pytest.mark.timeout(900)
def test_c3_slice_wide_scan(vb):
    # K = 50 (two-warp Gram teams, two column chunks of the fused pass), D = 104 + 2 * 140k > 256k:
    # the three-phase scan of the CSR row pointers; ragged groups of ~11 observations
    case = make_case(N=1_500_000, K=50, G=140_000, Q=8, seed=3000, ragged=True)
    _check(vb, case)


@pytest.mark.timeout(900)
def test_k200(vb):
    # BASELINE configs[3] width: Dg = 404
    case = make_case(N=30_011, K=200, G=60, Q=8, seed=4000, weights=True)
    oracle, model, obj, H = _check(vb, case)
    x = case["free"]
    Hd = oracle.kl_hessian_dense(x)
    assert_close(H.toarray(), Hd, scale=np.abs(Hd).max() * 1e-6, what="dense hessian")


@pytest.mark.timeout(900)
@pytest.mark.parametrize("name,kw", [
    ("c1", dict(N=5000, K=5, G=100, Q=4, seed=1001)),
    ("k20", dict(N=20000, K=20, G=200, Q=8, seed=2001)),
    ("k50", dict(N=6000, K=50, G=60, Q=8, seed=3001)),
    ("k130", dict(N=4000, K=130, G=20, Q=4, seed=23)),
    ("k200", dict(N=12_000, K=200, G=30, Q=8, seed=4001)),
])
def test_lrvb_and_cg_at_the_optimum(vb, name, kw):
    """LRVB is defined at the optimum: device Newton to |grad| < 1e-7, then every
    covariance / solve against the dense inverse of the ORACLE's Hessian at that point -- nothing is
    skipped for lack of positive definiteness."""
    case = make_case(**kw)
    oracle = make_oracle(case)
    model = make_model(vb, case)
    obj = vb.Objective(model.glmm_par, model)
    xo, res = vb.OptimizationUtils.minimize_objective_newton(obj, case["free"], maxiter=60, gtol=1e-7)
    assert res.success, res.message
    assert np.abs(oracle.kl_grad(xo)).max() < 1e-6
    Hd = oracle.kl_hessian_dense(xo)
    assert np.linalg.eigvalsh(Hd).min() > 0.0, "the Hessian at the optimum must be positive definite"
    Hinv = np.linalg.inv(Hd)
    Dg, G, D = model.Dg, model.G, model.D
    lr = vb.LinearResponseCovariances(obj, xo, validate_optimum=True, grad_tol=1e-6)
    assert_close(lr.get_global_covariance(), Hinv[:Dg, :Dg], rtol=1e-8,
                 scale=np.abs(Hinv[:Dg, :Dg]).max() * 1e-3, what="global covariance (Schur)")
    um, ui = np.arange(Dg, Dg + G), np.arange(Dg + G, Dg + 2 * G)
    ref = np.stack([Hinv[um, um], Hinv[um, ui], Hinv[ui, ui]], axis=1)
    assert_close(lr.get_local_covariances(), ref, rtol=1e-8, scale=np.abs(ref).max() * 1e-3,
                 what="local covariances")
    J = model.moment_jacobian(xo)
    refm = J @ Hinv @ J.T
    assert_close(lr.get_lr_covariance(), refm, rtol=1e-8, scale=np.abs(refm).max() * 1e-3,
                 what="moment covariance")
    fac = lr.get_lr_covariance_factors(J)
    assert_close(fac.diagonal(), np.diag(refm), rtol=1e-8, scale=np.abs(refm).max() * 1e-3, what="factored diagonal")
    rows = np.array([0, 1, 2 + model.K // 2, refm.shape[0] - 1])
    assert_close(fac.block(rows, slice(None)), refm[rows], rtol=1e-8, scale=np.abs(refm).max() * 1e-3,
                 what="factored block")
    vv = np.random.default_rng(8).standard_normal(refm.shape[0])
    assert_close(fac.matvec(vv), refm @ vv, rtol=1e-8, scale=np.abs(refm @ vv).max() * 1e-3, what="factored matvec")
    b = np.random.default_rng(6).standard_normal(D)
    xe = Hinv @ b
    solver = vb.ConjugateGradientSolver(obj.fun_free_hvp, xo)
    solver.tol = 1e-11
    for pre in (None, "block_jacobi", "schur"):
        solver.preconditioner = pre
        xs, info = solver.get_hinv_vec(b)
        assert info == 0, (pre, info)
        assert np.max(np.abs(xs - xe)) < 1e-8 * max(1.0, np.abs(xe).max()), pre
    # the CG route to the global covariance (BASELINE configs[3]): Dg solves with device HVPs
    lrc = vb.LinearResponseCovariances(obj, xo, method="cg", cg_tol=1e-11, cg_preconditioner="schur")
    cov_cg = lrc.get_global_covariance()
    assert_close(cov_cg, Hinv[:Dg, :Dg], rtol=1e-8, scale=np.abs(Hinv[:Dg, :Dg]).max() * 1e-3,
                 what="global covariance (CG)")


@pytest.mark.timeout(900)
def test_moment_covariance_at_scale(vb):
    """The moment covariance of a C2-sized model (m = 2 + K + G = 10,022 moments, D = 20,044): never a dense
    (m, D) right-hand side block nor the (m, m) result -- diagonal and rows from the factored form against
    direct arrowhead solves of single Jacobian rows."""
    case = make_case(N=400_000, K=20, G=10_000, Q=8, seed=2100)
    model = make_model(vb, case)
    obj = vb.Objective(model.glmm_par, model)
    xo, res = vb.OptimizationUtils.minimize_objective_newton(obj, case["free"], maxiter=60, gtol=1e-7)
    assert res.success, res.message
    lr = vb.LinearResponseCovariances(obj, xo)
    J = model.moment_jacobian(xo)
    with pytest.raises(ValueError):
        lr.get_lr_covariance_from_jacobians(J, J)          # 10,022^2 entries: refused as a dense matrix
    fac = lr.get_lr_covariance_factors(J)
    diag = fac.diagonal()
    rows = np.array([0, 1, 5, 22, 23, 5000, J.shape[0] - 1])
    blk = fac.block(rows, slice(None))
    for k, r in enumerate(rows):
        hr = lr.hinv(J[r].toarray().reshape(-1)).cpu().numpy()
        ref_row = J @ hr
        assert_close(blk[k], ref_row, rtol=1e-8, scale=np.abs(ref_row).max() * 1e-3, what="row %d" % r)
        assert abs(diag[r] - ref_row[r]) <= 1e-8 * abs(ref_row[r]) + 1e-14
    assert (diag > 0).all()


def test_group_pass_beside_the_wide_gram_is_bitwise_the_serial_order(vb):
    """Under k_gram_wide the per-group pass (k_group) runs on the handle's side stream beside the Gram kernel
    (launch_eval); LRVB_GROUP_OVERLAP=0 restores the serial order.  Same kernels, same inputs: the Hessian,
    gradient and KL must be bitwise identical, over several evaluations at different points (the fork / join
    events are reused)."""
    import os
    case = make_case(N=20_003, K=200, G=40, Q=8, seed=4002)
    m1 = make_model(vb, case)
    os.environ["LRVB_GROUP_OVERLAP"] = "0"
    try:
        m0 = make_model(vb, case)
    finally:
        os.environ.pop("LRVB_GROUP_OVERLAP", None)
    o1, o0 = vb.Objective(m1.glmm_par, m1), vb.Objective(m0.glmm_par, m0)
    for shift in (0.0, 0.01, -0.02):
        x = case["free"] + shift
        H1, H0 = o1.fun_free_hessian(x), o0.fun_free_hessian(x)
        np.testing.assert_array_equal(H1.indices, H0.indices)
        np.testing.assert_array_equal(H1.data, H0.data)
        np.testing.assert_array_equal(o1.fun_free_grad(x), o0.fun_free_grad(x))
        assert o1.fun_free(x) == o0.fun_free(x)
