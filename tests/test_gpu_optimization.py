"""GPU: the optimisers of OptimizationUtils over the device objective reach the optimum the CPU
oracle finds (the caller on the other side of the hot path)."""
import numpy as np
import pytest
import scipy.optimize

from helpers import make_case, make_model, make_oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup(vb):
    case = make_case(N=3000, K=4, G=25, Q=8, seed=51, ragged=True, weights=True)
    oracle = make_oracle(case)
    model = make_model(vb, case)
    obj = vb.Objective(model.glmm_par, model)
    x0 = case["free"]
    ref = scipy.optimize.minimize(oracle.kl, x0, jac=oracle.kl_grad,
                                  hessp=lambda x, v: oracle.kl_hvp(x, v), method="trust-ncg",
                                  options=dict(gtol=1e-9, maxiter=200))
    assert np.abs(oracle.kl_grad(ref.x)).max() < 1e-6
    return case, oracle, model, obj, x0, ref


def test_device_newton_reaches_the_oracle_optimum(setup, vb):
    case, oracle, model, obj, x0, ref = setup
    x, res = vb.OptimizationUtils.minimize_objective_newton(obj, x0, gtol=1e-9, maxiter=60)
    assert res.success, res.message
    assert np.abs(oracle.kl_grad(x)).max() < 1e-7
    assert abs(res.fun - ref.fun) <= 1e-9 * abs(ref.fun)
    assert np.max(np.abs(x - ref.x)) < 1e-5
    # par holds the optimum afterwards, as with the reference's optimisers
    np.testing.assert_allclose(model.glmm_par.get_free(), x)
    # tensor in, tensor out, nothing leaves the device
    import torch
    xt, rest = vb.OptimizationUtils.minimize_objective_newton(
        obj, torch.from_numpy(x0).cuda(), gtol=1e-9, maxiter=60)
    assert xt.is_cuda and np.max(np.abs(xt.cpu().numpy() - x)) < 1e-9


def test_trust_ncg_and_repeated_optimisation(setup, vb):
    case, oracle, model, obj, x0, ref = setup
    ou = vb.OptimizationUtils
    x, res = ou.minimize_objective_trust_ncg(obj, x0, precondition=False, maxiter=200, gtol=1e-8,
                                             disp=False)
    assert abs(res.fun - ref.fun) <= 1e-8 * abs(ref.fun)
    out = ou.repeatedly_optimize(
        obj, lambda s: ou.minimize_objective_trust_ncg(obj, s, False, maxiter=50, gtol=1e-9, disp=False),
        x0, max_iter=5, gtol=1e-6)
    new_x, converged = out[0], out[1]
    assert converged
    assert np.max(np.abs(new_x - ref.x)) < 1e-4
    # the dense preconditioner helpers on the Hessian at the optimum
    H = obj.fun_free_hessian(ref.x)
    hess, inv_sqrt, corrected = ou.set_objective_preconditioner(obj, hessian=H, ev_min=1e-6)
    Hd = H.toarray()
    np.testing.assert_allclose(inv_sqrt @ Hd @ inv_sqrt, np.eye(Hd.shape[0]), atol=1e-6)
    xp, resp = ou.minimize_objective_trust_ncg(obj, x0, precondition=True, maxiter=100, gtol=1e-8,
                                               disp=False)
    assert abs(resp.fun - ref.fun) <= 1e-8 * abs(ref.fun)
