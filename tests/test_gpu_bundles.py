"""GPU: the remaining parameter bundles of the reference -- MVNParam (NormalParams.py:6-23),
UVNMomentParamArray (:108-147), MVNArray (:150-162), DirichletParamArray (DirichletParams.py:11-26),
WishartParam (WishartParams.py:6-35) -- run through the parameter protocol of
test_variational_bayes.py:74-106 and checked against closed forms / scipy.stats as the reference's own
tests do (test_exponential_families.py:25-31, 40-58)."""
import numpy as np
import pytest
import scipy.special
import scipy.stats

pytestmark = pytest.mark.gpu


def protocol(par):
    par.names()
    par.dictval()
    free = par.get_free()
    par.set_free(free)
    vec = par.get_vector()
    par.set_vector(vec)
    assert free.ndim == 1 and vec.ndim == 1
    assert free.size == par.free_size() and vec.size == par.vector_size()
    np.testing.assert_allclose(par.get_free(), free, rtol=1e-12, atol=1e-12)
    jac = par.free_to_vector_jac(free)
    hess = par.free_to_vector_hess(free)
    assert jac.shape == (par.vector_size(), par.free_size()) and len(hess) == par.vector_size()
    str(par)


def test_mvn_param(vb):
    rng = np.random.default_rng(3)
    par = vb.MVNParam("x", dim=3, min_info=0.1)
    a = rng.standard_normal((3, 3))
    info = a @ a.T + 0.5 * np.eye(3)
    par["mean"].set(rng.standard_normal(3))
    par["info"].set(info)
    protocol(par)
    cov = np.linalg.inv(info)
    np.testing.assert_allclose(par.cov(), cov, rtol=1e-10)
    np.testing.assert_allclose(par.e_outer(), np.outer(par.e(), par.e()) + cov, rtol=1e-10)
    ref = scipy.stats.multivariate_normal(mean=par.e(), cov=cov).entropy()
    assert abs(float(par.entropy()) - ref) < 1e-10
    # free round trip through the log-Cholesky packing on the device
    f = par.get_free()
    par.set_free(f + 0.1)
    par.set_free(f)
    np.testing.assert_allclose(par["info"].get(), info, rtol=1e-10)


def test_moment_and_mvn_arrays(vb):
    rng = np.random.default_rng(4)
    uvn = vb.UVNParamArray("u", shape=(2, 3), min_info=0.0)
    uvn["mean"].set(rng.standard_normal((2, 3)))
    uvn["info"].set(rng.uniform(0.5, 2.0, (2, 3)))
    mom = vb.UVNMomentParamArray("m", shape=(2, 3))
    mom.set_from_uvn_param_array(uvn)
    protocol(mom)
    np.testing.assert_allclose(mom.var(), uvn.var(), rtol=1e-12)
    np.testing.assert_allclose(mom.e_exp(), uvn.e_exp(), rtol=1e-12)
    np.testing.assert_allclose(mom.var_exp(), uvn.var_exp(), rtol=1e-10)
    assert abs(float(mom.entropy()) - float(uvn.entropy())) < 1e-10
    arr = vb.MVNArray("a", shape=(4, 2), min_info=0.1)
    arr["mean"].set(rng.standard_normal((4, 2)))
    arr["info"].set(rng.uniform(0.5, 2.0, 4))
    protocol(arr)
    np.testing.assert_allclose(arr.e2(), arr.e() ** 2 + (1 / arr["info"].get())[:, None])


def test_dirichlet_param_array(vb):
    alpha = np.array([[23.0, 1.5], [4.0, 2.5], [5.0, 0.7], [6.0, 3.0], [7.0, 9.0]])   # simplex dimension first
    par = vb.DirichletParamArray("d", shape=alpha.shape, val=alpha)
    protocol(par)
    for j in range(alpha.shape[1]):
        ref = scipy.stats.dirichlet(alpha[:, j])
        assert abs(np.asarray(par.entropy())[j] - ref.entropy()) < 1e-10
        np.testing.assert_allclose(np.asarray(par.e())[:, j], ref.mean(), rtol=1e-12)
    np.testing.assert_allclose(np.asarray(par.e_log()),
                               scipy.special.digamma(alpha) - scipy.special.digamma(alpha.sum(0))[None, :], rtol=1e-11)


def test_wishart_param(vb):
    par = vb.WishartParam("w", size=3, diag_lb=0.0)
    v = np.eye(3) + np.full((3, 3), 0.1)
    par["df"].set(4.3)
    par["v"].set(v)
    protocol(par)
    ref = scipy.stats.wishart(df=4.3, scale=v)
    assert abs(float(par.entropy()) - ref.entropy()) < 1e-9
    np.testing.assert_allclose(par.e(), 4.3 * v)
    np.testing.assert_allclose(par.e_inv(), 4.3 * np.linalg.inv(v), rtol=1e-12)
    eld = sum(scipy.special.digamma(0.5 * (4.3 - j)) for j in range(3)) + 3 * np.log(2) + np.linalg.slogdet(v)[1]
    assert abs(float(par.e_log_det()) - eld) < 1e-10
    assert np.isfinite(float(par.e_log_lkj_inv_prior(2.0)))
