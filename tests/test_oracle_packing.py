"""CPU: the packing oracle (oracle/packing_oracle.py) against the golden vectors produced by the
UNMODIFIED reference (tests/golden/packing.npz, generator tests/golden/make_golden.py) and against
torch-fp64 autodiff; host-side logic of the parameter classes that needs no device."""
import os

import numpy as np
import pytest

from oracle import packing_oracle as po

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "packing.npz"))


@pytest.mark.parametrize("k", [1, 2, 3, 4])
@pytest.mark.parametrize("lbtag,lb", [("0", 0.0), ("p3", 0.3)])
def test_posdef_oracle_matches_reference(k, lbtag, lb):
    tag = "pd_k%d_lb%s" % (k, lbtag)
    free = GOLD[tag + "_free"]
    for m in range(free.shape[0]):
        mat = po.unpack_posdef_matrix(free[m], lb)
        np.testing.assert_allclose(mat, GOLD[tag + "_mat"][m], rtol=1e-14, atol=1e-15)
        np.testing.assert_allclose(po.pack_posdef_matrix(GOLD[tag + "_mat"][m], lb), GOLD[tag + "_pack"][m],
                                   rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(po.pack_posdef_matrix(mat, lb), free[m], rtol=1e-10, atol=1e-11)
        np.testing.assert_allclose(po.pos_def_matrix_free_to_vector(free[m], lb), GOLD[tag + "_vec"][m],
                                   rtol=1e-14, atol=1e-15)
        # derivatives: extrapolated differences of the reference's forward map
        J = po.pos_def_matrix_free_to_vector_jac(free[m], lb)
        np.testing.assert_allclose(J, GOLD[tag + "_jac_fd"][m], rtol=1e-8, atol=1e-9)
        if tag + "_hess_fd" in GOLD:
            H = po.pos_def_matrix_free_to_vector_hess(free[m], lb)
            np.testing.assert_allclose(H, GOLD[tag + "_hess_fd"][m], rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5])
def test_posdef_closed_forms_match_autodiff(k):
    rng = np.random.default_rng(k)
    f = rng.normal(scale=0.8, size=k * (k + 1) // 2)
    J, H = po.pos_def_autodiff(f, 0.1)
    np.testing.assert_allclose(po.pos_def_matrix_free_to_vector_jac(f, 0.1), J, rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(po.pos_def_matrix_free_to_vector_hess(f, 0.1), H, rtol=1e-13, atol=1e-13)


@pytest.mark.parametrize("d", [2, 3, 5])
def test_simplex_oracle_matches_reference(d):
    tag = "sx_d%d" % d
    free = GOLD[tag + "_free"]
    z = po.constrain_simplex_matrix(free)
    np.testing.assert_allclose(z, GOLD[tag + "_z"], rtol=1e-13, atol=1e-300)
    np.testing.assert_allclose(po.unconstrain_simplex_matrix(GOLD[tag + "_z"]), GOLD[tag + "_unc"],
                               rtol=1e-13, atol=1e-13)
    for m in range(free.shape[0]):
        np.testing.assert_allclose(po.constrain_grad_from_moment(z[m]), GOLD[tag + "_jac"][m],
                                   rtol=1e-12, atol=1e-14)   # z - z*z vs z (1 - z) near a vertex
        np.testing.assert_allclose(po.constrain_hess_from_moment(z[m]), GOLD[tag + "_hess"][m],
                                   rtol=1e-11, atol=1e-14)


def test_pdvec_reference_round_trip():
    fr = GOLD["pdvec_free"]
    mats = np.array([po.unpack_posdef_matrix(fr[6 * i:6 * i + 6], 0.2) for i in range(4)])
    np.testing.assert_allclose(mats, GOLD["pdvec_val"], rtol=1e-14)
    np.testing.assert_allclose(np.hstack([po.vectorize_ld_matrix(m) for m in mats]), GOLD["pdvec_vector"],
                               rtol=1e-14)
    np.testing.assert_allclose(GOLD["pdvec_free_back"], fr, rtol=1e-10, atol=1e-12)


def test_index_helpers_host_logic():
    import lrvb_b200 as vb
    mp = vb.MatrixParameters
    assert [mp.SymIndex(a, b) for a, b in [(0, 0), (1, 0), (0, 1), (1, 1), (2, 0), (2, 2)]] == [0, 1, 1, 2, 3, 5]
    m = np.arange(9.0).reshape(3, 3)
    np.testing.assert_array_equal(mp.vectorize_ld_matrix(m), po.vectorize_ld_matrix(m))
    v = np.arange(6.0)
    np.testing.assert_array_equal(mp.unvectorize_ld_matrix(v), po.unvectorize_ld_matrix(v))
    s = mp.unvectorize_symmetric_matrix(v)
    np.testing.assert_array_equal(s, s.T)
    np.testing.assert_array_equal(mp.vectorize_ld_matrix(s), v)
    with pytest.raises(ValueError):
        mp.unvectorize_ld_matrix(np.arange(5.0))
    with pytest.raises(ValueError):
        mp.vectorize_ld_matrix(np.zeros((2, 3)))
    p = vb.PosDefMatrixParam("p", size=3, diag_lb=0.5)
    assert p.free_size() == 6 and p.vector_size() == 6 and p.size() == 3
    np.testing.assert_array_equal(p.get(), 1.5 * np.eye(3))
    with pytest.raises(ValueError):
        p.set(np.eye(2))
    with pytest.raises(ValueError):
        p.set(np.array([[1.0, 0.2, 0], [0.1, 1, 0], [0, 0, 1]]))
    pv = vb.PosDefMatrixParamVector("pv", length=4, matrix_size=2)
    assert pv.free_size() == 12 and pv.length() == 4 and pv.free_obs_slice(2) == slice(6, 9)
    pa = vb.PosDefMatrixParamArray("pa", array_shape=(2, 3), matrix_size=2)
    assert pa.stacked_obs_slice((1, 1)) == slice(12, 15)
    sp = vb.SimplexParam("s", shape=(5, 4))
    assert sp.free_size() == 15 and sp.vector_size() == 20 and sp.free_shape() == (5, 3)
    np.testing.assert_array_equal(sp.get_vector_indices(2), np.arange(8, 12))
    with pytest.raises(ValueError):
        sp.set(np.zeros((4, 4)))
