"""GPU: batched matrix / simplex parameter packing (csrc/packing.cu through MatrixParameters.py and
SimplexParams.py) against the oracle, the reference's golden vectors, and size-independent
properties at a million parameters.  Tolerance: 1e-9 rel / 1e-12 abs (fp64)."""
import os

import numpy as np
import pytest
import torch

from oracle import packing_oracle as po

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-9, 1e-12
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "packing.npz"))


def close(a, b, rtol=RTOL, atol=ATOL):
    a = a.detach().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol)


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 8])
@pytest.mark.parametrize("lb", [0.0, 0.3])
def test_posdef_maps_match_oracle(k, lb):
    import lrvb_b200 as vb
    mp = vb.MatrixParameters
    rng = np.random.default_rng(100 + k)
    v = k * (k + 1) // 2
    M = 37
    free = rng.normal(scale=0.6, size=(M, v))
    mats = np.array([po.unpack_posdef_matrix(f, lb) for f in free])
    close(mp.unpack_posdef_matrix(free, diag_lb=lb), mats)
    close(mp.pos_def_matrix_free_to_vector(free, diag_lb=lb), np.array([po.vectorize_ld_matrix(m) for m in mats]))
    close(mp.pack_posdef_matrix(mats, diag_lb=lb), free)
    J = mp.pos_def_matrix_free_to_vector_jac(free, diag_lb=lb)
    close(J, np.array([po.pos_def_matrix_free_to_vector_jac(f, lb) for f in free]))
    if k <= 5:
        H = mp.pos_def_matrix_free_to_vector_hess(free, diag_lb=lb)
        close(H, np.array([po.pos_def_matrix_free_to_vector_hess(f, lb) for f in free]))
    # device tensors stay on the device
    out = mp.unpack_posdef_matrix(torch.from_numpy(free).cuda(), diag_lb=lb)
    assert torch.is_tensor(out) and out.is_cuda
    close(out, mats)


@pytest.mark.parametrize("k", [1, 2, 3, 4])
@pytest.mark.parametrize("lbtag,lb", [("0", 0.0), ("p3", 0.3)])
def test_posdef_maps_match_reference_golden(k, lbtag, lb):
    import lrvb_b200 as vb
    mp = vb.MatrixParameters
    tag = "pd_k%d_lb%s" % (k, lbtag)
    free = GOLD[tag + "_free"]
    close(mp.unpack_posdef_matrix(free, diag_lb=lb), GOLD[tag + "_mat"])
    close(mp.pack_posdef_matrix(GOLD[tag + "_mat"], diag_lb=lb), GOLD[tag + "_pack"])
    close(mp.pos_def_matrix_free_to_vector(free, diag_lb=lb), GOLD[tag + "_vec"])
    close(mp.unvectorize_symmetric_matrix(mp.vectorize_ld_matrix(GOLD[tag + "_mat"])), GOLD[tag + "_sym"])
    # derivative goldens are extrapolated differences of the reference's forward map
    close(mp.pos_def_matrix_free_to_vector_jac(free, diag_lb=lb), GOLD[tag + "_jac_fd"], 1e-8, 1e-9)
    if tag + "_hess_fd" in GOLD:
        close(mp.pos_def_matrix_free_to_vector_hess(free, diag_lb=lb), GOLD[tag + "_hess_fd"], 1e-6, 1e-6)


def test_posdef_param_classes():
    import lrvb_b200 as vb
    fr = GOLD["pdvec_free"]
    pv = vb.PosDefMatrixParamVector("pv", length=4, matrix_size=3, diag_lb=0.2)
    pv.set_free(fr)
    close(pv.get(), GOLD["pdvec_val"])
    close(pv.get_vector(), GOLD["pdvec_vector"])
    close(pv.get_free(), fr, 1e-9, 1e-11)
    close(pv.free_to_vector(fr), GOLD["pdvec_vector"])
    pv2 = vb.PosDefMatrixParamVector("pv2", length=4, matrix_size=3, diag_lb=0.2)
    pv2.set_vector(GOLD["pdvec_vector"])
    close(pv2.get(), GOLD["pdvec_val"])
    # sparse Jacobian / Hessians: block diagonal, blocks = the single-matrix maps
    J = pv.free_to_vector_jac(fr)
    assert J.shape == (24, 24) and J.nnz == 4 * 36
    Jd = J.toarray()
    hs = pv.free_to_vector_hess(fr)
    assert len(hs) == 24 and hs[0].shape == (24, 24)
    for m in range(4):
        blk = po.pos_def_matrix_free_to_vector_jac(fr[6 * m:6 * m + 6], 0.2)
        close(Jd[6 * m:6 * m + 6, 6 * m:6 * m + 6], blk)
        Hm = po.pos_def_matrix_free_to_vector_hess(fr[6 * m:6 * m + 6], 0.2)
        for r in range(6):
            full = hs[6 * m + r].toarray()
            close(full[6 * m:6 * m + 6, 6 * m:6 * m + 6], Hm[r])
            full[6 * m:6 * m + 6, 6 * m:6 * m + 6] = 0
            assert not full.any()
    Jd[np.kron(np.eye(4), np.ones((6, 6))) > 0] = 0
    assert not Jd.any()
    # single matrix
    p = vb.PosDefMatrixParam("p", size=3, diag_lb=0.2)
    p.set_free(fr[:6])
    close(p.get(), GOLD["pdvec_val"][0])
    close(p.get_free(), fr[:6], 1e-9, 1e-11)
    close(p.free_to_vector_jac(fr[:6]).toarray(), po.pos_def_matrix_free_to_vector_jac(fr[:6], 0.2))
    assert len(p.free_to_vector_hess(fr[:6])) == 6
    # array of matrices
    pa = vb.PosDefMatrixParamArray("pa", array_shape=(2, 2), matrix_size=3, diag_lb=0.2)
    pa.set_free(fr)
    close(pa.get().reshape(4, 3, 3), GOLD["pdvec_val"])
    close(pa.get_free(), fr, 1e-9, 1e-11)
    with pytest.raises(ValueError):
        pv.set_free(fr[:-1])
    with pytest.raises(np.linalg.LinAlgError):
        vb.MatrixParameters.pack_posdef_matrix(np.array([[1.0, 2.0], [2.0, 1.0]]))
    with pytest.raises(ValueError):
        vb.MatrixParameters.unpack_posdef_matrix(np.zeros(45))     # k = 9 > 8


@pytest.mark.parametrize("d", [2, 3, 5, 17, 64])
def test_simplex_maps_match_oracle(d):
    import lrvb_b200 as vb
    sx = vb.SimplexParams
    rng = np.random.default_rng(200 + d)
    M = 41
    free = rng.normal(scale=2.0, size=(M, d - 1))
    free[0] *= 30.0
    z = po.constrain_simplex_matrix(free)
    close(sx.constrain_simplex_matrix(free), z, 1e-12, 1e-300)
    close(sx.unconstrain_simplex_matrix(z[1:]), free[1:], 1e-9, 1e-10)
    close(sx.constrain_jac_matrix(free), np.array([po.constrain_grad_from_moment(r) for r in z]))
    if d <= 17:
        close(sx.constrain_hess_matrix(free), np.array([po.constrain_hess_from_moment(r) for r in z]))
    close(sx.constrain_simplex_vector(free[3]), z[3])


@pytest.mark.parametrize("d", [2, 3, 5])
def test_simplex_matches_reference_golden(d):
    import lrvb_b200 as vb
    sx = vb.SimplexParams
    tag = "sx_d%d" % d
    free = GOLD[tag + "_free"]
    close(sx.constrain_simplex_matrix(free), GOLD[tag + "_z"], 1e-12, 1e-300)
    close(sx.unconstrain_simplex_matrix(GOLD[tag + "_z"]), GOLD[tag + "_unc"], 1e-12, 1e-12)
    close(sx.constrain_jac_matrix(free), GOLD[tag + "_jac"])
    close(sx.constrain_hess_matrix(free), GOLD[tag + "_hess"])
    par = vb.SimplexParam("s", shape=(6, d))
    par.set_free(free.flatten())
    close(par.get(), GOLD[tag + "_z"], 1e-12, 1e-300)
    close(par.get_free(), GOLD[tag + "_unc"].flatten(), 1e-12, 1e-12)
    close(par.free_to_vector_jac(free.flatten()).toarray(), GOLD[tag + "_jac_sparse"])
    hs = par.free_to_vector_hess(free.flatten())
    assert len(hs) == 6 * d
    close(np.array([h.toarray() for h in hs]), GOLD[tag + "_hess_sparse"])
    with pytest.raises(ValueError):
        par.set_free(free.flatten()[:-1])


def test_million_parameters_properties():
    """Full-size properties: pack(unpack(f)) = f, rows of z sum to one, Jacobian columns of a simplex
    sum to zero, unpacked matrices are symmetric with positive pivots."""
    import lrvb_b200 as vb
    mp, sx = vb.MatrixParameters, vb.SimplexParams
    M = 1000000
    gen = torch.Generator(device="cuda").manual_seed(5)
    free = torch.randn(M, 6, generator=gen, device="cuda", dtype=torch.float64) * 0.7
    mats = mp.unpack_posdef_matrix(free, diag_lb=0.1)
    assert torch.equal(mats, mats.transpose(-1, -2))
    back = mp.pack_posdef_matrix(mats, diag_lb=0.1)
    assert float((back - free).abs().max()) < 1e-9
    vec = mp.pos_def_matrix_free_to_vector(free, diag_lb=0.1)
    assert torch.equal(vec, mp.vectorize_ld_matrix(mats))
    J = mp.pos_def_matrix_free_to_vector_jac(free[:200000])
    # d vec / d free applied to a direction = finite difference of the map
    dirn = torch.randn(200000, 6, generator=gen, device="cuda", dtype=torch.float64)
    h = 1e-6
    fd = (mp.pos_def_matrix_free_to_vector(free[:200000] + h * dirn)
          - mp.pos_def_matrix_free_to_vector(free[:200000] - h * dirn)) / (2 * h)
    lin = torch.einsum("mrc,mc->mr", J, dirn)
    assert float((fd - lin).abs().max() / lin.abs().max()) < 1e-7
    fs = torch.randn(M, 4, generator=gen, device="cuda", dtype=torch.float64) * 3
    z = sx.constrain_simplex_matrix(fs)
    assert float((z.sum(1) - 1).abs().max()) < 1e-14 and float(z.min()) > 0
    assert float((sx.unconstrain_simplex_matrix(z) - fs).abs().max()) < 1e-8
    Js = sx.constrain_jac_matrix(fs[:200000])
    assert float(Js.sum(1).abs().max()) < 1e-14
    Hs = sx.constrain_hess_matrix(fs[:200000])
    assert float(Hs.sum(1).abs().max()) < 1e-14
    assert torch.equal(Hs, Hs.transpose(-1, -2))


def test_empty_batches():
    import lrvb_b200 as vb
    assert vb.MatrixParameters.unpack_posdef_matrix(np.zeros((0, 3))).shape == (0, 2, 2)
    assert vb.SimplexParams.constrain_simplex_matrix(np.zeros((0, 3))).shape == (0, 4)
