"""CPU restatement of the reference's matrix / simplex parameter packing.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): imported by tests/ and smoke() as the checker,
never by the product path.  Parity pin: tests/golden/packing.npz holds outputs of the UNMODIFIED
reference functions (run under oracle/ref_shim.py by tests/golden/make_golden.py) for the value maps
and the simplex derivative maps, and Richardson-extrapolated central differences of the reference's
``pos_def_matrix_free_to_vector`` for the log-Cholesky Jacobian / Hessian (the reference gets those
from autograd, which is not installable here); tests/test_oracle_golden.py checks this file against
them, and the closed forms below against torch-fp64 autodiff.

Every function cites the reference lines it follows
(/root/reference/LinearResponseVariationalBayes/...).
"""
import numpy as np


# ---- MatrixParameters.py ---------------------------------------------------------------------------
def vectorize_ld_matrix(mat):
    """MatrixParameters.py:39-42."""
    return mat[np.tril_indices(mat.shape[0])]


def unvectorize_ld_matrix(vec):
    """MatrixParameters.py:60-68 (loop form kept: k1 rows, k2 <= k1 columns, SymIndex order)."""
    n = int(0.5 * (np.sqrt(1 + 8 * vec.size) - 1))
    assert n * (n + 1) // 2 == vec.size
    mat = np.zeros((n, n), dtype=vec.dtype)
    for k1 in range(n):
        for k2 in range(k1 + 1):
            mat[k1, k2] = vec[k2 + k1 * (k1 + 1) // 2]
    return mat


def unpack_posdef_matrix(free_vec, diag_lb=0.0):
    """MatrixParameters.py:122-127: chol = exp-diagonal of the unvectorized free vector,
    mat = chol chol^T + diag_lb I."""
    ld = unvectorize_ld_matrix(free_vec)
    chol = ld - np.diag(np.diag(ld)) + np.diag(np.exp(np.diag(ld)))      # :86-93
    return chol @ chol.T + diag_lb * np.eye(ld.shape[0])


def pack_posdef_matrix(mat, diag_lb=0.0):
    """MatrixParameters.py:114-119."""
    k = mat.shape[0]
    chol = np.linalg.cholesky(mat - diag_lb * np.eye(k))
    logd = chol - np.diag(np.diag(chol)) + np.diag(np.log(np.diag(chol)))  # :96-103
    return vectorize_ld_matrix(logd)


def pos_def_matrix_free_to_vector(free_val, diag_lb=0.0):
    """MatrixParameters.py:145-147."""
    return vectorize_ld_matrix(unpack_posdef_matrix(free_val, diag_lb))


def _ld_index(v):
    n = int(0.5 * (np.sqrt(1 + 8 * v) - 1))
    return [(i, j) for i in range(n) for j in range(i + 1)]


def pos_def_matrix_free_to_vector_jac(free_val, diag_lb=0.0):
    """Closed form of autograd.jacobian(pos_def_matrix_free_to_vector) (MatrixParameters.py:149-150):
    A_ab = sum_c L_ac L_bc  =>  dA_ab/df_ij = D_ij ([a=i] L_bj + [b=i] L_aj)."""
    v = free_val.size
    idx = _ld_index(v)
    ld = unvectorize_ld_matrix(free_val)
    L = ld - np.diag(np.diag(ld)) + np.diag(np.exp(np.diag(ld)))
    J = np.zeros((v, v))
    for r, (a, b) in enumerate(idx):
        for c, (i, j) in enumerate(idx):
            D = L[i, i] if i == j else 1.0
            s = 0.0
            if a == i:
                s += L[b, j]
            if b == i:
                s += L[a, j]
            J[r, c] = D * s
    return J


def pos_def_matrix_free_to_vector_hess(free_val, diag_lb=0.0):
    """Closed form of autograd.hessian(...) (MatrixParameters.py:151-152), (v, v, v)."""
    v = free_val.size
    idx = _ld_index(v)
    ld = unvectorize_ld_matrix(free_val)
    L = ld - np.diag(np.diag(ld)) + np.diag(np.exp(np.diag(ld)))
    H = np.zeros((v, v, v))
    for r, (a, b) in enumerate(idx):
        for c1, (i, j) in enumerate(idx):
            for c2, (p, q) in enumerate(idx):
                s = 0.0
                if j == q:
                    D1 = L[i, i] if i == j else 1.0
                    D2 = L[p, p] if p == q else 1.0
                    t = (1.0 if (a == i and b == p) else 0.0) + (1.0 if (b == i and a == p) else 0.0)
                    s += D1 * D2 * t
                if c1 == c2 and i == j:
                    t = (L[b, j] if a == i else 0.0) + (L[a, j] if b == i else 0.0)
                    s += L[i, i] * t
                H[r, c1, c2] = s
    return H


def pos_def_autodiff(free_val, diag_lb=0.0):
    """torch-fp64 autodiff of the same composition: (jac (v, v), hess (v, v, v))."""
    import torch
    v = free_val.size
    idx = _ld_index(v)
    n = idx[-1][0] + 1
    rows = torch.tensor([i for i, _ in idx])
    cols = torch.tensor([j for _, j in idx])

    def f(x):
        ld = torch.zeros(n, n, dtype=torch.float64).index_put((rows, cols), x)
        d = torch.diagonal(ld)
        chol = ld - torch.diag(d) + torch.diag(torch.exp(d))
        mat = chol @ chol.T + diag_lb * torch.eye(n, dtype=torch.float64)
        return mat[rows, cols]
    x = torch.tensor(np.asarray(free_val, dtype=np.float64))
    jac = torch.autograd.functional.jacobian(f, x)
    hess = torch.stack([torch.autograd.functional.hessian(lambda t, r=r: f(t)[r], x) for r in range(v)])
    return jac.numpy(), hess.numpy()


# ---- SimplexParams.py ---------------------------------------------------------------------------------
def constrain_simplex_matrix(free_mat):
    """SimplexParams.py:11-18."""
    aug = np.hstack([np.zeros((free_mat.shape[0], 1)), free_mat])
    mx = aug.max(axis=1, keepdims=True)
    log_norm = mx + np.log(np.exp(aug - mx).sum(axis=1, keepdims=True))      # logsumexp
    return np.exp(aug - log_norm)


def unconstrain_simplex_matrix(simplex_mat):
    """SimplexParams.py:21-23."""
    return np.log(simplex_mat[:, 1:]) - np.log(simplex_mat[:, :1])


def constrain_grad_from_moment(z):
    """SimplexParams.py:33-38."""
    z_last = z[1:]
    jac = -np.outer(z, z_last)
    for k in range(1, len(z)):
        jac[k, k - 1] += z[k]
    return jac


def constrain_hess_from_moment(z):
    """SimplexParams.py:42-63, restated from the softmax second derivative
    d2 z_k = z_k (([k=a+1] - z_{a+1}) ([k=b+1] - z_{b+1}) - z_{a+1} ([a=b] - z_{b+1}))."""
    d = len(z)
    H = np.zeros((d, d - 1, d - 1))
    for k in range(d):
        for a in range(d - 1):
            for b in range(d - 1):
                ta = (1.0 if k == a + 1 else 0.0) - z[a + 1]
                tb = (1.0 if k == b + 1 else 0.0) - z[b + 1]
                H[k, a, b] = z[k] * (ta * tb - z[a + 1] * ((1.0 if a == b else 0.0) - z[b + 1]))
    return H
