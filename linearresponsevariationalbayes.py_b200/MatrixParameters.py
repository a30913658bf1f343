"""Positive-definite-matrix parameters: log-Cholesky packing on the device, batched.

Mirror of the reference's ``MatrixParameters.py`` (file:line below) with the same names and
protocol (``get/set``, ``get_free/set_free``, ``get_vector/set_vector``, ``free_to_vector``,
``free_to_vector_jac`` (scipy COO) and ``free_to_vector_hess`` (list of COO, one per vector
entry)).  The reference unpacks, packs and differentiates ONE k x k matrix per Python iteration
(MatrixParameters.py:236-297, 368-460) and obtains Jacobian / Hessian from autograd
(:149-152); here a parameter vector / array is ONE batched kernel per map (``csrc/packing.cu``,
closed-form derivatives).  Arithmetic always runs in the CUDA library -- numpy input is moved to the
device and back, CUDA tensors stay there; there is no CPU fallback.  The pure index shuffles
(``vectorize_ld_matrix`` and friends) do no arithmetic and work on numpy arrays or tensors directly.
"""
import math

import numpy as np
from scipy.sparse import coo_matrix

from . import _native as nat
from ._tensors import is_torch, like_input, to_device

MAX_MATRIX_SIZE = 8     # kPdMaxK of csrc/packing.cu


def SymIndex(k1, k2):
    """Packed index of entry (k1, k2) of a symmetric matrix, 0-based (MatrixParameters.py:15-23)."""
    if k2 > k1:
        k1, k2 = k2, k1
    return int(k2 + k1 * (k1 + 1) // 2)


def _matrix_size_from_vec(n):
    k = int(0.5 * (math.sqrt(1 + 8 * n) - 1))
    if k * (k + 1) // 2 != n:
        raise ValueError('Vector is an impossible size')
    return k


def vectorize_ld_matrix(mat):
    """(k, k) -> packed lower triangle [x11, x21, x22, x31, ...] (MatrixParameters.py:39-42)."""
    nrow, ncol = mat.shape[-2], mat.shape[-1]
    if nrow != ncol:
        raise ValueError('mat must be square')
    r, c = np.tril_indices(nrow)
    if is_torch(mat):
        import torch
        return mat[..., torch.as_tensor(r, device=mat.device), torch.as_tensor(c, device=mat.device)]
    return np.asarray(mat)[..., r, c]


def unvectorize_ld_matrix(vec):
    """Packed vector -> lower-triangular (k, k) matrix, zeros above the diagonal
    (MatrixParameters.py:60-68)."""
    k = _matrix_size_from_vec(vec.shape[-1])
    r, c = np.tril_indices(k)
    if is_torch(vec):
        import torch
        mat = torch.zeros(tuple(vec.shape[:-1]) + (k, k), dtype=vec.dtype, device=vec.device)
        mat[..., torch.as_tensor(r, device=vec.device), torch.as_tensor(c, device=vec.device)] = vec
        return mat
    vec = np.asarray(vec)
    mat = np.zeros(vec.shape[:-1] + (k, k), dtype=vec.dtype)
    mat[..., r, c] = vec
    return mat


def unvectorize_symmetric_matrix(vec_val):
    """Packed lower triangle -> full symmetric matrix (MatrixParameters.py:132-141)."""
    ld = unvectorize_ld_matrix(vec_val)
    if is_torch(ld):
        import torch
        return ld + ld.transpose(-1, -2) - torch.diag_embed(torch.diagonal(ld, dim1=-2, dim2=-1))
    diag = np.zeros_like(ld)
    idx = np.arange(ld.shape[-1])
    diag[..., idx, idx] = ld[..., idx, idx]
    return ld + np.swapaxes(ld, -1, -2) - diag


def _batched(fn_name, x, k, diag_lb, out_tail, in_tail, extra=()):
    """Runs one lrvb_posdef_* kernel over the leading dimensions of ``x``."""
    torch = nat.require_cuda()
    lib = nat.load()
    if k < 1 or k > MAX_MATRIX_SIZE:
        raise ValueError('matrix size {} outside [1, {}]'.format(k, MAX_MATRIX_SIZE))
    xd = to_device(x)
    lead = tuple(xd.shape[:xd.dim() - len(in_tail)])
    if tuple(xd.shape[xd.dim() - len(in_tail):]) != tuple(in_tail):
        raise ValueError('Wrong trailing shape {} (expected {})'.format(tuple(xd.shape), tuple(in_tail)))
    M = int(np.prod(lead)) if lead else 1
    out = torch.empty(lead + tuple(out_tail), dtype=torch.float64, device=xd.device)
    nat.check(getattr(lib, fn_name)(nat.ptr(xd), int(k), M, float(diag_lb), nat.ptr(out), *extra,
                                    nat.stream_ptr()))
    return out


def unpack_posdef_matrix(free_vec, diag_lb=0.0):
    """free (..., v) -> L L^T + diag_lb I (..., k, k), L = exp-diagonal lower factor
    (MatrixParameters.py:122-127), batched over the leading dimensions."""
    k = _matrix_size_from_vec(free_vec.shape[-1])
    v = k * (k + 1) // 2
    return like_input(_batched("lrvb_posdef_unpack", free_vec, k, diag_lb, (k, k), (v,)), free_vec)


def pack_posdef_matrix(mat, diag_lb=0.0):
    """(..., k, k) -> log-Cholesky free vector (..., v) of mat - diag_lb I
    (MatrixParameters.py:114-119).  Raises ``numpy.linalg.LinAlgError`` like
    ``numpy.linalg.cholesky`` when a matrix is not positive definite."""
    torch = nat.require_cuda()
    k = int(mat.shape[-1])
    if mat.shape[-2] != k:
        raise ValueError('mat must be square')
    bad = torch.zeros(1, dtype=torch.int32, device="cuda")
    out = _batched("lrvb_posdef_pack", mat, k, diag_lb, (k * (k + 1) // 2,), (k, k), extra=(nat.ptr(bad),))
    nbad = int(bad.item())
    if nbad:
        raise np.linalg.LinAlgError('{} matri{} not positive definite'.format(
            nbad, 'x is' if nbad == 1 else 'ces are'))
    return like_input(out, mat)


def pos_def_matrix_free_to_vector(free_val, diag_lb=0.0):
    """free (..., v) -> packed lower triangle of the unpacked matrix (MatrixParameters.py:145-147)."""
    k = _matrix_size_from_vec(free_val.shape[-1])
    v = k * (k + 1) // 2
    return like_input(_batched("lrvb_posdef_free_to_vector", free_val, k, diag_lb, (v,), (v,)), free_val)


def pos_def_matrix_free_to_vector_jac(free_val, diag_lb=0.0):
    """d vec / d free, (..., v, v) (autograd.jacobian in the reference, MatrixParameters.py:149-150)."""
    k = _matrix_size_from_vec(free_val.shape[-1])
    v = k * (k + 1) // 2
    return like_input(_batched("lrvb_posdef_free_to_vector_jac", free_val, k, diag_lb, (v, v), (v,)),
                      free_val)


def pos_def_matrix_free_to_vector_hess(free_val, diag_lb=0.0):
    """d2 vec_r / d free d free, (..., v, v, v) (autograd.hessian, MatrixParameters.py:151-152)."""
    k = _matrix_size_from_vec(free_val.shape[-1])
    v = k * (k + 1) // 2
    return like_input(_batched("lrvb_posdef_free_to_vector_hess", free_val, k, diag_lb, (v, v, v), (v,)),
                      free_val)


def _block_diagonal_jac(blocks):
    """(M, v, v) dense blocks -> COO (M v, M v), entries in the reference's emission order
    (block, vector row, free column: MatrixParameters.py:257-266)."""
    M, v, _ = blocks.shape
    base = (np.arange(M) * v)[:, None, None]
    rows = np.broadcast_to(base + np.arange(v)[None, :, None], (M, v, v))
    cols = np.broadcast_to(base + np.arange(v)[None, None, :], (M, v, v))
    return coo_matrix((blocks.reshape(-1), (rows.reshape(-1), cols.reshape(-1))), (M * v, M * v))


def _block_diagonal_hess(blocks):
    """(M, v, v, v) -> list of M v COO matrices (M v, M v): entry [m v + r] holds block (m, r) on the
    diagonal block m (MatrixParameters.py:273-297)."""
    M, v = blocks.shape[0], blocks.shape[1]
    n = M * v
    out = []
    r1 = np.repeat(np.arange(v), v)
    c1 = np.tile(np.arange(v), v)
    for m in range(M):
        for r in range(v):
            out.append(coo_matrix((blocks[m, r].reshape(-1), (m * v + r1, m * v + c1)), (n, n)))
    return out


def _host(x):
    return x.detach().cpu().numpy() if is_torch(x) else np.asarray(x)


class PosDefMatrixParam(object):
    """One symmetric positive-definite (size, size) matrix (MatrixParameters.py:155-208)."""

    def __init__(self, name='', size=2, diag_lb=0.0, val=None):
        self.name = name
        self.__size = int(size)
        self.__vec_size = int(size * (size + 1) // 2)
        self.__diag_lb = diag_lb
        assert diag_lb >= 0
        if val is None:
            self.__val = np.diag(np.full(self.__size, diag_lb + 1.0))
        else:
            self.set(val)

    def __str__(self):
        return self.name + ':\n' + str(self.__val)

    def names(self):
        return [self.name]

    def dictval(self):
        return _host(self.__val).tolist()

    def set(self, val):
        nrow, ncol = val.shape
        if nrow != self.__size or ncol != self.__size:
            raise ValueError('Matrix is a different size')
        if not bool((val.T == val).all()):
            raise ValueError('Matrix is not symmetric')
        self.__val = val

    def get(self):
        return self.__val

    def set_free(self, free_val):
        if int(np.prod(free_val.shape)) != self.__vec_size:
            raise ValueError('Free value is the wrong length')
        self.__val = unpack_posdef_matrix(free_val.reshape(-1), diag_lb=self.__diag_lb)

    def get_free(self):
        return pack_posdef_matrix(self.__val, diag_lb=self.__diag_lb)

    def free_to_vector(self, free_val):
        self.set_free(free_val)
        return self.get_vector()

    def free_to_vector_jac_dense(self, free_val):
        return pos_def_matrix_free_to_vector_jac(free_val.reshape(-1), diag_lb=self.__diag_lb)

    def free_to_vector_hess_dense(self, free_val):
        return pos_def_matrix_free_to_vector_hess(free_val.reshape(-1), diag_lb=self.__diag_lb)

    def free_to_vector_jac(self, free_val):
        return coo_matrix(_host(self.free_to_vector_jac_dense(free_val)))

    def free_to_vector_hess(self, free_val):
        hess_dense = _host(self.free_to_vector_hess_dense(free_val))
        return [coo_matrix(hess_dense[ind, :, :]) for ind in range(hess_dense.shape[0])]

    def set_vector(self, vec_val):
        if int(np.prod(vec_val.shape)) != self.__vec_size:
            raise ValueError('Vector value is the wrong length')
        self.__val = unvectorize_symmetric_matrix(vec_val.reshape(-1))

    def get_vector(self):
        return vectorize_ld_matrix(self.__val)

    def size(self):
        return self.__size

    def free_size(self):
        return self.__vec_size

    def vector_size(self):
        return self.__vec_size


class PosDefMatrixParamArray(object):
    """An array of positive-definite matrices, the last two indices are the matrix
    (MatrixParameters.py:320-470).  Every map is one batched kernel over the array."""

    def __init__(self, name='', array_shape=(1,), matrix_size=2, diag_lb=0.0, val=None):
        self.name = name
        self.__matrix_size = int(matrix_size)
        if isinstance(array_shape, (int, np.integer)):
            array_shape = (int(array_shape),)
        self.__array_shape = tuple(int(s) for s in array_shape)
        self.__array_length = int(np.prod(self.__array_shape))
        self.__shape = self.__array_shape + (self.__matrix_size, self.__matrix_size)
        self.__vec_size = int(matrix_size * (matrix_size + 1) // 2)
        self.__diag_lb = diag_lb
        assert diag_lb >= 0
        if val is None:
            default_val = np.diag(np.full(self.__matrix_size, diag_lb + 1.0))
            self.__val = np.broadcast_to(default_val, self.__shape)
        else:
            self.set(val)

    def __str__(self):
        return self.name + ':\n' + str(self.__val)

    def names(self):
        return [self.name]

    def dictval(self):
        return _host(self.__val).tolist()

    def set(self, val):
        if tuple(val.shape) != self.__shape:
            raise ValueError('Array is the wrong size')
        self.__val = val

    def get(self):
        return self.__val

    def stacked_obs_slice(self, obs):
        """Slice of the free / vector representation holding array element ``obs`` (a tuple)
        (MatrixParameters.py:363-368)."""
        assert len(obs) == len(self.__array_shape)
        linear_obs = int(np.ravel_multi_index(obs, self.__array_shape)) * self.__vec_size
        return slice(linear_obs, linear_obs + self.__vec_size)

    def _packed(self, flat, what):
        if int(np.prod(flat.shape)) != self.free_size():
            raise ValueError('{} value is the wrong length'.format(what))
        return flat.reshape(self.__array_shape + (self.__vec_size,))

    def set_free(self, free_val):
        self.__val = unpack_posdef_matrix(self._packed(free_val, 'Free'), diag_lb=self.__diag_lb)

    def get_free(self):
        return pack_posdef_matrix(self.__val, diag_lb=self.__diag_lb).reshape(-1)

    def apply_matrix_function(self, mat_func):
        import itertools
        res = np.array([mat_func(self.__val[obs])
                        for obs in itertools.product(*[range(t) for t in self.__array_shape])])
        return np.reshape(res, self.__array_shape + res[0].shape)

    def set_vector(self, vec_val):
        self.__val = unvectorize_symmetric_matrix(self._packed(vec_val, 'Vector'))

    def get_vector(self):
        return vectorize_ld_matrix(self.__val).reshape(-1)

    def free_to_vector(self, free_val):
        self.set_free(free_val)
        return self.get_vector()

    def free_to_vector_jac_blocks(self, free_val):
        """(M, v, v) Jacobian blocks, device in -> device out (M = number of matrices)."""
        f = self._packed(free_val, 'Free').reshape(self.__array_length, self.__vec_size)
        return pos_def_matrix_free_to_vector_jac(f, diag_lb=self.__diag_lb)

    def free_to_vector_hess_blocks(self, free_val):
        """(M, v, v, v) Hessian blocks."""
        f = self._packed(free_val, 'Free').reshape(self.__array_length, self.__vec_size)
        return pos_def_matrix_free_to_vector_hess(f, diag_lb=self.__diag_lb)

    def free_to_vector_jac(self, free_val):
        return _block_diagonal_jac(_host(self.free_to_vector_jac_blocks(free_val)))

    def free_to_vector_hess(self, free_val):
        return _block_diagonal_hess(_host(self.free_to_vector_hess_blocks(free_val)))

    def array_shape(self):
        return self.__array_shape

    def matrix_size(self):
        return self.__matrix_size

    def free_size(self):
        return self.__vec_size * self.__array_length

    def vector_size(self):
        return self.__vec_size * self.__array_length


class PosDefMatrixParamVector(PosDefMatrixParamArray):
    """A vector of ``length`` positive-definite matrices (MatrixParameters.py:211-316)."""

    def __init__(self, name='', length=1, matrix_size=2, diag_lb=0.0, val=None):
        PosDefMatrixParamArray.__init__(self, name=name, array_shape=(int(length),),
                                        matrix_size=matrix_size, diag_lb=diag_lb, val=val)
        self.__length = int(length)
        self.__vec = int(matrix_size * (matrix_size + 1) // 2)

    def free_obs_slice(self, obs):
        assert obs < self.__length
        return slice(self.__vec * obs, self.__vec * (obs + 1))

    def length(self):
        return self.__length
