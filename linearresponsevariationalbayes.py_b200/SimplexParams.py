"""Simplex-valued parameters: softmax packing with a reference category, batched on the device.

Mirror of the reference's ``SimplexParams.py`` (file:line below): the same function names and the
``SimplexParam`` protocol.  The first index runs over simplexes, the second within a simplex; the
first entry of every simplex is the reference value (free value 0).  The reference computes the
Jacobian / Hessian one row at a time with Python loops over the entries (SimplexParams.py:33-63,
105-150); here each map over all rows is one kernel of ``csrc/packing.cu``.  Arithmetic always runs
in the CUDA library (numpy in -> numpy out through the device, CUDA tensors stay on the device);
there is no CPU fallback.
"""
import numpy as np
from scipy.sparse import coo_matrix

from . import _native as nat
from ._tensors import is_torch, like_input, to_device

MAX_SIMPLEX_SIZE = 64    # kSxMaxD of csrc/packing.cu


def _rows(fn_name, x, d_in, d, out_tail):
    torch = nat.require_cuda()
    lib = nat.load()
    if d < 2 or d > MAX_SIMPLEX_SIZE:
        raise ValueError('simplex size {} outside [2, {}]'.format(d, MAX_SIMPLEX_SIZE))
    xd = to_device(x)
    if xd.dim() != 2 or xd.shape[1] != d_in:
        raise ValueError('Wrong shape {} (expected (M, {}))'.format(tuple(xd.shape), d_in))
    M = int(xd.shape[0])
    out = torch.empty((M,) + tuple(out_tail), dtype=torch.float64, device=xd.device)
    nat.check(getattr(lib, fn_name)(nat.ptr(xd), M, int(d), nat.ptr(out), nat.stream_ptr()))
    return out


def constrain_simplex_matrix(free_mat):
    """(M, d-1) free -> (M, d) rows on the simplex, softmax of [0, free] (SimplexParams.py:11-18)."""
    d = int(free_mat.shape[1]) + 1
    return like_input(_rows("lrvb_simplex_constrain", free_mat, d - 1, d, (d,)), free_mat)


def unconstrain_simplex_matrix(simplex_mat):
    """(M, d) -> (M, d-1): log z[:, 1:] - log z[:, :1] (SimplexParams.py:21-23)."""
    d = int(simplex_mat.shape[1])
    return like_input(_rows("lrvb_simplex_unconstrain", simplex_mat, d, d, (d - 1,)), simplex_mat)


def constrain_simplex_vector(free_vec):
    """One simplex (SimplexParams.py:26-27)."""
    return constrain_simplex_matrix(free_vec.reshape(1, -1)).reshape(-1)


def constrain_jac_matrix(free_mat):
    """d z / d free for every row: (M, d, d-1) -- ``constrain_grad_from_moment`` of each constrained
    row (SimplexParams.py:33-38)."""
    d = int(free_mat.shape[1]) + 1
    return like_input(_rows("lrvb_simplex_jac", free_mat, d - 1, d, (d, d - 1)), free_mat)


def constrain_hess_matrix(free_mat):
    """d2 z_k / d free d free for every row: (M, d, d-1, d-1) -- ``constrain_hess_from_moment``
    (SimplexParams.py:42-63)."""
    d = int(free_mat.shape[1]) + 1
    return like_input(_rows("lrvb_simplex_hess", free_mat, d - 1, d, (d, d - 1, d - 1)), free_mat)


def _host(x):
    return x.detach().cpu().numpy() if is_torch(x) else np.asarray(x)


class SimplexParam(object):
    """A vector of simplexes, shape (number of simplexes, entries per simplex)
    (SimplexParams.py:69-176)."""

    def __init__(self, name='', shape=(1, 2), val=None):
        self.name = name
        self.__shape = (int(shape[0]), int(shape[1]))
        self.__free_shape = (self.__shape[0], self.__shape[1] - 1)
        if val is not None:
            self.set(val)
        else:
            self.set(np.full(self.__shape, 1. / self.__shape[1]))

    def __str__(self):
        return self.name + ': ' + str(self.__val)

    def names(self):
        return [self.name]

    def dictval(self):
        return _host(self.__val).tolist()

    def set(self, val):
        if tuple(val.shape) != self.__shape:
            raise ValueError('Wrong shape for SimplexParam ' + self.name)
        self.__val = val

    def get(self):
        return self.__val

    def set_free(self, free_val):
        if int(np.prod(free_val.shape)) != self.free_size():
            raise ValueError('Wrong free size for SimplexParam ' + self.name)
        self.set(constrain_simplex_matrix(free_val.reshape(self.__free_shape)))

    def get_free(self):
        return unconstrain_simplex_matrix(self.__val).reshape(-1)

    def free_to_vector(self, free_val):
        self.set_free(free_val)
        return self.get_vector()

    def free_to_vector_jac_blocks(self, free_val):
        """(M, d, d-1) per-simplex Jacobians (device in -> device out)."""
        return constrain_jac_matrix(free_val.reshape(self.__free_shape))

    def free_to_vector_hess_blocks(self, free_val):
        """(M, d, d-1, d-1) per-simplex Hessians."""
        return constrain_hess_matrix(free_val.reshape(self.__free_shape))

    def free_to_vector_jac(self, free_val):
        """Sparse (vector_size, free_size) Jacobian, rows = vector entries, columns = free values;
        entries in the reference's order (row, vector column, free column: SimplexParams.py:105-127)."""
        M, d = self.__shape
        blocks = _host(self.free_to_vector_jac_blocks(free_val))
        rows = np.broadcast_to((np.arange(M) * d)[:, None, None] + np.arange(d)[None, :, None],
                               (M, d, d - 1))
        cols = np.broadcast_to((np.arange(M) * (d - 1))[:, None, None] + np.arange(d - 1)[None, None, :],
                               (M, d, d - 1))
        return coo_matrix((blocks.reshape(-1), (rows.reshape(-1), cols.reshape(-1))),
                          (self.vector_size(), self.free_size()))

    def free_to_vector_hess(self, free_val):
        """One sparse (free_size, free_size) Hessian per vector entry, in vector order
        (SimplexParams.py:130-157)."""
        M, d = self.__shape
        blocks = _host(self.free_to_vector_hess_blocks(free_val))
        n = self.free_size()
        r1 = np.repeat(np.arange(d - 1), d - 1)
        c1 = np.tile(np.arange(d - 1), d - 1)
        hesses = []
        for m in range(M):
            off = m * (d - 1)
            for k in range(d):
                hesses.append(coo_matrix((blocks[m, k].reshape(-1), (off + r1, off + c1)), (n, n)))
        return hesses

    def set_vector(self, vec_val):
        if int(np.prod(vec_val.shape)) != self.vector_size():
            raise ValueError('Wrong vector size for SimplexParam ' + self.name)
        self.set(vec_val.reshape(self.__shape))

    def get_vector(self):
        return self.__val.reshape(-1)

    def get_vector_indices(self, row):
        return np.ravel_multi_index([[row], range(self.__shape[1])], self.__shape)

    def shape(self):
        return self.__shape

    def free_shape(self):
        return self.__free_shape

    def free_size(self):
        return int(np.prod(self.__free_shape))

    def vector_size(self):
        return int(np.prod(self.__shape))
