"""Constrained <-> unconstrained ("vector" <-> "free") parameter packing.

Host-side mirror of the reference's parameter protocol
(/root/reference/LinearResponseVariationalBayes/Parameters.py:31-61 transforms, :82-150
ScalarParam, :154-231 VectorParam, :234-322 ArrayParam, :326-424 offset helpers and
convert_vector_to_free_hessian).  The protocol -- ``get/set``, ``get_free/set_free``,
``get_vector/set_vector``, ``free_size/vector_size``, ``free_to_vector[_jac|_hess]``, ``names``,
``dictval`` -- is what the reference's tests exercise (test_variational_bayes.py:74-106).

Differences by design: the diagonal Jacobian / second derivative of the elementwise transforms
are closed-form numpy expressions (the reference calls autograd once per element,
Parameters.py:200-218); everything is numpy-2 / scipy-1.18 clean.  The GLMM hot path does not run
through this module at all -- its transforms are fused into the CUDA kernels -- this is the
packing layer users keep.
"""
import copy
import numbers

import numpy as np
from scipy.sparse import coo_matrix

_INF = float("inf")


def _check_bounds(lb, ub):
    if ub <= lb:
        raise ValueError("Upper bound must be greater than lower bound")


def unconstrain(vec, lb, ub):
    """vector -> free (Parameters.py:31-44)."""
    _check_bounds(lb, ub)
    lower, upper = lb > -_INF, ub < _INF
    if lower and upper:
        return np.log(vec - lb) - np.log(ub - vec)
    if lower:
        return np.log(vec - lb)
    if upper:
        return -1 * np.log(ub - vec)
    return copy.copy(vec)


def constrain(free_vec, lb, ub):
    """free -> vector (Parameters.py:47-61)."""
    _check_bounds(lb, ub)
    lower, upper = lb > -_INF, ub < _INF
    if lower and upper:
        ex = np.exp(free_vec)
        return (ub - lb) * ex / (1 + ex) + lb
    if lower:
        return np.exp(free_vec) + lb
    if upper:
        return ub - np.exp(-1 * free_vec)
    return copy.copy(free_vec)


def constrain_scalar_jac(free_val, lb, ub):
    """d constrain / d free, elementwise (closed form of Parameters.py:63)."""
    _check_bounds(lb, ub)
    free_val = np.asarray(free_val, dtype=np.float64)
    lower, upper = lb > -_INF, ub < _INF
    if lower and upper:
        s = 1.0 / (1.0 + np.exp(-free_val))
        return (ub - lb) * s * (1 - s)
    if lower:
        return np.exp(free_val)
    if upper:
        return np.exp(-free_val)
    return np.ones_like(free_val)


def constrain_scalar_hess(free_val, lb, ub):
    """d2 constrain / d free2, elementwise (closed form of Parameters.py:64)."""
    _check_bounds(lb, ub)
    free_val = np.asarray(free_val, dtype=np.float64)
    lower, upper = lb > -_INF, ub < _INF
    if lower and upper:
        s = 1.0 / (1.0 + np.exp(-free_val))
        return (ub - lb) * s * (1 - s) * (1 - 2 * s)
    if lower:
        return np.exp(free_val)
    if upper:
        return -np.exp(-free_val)
    return np.zeros_like(free_val)


def unconstrain_array(vec, lb, ub):
    vec = np.asarray(vec)
    if not (vec <= ub).all():
        raise ValueError("Elements larger than the upper bound")
    if not (vec >= lb).all():
        raise ValueError("Elements smaller than the lower bound")
    return np.asarray(unconstrain(vec, lb, ub)).flatten()


def unconstrain_scalar(val, lb, ub):
    if not val <= ub:
        raise ValueError("Value larger than the upper bound")
    if not val >= lb:
        raise ValueError("Value smaller than the lower bound")
    return unconstrain(val, lb, ub)


def get_inbounds_value(lb, ub):
    """A default value strictly inside (lb, ub) (Parameters.py:66-79)."""
    assert lb < ub
    if lb > -_INF and ub < _INF:
        return 0.5 * (ub - lb)
    if lb > -_INF:
        return lb + 1.0
    if ub < _INF:
        return ub - 1.0
    return 0.0


class _ElementwiseParam(object):
    """Shared machinery of the three elementwise-transformed parameter kinds."""

    def __init__(self, name, lb, ub):
        if lb >= ub:
            raise ValueError("Upper bound must strictly exceed lower bound")
        assert lb >= -_INF and ub <= _INF
        self.name = name
        self._lb, self._ub = lb, ub

    # bounds are read by the GLMM model to hand the transforms to the CUDA kernels
    def bounds(self):
        return self._lb, self._ub

    def get(self):
        return self._val

    def free_to_vector(self, free_val):
        self.set_free(free_val)
        return self.get_vector()

    def free_to_vector_jac(self, free_val):
        free_val = np.atleast_1d(np.asarray(free_val, dtype=np.float64)).ravel()
        idx = np.arange(self.vector_size())
        d1 = constrain_scalar_jac(free_val, self._lb, self._ub)
        return coo_matrix((d1, (idx, idx)), (self.vector_size(), self.free_size()))

    def free_to_vector_hess(self, free_val):
        free_val = np.atleast_1d(np.asarray(free_val, dtype=np.float64)).ravel()
        d2 = constrain_scalar_hess(free_val, self._lb, self._ub)
        shape = (self.free_size(), self.free_size())
        return [coo_matrix(([d2[i]], ([i], [i])), shape) for i in range(self.vector_size())]


class ScalarParam(_ElementwiseParam):
    def __init__(self, name="", lb=-_INF, ub=_INF, val=None):
        super().__init__(name, lb, ub)
        self.set(get_inbounds_value(lb, ub) if val is None else val)

    def __str__(self):
        return self.name + ": " + str(self._val)

    def names(self):
        return [self.name]

    def dictval(self):
        return self._val if isinstance(self._val, numbers.Number) else np.asarray(self._val).tolist()

    def set(self, val):
        self._val = val

    def set_free(self, free_val):
        self.set(constrain(free_val, self._lb, self._ub))

    def get_free(self):
        return np.reshape(unconstrain_scalar(self._val, self._lb, self._ub), 1)

    def set_vector(self, val):
        self.set(val)

    def get_vector(self):
        return np.reshape(self._val, 1)

    def size(self):
        return 1

    free_size = size
    vector_size = size


class VectorParam(_ElementwiseParam):
    def __init__(self, name="", size=1, lb=-_INF, ub=_INF, val=None):
        super().__init__(name, lb, ub)
        self._size = int(size)
        self.set(np.full(self._size, get_inbounds_value(lb, ub)) if val is None else val)

    def __str__(self):
        return self.name + ":\n" + str(self._val)

    def names(self):
        return [self.name + "_" + str(k) for k in range(self._size)]

    def dictval(self):
        return np.asarray(self._val).tolist()

    def set(self, val):
        if val.size != self._size:
            raise ValueError("Wrong size for vector " + self.name + ".  Expected: " +
                             str(self._size) + ", got " + str(val.size))
        self._val = val

    def set_free(self, free_val):
        if free_val.size != self._size:
            raise ValueError("Wrong size for vector " + self.name)
        self.set(constrain(free_val, self._lb, self._ub))

    def get_free(self):
        return unconstrain_array(self._val, self._lb, self._ub)

    def set_vector(self, val):
        self.set(val)

    def get_vector(self):
        return self._val

    def size(self):
        return self._size

    free_size = size
    vector_size = size


class ArrayParam(_ElementwiseParam):
    def __init__(self, name="", shape=(1, 1), lb=-_INF, ub=_INF, val=None):
        super().__init__(name, lb, ub)
        self._shape = tuple(shape)
        self.set(np.full(self._shape, get_inbounds_value(lb, ub)) if val is None else val)

    def __str__(self):
        return self.name + ":\n" + str(self._val)

    def names(self):
        return self.name  # a str, as in the reference (Parameters.py:253-254)

    def dictval(self):
        return np.asarray(self._val).tolist()

    def set(self, val):
        if tuple(val.shape) != self._shape:
            raise ValueError("Wrong size for array " + self.name + " Expected shape: " +
                             str(self._shape) + " Got shape: " + str(val.shape))
        self._val = val

    def _check(self, n, what):
        if n != self.free_size():
            raise ValueError("Wrong size for array {}.  Expected {}, got {}".format(
                self.name, str(self.free_size()), str(n)))

    def set_free(self, free_val):
        self._check(free_val.size, "free")
        self.set(np.reshape(constrain(free_val, self._lb, self._ub), self._shape))

    def get_free(self):
        return unconstrain_array(self._val, self._lb, self._ub)

    def set_vector(self, val):
        self._check(val.size, "vector")
        self.set(np.reshape(val, self._shape))

    def get_vector(self):
        return np.asarray(self._val).flatten()  # C order

    def shape(self):
        return self._shape

    def free_size(self):
        return int(np.prod(self._shape))

    vector_size = free_size


# ---- offset helpers used by ModelParamsDict (Parameters.py:326-387) ---------------------------

def set_free_offset(param, free_vec, offset):
    n = param.free_size()
    param.set_free(free_vec[offset:offset + n])
    return offset + n


def get_free_offset(param, vec, offset):
    n = param.free_size()
    vec[offset:offset + n] = param.get_free()
    return offset + n


def set_vector_offset(param, vec, offset):
    n = param.vector_size()
    param.set_vector(vec[offset:offset + n])
    return offset + n


def get_vector_offset(param, vec, offset):
    n = param.vector_size()
    vec[offset:offset + n] = param.get_vector()
    return offset + n


def free_to_vector_jac_offset(param, free_vec, free_offset, vec_offset):
    nf = param.free_size()
    jac = param.free_to_vector_jac(free_vec[free_offset:free_offset + nf])
    return free_offset + nf, vec_offset + param.vector_size(), jac


def offset_sparse_matrix(spmat, offset_shape, full_shape):
    spmat = coo_matrix(spmat)
    return coo_matrix((spmat.data, (spmat.row + offset_shape[0], spmat.col + offset_shape[1])),
                      shape=full_shape)


def free_to_vector_hess_offset(param, free_vec, hessians, free_offset, full_shape):
    nf = param.free_size()
    for h in param.free_to_vector_hess(free_vec[free_offset:free_offset + nf]):
        hessians.append(offset_sparse_matrix(h, (free_offset, free_offset), full_shape))
    return free_offset + nf


def convert_vector_to_free_hessian(param, free_val, vector_grad, vector_hess):
    """H_free = J^T H_vec J + sum_i grad_vec[i] * d2 vec_i / d free2   (Parameters.py:397-424).

    ``vector_hess`` may be dense or scipy-sparse; the result is sparse (CSR) when it is sparse.
    The CUDA path applies the same rule for the GLMM's diagonal transforms inside the kernels
    (csrc/glmm_eval.cu k_local / k_border / k_gram_finish / k_global)."""
    import scipy.sparse as sps
    param.set_free(copy.deepcopy(free_val))
    jac = sps.csr_matrix(param.free_to_vector_jac(free_val))
    param.set_free(copy.deepcopy(free_val))
    second = param.free_to_vector_hess(free_val)
    n = param.free_size()
    rows, cols, vals = [np.zeros(0, dtype=np.int64)], [np.zeros(0, dtype=np.int64)], [np.zeros(0)]
    for i, h in enumerate(second):
        h = coo_matrix(h)
        rows.append(h.row)
        cols.append(h.col)
        vals.append(h.data * vector_grad[i])
    curvature = coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                           (n, n))
    if sps.issparse(vector_hess):
        return sps.csr_matrix(curvature + jac.T @ sps.csr_matrix(vector_hess) @ jac)
    return curvature.toarray() + jac.T @ (np.asarray(vector_hess) @ jac.toarray())
