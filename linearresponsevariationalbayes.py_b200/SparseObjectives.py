"""Objective wrappers: the differentiation boundary of the library, backed by the CUDA path.

Mirror of /root/reference/LinearResponseVariationalBayes/SparseObjectives.py: ``Objective``
(:95-240, same method names and argument orders, including the HVP convention "vector is the
LAST positional argument", :183-193, and the preconditioned ``*_cond`` family, :202-240), the
sparse-Hessian helpers (:581-634), ``Logger`` / ``Timer`` (:35-87), ``safe_matmul`` (:21-25).

The reference builds its derivative callables with autograd over an arbitrary Python ``fun``.
Here ``Objective(par, model)`` takes a *model object* (``GLMM.LogisticGLMM`` or
``distributed.ShardedLogisticGLMM``) whose value / gradient / Hessian / HVP are hand-written
sm_100a kernels; a bare Python callable cannot be differentiated by this library and is
rejected -- there is no CPU or autodiff fallback.

Return kinds: numpy in -> Python float / numpy array / ``scipy.sparse.csr_matrix`` out;
CUDA torch tensor in -> torch tensors / ``GLMM.DeviceCSR`` out (nothing leaves the device).
"""
import time
from copy import deepcopy

import numpy as np
import scipy as sp
import scipy.sparse

from ._tensors import is_torch


def safe_matmul(x, y):
    """:21-25 -- sparse-aware product."""
    if sp.sparse.issparse(x) or sp.sparse.issparse(y):
        return x @ y
    if is_torch(x) or is_torch(y):
        import torch
        return torch.matmul(x, y)
    return np.matmul(x, y)


def compress(x):
    if sp.sparse.issparse(x):
        return np.squeeze(np.asarray(x.todense()))
    return np.squeeze(np.asarray(x))


class Timer(object):
    """Wall-clock dictionary (:35-45)."""

    def __init__(self):
        self.time_dict = {}

    def tic(self):
        self.tic_time = time.time()

    def toc(self, time_name, verbose=True):
        self.time_dict[time_name] = time.time() - self.tic_time
        if verbose:
            print("{}: {} seconds".format(time_name, self.time_dict[time_name]))

    def __str__(self):
        return str(self.time_dict)


class Logger(object):
    """Iteration log used by ``fun_free(..., verbose=True)`` (:48-87)."""

    def __init__(self, print_every=1):
        self.print_every = print_every
        self.print_x_diff = True
        self.callback = None
        self.initialize()

    def initialize(self):
        self.iter = 0
        self.last_x = self.x = None
        self.value = self.last_value = None
        self.x_array, self.val_array = [], []

    def print_message(self):
        print("Iter ", self.iter, " value: ", self.value)

    def log(self, value, x):
        self.value, self.x = value, x
        self.x_array.append(x)
        self.val_array.append(value)
        self.last_x, self.last_value = x, value
        if self.iter % self.print_every == 0:
            if self.callback is None:
                self.print_message()
            else:
                self.callback(self)
        self.iter += 1


def _host(x):
    return x.detach().cpu().numpy() if is_torch(x) else np.asarray(x, dtype=np.float64)


class Objective(object):
    """``Objective(par, model)``: value / gradient / Hessian / HVP in free or vector coordinates.

    ``par`` is the model's ModelParamsDict (``model.glmm_par``); after every call it holds plain
    numeric values equal to the evaluation point, as in the reference (:142-150).
    """

    def __init__(self, par, fun):
        if not getattr(fun, "_lrvb_device_model", False):
            raise TypeError(
                "lrvb_b200.Objective differentiates device models only (e.g. GLMM.LogisticGLMM); "
                "got %r. Arbitrary Python callables need autograd, which this CUDA library "
                "does not provide (no CPU fallback)." % (fun,))
        self._par = par
        self._pending = None        # (device copy of the point, coords): par is filled in on first access
        self.fun = fun
        self.model = fun
        self.preconditioner = None
        self.logger = Logger()
        self._par_key, self._par_coords = None, None

    @property
    def par(self):
        """The ModelParamsDict, holding the last evaluation point (:142-150).  For CUDA-tensor input the point
        stays on the device (a device-side copy is kept) and is brought to the host only when ``par`` is
        read: an optimiser or CG loop over device tensors never synchronises for it."""
        if self._pending is not None:
            x, coords = self._pending
            self._pending = None
            xh = _host(x).reshape(-1)
            if coords == "free":
                self._par.set_free(xh)
            else:
                self._par.set_vector(xh)
        return self._par

    @par.setter
    def par(self, value):
        self._par = value
        self._pending = None

    def invalidate(self):
        """Forget the cached evaluation.  The "already evaluated at this point?" check compares numpy input
        by value, but a torch tensor by identity (storage address, shape, strides and torch's in-place
        version counter, ``_tensors.PointKey``): a write that bypasses the version counter -- a raw-pointer
        write by other CUDA code, ``x.data`` updates, a CPU tensor changed through its shared ``.numpy()``
        view -- is invisible to it.  Call this (or pass a fresh tensor) after such a write."""
        self.model.invalidate()
        self._par_key, self._par_coords = None, None

    # ---- helpers ----
    def _set_par(self, x, coords):
        # ``par`` follows the evaluation point (:142-150).  The model keeps one key object per
        # evaluated point; while that object is unchanged ``par`` already holds this point (value,
        # gradient and Hessian requested at the same x set it once, not three times).
        key = getattr(self.model, "_cache", {}).get("x")
        if key is not None and self._par_key is key and self._par_coords == coords:
            return
        if is_torch(x) and x.is_cuda:
            self._pending = (x.detach().clone(), coords)
        else:
            self._pending = None
            xh = _host(x).reshape(-1)
            if coords == "free":
                self._par.set_free(xh)
            else:
                self._par.set_vector(xh)
        self._par_key, self._par_coords = key, coords

    def _value(self, x, coords):
        self.model.evaluate(x, 0, coords)
        self._set_par(x, coords)
        if is_torch(x):
            return self.model.kl_tensor().clone()
        if hasattr(self.model, "kl_host"):
            return self.model.kl_host()
        return float(self.model.kl_tensor().item())

    def _grad(self, x, coords):
        self.model.evaluate(x, 1, coords)
        self._set_par(x, coords)
        if is_torch(x):
            return self.model.grad_tensor()
        if hasattr(self.model, "grad_host"):
            return self.model.grad_host()
        return self.model.grad_tensor().cpu().numpy()

    def _hessian(self, x, coords):
        self.model.evaluate(x, 2, coords)
        self._set_par(x, coords)
        if is_torch(x):
            return self.model.hessian_csr()      # device CSR (a sharded model: this rank's part)
        if hasattr(self.model, "hessian_csr_global"):
            return self.model.hessian_csr_global()   # sharded model: gather the full matrix
        if hasattr(self.model, "hessian_scipy"):
            return self.model.hessian_scipy()
        return self.model.hessian_csr().to_scipy()

    def _hvp(self, x, vec, coords):
        self.model.evaluate(x, 2, coords)
        self._set_par(x, coords)
        out = self.model.hvp(vec)
        return out if is_torch(vec) else out.cpu().numpy()

    @staticmethod
    def _no_extra(argv, argk):
        if argv or argk:
            raise TypeError("the GLMM objective takes no extra arguments")

    # ---- free coordinates (:120-125, :152-158, :183-187) ----
    def fun_free(self, free_val, *argv, verbose=False, **argk):
        self._no_extra(argv, argk)
        val = self._value(free_val, "free")
        if verbose:
            self.logger.log(val, free_val)
        return val

    def fun_free_grad(self, free_val, *argv, **argk):
        self._no_extra(argv, argk)
        return self._grad(free_val, "free")

    def fun_free_hessian(self, free_val, *argv, **argk):
        self._no_extra(argv, argk)
        return self._hessian(free_val, "free")

    def fun_free_jacobian(self, free_val, *argv, **argk):
        # the objective is scalar: its Jacobian is the gradient (:160-162)
        return self.fun_free_grad(free_val, *argv, **argk)

    def fun_free_hvp(self, *argv, **argk):
        args, vec = argv[:-1], argv[-1]
        self._no_extra(args[1:], argk)
        return self._hvp(args[0], vec, "free")

    # ---- vector coordinates (:127-129, :164-181, :189-193) ----
    def fun_vector(self, vec_val, *argv, **argk):
        self._no_extra(argv, argk)
        return self._value(vec_val, "vector")

    def fun_vector_grad(self, vec_val, *argv, **argk):
        self._no_extra(argv, argk)
        return self._grad(vec_val, "vector")

    def fun_vector_hessian(self, vec_val, *argv, **argk):
        self._no_extra(argv, argk)
        return self._hessian(vec_val, "vector")

    def fun_vector_jacobian(self, vec_val, *argv, **argk):
        return self.fun_vector_grad(vec_val, *argv, **argk)

    def fun_vector_hvp(self, *argv, **argk):
        args, vec = argv[:-1], argv[-1]
        self._no_extra(args[1:], argk)
        return self._hvp(args[0], vec, "vector")

    # ---- preconditioned variants: free_val = P x  (:202-240) ----
    def get_conditioned_x(self, free_val):
        return safe_matmul(self.preconditioner, free_val)

    def fun_free_cond(self, free_val, *argv, verbose=False, **argk):
        assert self.preconditioner is not None
        y = self.get_conditioned_x(free_val)
        return self.fun_free(y, *argv, verbose=verbose, **argk)

    def fun_free_grad_cond(self, free_val, *argv, **argk):
        assert self.preconditioner is not None
        y = self.get_conditioned_x(free_val)
        return safe_matmul(self.preconditioner.T, self.fun_free_grad(y, *argv, **argk))

    def fun_free_hessian_cond(self, free_val, *argv, **argk):
        assert self.preconditioner is not None
        y = self.get_conditioned_x(free_val)
        hess = self.fun_free_hessian(y, *argv, **argk)
        return safe_matmul(self.preconditioner.T, safe_matmul(hess, self.preconditioner))

    def fun_free_hvp_cond(self, *argv, **argk):
        assert self.preconditioner is not None
        args, vec = argv[1:-1], argv[-1]
        y = self.get_conditioned_x(argv[0])
        return safe_matmul(
            self.preconditioner.T,
            self.fun_free_hvp(y, *args, safe_matmul(self.preconditioner, vec), **argk))

    def uncondition_x(self, cond_x):
        return safe_matmul(self.preconditioner, cond_x)


# ---- sparse-Hessian helpers (:581-634) ----------------------------------------------------------

def make_index_param(param):
    """A copy of ``param`` whose entries hold their own positions in the flat vector (:581-584)."""
    index_param = deepcopy(param)
    index_param.set_vector(np.arange(0, index_param.vector_size()))
    return index_param


def get_sparse_sub_matrix(sub_matrix, row_indices, col_indices, row_dim, col_dim):
    """Places the EXACTLY-nonzero entries of a dense block at (row_indices x col_indices) of a
    (row_dim, col_dim) CSR matrix (:604-619; duplicates sum, columns sorted, int32).
    Vectorised host helper; the GLMM Hessian is assembled on the device instead
    (csrc/csr.cu), with the same canonical form."""
    sub_matrix = np.asarray(sub_matrix)
    r, c = np.nonzero(sub_matrix)
    rows = np.asarray(row_indices)[r].astype(np.int64)
    cols = np.asarray(col_indices)[c].astype(np.int64)
    return sp.sparse.csr_matrix((sub_matrix[r, c], (rows, cols)), (row_dim, col_dim))


def get_sparse_sub_hessian(sub_hessian, full_indices, full_hess_dim):
    """:591-597."""
    return get_sparse_sub_matrix(sub_hessian, full_indices, full_indices,
                                 full_hess_dim, full_hess_dim)


def pack_csr_matrix(sp_mat):
    """:624-629."""
    sp_mat = sp.sparse.csr_matrix(sp_mat)
    return {"data": sp_mat.data, "indices": sp_mat.indices, "indptr": sp_mat.indptr,
            "shape": sp_mat.shape}


def unpack_csr_matrix(sp_mat_dict):
    """:631-634."""
    return sp.sparse.csr_matrix(
        (sp_mat_dict["data"], sp_mat_dict["indices"], sp_mat_dict["indptr"]),
        shape=sp_mat_dict["shape"])


# ---- JSON wire format of a CSR matrix (:639-657) -------------------------------------------------
# The reference serialises the three arrays with ``json_tricks.dumps`` (an optional dependency that
# is not needed here): each becomes the JSON text of
#   {"__ndarray__": [...], "dtype": "float64", "shape": [n], "Corder": true}
# These two helpers write and read exactly that text with the standard library, so dictionaries
# produced by either code base can be read by the other.
def _ndarray_to_json(a):
    import json
    a = np.ascontiguousarray(a)
    return json.dumps({"__ndarray__": a.tolist(), "dtype": str(a.dtype), "shape": list(a.shape),
                       "Corder": True})


def _ndarray_from_json(text):
    import json
    d = json.loads(text) if isinstance(text, str) else text
    if isinstance(d, dict) and "__ndarray__" in d:
        a = np.asarray(d["__ndarray__"], dtype=np.dtype(d.get("dtype", "float64")))
        return a.reshape(d["shape"]) if "shape" in d else a
    return np.asarray(d)


def json_pack_csr_matrix(sp_mat):
    """:639-647 -- a json-serialisable dict of a CSR matrix (host scipy matrix or DeviceCSR)."""
    if hasattr(sp_mat, "to_scipy"):
        sp_mat = sp_mat.to_scipy()
    assert sp.sparse.isspmatrix_csr(sp_mat)
    sp_mat = sp.sparse.csr_matrix(sp_mat)
    return {"data": _ndarray_to_json(sp_mat.data), "indices": _ndarray_to_json(sp_mat.indices),
            "indptr": _ndarray_to_json(sp_mat.indptr), "shape": tuple(int(v) for v in sp_mat.shape),
            "type": "csr_matrix"}


def json_unpack_csr_matrix(sp_mat_dict):
    """:651-657 -- the inverse of ``json_pack_csr_matrix``."""
    assert sp_mat_dict["type"] == "csr_matrix"
    data = _ndarray_from_json(sp_mat_dict["data"])
    indices = _ndarray_from_json(sp_mat_dict["indices"])
    indptr = _ndarray_from_json(sp_mat_dict["indptr"])
    return sp.sparse.csr_matrix((data, indices, indptr), shape=tuple(sp_mat_dict["shape"]))

