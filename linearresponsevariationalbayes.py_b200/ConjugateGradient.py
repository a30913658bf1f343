"""H^{-1} v by conjugate gradient, with the iteration on the device for GLMM objectives.

Mirror of /root/reference/LinearResponseVariationalBayes/ConjugateGradient.py: mask helpers
(:19-60) and ``ConjugateGradientSolver`` (:63-105: ``get_hinv_vec`` ->
``scipy.sparse.linalg.cg(A, b, x0, tol, M)``; ``get_hinv_vec_subsets`` one solve per mask,
recording vecs / hinv_vecs / masks / times / cg_infos).

``ConjugateGradientSolver(objective.fun_free_hvp, x0)`` (the reference's call,
test_objectives.py:541) is recognised when the callable is a bound method of a device
``Objective``: the Hessian blocks are evaluated once at ``x0`` and the whole CG loop (HVP, dots,
axpys, convergence test; csrc/solve.cu) runs on the GPU with scipy's stopping rule
``||r|| < tol * ||b||``.  ``tol`` keeps the reference attribute name; scipy >= 1.14 spells it
``rtol`` (SURVEY.md 8b "CG API drift").  ``preconditioner`` is scipy's ``M=`` (:84): None, a
scipy sparse matrix or a dense array (applied on the device by SpMV / GEMV), or the strings
``"block_jacobi"`` (inverse of the block diagonal of H) and ``"schur"`` (H^-1 itself through the
Schur complement of the local blocks); for a generic Python HVP callable (host
control path, exactly the reference's scipy call) anything scipy accepts.
"""
import time

import numpy as np
import scipy as sp
import scipy.sparse.linalg
from scipy.sparse.linalg import LinearOperator

from ._tensors import is_torch


def split_vector(vec):
    """Two boolean vectors holding roughly half of the True entries each (:19-31)."""
    vec = np.asarray(vec)
    a = np.full(len(vec), False)
    b = np.full(len(vec), False)
    true_inds = np.flatnonzero(vec)
    half = int(len(true_inds) / 2)
    a[true_inds[:half]] = True
    b[true_inds[half:]] = True
    return a, b


def recursive_split(mask, results=None, terminate_len=10):
    """Split ``mask`` until every piece has at most ``terminate_len`` True values (:36-43).
    (The reference's mutable default ``results=[]`` is not reproduced.)"""
    if results is None:
        results = []
    if np.sum(mask) > terminate_len:
        m1, m2 = split_vector(mask)
        recursive_split(m1, results=results, terminate_len=terminate_len)
        recursive_split(m2, results=results, terminate_len=terminate_len)
    else:
        results.append(mask)
    return results


def get_masks(full_len, min_mask_len):
    """Contiguous boolean masks of length ``min_mask_len`` partitioning range(full_len) (:46-57)."""
    assert min_mask_len > 0
    assert min_mask_len < full_len
    masks = []
    for start in range(0, full_len, min_mask_len):
        mask = np.full(full_len, False)
        mask[start:min(start + min_mask_len, full_len)] = True
        masks.append(mask)
    return masks


def _device_objective_of(hvp_callable):
    obj = getattr(hvp_callable, "__self__", None)
    name = getattr(hvp_callable, "__name__", "")
    if obj is not None and getattr(getattr(obj, "model", None), "_lrvb_device_model", False) \
            and name in ("fun_free_hvp", "fun_vector_hvp"):
        return obj, ("free" if name == "fun_free_hvp" else "vector")
    return None, None


class ConjugateGradientSolver(object):
    def __init__(self, eval_hessian_vector_product, x0):
        self.dim = len(x0)
        self.x0 = x0
        self._objective, self._coords = _device_objective_of(eval_hessian_vector_product)
        if self._objective is None:
            # generic callable: the reference's own scipy path (host control flow)
            self.ObjHessVecProdLO = LinearOperator(
                (self.dim, self.dim), lambda vec: eval_hessian_vector_product(x0, vec))
        else:
            self.ObjHessVecProdLO = LinearOperator(
                (self.dim, self.dim), lambda vec: eval_hessian_vector_product(x0, vec))
        self.preconditioner = None
        self.tol = 1e-8
        self.maxiter = None
        self.last_iterations = None
        self.initialize()

    def initialize(self):
        self.vecs = []
        self.hinv_vecs = []
        self.masks = []
        self.times = []
        self.cg_infos = []

    def get_hinv_vec(self, vec, x0=None):
        """Returns (H^{-1} vec, info); info 0 = converged, > 0 = iteration limit (scipy)."""
        if self._objective is not None:
            return self._device_solve(vec, x0)
        hinv_vec, cg_info = sp.sparse.linalg.cg(
            self.ObjHessVecProdLO, vec, x0=x0, rtol=self.tol, atol=0.0, M=self.preconditioner,
            maxiter=self.maxiter)
        return hinv_vec, cg_info

    def _device_solve(self, vec, x0):
        model = self._objective.model
        # None, "block_jacobi", "schur", or any matrix scipy's M= accepts (sparse / dense): the model
        # turns it into a device operator (GLMM.make_preconditioner); a sharded model takes the
        # first two
        precond = 0 if self.preconditioner is None else self.preconditioner
        model.evaluate(self.x0, 2, self._coords)  # cached after the first solve
        n = vec.numel() if is_torch(vec) else np.asarray(vec).size
        if n != self.dim:
            raise ValueError("Wrong size for CG right-hand side.  Expected {}, got {}".format(
                self.dim, n))
        x, info, iters = model.cg(vec, x0, precond=precond, rtol=self.tol,
                                  maxiter=self.maxiter or 0)
        self.last_iterations = iters
        return (x if is_torch(vec) else x.cpu().numpy()), info

    def get_hinv_vec_subsets(self, vec, masks, verbose=False, print_every=10):
        """One solve per boolean mask with the unmasked entries zeroed (:89-105)."""
        num_masks = len(masks)
        vec = np.asarray(vec)
        for ind, mask in enumerate(masks, start=1):
            if verbose and ind % print_every == 0:
                print("{} of {}\n".format(ind, num_masks))
            vec_masked = np.zeros(len(vec))
            vec_masked[mask] = vec[mask]
            tic = time.time()
            hinv_vec, cg_info = self.get_hinv_vec(vec_masked)
            self.times.append(time.time() - tic)
            self.vecs.append(vec_masked)
            self.masks.append(mask)
            self.hinv_vecs.append(hinv_vec)
            self.cg_infos.append(cg_info)
