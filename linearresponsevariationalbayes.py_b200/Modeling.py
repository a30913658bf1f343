"""Gauss-Hermite expectation of the logistic log-partition term on the device.

Mirror of /root/reference/LinearResponseVariationalBayes/Modeling.py:35-52
(``get_e_logistic_term_guass_hermite`` -- the reference's spelling is kept).  ``gh_x, gh_w`` are
``numpy.polynomial.hermite.hermgauss(Q)`` nodes and weights supplied by the caller, as in the
reference.  The function is evaluated by csrc/ef.cu ``k_gh_logistic`` in the stable form
``max(t,0) + log1p(exp(-|t|))`` -- equal to the reference's ``log1p(exp(t))`` (:48) on its finite
range and finite beyond it.
"""
import numpy as np

from . import _native as nat
from ._tensors import like_input, to_device
from .ExponentialFamilies import _dsum


def get_e_logistic_term_guass_hermite(z_mean, z_sd, gh_x, gh_w, aggregate_all=True):
    zm, zs = to_device(z_mean), to_device(z_sd)
    assert zm.shape == zs.shape  # Modeling.py:38
    torch = nat.require_cuda()
    gx = np.ascontiguousarray(np.asarray(gh_x, dtype=np.float64))
    gw = np.ascontiguousarray(np.asarray(gh_w, dtype=np.float64))
    assert gx.shape == gw.shape and gx.ndim == 1
    out = torch.empty(zm.numel(), dtype=torch.float64, device=zm.device)
    nat.check(nat.load().lrvb_gh_logistic_term(
        nat.ptr(zm.reshape(-1)), nat.ptr(zs.reshape(-1)), zm.numel(), nat.darray(gx),
        nat.darray(gw), gx.size, nat.ptr(out), nat.stream_ptr()))
    if aggregate_all:
        return like_input(_dsum(out), z_mean, z_sd)
    return like_input(out.reshape(zm.shape), z_mean, z_sd)


get_e_logistic_term_gauss_hermite = get_e_logistic_term_guass_hermite
