"""Exponential-family entropies and expectations, evaluated by the batched CUDA kernels of
csrc/ef.cu through the C ABI (``lrvb_ef_*``).

Same names, argument meaning and aggregation as
/root/reference/LinearResponseVariationalBayes/ExponentialFamilies.py:5-120 and :186-204 (which
functions sum over factors and which return one value per factor follows the reference, e.g.
``gamma_entropy`` sums (:33-35) while ``dirichlet_entropy`` does not (:43-52)).  Inputs may be
numpy arrays / Python scalars (results come back as numpy / float) or CUDA torch tensors (results
stay on the device).  ``*_batched`` variants return the per-factor values.
"""
import math

import numpy as np

from . import _native as nat
from ._tensors import is_torch, like_input, to_device


def _empty(n, like):
    torch = nat.require_cuda()
    return torch.empty(n, dtype=torch.float64, device=like.device)


def _dsum(t):
    torch = nat.require_cuda()
    out = torch.empty((), dtype=torch.float64, device=t.device)
    nat.check(nat.load().lrvb_sum(nat.ptr(t), t.numel(), nat.ptr(out), nat.stream_ptr()))
    return out


# ---- multivariate special functions (:5-13) ------------------------------------------------

def multivariate_digamma(x, size):
    xs = np.asarray(x, dtype=np.float64) - 0.5 * np.linspace(0, size - 1.0, int(size))
    # sum_j digamma(x - j/2) = sum_j E[log Gamma(x - j/2, 1)]
    d = to_device(xs)
    out = _empty(d.numel(), d)
    nat.check(nat.load().lrvb_ef_e_log_gamma(nat.ptr(d), nat.ptr(to_device(np.ones(d.numel()))),
                                             d.numel(), nat.ptr(out), nat.stream_ptr()))
    return like_input(_dsum(out), x)


def multivariate_gammaln(x, size):
    # through the Wishart kernel identity is not needed: lgamma is in libm; keep the reference's
    # closed form on the host for this scalar helper (:10-13)
    xs = np.asarray(x, dtype=np.float64) - 0.5 * np.linspace(0, size - 1.0, int(size))
    return float(np.sum([math.lgamma(v) for v in np.atleast_1d(xs)])
                 + 0.25 * np.log(np.pi) * size * (size - 1.0))


# ---- entropies ---------------------------------------------------------------------------------

def multinoulli_entropy(p, min_prob=1e-16):
    """:20-21; p (M, d) -> (M,)."""
    d = to_device(p)
    assert d.dim() == 2
    out = _empty(d.shape[0], d)
    nat.check(nat.load().lrvb_ef_multinoulli_entropy(nat.ptr(d), d.shape[1], d.shape[0],
                                                     float(min_prob), nat.ptr(out),
                                                     nat.stream_ptr()))
    return like_input(out, p)


def univariate_normal_entropy_batched(info_obs):
    d = to_device(info_obs).reshape(-1)
    out = _empty(d.numel(), d)
    nat.check(nat.load().lrvb_ef_uvn_entropy(nat.ptr(d), d.numel(), nat.ptr(out), nat.stream_ptr()))
    return like_input(out, info_obs)


def univariate_normal_entropy(info_obs):
    """:23-25 (sums over factors)."""
    d = to_device(info_obs).reshape(-1)
    out = _empty(d.numel(), d)
    nat.check(nat.load().lrvb_ef_uvn_entropy(nat.ptr(d), d.numel(), nat.ptr(out), nat.stream_ptr()))
    return like_input(_dsum(out), info_obs)


def multivariate_normal_entropy(info_obs):
    """:27-31, via the batched Wishart kernel's Cholesky log-determinant."""
    v = to_device(info_obs)
    k = v.shape[0]
    assert v.shape == (k, k)
    torch = nat.require_cuda()
    df = torch.full((1,), float(k + 1), dtype=torch.float64, device=v.device)
    eld = _empty(1, v)
    nat.check(nat.load().lrvb_ef_wishart(nat.ptr(df), nat.ptr(v), k, 1, nat.ptr(None), nat.ptr(eld),
                                         nat.ptr(None), nat.stream_ptr()))
    # e_log_det = mv_digamma(df/2, k) + k log 2 + log det v
    mvd = multivariate_digamma(0.5 * (k + 1), k)
    logdet = eld[0] - mvd - k * math.log(2.0)
    ent = 0.5 * (-1 * logdet + k + k * math.log(2 * math.pi))
    return like_input(ent, info_obs)


def gamma_entropy_batched(shape, rate):
    a, b = to_device(shape).reshape(-1), to_device(rate).reshape(-1)
    assert a.numel() == b.numel()
    out = _empty(a.numel(), a)
    nat.check(nat.load().lrvb_ef_gamma_entropy(nat.ptr(a), nat.ptr(b), a.numel(), nat.ptr(out),
                                               nat.stream_ptr()))
    return like_input(out, shape, rate)


def gamma_entropy_and_e_log(shape, rate):
    """:33-35 and :111-112 per factor from ONE launch (they share digamma(shape) and log(rate)):
    returns (entropy, e_log), both shaped like ``shape``."""
    a, b = to_device(shape), to_device(rate)
    shp = a.shape
    assert a.numel() == b.numel()
    ent, el = _empty(a.numel(), a), _empty(a.numel(), a)
    nat.check(nat.load().lrvb_ef_gamma_terms(nat.ptr(a.reshape(-1)), nat.ptr(b.reshape(-1)), a.numel(),
                                             nat.ptr(ent), nat.ptr(el), nat.stream_ptr()))
    return like_input(ent.reshape(shp), shape, rate), like_input(el.reshape(shp), shape, rate)


def dirichlet_entropy_and_e_log(alpha):
    """:43-52 and :118-120 from ONE launch; alpha (d, ...): returns (entropy (...), e_log (d, ...))."""
    d = to_device(alpha)
    dim, rest = d.shape[0], tuple(d.shape[1:])
    flat = d.reshape(dim, -1).contiguous()
    M = flat.shape[1]
    ent, el = _empty(M, d), _empty(dim * M, d)
    nat.check(nat.load().lrvb_ef_dirichlet_terms(nat.ptr(flat), dim, M, nat.ptr(ent), nat.ptr(el),
                                                 nat.stream_ptr()))
    return like_input(ent.reshape(rest), alpha), like_input(el.reshape(d.shape), alpha)


def gamma_entropy(shape, rate):
    """:33-35 (sums over factors)."""
    a, b = to_device(shape).reshape(-1), to_device(rate).reshape(-1)
    assert a.numel() == b.numel()
    out = _empty(a.numel(), a)
    nat.check(nat.load().lrvb_ef_gamma_entropy(nat.ptr(a), nat.ptr(b), a.numel(), nat.ptr(out),
                                               nat.stream_ptr()))
    return like_input(_dsum(out), shape, rate)


def dirichlet_entropy(alpha):
    """:43-52; alpha (d, ...) with the simplex dimension on axis 0 -> entropies of shape (...)."""
    d = to_device(alpha)
    dim, rest = d.shape[0], tuple(d.shape[1:])
    flat = d.reshape(dim, -1).contiguous()
    M = flat.shape[1]
    out = _empty(M, d)
    nat.check(nat.load().lrvb_ef_dirichlet_entropy(nat.ptr(flat), dim, M, nat.ptr(out),
                                                   nat.stream_ptr()))
    return like_input(out.reshape(rest), alpha)


def beta_entropy_batched(tau):
    d = to_device(tau)
    assert d.dim() == 2 and d.shape[1] == 2
    out = _empty(d.shape[0], d)
    nat.check(nat.load().lrvb_ef_beta_entropy(nat.ptr(d), d.shape[0], nat.ptr(out), nat.stream_ptr()))
    return like_input(out, tau)


def beta_entropy(tau):
    """:54-69 (sums over rows of tau (M,2))."""
    d = to_device(tau)
    assert d.dim() == 2 and d.shape[1] == 2
    out = _empty(d.shape[0], d)
    nat.check(nat.load().lrvb_ef_beta_entropy(nat.ptr(d), d.shape[0], nat.ptr(out), nat.stream_ptr()))
    return like_input(_dsum(out), tau)


def _wishart(df, v, want):
    vd = to_device(v)
    if vd.dim() == 2:
        vd = vd.unsqueeze(0)
    M, k = vd.shape[0], vd.shape[1]
    assert vd.shape[2] == k
    dfd = to_device(df).reshape(-1)
    assert dfd.numel() == M
    ent = _empty(M, vd) if want == "entropy" else None
    eld = _empty(M, vd) if want == "e_log_det" else None
    eid = _empty(M * k, vd) if want == "e_log_inv_diag" else None
    nat.check(nat.load().lrvb_ef_wishart(nat.ptr(dfd), nat.ptr(vd), k, M, nat.ptr(ent), nat.ptr(eld),
                                         nat.ptr(eid), nat.stream_ptr()))
    if want == "e_log_inv_diag":
        return eid.reshape(M, k)
    return ent if want == "entropy" else eld


def _squeeze_single(res, v):
    single = (np.ndim(v) if not is_torch(v) else v.dim()) == 2
    return res[0] if single else res


def wishart_entropy(df, v):
    """:72-82; v (k,k) with scalar df as in the reference, or batched v (M,k,k), df (M,)."""
    return like_input(_squeeze_single(_wishart(df, v, "entropy"), v), df, v)


def e_log_det_wishart(df, v):
    """:88-94."""
    return like_input(_squeeze_single(_wishart(df, v, "e_log_det"), v), df, v)


def e_log_inv_wishart_diag(df, v):
    """:97-102."""
    return like_input(_squeeze_single(_wishart(df, v, "e_log_inv_diag"), v), df, v)


# ---- expectations --------------------------------------------------------------------------------

def _xp(x):
    if is_torch(x):
        import torch
        return torch
    return np


def get_e_lognormal(mu, sigma_sq):
    """:104-105 (plain elementwise; stays in the caller's array library)."""
    return _xp(mu).exp(mu + 0.5 * sigma_sq)


def get_var_lognormal(mu, sigma_sq):
    """:107-109."""
    e = get_e_lognormal(mu, sigma_sq)
    return (_xp(sigma_sq).exp(sigma_sq) - 1) * (e ** 2)


def get_e_log_gamma(shape, rate):
    """:111-112, elementwise."""
    a, b = to_device(shape), to_device(rate)
    shp = a.shape
    out = _empty(a.numel(), a)
    nat.check(nat.load().lrvb_ef_e_log_gamma(nat.ptr(a.reshape(-1)), nat.ptr(b.reshape(-1)),
                                             a.numel(), nat.ptr(out), nat.stream_ptr()))
    return like_input(out.reshape(shp), shape, rate)


def get_e_dirichlet(alpha):
    """:114-116."""
    if is_torch(alpha):
        return alpha / alpha.sum(0, keepdim=True)
    alpha = np.asarray(alpha)
    return alpha / np.sum(alpha, 0, keepdims=True)


def get_e_log_dirichlet(alpha):
    """:118-120."""
    d = to_device(alpha)
    dim = d.shape[0]
    flat = d.reshape(dim, -1).contiguous()
    M = flat.shape[1]
    out = _empty(dim * M, d)
    nat.check(nat.load().lrvb_ef_e_log_dirichlet(nat.ptr(flat), dim, M, nat.ptr(out),
                                                 nat.stream_ptr()))
    return like_input(out.reshape(d.shape), alpha)


# ---- priors (:186-195): closed-form arithmetic on expectations -----------------------------------

def mvn_prior(prior_mean, prior_info, e_obs, cov_obs):
    obs_diff = e_obs - prior_mean
    xp = _xp(e_obs)
    return -0.5 * (xp.dot(obs_diff, xp.matmul(prior_info, obs_diff)) + xp.trace(
        xp.matmul(prior_info, cov_obs)))


def uvn_prior(prior_mean, prior_info, e_obs, var_obs):
    return -0.5 * (prior_info * ((e_obs - prior_mean) ** 2 + var_obs))


def gamma_prior(prior_shape, prior_rate, e_obs, e_log_obs):
    return (prior_shape - 1) * e_log_obs - prior_rate * e_obs


def exponential_prior(lambda_par, e_obs):
    """:197-198."""
    return -1 * lambda_par * e_obs


def dirichlet_prior(alpha, log_e_obs):
    """:200-204 (the argument is E log of the observation, as the reference's own comment says)."""
    xp = _xp(alpha)
    assert tuple(alpha.shape) == tuple(log_e_obs.shape), "shape of alpha and log_e_obs do not match"
    return xp.dot(alpha - 1, log_e_obs) if xp is np else ((alpha - 1) * log_e_obs).sum()


def expected_ljk_prior(lkj_param, df, v):
    """:206-211: if Sigma^-1 ~ Wishart(v, df), the expected LKJ prior (lkj_param - 1) E log |R|."""
    e_log_r = -1 * e_log_det_wishart(df, v) - e_log_inv_wishart_diag(df, v).sum()
    return (lkj_param - 1) * e_log_r


# ---- numeric integration (:123-172) --------------------------------------------------------------
def get_e_fun_normal(means, infos, gh_loc, gh_weights, fun):
    """E fun(X), X an array of normals (means, infos), by Gauss-Hermite quadrature (:125-141); ``fun`` is any
    elementwise callable of the caller's array library (numpy or torch)."""
    xp = _xp(means)
    assert tuple(means.shape) == tuple(infos.shape)
    if xp is np:
        loc = np.sqrt(2) * np.asarray(gh_loc) / np.sqrt(np.expand_dims(infos, means.ndim)) \
            + np.expand_dims(means, means.ndim)
        return np.sum(1 / np.sqrt(np.pi) * np.asarray(gh_weights) * fun(loc), axis=means.ndim)
    import torch
    gl = torch.as_tensor(gh_loc, dtype=means.dtype, device=means.device)
    gw = torch.as_tensor(gh_weights, dtype=means.dtype, device=means.device)
    loc = math.sqrt(2) * gl / torch.sqrt(infos.unsqueeze(-1)) + means.unsqueeze(-1)
    return torch.sum(1 / math.sqrt(math.pi) * gw * fun(loc), dim=-1)


def get_e_logitnormal(lognorm_means, lognorm_infos, gh_loc, gh_weights):
    """:143-148."""
    xp = _xp(lognorm_means)
    expit = (lambda x: 1.0 / (1.0 + np.exp(-x))) if xp is np else (lambda x: xp.sigmoid(x))
    return get_e_fun_normal(lognorm_means, lognorm_infos, gh_loc, gh_weights, expit)


def get_e_log_logitnormal(lognorm_means, lognorm_infos, gh_loc, gh_weights):
    """:150-172: (E log X, E log(1 - X)) for a logit-normal X.  log(expit(v)) is evaluated as
    min(v, 0) - log1p(exp(-|v|)): the same function as the reference's guarded two-branch expression
    (-1e16 / x <= -100 cut-offs) wherever that one is finite, and finite everywhere."""
    xp = _xp(lognorm_means)
    if xp is np:
        log_v = lambda x: np.minimum(x, 0.0) - np.log1p(np.exp(-np.abs(x)))   # noqa: E731
    else:
        log_v = lambda x: xp.clamp(x, max=0.0) - xp.log1p(xp.exp(-xp.abs(x)))   # noqa: E731
    e_log_v = get_e_fun_normal(lognorm_means, lognorm_infos, gh_loc, gh_weights, log_v)
    return e_log_v, -lognorm_means + e_log_v


def get_uvn_from_natural_parameters(e_term, e2_term):
    """:179-182: mean and info of the normal with log p(x) = e_term x + e2_term x^2 + C."""
    x_info = -2.0 * e2_term
    return e_term / x_info, x_info


def get_e_dp_prior_logitnorm_approx(alpha, lognorm_means, lognorm_infos, gh_loc, gh_weights):
    """:213-221: expected Dirichlet-process prior with logit-normal sticks."""
    _, e_log_1mv = get_e_log_logitnormal(lognorm_means, lognorm_infos, gh_loc, gh_weights)
    return (alpha - 1) * e_log_1mv
