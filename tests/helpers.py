"""Shared builders for the parity tests: seeded synthetic GLMM data, the CPU
oracle and the device model on the same inputs."""
import numpy as np

from oracle import glmm_oracle as go

RTOL, ATOL = 1e-9, 1e-12   # north_star tolerance: 1e-9 relative / 1e-12 absolute in fp64


def assert_close(actual, expected, rtol=RTOL, atol=ATOL, scale=None, what=""):
    """|a - e| <= atol + rtol * max(|e|, scale): ``scale`` lets entries of a matrix that are
    tiny by cancellation be judged against the matrix's magnitude."""
    a, e = np.asarray(actual, dtype=np.float64), np.asarray(expected, dtype=np.float64)
    assert a.shape == e.shape, (what, a.shape, e.shape)
    ref = np.abs(e)
    if scale is not None:
        ref = np.maximum(ref, scale)
    err = np.abs(a - e)
    bad = err > atol + rtol * ref
    assert not bad.any(), "%s: %d entries differ, max err %.3e (ref scale %.3e)" % (
        what, int(bad.sum()), float(err.max()), float(ref.max()))


def make_case(N, K, G, Q, seed, ragged=False, shuffle=False, weights=False, bounds=0.0,
              intercept=False, empty_groups=0):
    rng = np.random.default_rng(seed)
    X, y, g = go.make_glmm_data(N, K, G - empty_groups, seed, intercept=intercept)
    if ragged:
        # random group sizes, some groups possibly empty
        g = np.sort(rng.integers(0, G - empty_groups, size=N)).astype(np.int64)
    if empty_groups:
        # leave `empty_groups` ids unused, spread over the id range
        used = np.sort(rng.choice(G, size=G - empty_groups, replace=False))
        g = used[g]
    w = rng.uniform(0.5, 1.5, size=N) if weights else None
    if shuffle:
        perm = rng.permutation(N)
        X, y, g = X[perm], y[perm], g[perm]
        if w is not None:
            w = w[perm]
    gh_x, gh_w = np.polynomial.hermite.hermgauss(Q)
    D = 4 + 2 * K + 2 * G
    free = go.make_free(D, seed)
    return dict(X=X, y=y, g=g, w=w, gh_x=gh_x, gh_w=gh_w, G=G, K=K, N=N, free=free, bounds=bounds)


def make_oracle(case):
    order = np.argsort(case["g"], kind="stable")
    b = case["bounds"]
    return go.GLMMOracle(
        case["X"][order], case["y"][order], case["g"][order], case["gh_x"], case["gh_w"],
        weights=None if case["w"] is None else case["w"][order], G=case["G"],
        bounds=go.GLMMBounds(b, b, b, b, b))


def make_model(vb, case, **kw):
    b = case["bounds"]
    return vb.LogisticGLMM(case["X"], case["y"], case["g"], gh_x=case["gh_x"], gh_w=case["gh_w"],
                           weights=case["w"], num_groups=case["G"], min_info=b, min_shape=b,
                           min_rate=b, **kw)
