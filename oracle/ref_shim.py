"""Import the UNMODIFIED reference from /root/reference under a ``sys.modules`` shim.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Only usable in the build
container (``/root/reference`` does not exist on the GPU box); used by
``tests/golden/make_golden.py`` to generate the committed golden vectors.

The reference needs ``autograd`` (absent).  Its *forward* code only uses
``autograd.numpy`` / ``autograd.scipy`` as drop-in numpy / scipy, so those names are
aliased to the real numpy / scipy; every differentiation entry point
(``autograd.grad`` etc.) is replaced by a stub that raises if it is ever called.
"""
import os
import sys
import types

REFERENCE_ROOT = "/root/reference"


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "LinearResponseVariationalBayes"))


def _stub_factory(*_a, **_k):
    def _raise(*_aa, **_kk):
        raise NotImplementedError("autograd is not installed: differentiation stub called")
    return _raise


def import_reference():
    """Returns the reference package (``import LinearResponseVariationalBayes``)."""
    if not reference_available():
        raise RuntimeError("/root/reference is not present on this machine")
    import numpy
    import scipy
    import scipy.special, scipy.stats, scipy.linalg, scipy.sparse  # noqa: F401,E401
    import scipy.sparse.linalg  # noqa: F401

    if "autograd" not in sys.modules:
        ag = types.ModuleType("autograd")
        ag.numpy = numpy
        ag.scipy = scipy
        for name in ("grad", "jacobian", "hessian", "hessian_vector_product",
                     "make_jvp", "make_vjp", "elementwise_grad"):
            setattr(ag, name, _stub_factory)
        core = types.ModuleType("autograd.core")
        core.primitive = lambda f: f
        core.defvjp = lambda *a, **k: None
        core.defjvp = lambda *a, **k: None
        ext = types.ModuleType("autograd.extend")
        ext.primitive = core.primitive
        ext.defvjp = core.defvjp
        ext.defjvp = core.defjvp
        tu = types.ModuleType("autograd.test_util")
        tu.check_grads = _stub_factory
        ag.core, ag.extend = core, ext
        sys.modules.update({
            "autograd": ag, "autograd.numpy": numpy, "autograd.scipy": scipy,
            "autograd.core": core, "autograd.extend": ext, "autograd.test_util": tu,
            "autograd.numpy.random": numpy.random,
        })
    if "json_tricks" not in sys.modules:
        sys.modules["json_tricks"] = types.ModuleType("json_tricks")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import LinearResponseVariationalBayes as vb
        import LinearResponseVariationalBayes.Modeling  # noqa: F401
        import LinearResponseVariationalBayes.ExponentialFamilies  # noqa: F401
        import LinearResponseVariationalBayes.SparseObjectives  # noqa: F401
        import LinearResponseVariationalBayes.ConjugateGradient  # noqa: F401
    return vb
