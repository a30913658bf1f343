"""Normal variational factors.

Host-side mirror of /root/reference/LinearResponseVariationalBayes/NormalParams.py: the univariate
bundles ``UVNParam`` / ``UVNParamVector`` / ``UVNParamArray`` (:26-104; ``info`` lower-bounded by
``min_info``, ``var`` = ``1/info``) that the GLMM is made of, and the remaining ones over the packing
types of this library: ``MVNParam`` (:6-23, a ``PosDefMatrixParam`` information matrix, log-Cholesky
packing on the device), ``UVNMomentParamArray`` (:108-147), ``MVNArray`` (:150-162).
"""
import numpy as np

from . import ExponentialFamilies as ef
from .MatrixParameters import PosDefMatrixParam
from .ParameterDictionary import ModelParamsDict
from .Parameters import ArrayParam, ScalarParam, VectorParam


class _UVNBase(ModelParamsDict):
    def e(self):
        return self["mean"].get()

    def var(self):
        return 1.0 / self["info"].get()

    def e_outer(self):
        return self["mean"].get() ** 2 + 1 / self["info"].get()

    def e_exp(self):
        return ef.get_e_lognormal(self["mean"].get(), 1.0 / self["info"].get())

    def var_exp(self):
        return ef.get_var_lognormal(self["mean"].get(), 1.0 / self["info"].get())

    def e2_exp(self):
        return self.e_exp() ** 2 + self.var_exp()

    def entropy(self):
        return np.sum(ef.univariate_normal_entropy(self["info"].get()))


class UVNParam(_UVNBase):
    def __init__(self, name="", min_info=0.0):
        super().__init__(name=name)
        self.push_param(ScalarParam("mean"))
        self.push_param(ScalarParam("info", lb=min_info))


class UVNParamVector(_UVNBase):
    def __init__(self, name="", length=2, min_info=0.0):
        super().__init__(name=name)
        self._length = int(length)
        self.push_param(VectorParam("mean", length))
        self.push_param(VectorParam("info", length, lb=min_info))

    def size(self):
        return self._length


class UVNParamArray(_UVNBase):
    def __init__(self, name="", shape=(1, 1), min_info=0.0):
        super().__init__(name=name)
        self._shape = tuple(shape)
        self.push_param(ArrayParam("mean", shape=shape))
        self.push_param(ArrayParam("info", shape=shape, lb=min_info))

    def shape(self):
        return self._shape


class MVNParam(ModelParamsDict):
    """Multivariate normal, mean + information matrix (:6-23)."""

    def __init__(self, name="", dim=2, min_info=0.0):
        super().__init__(name=name)
        self._dim = int(dim)
        self.push_param(VectorParam("mean", dim))
        self.push_param(PosDefMatrixParam("info", dim, diag_lb=min_info))

    def e(self):
        return self["mean"].get()

    def cov(self):
        return np.linalg.inv(self["info"].get())

    def e_outer(self):
        mean = self["mean"].get()
        e_outer = np.outer(mean, mean) + self.cov()
        return 0.5 * (e_outer + e_outer.transpose())

    def entropy(self):
        return ef.multivariate_normal_entropy(self["info"].get())

    def dim(self):
        return self._dim


class UVNMomentParamArray(ModelParamsDict):
    """Array of univariate normals in the moment parameterisation (E x, E x^2) (:108-147)."""

    def __init__(self, name="", shape=(2, 3), min_info=0.0):
        super().__init__(name=name)
        self._shape = tuple(shape)
        self.push_param(ArrayParam("e", shape))
        self.push_param(ArrayParam("e2", shape, lb=min_info))

    def e(self):
        return self["e"].get()

    def e_outer(self):
        return self["e2"].get()

    def var(self):
        return self["e2"].get() - self["e"].get() ** 2

    def e_exp(self):
        return ef.get_e_lognormal(self["e"].get(), self.var())

    def var_exp(self):
        # the reference returns get_e_lognormal here too (:123-124); the variance of the
        # log-normal is what the name says and what e2_exp needs
        return ef.get_var_lognormal(self["e"].get(), self.var())

    def e2_exp(self):
        return self.e_exp() ** 2 + self.var_exp()

    def entropy(self):
        return np.sum(ef.univariate_normal_entropy(1.0 / self.var()))

    def shape(self):
        return self._shape

    def set_from_uvn_param_array(self, uvn_par):
        assert tuple(uvn_par.shape()) == self.shape()
        self["e"].set(uvn_par.e())
        self["e2"].set(uvn_par.e_outer())

    def set_from_constant(self, scalar_array_par):
        """From a plain array parameter, assuming zero variance."""
        assert tuple(scalar_array_par.shape()) == self.shape()
        self["e"].set(scalar_array_par.get())
        self["e2"].set(scalar_array_par.get() ** 2)


class MVNArray(ModelParamsDict):
    """Rows of independent normals with one information value per row (:150-162)."""

    def __init__(self, name="", shape=(2, 2), min_info=0.0):
        super().__init__(name=name)
        self._shape = tuple(shape)
        self.push_param(ArrayParam("mean", shape=shape))
        self.push_param(VectorParam("info", size=shape[0], lb=min_info))

    def e(self):
        return self["mean"].get()

    def e2(self):
        var = 1 / self["info"].get()
        return self["mean"].get() ** 2 + var[:, None]

    def shape(self):
        return self._shape
