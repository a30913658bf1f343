"""Univariate-normal variational factors (mean / info parameterisation).

Host-side mirror of /root/reference/LinearResponseVariationalBayes/NormalParams.py:26-104
(UVNParam, UVNParamVector, UVNParamArray).  ``info`` is lower-bounded by ``min_info``; ``var`` is
``1/info``.  The multivariate (PosDefMatrix) bundles of the reference are outside the GLMM hot
path ("OUT OF SCOPE").
"""
import numpy as np

from . import ExponentialFamilies as ef
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
