"""Dirichlet variational factors, mirror of
/root/reference/LinearResponseVariationalBayes/DirichletParams.py:11-26: an ``ArrayParam`` ``alpha`` whose
FIRST axis is the simplex dimension; entropy and E-log come from the batched device kernels
(csrc/ef.cu k_dirichlet_terms)."""
from . import ExponentialFamilies as ef
from .ParameterDictionary import ModelParamsDict
from .Parameters import ArrayParam


class DirichletParamArray(ModelParamsDict):
    def __init__(self, name="", shape=(1, 2), min_alpha=0.0, val=None):
        super().__init__(name=name)
        self._shape = tuple(shape)
        assert min_alpha >= 0, "alpha parameter must be non-negative"
        self.push_param(ArrayParam("alpha", shape=shape, lb=min_alpha, val=val))

    def e(self):
        return ef.get_e_dirichlet(self["alpha"].get())

    def e_log(self):
        return ef.get_e_log_dirichlet(self["alpha"].get())

    def entropy(self):
        return ef.dirichlet_entropy(self["alpha"].get())

    def shape(self):
        return self._shape
