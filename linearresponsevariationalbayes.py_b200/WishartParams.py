"""Wishart variational factor, mirror of
/root/reference/LinearResponseVariationalBayes/WishartParams.py:6-35: degrees of freedom ``df`` (lower bound
``size - 1``) and a positive-definite scale ``v`` (log-Cholesky packing on the device, csrc/packing.cu);
entropy / E log det from the batched Wishart kernel (csrc/ef.cu)."""
import numpy as np

from . import ExponentialFamilies as ef
from .MatrixParameters import PosDefMatrixParam
from .ParameterDictionary import ModelParamsDict
from .Parameters import ScalarParam


class WishartParam(ModelParamsDict):
    def __init__(self, name="", size=2, diag_lb=0.0, min_df=None):
        super().__init__(name=name)
        self._size = int(size)
        if not min_df:
            min_df = size - 1
        assert min_df >= size - 1
        self.push_param(ScalarParam("df", lb=min_df))
        # the reference builds v with the default size whatever `size` is (:14-15);
        # here v is size x size
        self.push_param(PosDefMatrixParam("v", self._size, diag_lb=diag_lb))

    def e(self):
        return self["df"].get() * self["v"].get()

    def e_log_det(self):
        return ef.e_log_det_wishart(self["df"].get(), self["v"].get())

    def e_inv(self):
        return self["df"].get() * np.linalg.inv(self["v"].get())

    def entropy(self):
        return ef.wishart_entropy(self["df"].get(), self["v"].get())

    def e_log_lkj_inv_prior(self, lkj_param):
        """Expected LKJ log-prior on the INVERSE of the Wishart-distributed matrix (:29-35)."""
        return ef.expected_ljk_prior(lkj_param, self["df"].get(), self["v"].get())

    def size(self):
        return self._size
