"""Ordered dictionary of parameters with a flat free / vector layout.

Host-side mirror of /root/reference/LinearResponseVariationalBayes/ParameterDictionary.py:19-124.
The packing order is the ``push_param`` order (:39-46); ``free_indices_dict`` /
``vector_indices_dict`` give each entry's range in the flat vectors; a ModelParamsDict is itself a
parameter, so dictionaries nest (UVNParam, GammaParam ... are ModelParamsDicts).  The GLMM CUDA
path relies on exactly this layout: [mu.mean, mu.info, tau.shape, tau.rate,
beta.mean[K], beta.info[K], u.mean[G], u.info[G]].
"""
from collections import OrderedDict

import numpy as np
from scipy.sparse import block_diag

from . import Parameters as par


class ModelParamsDict(object):
    def __init__(self, name="ModelParamsDict"):
        self.name = name
        self.param_dict = OrderedDict()
        self.free_indices_dict = OrderedDict()
        self.vector_indices_dict = OrderedDict()
        self._n_free = 0
        self._n_vec = 0
        self.values = ModelParamsDictValues(self)

    def __str__(self):
        return self.name + ":\n" + "\n".join("\t" + str(p) for p in self.param_dict.values())

    def __getitem__(self, key):
        return self.param_dict[key]

    def push_param(self, param):
        nf, nv = param.free_size(), param.vector_size()
        self.param_dict[param.name] = param
        self.free_indices_dict[param.name] = range(self._n_free, self._n_free + nf)
        self.vector_indices_dict[param.name] = range(self._n_vec, self._n_vec + nv)
        self._n_free += nf
        self._n_vec += nv

    def set_name(self, name):
        self.name = name

    def dictval(self):
        return {p.name: p.dictval() for p in self.param_dict.values()}

    def _size_error(self, expected, got):
        return ValueError("Wrong size for parameter {}.  Expected {}, got {}".format(
            self.name, str(expected), str(got)))

    # ---- free ----
    def set_free(self, vec):
        if vec.size != self._n_free:
            raise self._size_error(self._n_free, vec.size)
        offset = 0
        for p in self.param_dict.values():
            offset = par.set_free_offset(p, vec, offset)

    def get_free(self):
        return np.hstack([p.get_free() for p in self.param_dict.values()])

    def free_to_vector(self, free_val):
        self.set_free(free_val)
        return self.get_vector()

    def free_to_vector_jac(self, free_val):
        fo, vo, blocks = 0, 0, []
        for p in self.param_dict.values():
            fo, vo, jac = par.free_to_vector_jac_offset(p, free_val, fo, vo)
            blocks.append(jac)
        return block_diag(blocks)

    def free_to_vector_hess(self, free_val):
        shape = (self._n_free, self._n_free)
        out, fo = [], 0
        for p in self.param_dict.values():
            fo = par.free_to_vector_hess_offset(p, free_val, out, fo, shape)
        return out

    # ---- vector ----
    def set_vector(self, vec):
        if vec.size != self._n_vec:
            raise self._size_error(self._n_vec, vec.size)
        offset = 0
        for p in self.param_dict.values():
            offset = par.set_vector_offset(p, vec, offset)

    def get_vector(self):
        return np.hstack([p.get_vector() for p in self.param_dict.values()])

    def names(self):
        return np.concatenate([np.atleast_1d(p.names()) for p in self.param_dict.values()])

    def free_size(self):
        return self._n_free

    def vector_size(self):
        return self._n_vec

    def get(self):
        return self.values


class ModelParamsDictValues(object):
    """Attribute-free ``par.values['name']`` view (ParameterDictionary.py:115-124)."""

    def __init__(self, owner):
        self.param_dict = owner

    def __getitem__(self, key):
        return self.param_dict[key].get()

    def __setitem__(self, key, val):
        return self.param_dict[key].set(val)
