"""Small helpers moving caller data to the device and results back in the caller's kind."""
import numpy as np

from . import _native as nat


def is_torch(x):
    return type(x).__module__.startswith("torch")


def to_device(x, dtype=None):
    """numpy / scalar / torch tensor -> contiguous CUDA tensor (fp64 by default)."""
    torch = nat.require_cuda()
    dtype = dtype or torch.float64
    if is_torch(x):
        t = x
        if not t.is_cuda:
            t = t.cuda()
        return t.to(dtype).contiguous()
    arr = np.ascontiguousarray(np.asarray(x), dtype={torch.float64: np.float64,
                                                     torch.int32: np.int32,
                                                     torch.int64: np.int64}[dtype])
    return torch.from_numpy(arr).cuda()


def like_input(t, *inputs):
    """Returns ``t`` (a CUDA tensor) in the kind of the inputs: torch stays torch, otherwise a
    numpy array (0-d results become Python floats)."""
    if any(is_torch(x) for x in inputs):
        return t
    out = t.detach().cpu().numpy()
    if out.ndim == 0:
        return float(out)
    return out


class PointKey(object):
    """Identity of an evaluation point for the "already evaluated here?" check of the models.

    numpy input: a private copy, compared by value (host work only).  torch input: the tensor
    itself (kept alive) with its storage address, shape, strides and torch's in-place version
    counter -- comparing VALUES of device tensors would cost a kernel, a device-to-host read and
    a stream synchronisation on every call.  A tensor that is modified in place through torch
    bumps its version and is re-evaluated; an equal-valued copy is simply evaluated again."""

    __slots__ = ("arr", "ref", "sig")

    def __init__(self, x):
        if is_torch(x):
            self.arr, self.ref, self.sig = None, x, self._sig(x)
        else:
            self.arr, self.ref, self.sig = np.array(x, dtype=np.float64).reshape(-1), None, None

    @staticmethod
    def _sig(x):
        return (x.data_ptr(), tuple(x.shape), tuple(x.stride()), x.dtype, x.device, x._version)

    def matches(self, x):
        if is_torch(x):
            return self.ref is not None and self.sig == self._sig(x)
        if self.arr is None:
            return False
        return np.array_equal(self.arr, np.asarray(x, dtype=np.float64).reshape(-1))
