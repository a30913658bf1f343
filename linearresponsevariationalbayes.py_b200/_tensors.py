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
