"""Timing of lrvb_spd_inverse for several n."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes, torch
from lrvb_b200 import _native as nat
lib = nat.load()
for n in (12, 44, 64, 65, 104, 128, 129, 144, 167, 204, 212, 234, 264, 300, 404, 516):
    A = torch.randn(n, n, dtype=torch.float64, device="cuda")
    S0 = A @ A.T + n * torch.eye(n, dtype=torch.float64, device="cuda")
    info = ctypes.c_int32()
    ts = []
    for r in range(6):
        S = S0.clone()
        torch.cuda.synchronize()
        t = time.perf_counter()
        nat.check(lib.lrvb_spd_inverse(nat.ptr(S), n, ctypes.byref(info), nat.stream_ptr()))
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t)
    err = float((S @ S0 - torch.eye(n, dtype=torch.float64, device="cuda")).abs().max())
    print("n=%3d: %.1f us (wall, incl. the info read-back)  |S^-1 S - I|max %.2e info %d" % (n, 1e6 * sorted(ts)[2], err, info.value))
