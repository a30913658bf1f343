"""Where the end-to-end step (host numpy in, scipy CSR + numpy gradient + float out) spends its time at C2."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import lrvb_b200 as vb


def timeit(f, reps=50):
    for _ in range(5):
        f()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


def main():
    N, K, G, Q = 1000000, 20, 10000, 8
    X = torch.randn(N, K, dtype=torch.float64, device="cuda")
    g = torch.repeat_interleave(torch.arange(G), N // G).cuda()
    y = (torch.rand(N, device="cuda") < 0.5).double()
    model = vb.LogisticGLMM(X, y, g, num_gh_points=Q, num_groups=G)
    obj = vb.Objective(model.glmm_par, model)
    rng = np.random.default_rng(0)
    xs = [rng.normal(size=model.D) * 0.1 for _ in range(8)]
    it = [0]

    def nx():
        it[0] += 1
        return xs[it[0] % len(xs)]

    print("evaluate(order 2) host x      %.3f ms" % timeit(lambda: model.evaluate(nx(), 2)))
    def ev_csr():
        model.evaluate(nx(), 2)
        return model.hessian_csr()
    print(" + hessian_csr (device)       %.3f ms" % timeit(ev_csr))
    csr = ev_csr()
    print("to_scipy alone                %.3f ms  (nnz %d)" % (timeit(lambda: csr.to_scipy()), csr.nnz))
    v = csr.values
    hv = torch.empty(v.shape, dtype=v.dtype, pin_memory=True)
    def d2h():
        hv.copy_(v, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    ms = timeit(d2h)
    print("D2H values only (pinned, reused) %.3f ms  = %.1f GB/s" % (ms, v.numel() * 8 / ms / 1e6))
    def pin_alloc():
        return torch.empty(v.shape, dtype=v.dtype, pin_memory=True)
    print("pinned alloc from cache       %.3f ms" % timeit(pin_alloc))
    def full():
        x = nx()
        H = obj.fun_free_hessian(x)
        gr = obj.fun_free_grad(x)
        kl = obj.fun_free(x)
        return H, gr, kl
    print("full e2e step                 %.3f ms" % timeit(full))
    x = nx()
    print("fun_free_hessian              %.3f ms" % timeit(lambda: obj.fun_free_hessian(nx())))
    obj.fun_free_hessian(x)
    print("fun_free_grad (cached point)  %.3f ms" % timeit(lambda: obj.fun_free_grad(x)))
    print("fun_free (cached point)       %.3f ms" % timeit(lambda: obj.fun_free(x)))


if __name__ == "__main__":
    main()
