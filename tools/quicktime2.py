"""Scratch timing of the evaluation at C2 / C3 (orders 0,1,2), smaller than tools_quicktime."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import sys
import torch
import lrvb_b200 as vb

def run(N, K, G, Q, reps=10):
    X = torch.randn(N, K, dtype=torch.float64, device="cuda")
    base, rem = divmod(N, G)
    counts = torch.full((G,), base, dtype=torch.int64); counts[:rem] += 1
    g = torch.repeat_interleave(torch.arange(G), counts).cuda()
    y = (torch.rand(N, device="cuda") < 0.5).double()
    model = vb.LogisticGLMM(X, y, g, num_gh_points=Q, num_groups=G)
    x = torch.randn(model.D, dtype=torch.float64, device="cuda") * 0.1
    out = []
    for order in (0, 1, 2):
        for _ in range(3):
            model.evaluate(x, order, force=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            model.evaluate(x, order, force=True)
        e1.record(); torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1) / reps)
    keep = None
    def step():
        model.evaluate(x, 2, force=True)
        return model.hessian_csr()
    for _ in range(3):
        keep = step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        keep = step()
    e1.record(); torch.cuda.synchronize()
    out.append(e0.elapsed_time(e1) / reps)
    print("N=%d K=%d G=%d Q=%d: order0 %.3f  order1 %.3f  order2 %.3f  order2+csr %.3f ms" % ((N, K, G, Q) + tuple(out)), flush=True)

if __name__ == "__main__":
    run(1000000, 20, 10000, 8)
    if len(sys.argv) > 1:
        run(10000000, 50, 100000, 8, reps=3)
