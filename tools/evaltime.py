"""CUDA-event time of one order-2 evaluation at a named config (tools/profile_eval.py CFG): python tools/evaltime.py c4s [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lrvb_b200 as vb
from profile_eval import CFG

def main(name, reps=10):
    N, K, G, Q = CFG[name]
    torch.manual_seed(0)
    X = torch.randn(N, K, dtype=torch.float64, device="cuda")
    base, rem = divmod(N, G)
    counts = torch.full((G,), base, dtype=torch.int64); counts[:rem] += 1
    g = torch.repeat_interleave(torch.arange(G), counts).cuda()
    y = (torch.rand(N, device="cuda") < 0.5).double()
    model = vb.LogisticGLMM(X, y, g, num_gh_points=Q, num_groups=G)
    x = torch.randn(model.D, dtype=torch.float64, device="cuda") * 0.1
    ts = []
    for i in range(reps + 3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); model.evaluate(x, 2, force=True); b.record(); torch.cuda.synchronize()
        if i >= 3: ts.append(a.elapsed_time(b))
    ts.sort()
    H = model.hessian_csr()
    chk = float(H.values.sum())
    print("%s: eval median %.4f ms  min %.4f  (KL %.10g, sum(H) %.10g)" % (name, ts[len(ts) // 2], ts[0], float(model.kl_tensor()), chk), flush=True)

if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 10)
