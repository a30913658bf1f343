"""Which entries of the global block differ from the oracle for a set of K (debug of the one-pass kernel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import lrvb_b200 as vb
from helpers import make_case, make_model, make_oracle

for K, N, G, wts in ((33, 4100, 41, True), (33, 4096, 41, False), (36, 2501, 25, False), (40, 2503, 20, False),
                     (50, 6000, 60, False), (56, 1503, 15, False), (62, 2500, 25, False), (20, 20000, 200, False),
                     (28, 3000, 30, False), (32, 2500, 25, True)):
    case = make_case(N=N, K=K, G=G, Q=8, seed=11, weights=wts)
    o = make_oracle(case)
    m = make_model(vb, case)
    x = case["free"]
    m.evaluate(x, 2)
    A, B, L = [t.cpu().numpy() for t in m.blocks()]
    _, _, blk = o.kl_blocks(x)
    dA = np.abs(A - blk["A"])
    bad = np.argwhere(dA > 1e-7 * np.abs(blk["A"]).max())
    dB = np.abs(B - blk["B"]).max() / np.abs(blk["B"]).max()
    print("K=%d N=%d: A bad entries %d, B rel err %.2e" % (K, N, len(bad), dB))
    Dg = 4 + 2 * K
    for (r, c) in bad[:30]:
        if r <= c:
            pr, pc = r - 4, c - 4      # packed column index
            print("   (%d,%d) packed (%d,%d) tiles (%d,%d) got %.6g want %.6g ratio %.4f" % (
                r, c, pr, pc, pr // 8, pc // 8, A[r, c], blk["A"][r, c], A[r, c] / blk["A"][r, c]))
