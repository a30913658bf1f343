"""Throughput of the batched packing kernels (csrc/packing.cu) against the HBM roofline."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lrvb_b200 as vb

mp, sx = vb.MatrixParameters, vb.SimplexParams


def t(f, reps=20):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    try:
        peak = float(json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"])
    except Exception:
        peak = 6650.0
    M = 4000000
    print("M = %d parameters; HBM peak %.0f GB/s" % (M, peak))
    for k in (2, 3, 4):
        v = k * (k + 1) // 2
        free = torch.randn(M, v, device="cuda", dtype=torch.float64) * 0.5
        mats = mp.unpack_posdef_matrix(free, diag_lb=0.1)
        rows = [("unpack k=%d" % k, lambda: mp.unpack_posdef_matrix(free, diag_lb=0.1), 8 * (v + k * k)),
                ("pack k=%d" % k, lambda: mp.pack_posdef_matrix(mats, diag_lb=0.1), 8 * (v + k * k)),
                ("free_to_vector k=%d" % k, lambda: mp.pos_def_matrix_free_to_vector(free, diag_lb=0.1), 16 * v),
                ("jac k=%d" % k, lambda: mp.pos_def_matrix_free_to_vector_jac(free), 8 * (v + v * v))]
        if k <= 3:
            rows.append(("hess k=%d" % k, lambda: mp.pos_def_matrix_free_to_vector_hess(free), 8 * (v + v ** 3)))
        for name, f, bpp in rows:
            ms = t(f)
            gbs = M * bpp / ms / 1e6
            print("%-22s %8.3f ms  %7.0f GB/s  %.2f of HBM peak" % (name, ms, gbs, gbs / peak))
    for d in (3, 5, 10):
        free = torch.randn(M, d - 1, device="cuda", dtype=torch.float64) * 2
        z = sx.constrain_simplex_matrix(free)
        rows = [("simplex constrain d=%d" % d, lambda: sx.constrain_simplex_matrix(free), 8 * (2 * d - 1)),
                ("simplex unconstrain d=%d" % d, lambda: sx.unconstrain_simplex_matrix(z), 8 * (2 * d - 1)),
                ("simplex jac d=%d" % d, lambda: sx.constrain_jac_matrix(free), 8 * (d - 1 + d * (d - 1)))]
        if d <= 5:
            rows.append(("simplex hess d=%d" % d, lambda: sx.constrain_hess_matrix(free), 8 * (d - 1 + d * (d - 1) ** 2)))
        for name, f, bpp in rows:
            ms = t(f)
            gbs = M * bpp / ms / 1e6
            print("%-26s %8.3f ms  %7.0f GB/s  %.2f of HBM peak" % (name, ms, gbs, gbs / peak))


if __name__ == "__main__":
    main()
