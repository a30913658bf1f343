"""GPU: two ranks (gloo rendezvous, both on cuda:0 -- the test box has one GPU; NCCL needs one
device per rank) drive the REAL CUDA shards through lrvb_b200.distributed and must reproduce the
single-GPU result and the oracle: include_global_terms / include_A / set_global_block /
Schur-piece all-reduce paths of the C ABI."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import glmm_oracle as go

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import lrvb_b200 as vb
        from lrvb_b200.distributed import ShardedLogisticGLMM
        N, K, G, Q = 4000, 6, 37, 8
        X, y, g = go.make_glmm_data(N, K, G, seed=31)
        rng = np.random.default_rng(32)
        perm = rng.permutation(N)
        X, y, g = X[perm], y[perm], g[perm]
        w = rng.uniform(0.5, 1.5, N)
        model = ShardedLogisticGLMM.from_full(X, y, g, G, weights=w, num_gh_points=Q, min_info=0.02,
                                              min_shape=0.02, min_rate=0.02)
        obj = vb.Objective(model.glmm_par, model)
        order = np.argsort(g, kind="stable")
        gh_x, gh_w = np.polynomial.hermite.hermgauss(Q)
        b = 0.02
        oracle = go.GLMMOracle(X[order], y[order], g[order], gh_x, gh_w, weights=w[order], G=G,
                               bounds=go.GLMMBounds(b, b, b, b, b))
        x = go.make_free(oracle.lay.D, 31)
        res = {}
        res["kl"] = abs(obj.fun_free(x) - oracle.kl(x)) / abs(oracle.kl(x))
        ge = oracle.kl_grad(x)
        res["grad"] = np.abs(obj.fun_free_grad(x) - ge).max() / np.abs(ge).max()
        H = obj.fun_free_hessian(x)
        He = oracle.kl_hessian_csr(x)
        res["pattern"] = float(not (np.array_equal(H.indptr, He.indptr)
                                    and np.array_equal(H.indices, He.indices)))
        res["hess"] = np.abs(H.data - He.data).max() / np.abs(He.data).max()
        v = rng.standard_normal(x.size)
        hve = oracle.kl_hvp(x, v)
        res["hvp"] = np.abs(obj.fun_free_hvp(x, v) - hve).max() / np.abs(hve).max()
        Hd = He.toarray()
        xe = np.linalg.solve(Hd, v)
        solver = vb.ConjugateGradientSolver(obj.fun_free_hvp, x)
        solver.tol = 1e-11
        solver.preconditioner = "block_jacobi"
        xs, info = solver.get_hinv_vec(v)
        res["cg"] = float(info) + np.abs(xs - xe).max() / np.abs(xe).max()
        lr = vb.LinearResponseCovariances(obj, x)
        Hinv = np.linalg.inv(Hd)
        Dg = oracle.lay.Dg
        res["cov_g"] = np.abs(lr.get_global_covariance() - Hinv[:Dg, :Dg]).max() / np.abs(Hinv).max()
        res["solve"] = np.abs(lr.hinv(v).cpu().numpy() - xe).max() / np.abs(xe).max()
        out[rank] = res
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_rank_cuda_shards_match_oracle():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert len(out) == world
    for rank in range(world):
        res = out[rank]
        assert res["pattern"] == 0.0
        for key in ("kl", "grad", "hess", "hvp"):
            assert res[key] < 1e-9, (rank, key, res[key])
        for key in ("cg", "cov_g", "solve"):
            assert res[key] < 1e-8, (rank, key, res[key])
