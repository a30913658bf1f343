"""CPU, world_size 2, gloo: the sharded algebra of lrvb_b200.distributed (group-aligned
partition, packed all-reduce, HVP / CG / Schur solves, CSR gather) driven by an oracle-backed
local model must reproduce the single-shard oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import glmm_oracle as go


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_partition_groups_is_group_aligned_and_balanced():
    from lrvb_b200.distributed import local_index_map, partition_groups
    rng = np.random.default_rng(0)
    counts = rng.integers(0, 50, size=97)
    for world in (1, 2, 3, 4, 8):
        b = partition_groups(counts, world)
        assert b[0] == 0 and b[-1] == 97 and np.all(np.diff(b) >= 1)
        per = [counts[b[r]:b[r + 1]].sum() for r in range(world)]
        assert sum(per) == counts.sum()
        assert max(per) - min(per) <= 2 * counts.max() + counts.sum() // (4 * world)
    b = partition_groups(np.array([5, 5]), 4)  # more ranks than groups: some ranks stay empty
    assert b[0] == 0 and b[-1] == 2 and np.all(np.diff(b) >= 0)
    m = local_index_map(6, 2, 4, 10)
    np.testing.assert_array_equal(m, [0, 1, 2, 3, 4, 5, 8, 9, 18, 19])


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import lrvb_b200 as vb
        from lrvb_b200.distributed import ShardedLogisticGLMM
        from dist_double import OracleLocal
        N, K, G, Q = 900, 3, 14, 4
        X, y, g = go.make_glmm_data(N, K, G, seed=21)
        rng = np.random.default_rng(22)
        perm = rng.permutation(N)             # unsorted input: the shard builder must sort
        X, y, g = X[perm], y[perm], g[perm]
        gh_x, gh_w = np.polynomial.hermite.hermgauss(Q)
        model = ShardedLogisticGLMM.from_full(
            X, y, g, G, local_factory=lambda Xl, yl, gl, Gl, wl, inc, **k: OracleLocal(
                Xl, yl, gl, Gl, wl, inc, gh_x=gh_x, gh_w=gh_w))
        obj = vb.Objective(model.glmm_par, model)
        order = np.argsort(g, kind="stable")
        oracle = go.GLMMOracle(X[order], y[order], g[order], gh_x, gh_w, G=G)
        x = go.make_free(oracle.lay.D, 21)
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
        for pre in (None, "block_jacobi"):
            solver.preconditioner = pre
            xs, info = solver.get_hinv_vec(v)
            res["cg_%s" % pre] = float(info) + np.abs(xs - xe).max() / np.abs(xe).max()
        lr = vb.LinearResponseCovariances(obj, x)
        Hinv = np.linalg.inv(Hd)
        Dg = oracle.lay.Dg
        res["cov_g"] = np.abs(lr.get_global_covariance() - Hinv[:Dg, :Dg]).max() / np.abs(Hinv).max()
        res["solve"] = np.abs(lr.hinv(v).numpy() - xe).max() / np.abs(xe).max()
        res["ranges"] = model.group_ranges
        out[rank] = res
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gloo_matches_single_oracle():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert len(out) == world
    for rank in range(world):
        res = out[rank]
        assert res["pattern"] == 0.0
        for key in ("kl", "grad", "hess", "hvp", "cov_g", "solve"):
            assert res[key] < 1e-10, (rank, key, res[key])
        for key in ("cg_None", "cg_block_jacobi"):
            assert res[key] < 1e-8, (rank, key, res[key])
    assert out[0]["ranges"] == out[1]["ranges"] and out[0]["ranges"][0][1] == out[0]["ranges"][1][0]
