"""GPU (>= 2 devices): the NVLink peer-memory all-reduce (csrc/p2p.cu) against the exact sum, and
the sharded GLMM -- one process per GPU over NCCL, replicated blocks summed through the peer
windows -- against the oracle.  Skipped on a one-GPU box."""
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
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import lrvb_b200 as vb
        from lrvb_b200.distributed import PeerAllReduce, ShardedLogisticGLMM
        res = {}
        peer = PeerAllReduce.create(5000)
        res["created"] = peer is not None
        if peer is None:
            out[rank] = res
            return
        # exact sums: integers (any order of addition gives the same bits), many epochs, ragged sizes
        worst, same = 0.0, True
        for it, n in enumerate([1, 2, 511, 512, 513, 1981, 4999, 5000, 7, 1981, 1981, 3000]):
            gen = torch.Generator().manual_seed(100 + it)
            parts = torch.randint(-1000, 1000, (world, n), generator=gen).double()
            t = parts[rank].to(dev).contiguous()
            peer.all_reduce_(t)
            worst = max(worst, float((t.cpu() - parts.sum(0)).abs().max()))
        # non-integers: identical bits on every rank (fixed rank order of the additions)
        t = torch.randn(1981, generator=torch.Generator().manual_seed(7 + rank), dtype=torch.float64).to(dev)
        ref = t.clone()
        peer.all_reduce_(t)
        dist.all_reduce(ref)
        gathered = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(gathered, t)
        same = all(torch.equal(gathered[0], gr) for gr in gathered)
        res["sum_err"] = worst
        res["bitwise_same"] = bool(same)
        res["vs_nccl"] = float((t - ref).abs().max() / ref.abs().max())
        res["status"] = peer.status()
        peer.close()

        N, K, G, Q = 6000, 7, 41, 8
        X, y, g = go.make_glmm_data(N, K, G, seed=41)
        model = ShardedLogisticGLMM.from_full(X, y, g, G, num_gh_points=Q)
        res["model_peer"] = model._peer is not None
        obj = vb.Objective(model.glmm_par, model)
        order = np.argsort(g, kind="stable")
        gh_x, gh_w = np.polynomial.hermite.hermgauss(Q)
        oracle = go.GLMMOracle(X[order], y[order], g[order], gh_x, gh_w, G=G)
        x = go.make_free(oracle.lay.D, 41)
        res["kl"] = abs(obj.fun_free(x) - oracle.kl(x)) / abs(oracle.kl(x))
        ge = oracle.kl_grad(x)
        res["grad"] = np.abs(obj.fun_free_grad(x) - ge).max() / np.abs(ge).max()
        H = obj.fun_free_hessian(x)
        He = oracle.kl_hessian_csr(x)
        res["pattern"] = float(not (np.array_equal(H.indptr, He.indptr)
                                    and np.array_equal(H.indices, He.indices)))
        res["hess"] = np.abs(H.data - He.data).max() / np.abs(He.data).max()
        v = np.random.default_rng(42).standard_normal(x.size)
        hve = oracle.kl_hvp(x, v)
        res["hvp"] = np.abs(obj.fun_free_hvp(x, v) - hve).max() / np.abs(hve).max()
        # conjugate gradient over the shards: device-resident iteration with two peer all-reduces per
        # step (lrvb_glmm_cg_sharded) against the dense solve and against the torch / NCCL iteration
        Hd = He.toarray()
        xe = np.linalg.solve(Hd, v)
        for tag, pre in (("cg_jacobi", 1), ("cg_plain", 0)):
            xs, info, iters = model.cg(v, precond=pre, rtol=1e-11)
            res[tag] = float(info) + np.abs(xs.cpu().numpy() - xe).max() / np.abs(xe).max()
            res[tag + "_iters"] = iters
        xs0, info0, iters0 = model.cg(v, x0_full=xe * (1 + 1e-3), precond=1, rtol=1e-11)
        res["cg_x0"] = float(info0) + np.abs(xs0.cpu().numpy() - xe).max() / np.abs(xe).max()
        res["cg_x0_fewer"] = float(iters0 < res["cg_jacobi_iters"])
        peer = model._peer
        model._peer = None
        xs_t, info_t, iters_t = model.cg(v, precond=1, rtol=1e-11)
        model._peer = peer
        res["cg_vs_torch_iters"] = abs(iters_t - res["cg_jacobi_iters"])
        res["cg_torch"] = float(info_t) + np.abs(xs_t.cpu().numpy() - xe).max() / np.abs(xe).max()
        solver = vb.ConjugateGradientSolver(obj.fun_free_hvp, x)
        solver.tol = 1e-11
        solver.preconditioner = "block_jacobi"
        xs_s, info_s = solver.get_hinv_vec(v)
        res["cg_solver"] = float(info_s) + np.abs(xs_s - xe).max() / np.abs(xe).max()
        lr = vb.LinearResponseCovariances(obj, x)
        Hinv = np.linalg.inv(Hd)
        Dg = oracle.lay.Dg
        res["cov_g"] = np.abs(lr.get_global_covariance() - Hinv[:Dg, :Dg]).max() / np.abs(Hinv).max()
        # device Newton over the shards: every rank walks the same iterates (bitwise) to the optimum
        xo, r = vb.OptimizationUtils.minimize_objective_newton(obj, x, maxiter=30, gtol=1e-7)
        res["newton_ok"] = bool(r.success)
        res["newton_grad"] = float(np.abs(oracle.kl_grad(xo)).max())
        xt = torch.from_numpy(np.ascontiguousarray(xo)).to(dev)
        xs_all = [torch.empty_like(xt) for _ in range(world)]
        dist.all_gather(xs_all, xt)
        res["newton_same"] = all(torch.equal(xs_all[0], t) for t in xs_all)
        res["status2"] = model._peer.status() if model._peer is not None else 0
        out[rank] = res
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_peer_allreduce_and_sharded_model_all_gpus():
    world = min(torch.cuda.device_count(), 8)     # every GPU of the box, one process each
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert len(out) == world
    for rank in range(world):
        res = out[rank]
        assert res["created"], "peer windows could not be mapped on a multi-GPU box"
        assert res["sum_err"] == 0.0
        assert res["bitwise_same"]
        assert res["vs_nccl"] < 1e-14
        assert res["status"] == 0 and res["status2"] == 0
        assert res["model_peer"]
        assert res["pattern"] == 0.0
        for key in ("kl", "grad", "hess", "hvp"):
            assert res[key] < 1e-9, (rank, key, res[key])
        assert res["cov_g"] < 1e-8
        for key in ("cg_jacobi", "cg_plain", "cg_x0", "cg_torch", "cg_solver"):
            assert res[key] < 1e-8, (rank, key, res[key])
        assert res["newton_ok"] and res["newton_grad"] < 1e-6, (res["newton_ok"], res["newton_grad"])
        assert res["newton_same"]
        assert res["cg_x0_fewer"] == 1.0
        assert res["cg_vs_torch_iters"] <= 2, res["cg_vs_torch_iters"]


def _missing_peer_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from lrvb_b200.distributed import PeerAllReduce
        peer = PeerAllReduce.create(4096)
        res = {"created": peer is not None}
        if peer is None:
            out[rank] = res
            return
        peer.set_timeout(0.5)
        t = torch.full((100,), float(rank + 1), dtype=torch.float64, device=dev)
        peer.all_reduce_(t)                       # a healthy call first
        res["first_ok"] = bool((t == sum(range(1, world + 1))).all().item()) and peer.status() == 0
        dist.barrier()
        if rank != world - 1:
            # the last rank never issues this call: the others must come back with a STATUS, not a trap
            t2 = torch.ones(100, dtype=torch.float64, device=dev)
            peer.all_reduce_(t2)
            res["status"] = peer.status()          # syncs; 1 + (world - 1) = the missing rank
            res["nan"] = bool(torch.isnan(t2).all().item())
            # the CUDA context is alive: ordinary work still runs
            res["context_alive"] = float((torch.ones(8, device=dev) * 2).sum().item()) == 16.0
            try:
                peer.all_reduce_(t2)
                res["raises"] = False
            except RuntimeError as exc:
                res["raises"] = "did not arrive" in str(exc)
        dist.barrier()
        peer.close()
        out[rank] = res
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_missing_peer_is_a_status_not_a_trap():
    """a peer that never arrives within the deadline must not destroy the CUDA
    context; the status word (mapped pinned host memory) names the missing rank, the output is NaN and
    the communicator refuses further use."""
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_missing_peer_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert len(out) == world
    assert out[0]["created"] and out[0]["first_ok"] and out[1]["first_ok"]
    assert out[0]["status"] == world, out[0]
    assert out[0]["nan"] and out[0]["context_alive"] and out[0]["raises"] is True, dict(out[0])
