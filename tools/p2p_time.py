"""torchrun --nproc-per-node N tools/p2p_time.py: latency of the peer-memory all-reduce vs NCCL, and
where a sharded evaluation spends its time beyond the local evaluation."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist


def ev_time(f, reps, sync_ranks=True):
    for _ in range(10):
        f()
    torch.cuda.synchronize()
    if sync_ranks:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    os.dup2(2, 1)
    dist.init_process_group("nccl", device_id=dev)
    import lrvb_b200 as vb
    from lrvb_b200.distributed import PeerAllReduce, ShardedLogisticGLMM
    peer = PeerAllReduce.create(200000)
    out = []
    for n in (2, 1981, 10921, 163621):
        t = torch.randn(n, dtype=torch.float64, device=dev)
        a = ev_time(lambda: peer.all_reduce_(t), 500)
        t2 = torch.randn(n, dtype=torch.float64, device=dev)
        b = ev_time(lambda: dist.all_reduce(t2), 500)
        out.append("n=%6d  peer %.2f us   nccl %.2f us" % (n, a * 1e3, b * 1e3))
    st = peer.status()
    peer.close()

    N, K, G, Q = 1000000, 20, 10000, 8
    g = torch.repeat_interleave(torch.arange(G), N // G).to(dev)
    X = torch.randn(N, K, dtype=torch.float64, device=dev)
    y = (torch.rand(N, device=dev) < 0.5).double()
    model = ShardedLogisticGLMM.from_local_shard(X, y, g, G, num_gh_points=Q)
    x = torch.randn(model.D, dtype=torch.float64, device=dev) * 0.1
    loc = model.local

    def local_only():
        loc.evaluate(x, 2, force=True)
        return loc.hessian_csr()

    def sharded():
        model.evaluate(x, 2, force=True)
        return model.hessian_csr()
    keep = [local_only(), sharded()]
    out.append("local evaluate + csr (no comm)    %.4f ms" % ev_time(local_only, 200))
    out.append("sharded evaluate + csr (peer)     %.4f ms" % ev_time(sharded, 200))
    p = model._peer
    model._peer = None
    out.append("sharded evaluate + csr (nccl)     %.4f ms" % ev_time(sharded, 200))
    model._peer = p
    out.append("sharded evaluate only (peer)      %.4f ms" % ev_time(lambda: model.evaluate(x, 2, force=True), 200))
    out.append("local evaluate only               %.4f ms" % ev_time(lambda: loc.evaluate(x, 2, force=True), 200))
    import time
    v = torch.randn(model.D, dtype=torch.float64, device=dev)
    dist.broadcast(v, 0)
    for label in ("peer", "nccl"):
        if label == "nccl":
            model._peer = None
        model.cg(v, precond=1, rtol=1e-10, maxiter=16)
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter()
        xs, info, iters = model.cg(v, precond=1, rtol=1e-30, maxiter=96)   # fixed work: 96 iterations
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        out.append("sharded CG (%s): %d iterations, info %d, %.3f ms = %.1f us / iteration" % (
            label, iters, info, dt * 1e3, dt * 1e6 / max(iters, 1)))
    model._peer = p
    if rank == 0:
        print("world %d  status %d" % (world, st), file=sys.stderr)
        print("\n".join(out), file=sys.stderr)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
