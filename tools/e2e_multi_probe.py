"""(torchrun) Why the end-to-end step is slower with 8 ranks than with 1: per-rank D2H bandwidth of a 14 MB
pinned copy, alone and with all ranks copying at once, with and without binding the rank to its GPU's NUMA
node (LRVB_PROBE_BIND=1), and the CPU set / NUMA placement every rank sees."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
bind = os.environ.get("LRVB_PROBE_BIND", "0") == "1"
info = ""
try:
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(lr)
    before = sorted(os.sched_getaffinity(0))
    if bind:
        pynvml.nvmlDeviceSetCpuAffinity(h)
    after = sorted(os.sched_getaffinity(0))
    try:
        ideal = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        ideal = [hex(int(x)) for x in ideal]
    except Exception as e:
        ideal = repr(e)
    info = "cpus before %d [%s..%s] after %d [%s..%s] ideal %s ncpu %d" % (len(before), before[0], before[-1], len(after), after[0], after[-1], ideal, os.cpu_count())
except Exception as e:
    info = "nvml: %r" % (e,)
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n = 1741614
src = torch.randn(n, dtype=torch.float64, device="cuda")
dst = [torch.empty(n, dtype=torch.float64, pin_memory=True) for _ in range(2)]
def bw(reps=30):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(reps):
        dst[i & 1].copy_(src, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    return n * 8 * reps / (time.perf_counter() - t0) / 1e9
bw(5)
res = {}
# alone: ranks take turns
for r in range(world):
    dist.barrier()
    if r == rank:
        res["alone"] = bw()
dist.barrier()
res["together"] = bw()
dist.barrier()
out = [None] * world
dist.all_gather_object(out, (rank, info, res))
if rank == 0:
    for r, i, d in sorted(out):
        print("rank %d: D2H 14 MB alone %.1f GB/s, all ranks at once %.1f GB/s | %s" % (r, d["alone"], d["together"], i))
dist.destroy_process_group()
