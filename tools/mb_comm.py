"""NCCL latency probes at P ranks (torchrun): tiny all-reduce, 4 MB neighbour exchange; then bench phases in sync mode."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
def timed(fn, reps=200):
    for _ in range(20): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); e1.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item() * 1e3
small = torch.ones(136, dtype=torch.float64, device=dev)
t_ar = timed(lambda: dist.all_reduce(small))
buf_s = torch.ones(524288, dtype=torch.float64, device=dev); buf_r0 = torch.empty_like(buf_s); buf_r1 = torch.empty_like(buf_s)
def halo():
    ops = []
    if rank > 0: ops += [dist.P2POp(dist.isend, buf_s, rank - 1), dist.P2POp(dist.irecv, buf_r0, rank - 1)]
    if rank < world - 1: ops += [dist.P2POp(dist.isend, buf_s, rank + 1), dist.P2POp(dist.irecv, buf_r1, rank + 1)]
    for r in dist.batch_isend_irecv(ops): r.wait()
t_halo = timed(halo, reps=100)
if rank == 0: print("P=%d: allreduce(136 f64) %.1f us, neighbour exchange 4 MB each way %.1f us" % (world, t_ar, t_halo), flush=True)
dist.destroy_process_group()
