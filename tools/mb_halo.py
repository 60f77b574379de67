"""Halo-exchange cost at P ranks: MPK with s=1..8 on the 256^3 slab partition (torchrun)."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from ca_lanczos_b200 import _lib, api, gallery
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ctx = api.Context(local)
ids = [api.Context.comm_unique_id() if rank == 0 else None]; dist.broadcast_object_list(ids, src=0); ctx.init_comm(world, rank, ids[0])
m, s = int(sys.argv[1]) if len(sys.argv) > 1 else 256, 8
n, plane = m ** 3, m * m
lo, hi = (rank * n) // world, ((rank + 1) * n) // world
hl, hh = max(0, lo - s * plane), min(n, hi + s * plane)
dm = api.DeviceMatrix(gallery.laplace3d(m, row_lo=hl, row_hi=hh), s_max=s, layout=os.environ.get("MB_LAYOUT", "auto"), ctx=ctx, n_glob=n, row_begin=hl)
stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
v = torch.full((dm.n,), 1.0 / np.sqrt(n), dtype=torch.float64, device=dev); torch.cuda.synchronize()
re = np.ascontiguousarray(gallery.leja_points(0, 12, s))
Vp, ld = C.c_void_p(), C.c_int64()
for ss in (1, 2, 4, 8):
    def f(): _lib.check(ctx.lib.calz_mpk_inplace(dm.h, C.c_void_p(v.data_ptr()), ss, re.ctypes.data_as(_lib.c_dp), None, 1, 0, C.byref(Vp), C.byref(ld)), ctx.h)
    for _ in range(3): f()
    ctx.sync(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(10): f()
    e1.record(stream); e1.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 10], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0: print("P=%d s=%d: mpk %.3f ms (n_loc=%d n_own=%d ghosts=%d p2p_halo=%d p2p_ar=%d)" % (world, ss, t.item(), dm.info("n_loc"), dm.n, dm.info("n_ghost"), dm.info("p2p_halo"), dm.info("p2p_allreduce")), flush=True)
# all-reduce latency through the library: Gram of a tiny block = tsmm + all-reduce
import ctypes as C
X = torch.randn((8, 4096), dtype=torch.float64, device=dev); Cd = torch.zeros(64, dtype=torch.float64, device=dev); torch.cuda.synchronize()
for p2p in (1, 0):
    ctx.set_option("p2p", p2p) if False else None
def g(): ctx.lib.calz_gram(ctx.h, 4096, 8, X.data_ptr(), 4096, 8, X.data_ptr(), 4096, Cd.data_ptr())
for _ in range(20): g()
ctx.sync(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(200): g()
e1.record(stream); e1.synchronize()
if rank == 0: print("P=%d tiny gram + allreduce: %.1f us per call" % (world, e0.elapsed_time(e1) / 200 * 1e3), flush=True)
dist.destroy_process_group()
