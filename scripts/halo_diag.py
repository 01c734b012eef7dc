"""Per-rank timing of the slab-partitioned assembly (torchrun): element stage alone, serial halo exchange,
overlapped schedule.  Diagnostic only."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from folax_b200.distributed import SlabPartition, assemble_overlapped
import folax_b200.distributed as D
from folax_b200.loss_functions import MechanicalLoss3DHexa

world, rank, lr = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n = 128
part = SlabPartition(n, n, n * world, 1.0, 1.0, float(world), rank, world)
bc = {d: {"left": 0.0, "right": 0.1} for d in ("Ux", "Uy", "Uz")}
loss = MechanicalLoss3DHexa("b", {"dirichlet_bc_dict": bc, "material_dict": {"young_modulus": 1.0, "poisson_ratio": 0.3}}, part.mesh)
loss.Initialize()
K = torch.rand(loss._nn, device="cuda", dtype=torch.float64) + 0.1
u = torch.randn(loss.total_number_of_dofs, device="cuda", dtype=torch.float64) * 0.01
ke = torch.empty(loss._ne * 576, dtype=torch.float64, device="cuda")
comm = torch.cuda.Stream()

def timeit(fn, steps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): fn()
    e1.record(); e1.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()

def local_only(): loss._assemble(K, u, False, ke_out=ke)
def serial():
    _, R = loss._assemble(K, u, False, ke_out=ke); part.halo_sum(R, 3)
def halo_only():
    part.halo_sum(Rbuf, 3)
Rbuf = torch.zeros(loss.total_number_of_dofs, device="cuda", dtype=torch.float64)
for _ in range(20): part.halo_sum(Rbuf, 3)
torch.cuda.synchronize(); dist.barrier()
res = {"local": timeit(local_only, 40), "serial": timeit(serial, 40), "halo_only": timeit(halo_only, 40), "serial_again": timeit(serial, 40)}
for m in (0, 8, 16, 32):
    D.GRID_MARGIN_CTAS = m
    res[f"overlap_margin{m}"] = timeit(lambda: assemble_overlapped(loss, part, K, u, ke, comm), 40)
if rank == 0: print(json.dumps(res))
dist.barrier(); dist.destroy_process_group()
