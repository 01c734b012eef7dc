"""One GPU: the element-stage kernel in its plain and its interface-first / in-kernel-push form (no neighbour connected,
so the push is a plain gather), timed alternately in the same process after a long warm-up -- separates the cost of the
fused form from power-cap drift between runs.  N=128 python scripts/fused_ab.py"""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from folax_b200 import _lib
from folax_b200.distributed import SlabPartition, assemble_overlapped
from folax_b200.loss_functions import MechanicalLoss3DHexa

n = int(os.environ.get("N", 128))
part = SlabPartition(n, n, 2 * n, 1.0, 1.0, 2.0, 0, 2)
bc = {d: {"left": 0.0, "right": 0.1} for d in ("Ux", "Uy", "Uz")}
loss = MechanicalLoss3DHexa("ab", {"dirichlet_bc_dict": bc, "num_gp": 2, "material_dict": {"young_modulus": 1.0, "poisson_ratio": 0.3}}, part.mesh)
loss.Initialize()
g = torch.Generator(device="cuda").manual_seed(0)
K = torch.rand(loss._nn, generator=g, device="cuda", dtype=torch.float64) * 0.9 + 0.1
u = torch.randn(loss.total_number_of_dofs, generator=g, device="cuda", dtype=torch.float64) * 0.01
ke = torch.empty(loss._ne * 576, dtype=torch.float64, device="cuda")
h = C.c_void_p()
_lib.check(_lib.load().fol_halo_create(C.byref(h), loss._dt, part.plane_nodes * 3))
part._halo, part._halo_step = h, 0


def plain(ev=None):
    if ev: ev[0].record()
    loss._assemble(K, u, False, ke_out=ke)
    if ev: ev[1].record()


def fused(ev=None):
    assemble_overlapped(loss, part, K, u, ke, None, kernel_events=ev)


def timed(fn, steps, kernel_only):
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(steps):
        fn(evs[i] if kernel_only else None)
    b.record(); b.synchronize()
    return a.elapsed_time(b) / steps, (float(np.mean([x.elapsed_time(y) for x, y in evs])) if kernel_only else None)


out = {"n": n}
for _ in range(300):
    plain()
torch.cuda.synchronize()
for rnd in range(3):
    out[f"plain_step_ms_{rnd}"] = timed(plain, 50, False)[0]
    out[f"fused_step_ms_{rnd}"], out[f"fused_kernel_ms_{rnd}"] = timed(fused, 50, True)
# plane work alone: the fused step with the interface-first order but NO plane chunks handed out is not expressible
# through the API; instead time the fused step on a slab whose planes are tiny (same elements, 2 x 2 x (n^3/4) box)
print(json.dumps(out))
part.close_peer_halo()
