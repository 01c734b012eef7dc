"""GPU timing of the matrix-free Jacobian product vs the assembled path (128^3 hex elasticity f64)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import folax_b200
from folax_b200.loss_functions import MechanicalLoss3DHexa

n = int(os.environ.get("N", 128))
mesh = folax_b200.create_3D_box_mesh(n, n, n, 1.0, 1.0, 1.0)
bc = {d: {"left": 0.0, "right": 0.1} for d in ("Ux", "Uy", "Uz")}
loss = MechanicalLoss3DHexa("mv", {"dirichlet_bc_dict": bc, "num_gp": 2,
                                   "material_dict": {"young_modulus": 1.0, "poisson_ratio": 0.3}}, mesh)
loss.Initialize()
g = torch.Generator(device="cuda").manual_seed(0)
K = torch.rand(loss._nn, generator=g, device="cuda", dtype=torch.float64) * 0.9 + 0.1
u = 0.01 * torch.randn(loss.total_number_of_dofs, generator=g, device="cuda", dtype=torch.float64)
v = torch.randn(loss.total_number_of_dofs, generator=g, device="cuda", dtype=torch.float64)


def timeit(fn, steps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / steps


out = {"n": n, "elements": loss._ne}
out["apply_jacobian_ms"] = timeit(lambda: loss.ApplyJacobian(K, u, v))
out["apply_jacobian_T_ms"] = timeit(lambda: loss.ApplyJacobian(K, u, v, transpose_jacobian=True))
out["elements_per_s_matvec"] = loss._ne / (out["apply_jacobian_ms"] * 1e-3)
print(json.dumps(out))
