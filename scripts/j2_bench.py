"""GPU timing of the J2 elastoplastic assembly (hex, f64): one elastic-dominated and one fully plastic state."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import folax_b200
from folax_b200 import loss_functions as lf

n = int(os.environ.get("N", 64))
mesh = folax_b200.perturb_interior_nodes(folax_b200.create_3D_box_mesh(n, n, n, 1.0, 1.0, 1.0), 0.1)
bc = {d: {"left": 0.0, "right": 0.1} for d in ("Ux", "Uy", "Uz")}
mat = {"young_modulus": 3.0, "poisson_ratio": 0.3, "iso_hardening_parameter_1": 0.4, "iso_hardening_param_2": 10.0,
       "yield_limit": 0.2}
loss = lf.ElastoplasticityLoss3DHexa("j2", {"dirichlet_bc_dict": bc, "num_gp": 2, "material_dict": mat}, mesh)
loss.Initialize()
g = torch.Generator(device="cuda").manual_seed(0)
K = torch.ones(loss._nn, device="cuda", dtype=torch.float64)
ke = torch.empty(loss._ne * 576, dtype=torch.float64, device="cuda")
st = torch.zeros(loss.GetStateShape(), dtype=torch.float64, device="cuda")
st_out = torch.empty_like(st)
h = 1.0 / n
for amp in (0.02, 0.2, 2.0):
    u = amp * h * torch.randn(loss.total_number_of_dofs, generator=g, device="cuda", dtype=torch.float64)
    fn = lambda: loss._assemble(K, u, False, ke_out=ke, state_in=st, state_out=st_out)
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        fn()
    e1.record()
    e1.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(json.dumps({"n": n, "amp": amp, "plastic_fraction": float((st_out[..., -1] > 0).double().mean()), "ms": ms,
                      "elements_per_s": loss._ne / (ms * 1e-3), "checksum": float(ke.abs().sum())}))
