"""Tuning helper (GPU): time the batched loss+VJP kernel of the C3 workload (thermal 256x256 quads, B=1024, f64)
for the variant selected by FOL_ENERGY_VARIANT / FOL_ENERGY_MAX_ELEMS / FOL_ENERGY_TILE_NODES / FOL_ENERGY_WAVES /
FOL_ENERGY_V1 and print one line with a checksum, so variants can be compared run to run."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import folax_b200  # noqa: E402
from folax_b200.loss_functions import MechanicalLoss2DQuad, ThermalLoss2DQuad  # noqa: E402

B = int(os.environ.get("B", 1024))
DT = getattr(torch, os.environ.get("DTYPE", "float64"))
mesh = folax_b200.create_2D_square_mesh(1.0, 257)
MECH = os.environ.get("PHYS", "thermal") == "mech"     # PHYS=mech: MechanicalLoss2DQuad on the same mesh (2 dofs per node)
if MECH:
    loss = MechanicalLoss2DQuad("m", {"dirichlet_bc_dict": {"Ux": {"left": 0.0, "right": 0.05}, "Uy": {"left": 0.0, "right": 0.0}},
                                      "material_dict": {"young_modulus": 1.0, "poisson_ratio": 0.3},
                                      "dtype": os.environ.get("DTYPE", "float64")}, mesh)
else:
    loss = ThermalLoss2DQuad("t", {"dirichlet_bc_dict": {"T": {"left": 1.0, "right": 0.1}}, "beta": 2.0, "c": 4,
                                   "dtype": os.environ.get("DTYPE", "float64")}, mesh)
loss.Initialize()
nn = mesh.GetNumberOfNodes()
g = torch.Generator(device="cuda").manual_seed(0)
K = torch.rand((B, nn), generator=g, device="cuda", dtype=DT) * 0.9 + 0.1
u = torch.rand((B, nn * (2 if MECH else 1)), generator=g, device="cuda", dtype=DT) * (0.01 if MECH else 1.0)
for _ in range(3):
    e, gu, gk = loss._energy_and_grads(K, u)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
n = 10
ev[0].record()
for _ in range(n):
    e, gu, gk = loss._energy_and_grads(K, u)
ev[1].record()
torch.cuda.synchronize()
ms = ev[0].elapsed_time(ev[1]) / n
ep = loss._energy_plan()
print(json.dumps({"env": {k: v for k, v in os.environ.items() if k.startswith("FOL_")}, "ms": round(ms, 4),
                  "samples_per_s": round(B / ms * 1e3), "ntiles": ep["ntiles"], "ecap": ep["ecap"], "lcap": ep["lcap"],
                  "E": e.sum().item(), "gu": gu.abs().sum().item(), "gk": gk.abs().sum().item() if gk is not None else 0.0,
                  "hash": [float(e[17 % B]), float(gu[5 % B, 1234]), float(gk[1023 % B, 40000]) if gk is not None else 0.0]}))

# whole physics step through the public API (ComputeBatchLoss + backward), as bench.py's physics_only leg
def physics_only():
    uu = u.detach().requires_grad_(True)
    kk = K.detach().requires_grad_(True)
    mean, _ = loss.ComputeBatchLoss(kk, uu)
    mean.backward()


for _ in range(3):
    physics_only()
torch.cuda.synchronize()
ev[0].record()
for _ in range(n):
    physics_only()
ev[1].record()
torch.cuda.synchronize()
print(json.dumps({"physics_only_ms": round(ev[0].elapsed_time(ev[1]) / n, 4)}))
