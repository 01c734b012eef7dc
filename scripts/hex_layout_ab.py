"""One GPU, one process per layout: the tuned Hex8 element-stage kernels (elasticity: FOL_HEX_LAYOUT, J2: FOL_J2_LAYOUT)
timed in the two regimes of profiles/r2/hex_kernel_experiments.md -- "burst" (20 steps right after 8 untimed ones) and
"sustained" (3 x 50 steps after 300) -- plus a SHA-256 of the outputs, so runs of different layouts can be compared
bit for bit.   FOL_HEX_LAYOUT=2 N=128 python scripts/hex_layout_ab.py [mech|j2]"""
import hashlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import folax_b200
from folax_b200 import loss_functions as lf

which = sys.argv[1] if len(sys.argv) > 1 else "mech"
n = int(os.environ.get("N", 128))
mesh = folax_b200.perturb_interior_nodes(folax_b200.create_3D_box_mesh(n, n, n, 1.0, 1.0, 1.0), 0.1)
bc = {d: {"left": 0.0, "right": 0.1} for d in ("Ux", "Uy", "Uz")}
g = torch.Generator(device="cuda").manual_seed(0)
if which == "mech":
    loss = lf.MechanicalLoss3DHexa("ab", {"dirichlet_bc_dict": bc, "num_gp": 2,
                                          "material_dict": {"young_modulus": 1.0, "poisson_ratio": 0.3}}, mesh)
    loss.Initialize()
    K = torch.rand(loss._nn, generator=g, device="cuda", dtype=torch.float64) * 0.9 + 0.1
    u = torch.randn(loss.total_number_of_dofs, generator=g, device="cuda", dtype=torch.float64) * 0.01
    ke = torch.empty(loss._ne * 576, dtype=torch.float64, device="cuda")
    outs = [ke]
    step = lambda: loss._assemble(K, u, False, ke_out=ke)
else:
    mat = {"young_modulus": 3.0, "poisson_ratio": 0.3, "iso_hardening_parameter_1": 0.4, "iso_hardening_param_2": 10.0,
           "yield_limit": 0.2}
    loss = lf.ElastoplasticityLoss3DHexa("ab", {"dirichlet_bc_dict": bc, "num_gp": 2, "material_dict": mat}, mesh)
    loss.Initialize()
    K = torch.ones(loss._nn, device="cuda", dtype=torch.float64)
    u = 0.2 / n * torch.randn(loss.total_number_of_dofs, generator=g, device="cuda", dtype=torch.float64)
    ke = torch.empty(loss._ne * 576, dtype=torch.float64, device="cuda")
    st = torch.zeros(loss.GetStateShape(), dtype=torch.float64, device="cuda")
    # a non-trivial history: the state after one smaller step
    st1 = torch.empty_like(st)
    loss._assemble(K, 0.5 * u, False, ke_out=ke, state_in=st, state_out=st1)
    st_out = torch.empty_like(st)
    outs = [ke, st_out]
    step = lambda: loss._assemble(K, u, False, ke_out=ke, state_in=st1, state_out=st_out)


def timed(steps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        step()
    b.record(); b.synchronize()
    return a.elapsed_time(b) / steps


out = {"which": which, "n": n, "hex_layout": os.environ.get("FOL_HEX_LAYOUT", "default"), "j2_layout": os.environ.get("FOL_J2_LAYOUT", "default"),
       "diag": os.environ.get("FOL_HEX_DIAG", ""), "hint": os.environ.get("FOL_HEX_HINT", "0")}
for _ in range(8):
    step()
torch.cuda.synchronize()
out["burst_ms"] = timed(20)
for _ in range(300):
    step()
torch.cuda.synchronize()
out["sustained_ms"] = [timed(50) for _ in range(3)]
res = step()
torch.cuda.synchronize()
h = hashlib.sha256()
for t in outs:
    h.update(t.cpu().numpy().tobytes())
if isinstance(res, tuple):
    for t in res:
        if torch.is_tensor(t):
            h.update(t.cpu().numpy().tobytes())
out["sha256"] = h.hexdigest()
out["abs_sum"] = float(ke.abs().sum())
print(json.dumps(out))
