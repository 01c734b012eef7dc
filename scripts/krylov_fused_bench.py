"""ms per BiCGSTAB iteration: multi-launch host-scalar loop, device-scalar loop and the one-launch kernel
(csrc/krylov_fused.cu) on the Jacobian of a Tet4 elasticity box (N^3 Kuhn cells; N=70 is configs[3]'s size and
sparsity) and, with HEX=n, of an n^3 Hex8 box.  Fixed iteration count (tolerance out of reach), Jacobi preconditioner."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import folax_b200
from folax_b200 import linalg
from folax_b200.loss_functions import MechanicalLoss3DTetra, MechanicalLoss3DHexa

n = int(os.environ.get("N", 70))
iters = int(os.environ.get("ITERS", 200))
hexn = int(os.environ.get("HEX", 0))
bc = {d: {"left": 0.0, "right": 0.01} for d in ("Ux", "Uy", "Uz")}
mat = {"young_modulus": 1.0, "poisson_ratio": 0.3}
if hexn:
    mesh = folax_b200.create_3D_box_mesh(hexn, hexn, hexn, 1.0, 1.0, 1.0)
    loss = MechanicalLoss3DHexa("m", {"dirichlet_bc_dict": bc, "material_dict": mat}, mesh)
else:
    mesh = folax_b200.create_3D_tetra_box_mesh(n, n, n, 1.0, 1.0, 1.0)
    loss = MechanicalLoss3DTetra("m", {"dirichlet_bc_dict": bc, "material_dict": mat}, mesh)
loss.Initialize()
K = np.random.default_rng(0).uniform(0.5, 1.0, mesh.GetNumberOfNodes())
u0 = loss.ApplyDirichletBCOnDofVector(np.zeros(loss.total_number_of_dofs))
jac, R = loss.ComputeJacobianMatrixAndResidualVector(K, u0)
A = linalg.SellOperator(loss, jac)
del jac
diag, rhs = A.diagonal(), -R
out = {"mesh": f"hex{hexn}" if hexn else f"tet{n}", "dofs": int(loss.total_number_of_dofs),
       "stored_entries": int(A.plan["total"]), "iterations": iters}


def run(name, fn):
    fn()                                   # warm-up (allocations, first launch)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    x, k = fn()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out[name] = {"ms_per_iteration": 1e3 * dt / max(k, 1), "iterations": k}
    return x


kw = dict(x0=u0, tol=1e-30, atol=0.0, maxiter=iters, M_diagonal=diag)
xh = run("host_scalars", lambda: linalg.bicgstab(A, rhs, **kw))
xd = run("device_scalars_16", lambda: linalg.bicgstab_device(A, rhs, check_every=16, **kw))
xf = run("fused_one_launch", lambda: linalg.bicgstab_fused(A, rhs, **kw))
out["fused_vs_host_max_rel_diff"] = float((xf - xh).abs().max() / xh.abs().max())
# one product alone, for the split
y = torch.empty_like(rhs)
for _ in range(3):
    A.matvec(rhs, y)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    A.matvec(rhs, y)
e1.record()
torch.cuda.synchronize()
out["spmv_ms"] = e0.elapsed_time(e1) / 20
print(json.dumps(out))
