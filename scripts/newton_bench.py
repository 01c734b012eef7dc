"""configs[3] of BASELINE.json at full size, device-resident: 3-D Neo-Hooke on a Kuhn-split tetra box (70^3 cells,
2.06 M Tet4), incremental Newton-Raphson through FiniteElementNonLinearResidualBasedSolver with the Jacobian
re-assembled every iteration and the linear solve on the GPU (SELL SpMV + Jacobi-BiCGSTAB).  Prints one JSON line
with the time split (assembly / de-duplication + SELL conversion / Krylov) per Newton iteration.
Written without GPU time (round 1); for the next session:   N=70 python scripts/newton_bench.py"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import folax_b200
from folax_b200 import linalg
from folax_b200.loss_functions import NeoHookeMechanicalLoss3DTetra
from folax_b200.solvers import FiniteElementNonLinearResidualBasedSolver

n = int(os.environ.get("N", 30))
mesh = folax_b200.create_3D_tetra_box_mesh(n, n, n, 1.0, 1.0, 1.0)
# the boundary displacement per load step must stay a fraction of the element size h = 1/n: Newton without a line
# search (the reference's, fe_nonlinear_residual_based_solver.py:111-166) applies it to the undeformed interior at once
disp = float(os.environ.get("DISP", 0.02))
bc = {"Ux": {"left": 0.0, "right": disp}, "Uy": {"left": 0.0, "right": 0.2 * disp}, "Uz": {"left": 0.0, "right": -0.2 * disp}}
loss = NeoHookeMechanicalLoss3DTetra("nh", {"dirichlet_bc_dict": bc, "material_dict": {"young_modulus": 1.0,
                                                                                      "poisson_ratio": 0.3}}, mesh)
settings = {"linear_solver_settings": {"solver": "JAX-bicgstab", "tol": float(os.environ.get("TOL", 1e-8)), "atol": 0.0,
                                       "maxiter": int(os.environ.get("MAXITER", 2000)), "pre-conditioner": "jacobi",
                                       "device_scalars": bool(int(os.environ.get("DEVICE_SCALARS", 0))),
                                       "fused": bool(int(os.environ.get("FUSED", 0))),
                                       "use_graph": bool(int(os.environ.get("USE_GRAPH", 0))),
                                       "check_every": int(os.environ.get("CHECK_EVERY", 8))},
            "nonlinear_solver_settings": {"rel_tol": 1e-8, "abs_tol": 1e-8, "maxiter": 10,
                                          "load_incr": int(os.environ.get("LOAD_STEPS", 5))}}
solver = FiniteElementNonLinearResidualBasedSolver("nl", loss, settings)
loss.Initialize()
solver.Initialize()
K = np.random.default_rng(0).uniform(0.5, 1.0, mesh.GetNumberOfNodes())
split = {"assembly_s": 0.0, "operator_s": 0.0, "krylov_s": 0.0, "newton_iterations": 0, "krylov_iterations": 0}


def timed(fn, key):
    def wrapper(*a, **k):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn(*a, **k)
        torch.cuda.synchronize()
        split[key] += time.perf_counter() - t0
        return out
    return wrapper


loss.ComputeJacobianMatrixAndResidualVector = timed(loss.ComputeJacobianMatrixAndResidualVector, "assembly_s")
_Sell = linalg.SellOperator
linalg.SellOperator = timed(_Sell, "operator_s")
_bicg = linalg.bicgstab


def bicg(*a, **k):
    x, info = timed(_bicg, "krylov_s")(*a, **k)
    split["krylov_iterations"] += max(info, 0)
    print(f"bicgstab info {info}  |x| {float(torch.linalg.norm(x)):.3e}", file=sys.stderr)
    split["newton_iterations"] += 1
    return x, info


linalg.bicgstab = bicg
_bicg_dev = linalg.bicgstab_device


def bicg_dev(*a, **k):
    x, info = timed(_bicg_dev, "krylov_s")(*a, **k)
    split["krylov_iterations"] += max(info, 0)
    split["newton_iterations"] += 1
    return x, info


linalg.bicgstab_device = bicg_dev
_bicg_fused = linalg.bicgstab_fused


def bicg_fused(*a, **k):
    x, info = timed(_bicg_fused, "krylov_s")(*a, **k)
    split["krylov_iterations"] += max(info, 0)
    split["newton_iterations"] += 1
    return x, info


linalg.bicgstab_fused = bicg_fused
t0 = time.time()
plan_t0 = time.time()
loss._csr_plan(); loss._sell_plan()
split["host_plans_s"] = time.time() - plan_t0
u = solver.Solve(K, np.zeros(loss.GetTotalNumberOfDOFs()))
torch.cuda.synchronize()
out = {"n": n, "elements": loss._ne, "dofs": loss.total_number_of_dofs, "total_s": time.time() - t0, **split,
       "final_residual_norms": {s: h["res_norm"][-1] for s, h in solver.convergence_history.items()}}
if split["newton_iterations"]:
    it = split["newton_iterations"]
    out["per_newton_iteration_ms"] = {k[:-2]: 1e3 * split[k] / it for k in ("assembly_s", "operator_s", "krylov_s")}
print(json.dumps(out))
