"""GPU timing of the device-resident linear algebra (round-2 measurement; not run yet -- no GPU minutes were left
when it was written): SELL SpMV against its HBM roofline (12 B per stored entry + 16 B per row), CSR de-duplication,
SELL conversion, one BiCGSTAB solve, and the response / adjoint-sensitivity kernels, at N^3 Hex8 elasticity f64.

    N=128 python scripts/solver_bench.py        (about 45 GB of device memory at N=128; N=64 for a quick look)
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import folax_b200
from folax_b200 import linalg
from folax_b200.loss_functions import MechanicalLoss3DHexa
from folax_b200.responses import FiniteElementResponse, NodalControl

n = int(os.environ.get("N", 64))
mesh = folax_b200.create_3D_box_mesh(n, n, n, 1.0, 1.0, 1.0)
bc = {d: {"left": 0.0, "right": 0.1} for d in ("Ux", "Uy", "Uz")}
loss = MechanicalLoss3DHexa("sb", {"dirichlet_bc_dict": bc, "num_gp": 2,
                                   "material_dict": {"young_modulus": 1.0, "poisson_ratio": 0.3}}, mesh)
loss.Initialize()
g = torch.Generator(device="cuda").manual_seed(0)
K = torch.rand(loss._nn, generator=g, device="cuda", dtype=torch.float64) * 0.9 + 0.1
u = loss.ApplyDirichletBCOnDofVector(np.zeros(loss.total_number_of_dofs))
v = torch.randn(loss.total_number_of_dofs, generator=g, device="cuda", dtype=torch.float64)


def timeit(fn, steps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / steps


out = {"n": n, "elements": loss._ne, "dofs": loss.total_number_of_dofs}
jac, R = loss.ComputeJacobianMatrixAndResidualVector(K, u)
t0 = time.time()
loss._csr_plan()
out["csr_plan_host_s"] = time.time() - t0
t0 = time.time()
sp = loss._sell_plan()
out["sell_plan_host_s"] = time.time() - t0
out["nnz"], out["sell_entries"] = sp["nnz"], sp["total"]
out["csr_values_ms"] = timeit(lambda: loss.JacobianToCSR(jac), steps=5)
A = linalg.SellOperator(loss, jac)
y = torch.empty_like(v)
out["sell_spmv_ms"] = timeit(lambda: A.matvec(v, y), steps=20)
bytes_spmv = (8.0 + 4.0 / 3.0) * sp["total"] + 16.0 * sp["nrows"]          # block-column kernel, d = 3
out["sell_spmv_gbs"] = bytes_spmv / (out["sell_spmv_ms"] * 1e-3) / 1e9
A.use_block_kernel = False
out["sell_spmv_scalar_columns_ms"] = timeit(lambda: A.matvec(v, y), steps=20)
out["sell_spmv_scalar_columns_gbs"] = (12.0 * sp["total"] + 16.0 * sp["nrows"]) / (out["sell_spmv_scalar_columns_ms"] * 1e-3) / 1e9
A.use_block_kernel = True
out["apply_jacobian_ms"] = timeit(lambda: loss.ApplyJacobian(K, u, v))
vec = linalg._Vectors(loss._dt, v.numel(), loss.dtype, loss.device)
out["dot_ms"] = timeit(lambda: vec.dot_into(v, y, 0), steps=20)
out["axpby_ms"] = timeit(lambda: vec.axpby(1.0, v, 0.5, y, y), steps=20)
torch.cuda.synchronize()
t0 = time.time()
x, info = linalg.bicgstab(A, -R, x0=None, tol=1e-8, atol=0.0, maxiter=int(os.environ.get("MAXITER", 200)),
                          M_diagonal=A.diagonal())
torch.cuda.synchronize()
out["bicgstab_iterations"], out["bicgstab_s"] = info, time.time() - t0
if info > 0:
    out["bicgstab_ms_per_iteration"] = 1e3 * out["bicgstab_s"] / info
torch.cuda.synchronize()
t0 = time.time()
x2, info2 = linalg.bicgstab_device(A, -R, x0=None, tol=1e-8, atol=0.0, maxiter=int(os.environ.get("MAXITER", 200)),
                                   M_diagonal=A.diagonal(), check_every=8)
torch.cuda.synchronize()
out["bicgstab_device_scalars_iterations"], out["bicgstab_device_scalars_s"] = info2, time.time() - t0
out["bicgstab_device_scalars_identical"] = bool(info2 == info and torch.equal(x, x2))
try:
    torch.cuda.synchronize()
    t0 = time.time()
    x3, info3 = linalg.bicgstab_device(A, -R, x0=None, tol=1e-8, atol=0.0, maxiter=int(os.environ.get("MAXITER", 200)),
                                       M_diagonal=A.diagonal(), check_every=8, use_graph=True)
    torch.cuda.synchronize()
    out["bicgstab_graph_s"], out["bicgstab_graph_identical"] = time.time() - t0, bool(info3 == info and torch.equal(x, x3))
except Exception as ex:                                   # first run of the graph path: keep the other numbers
    out["bicgstab_graph_error"] = str(ex)[:200]

resp = FiniteElementResponse("r", "(E**2)*U[0]", loss, NodalControl("E", mesh))
resp.Initialize()
out["response_value_ms"] = timeit(lambda: resp.ComputeValue(K, u), steps=5)
out["adjoint_control_derivatives_ms"] = timeit(lambda: resp.ComputeAdjointNodalControlDerivatives(K, u, v), steps=5)
out["adjoint_shape_derivatives_ms"] = timeit(lambda: resp.ComputeAdjointNodalShapeDerivatives(K, u, v), steps=5)
print(json.dumps(out))
