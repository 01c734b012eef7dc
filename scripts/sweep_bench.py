"""Throughput sweep of the element-stage + gather step over the other BASELINE.json configurations and
dtypes (single GPU).  Prints one JSON line per case: elements/s, ms, achieved algorithmic GB/s and the
fraction of the measured HBM copy bandwidth.  Algorithmic bytes per element:
nd^2*s + 4a + rho*((3 + d + 1)*s + d*s) [+ 2*g*ns*s of Gauss-point state for J2]."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import folax_b200
from folax_b200 import loss_functions as lf

HBM = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists("MEASURED_PEAKS.json") else 6555.2
MAT = {"young_modulus": 1.0, "poisson_ratio": 0.3}
J2MAT = {"young_modulus": 3.0, "poisson_ratio": 0.3, "iso_hardening_parameter_1": 0.4, "iso_hardening_param_2": 10.0,
         "yield_limit": 0.2}


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


ONLY = os.environ.get("ONLY", "")      # substring filter on the case names


def case(name, cls, mesh, etype, settings, dtype, ufun, state=False):
    if ONLY and ONLY not in name:
        return
    loss = cls(name, {**settings, "dtype": dtype}, mesh)
    loss.Initialize()
    ne, nn, nd, d, a = loss._ne, loss._nn, loss._nd, loss.number_dofs_per_node, loss._nnode
    s = 8 if dtype == "float64" else 4
    tdt = torch.float64 if dtype == "float64" else torch.float32
    g = torch.Generator(device="cuda").manual_seed(0)
    K = (torch.rand(nn, generator=g, device="cuda", dtype=torch.float64) * 0.9 + 0.1).to(tdt)
    u = (ufun(torch.randn(loss.total_number_of_dofs, generator=g, device="cuda", dtype=torch.float64))).to(tdt)
    ke = torch.empty(ne * nd * nd, dtype=tdt, device="cuda")
    rho = nn / ne
    alg = nd * nd * s + 4 * a + rho * ((3 + d + 1) * s + d * s)
    if state:
        shp = loss.GetStateShape()
        st = torch.zeros(shp, dtype=tdt, device="cuda")
        st_out = torch.empty_like(st)
        alg += 2 * shp[1] * shp[2] * s
        fn = lambda: loss._assemble(K, u, False, ke_out=ke, state_in=st, state_out=st_out)
        fn()
        plastic = float((st_out[..., -1] > 0).double().mean())
    else:
        fn = lambda: loss._assemble(K, u, False, ke_out=ke)
        plastic = None
    ms = timeit(fn)
    out = {"case": name, "dtype": dtype, "elements": ne, "ms_per_step": ms, "elements_per_s": ne / (ms * 1e-3),
           "algorithmic_bytes_per_element": alg, "achieved_gbs": alg * ne / (ms * 1e-3) / 1e9,
           "frac_of_hbm_copy_bw": alg * ne / (ms * 1e-3) / 1e9 / HBM}
    if plastic is not None:
        out["plastic_point_fraction"] = plastic
    print(json.dumps(out), flush=True)
    del loss, ke


bc3 = {d: {"left": 0.0, "right": 0.1} for d in ("Ux", "Uy", "Uz")}
bc2 = {d: {"left": 0.0, "right": 0.1} for d in ("Ux", "Uy")}
small = lambda x: 0.01 * x
hex128 = folax_b200.create_3D_box_mesh(128, 128, 128, 1.0, 1.0, 1.0)
case("hex128_mech_f32", lf.MechanicalLoss3DHexa, hex128, "hexahedron", {"dirichlet_bc_dict": bc3, "material_dict": MAT}, "float32", small)
case("hex128_thermal_f64", lf.ThermalLoss3DHexa, hex128, "hexahedron", {"dirichlet_bc_dict": {"T": {"left": 1.0, "right": 0.1}}}, "float64", lambda x: 0.5 + 0.1 * x)
del hex128
quad = folax_b200.create_2D_square_mesh(1.0, 2049)
case("quad2048_mech_f64", lf.MechanicalLoss2DQuad, quad, "quad", {"dirichlet_bc_dict": bc2, "material_dict": MAT}, "float64", small)
del quad
tet = folax_b200.create_3D_tetra_box_mesh(70, 70, 70, 1.0, 1.0, 1.0)
case("tet70_neohooke_f64 (config 4)", lf.NeoHookeMechanicalLoss3DTetra, tet, "tetra", {"dirichlet_bc_dict": bc3, "material_dict": MAT}, "float64", lambda x: 0.02 / 70 * x)
case("tet70_mech_f64", lf.MechanicalLoss3DTetra, tet, "tetra", {"dirichlet_bc_dict": bc3, "material_dict": MAT}, "float64", small)
del tet
hex96 = folax_b200.create_3D_box_mesh(96, 96, 96, 1.0, 1.0, 1.0)
case("hex96_neohooke_f64", lf.NeoHookeMechanicalLoss3DHexa, hex96, "hexahedron", {"dirichlet_bc_dict": bc3, "material_dict": MAT}, "float64", lambda x: 0.02 / 96 * x)
case("hex96_j2_f64 (config 5 per-GPU share is 128^3)", lf.ElastoplasticityLoss3DHexa, hex96, "hexahedron", {"dirichlet_bc_dict": bc3, "material_dict": J2MAT}, "float64", lambda x: 0.15 / 96 * x, state=True)
del hex96
hex128 = folax_b200.create_3D_box_mesh(128, 128, 128, 1.0, 1.0, 1.0)
case("hex128_j2_f64 (config 5: per-GPU share of the 256^3 mesh on 8 GPUs)", lf.ElastoplasticityLoss3DHexa, hex128, "hexahedron",
     {"dirichlet_bc_dict": bc3, "material_dict": J2MAT}, "float64", lambda x: 0.15 / 128 * x, state=True)
