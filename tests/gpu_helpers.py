"""Shared builders for the GPU parity tests: seeded meshes / fields and loss construction."""
import numpy as np

import folax_b200
from folax_b200 import loss_functions as lf

MATERIAL = {"young_modulus": 1.0, "poisson_ratio": 0.3}

LOSS_CLASSES = {
    ("mechanical", "hexahedron"): lf.MechanicalLoss3DHexa, ("mechanical", "quad"): lf.MechanicalLoss2DQuad,
    ("mechanical", "tetra"): lf.MechanicalLoss3DTetra, ("mechanical", "triangle"): lf.MechanicalLoss2DTri,
    ("thermal", "hexahedron"): lf.ThermalLoss3DHexa, ("thermal", "quad"): lf.ThermalLoss2DQuad,
    ("thermal", "tetra"): lf.ThermalLoss3DTetra, ("thermal", "triangle"): lf.ThermalLoss2DTri,
    ("neohooke", "hexahedron"): lf.NeoHookeMechanicalLoss3DHexa, ("neohooke", "quad"): lf.NeoHookeMechanicalLoss2DQuad,
    ("neohooke", "tetra"): lf.NeoHookeMechanicalLoss3DTetra, ("neohooke", "triangle"): lf.NeoHookeMechanicalLoss2DTri,
    ("stvenant", "hexahedron"): lf.SaintVenantMechanicalLoss3DHexa, ("stvenant", "quad"): lf.SaintVenantMechanicalLoss2DQuad,
    ("stvenant", "tetra"): lf.SaintVenantMechanicalLoss3DTetra, ("stvenant", "triangle"): lf.SaintVenantMechanicalLoss2DTri,
}


def triangle_mesh(N, L=1.0):
    """Quad mesh split into two counter-clockwise triangles per cell."""
    q = folax_b200.create_2D_square_mesh(L, N)
    c = q.elements_nodes["quad"]
    tri = np.concatenate([c[:, [0, 1, 2]], c[:, [0, 2, 3]]], axis=1).reshape(-1, 3)
    q.elements_nodes = {"triangle": np.ascontiguousarray(tri, dtype=np.int32)}
    return q


def make_mesh(element_type, n, perturb=0.2, seed=0):
    if element_type == "hexahedron":
        m = folax_b200.create_3D_box_mesh(n, n + 1, n, 1.0, 1.2, 0.9)
    elif element_type == "tetra":
        m = folax_b200.create_3D_tetra_box_mesh(n, n, n + 1, 1.0, 1.0, 1.3)
    elif element_type == "quad":
        m = folax_b200.create_2D_square_mesh(1.0, n + 1)
    else:
        m = triangle_mesh(n + 1)
    if perturb:
        folax_b200.perturb_interior_nodes(m, perturb, seed)
    return m


def dofs_of(physics, element_type):
    if physics == "thermal":
        return ["T"]
    return ["Ux", "Uy", "Uz"] if element_type in ("hexahedron", "tetra") else ["Ux", "Uy"]


def bc_dict(physics, element_type):
    if physics == "thermal":
        return {"T": {"left": 1.0, "right": 0.1}}
    return {d: {"left": 0.0, "right": 0.1} for d in dofs_of(physics, element_type)}


def make_loss(physics, element_type, mesh, num_gp=None, dtype="float64", extra=None):
    settings = {"dirichlet_bc_dict": bc_dict(physics, element_type), "dtype": dtype}
    if physics != "thermal":
        settings["material_dict"] = dict(MATERIAL)
    if num_gp is not None:
        settings["num_gp"] = num_gp
    settings.update(extra or {})
    loss = LOSS_CLASSES[(physics, element_type)](f"{physics}_{element_type}", settings, mesh)
    loss.Initialize()
    return loss


def fields(physics, mesh, loss, seed=0, batch=None):
    rng = np.random.default_rng(seed)
    nn, ndof = mesh.GetNumberOfNodes(), loss.total_number_of_dofs
    shape_k = (nn,) if batch is None else (batch, nn)
    shape_u = (ndof,) if batch is None else (batch, ndof)
    K = rng.uniform(0.1, 1.0, shape_k)
    if physics == "thermal":
        u = rng.uniform(0.1, 1.0, shape_u)
    elif physics in ("neohooke", "stvenant"):
        h = 1.0 / max(2, round(nn ** (1.0 / loss.dim)))
        u = 0.02 * h * rng.standard_normal(shape_u)
    else:
        u = 0.01 * rng.standard_normal(shape_u)
    return K, u


def oracle_params(loss):
    p = {}
    if loss.physics != "thermal":
        md = loss.loss_settings["material_dict"]
        p.update(young_modulus=md["young_modulus"], poisson_ratio=md["poisson_ratio"])
        if "body_foce" in loss.loss_settings:
            p["body_force"] = np.asarray(loss.loss_settings["body_foce"], float).reshape(-1)
    else:
        p.update(beta=loss.thermal_loss_settings["beta"], c=loss.thermal_loss_settings["c"])
    return p
