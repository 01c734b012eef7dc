import sys; sys.path.insert(0, "/root/repo")
import numpy as np
from tests import gpu_helpers as gh
from tests.test_zz8_ad_variants_gpu import CLASSES
from oracle import assembly, losses
for law in ("neohooke_ad", "stvenant_ad"):
    for etype in ("hexahedron", "tetra", "quad", "triangle"):
        mesh = gh.make_mesh(etype, 3, perturb=0.2, seed=4)
        dofs = gh.dofs_of("mechanical", etype); d = len(dofs)
        body = [0.2, -0.4, 0.7][:d]
        loss = CLASSES[(law, etype)]("ad", {"dirichlet_bc_dict": {k: {"left": 0.0, "right": 0.1} for k in dofs},
               "material_dict": {"young_modulus": 1.0, "poisson_ratio": 0.3}, "body_foce": body, "dtype": "float32"}, mesh)
        loss.Initialize()
        coords, conn = np.asarray(mesh.GetNodesCoordinates()), mesh.GetElementsNodes(etype)
        nn = len(coords); rng = np.random.default_rng(8)
        K, u = rng.uniform(0.2, 1.0, nn), 0.03 * rng.standard_normal(nn * d)
        g = assembly.element_dof_ids(conn, d)
        bc = np.ones(nn * d); bc[loss.dirichlet_indices] = 0.0
        K32, u32, c32 = K.astype(np.float32).astype(float), u.astype(np.float32).astype(float), coords.astype(np.float32).astype(float)
        en_ref, re_ref, Ke_ref = losses.ad_variant_element(etype, loss.num_gp, c32[conn], K32[conn], u32[g], 0.3, np.array(body), law=law)
        jac, R = loss.ComputeJacobianMatrixAndResidualVector(K, u)
        re_m, Ke_m = assembly.apply_dirichlet(re_ref, Ke_ref, bc[g], False)
        R_ref = np.zeros(nn * d); np.add.at(R_ref, g.reshape(-1), re_m.reshape(-1))
        eK = np.abs(jac.data.cpu().numpy() - Ke_m.reshape(-1)).max() / np.abs(Ke_m).max()
        eR = np.abs(R.cpu().numpy() - R_ref).max() / np.abs(R_ref).max()
        print(law, etype, "Ke rel err %.2e" % eK, "R rel err %.2e" % eR)
