"""Second-order behaviour of ComputeBatchLoss, GPU."""
import pytest

from tests import gpu_helpers as H

pytestmark = pytest.mark.gpu


def test_second_order_semantics():
    """Differentiating THROUGH the gradient (create_graph=True; the latent-code steps of
    meta_implicit_parametric_operator_learning.py:95-105): where the reference keeps the element residual under
    stop_gradient (mechanical.py:116, thermal.py:31, 45-46) the first-order cotangent is a constant of the outer
    differentiation -- and is returned as one; for Neo-Hooke the Hessian terms are produced
    (next test); for the other true potentials the call must fail loudly."""
    import torch
    for physics in ("mechanical", "thermal"):
        mesh = H.make_mesh("quad", 4)
        loss = H.make_loss(physics, "quad", mesh, num_gp=2)
        K, u = H.fields(physics, mesh, loss, seed=1, batch=2)
        Kt = torch.tensor(K, device="cuda")
        ut = torch.tensor(u, device="cuda", requires_grad=True)
        w = torch.ones((), dtype=torch.float64, device="cuda", requires_grad=True)
        mean, _ = loss.ComputeBatchLoss(Kt, ut * w)                 # w stands for the network in front of the loss
        (g,) = torch.autograd.grad(mean, ut, create_graph=True)
        # the cotangent itself is a constant, its dependence on w comes from the chain rule in front of the loss only
        (gw,) = torch.autograd.grad((g * g.detach()).sum(), w)
        assert torch.isfinite(gw) and abs(float(gw) - float((g.detach() ** 2).sum())) <= 1e-12 * float((g.detach() ** 2).sum())
    mesh = H.make_mesh("quad", 4)
    loss = H.make_loss("stvenant", "quad", mesh, num_gp=2)      # its reference tangent is not the Hessian of psi
    K, u = H.fields("stvenant", mesh, loss, seed=1, batch=2)
    ut = torch.tensor(u, device="cuda", requires_grad=True)
    mean, _ = loss.ComputeBatchLoss(torch.tensor(K, device="cuda"), ut)
    with pytest.raises(NotImplementedError):
        torch.autograd.grad(mean, ut, create_graph=True)


@pytest.mark.parametrize("etype,num_gp", [("quad", 2), ("tetra", 1), ("hexahedron", 2)])
def test_neo_hooke_hessian_vector_products(etype, num_gp):
    """Neo-Hooke is a true potential whose tangent stiffness is its Hessian: differentiating through the gradient
    gives (1/B) mask K_T(u_b) mask w (in u), (1/B) w^T dF_int/dK (in K) and, through dE/dK, (1/B) mask F_int(u_b; V_b)
    -- against the oracle's assembled tangent, its complex-step control sensitivities and its internal force."""
    import numpy as np
    import torch
    from oracle import assembly, responses
    mesh = H.make_mesh(etype, 3, seed=2)
    loss = H.make_loss("neohooke", etype, mesh, num_gp=num_gp)
    nb = 3
    K, u = H.fields("neohooke", mesh, loss, seed=4, batch=nb)
    rng = np.random.default_rng(6)
    W, V = rng.standard_normal(u.shape), rng.standard_normal(K.shape)
    Kt = torch.tensor(K, device="cuda", requires_grad=True)
    ut = torch.tensor(u, device="cuda", requires_grad=True)
    mean, _ = loss.ComputeBatchLoss(Kt, ut)
    gK, gu = torch.autograd.grad(mean, (Kt, ut), create_graph=True)
    phi = (gu * torch.tensor(W, device="cuda")).sum() + (gK * torch.tensor(V, device="cuda")).sum()
    hK, hu = torch.autograd.grad(phi, (Kt, ut))
    coords, conn = np.asarray(mesh.GetNodesCoordinates()), mesh.GetElementsNodes(etype)
    par = H.oracle_params(loss)
    d = loss.number_dofs_per_node
    ndof, nn = loss.total_number_of_dofs, mesh.GetNumberOfNodes()
    didx = np.asarray(loss.dirichlet_indices, dtype=np.int64)
    g = assembly.element_dof_ids(conn, d)
    U = assembly.full_dof_vector(u, didx, loss.dirichlet_values)
    ref_u, ref_K = np.zeros_like(u), np.zeros_like(K)
    for b in range(nb):
        w = W[b].copy()
        w[didx] = 0.0
        _, _, Ke = assembly.compute_elements("neohooke", etype, num_gp, coords, conn, K[b], U[b], par)
        KT = np.zeros((ndof, ndof))
        np.add.at(KT, (g[:, :, None], g[:, None, :]), Ke)
        _, fint_V, _ = assembly.compute_elements("neohooke", etype, num_gp, coords, conn, V[b], U[b], par)
        fV = np.zeros(ndof)
        np.add.at(fV, g.reshape(-1), fint_V.reshape(-1))
        ref_u[b] = (KT @ w + fV) / nb
        ref_u[b][didx] = 0.0
        rK, _ = responses.residual_adjoint_grads("neohooke", etype, num_gp, coords[conn], K[b][conn], U[b][g], w[g], par)
        np.add.at(ref_K[b], conn.reshape(-1), rK.reshape(-1))
        ref_K[b] /= nb
    assert np.abs(hu.cpu().numpy() - ref_u).max() <= 1e-10 * np.abs(ref_u).max()
    assert np.abs(hK.cpu().numpy() - ref_K).max() <= 1e-10 * np.abs(ref_K).max()
