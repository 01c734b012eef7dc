"""Second-order behaviour of ComputeBatchLoss, GPU."""
import pytest

from tests import gpu_helpers as H

pytestmark = pytest.mark.gpu


def test_second_order_semantics():
    """Differentiating THROUGH the gradient (create_graph=True; the latent-code steps of
    meta_implicit_parametric_operator_learning.py:95-105): where the reference keeps the element residual under
    stop_gradient (mechanical.py:116, thermal.py:31, 45-46) the first-order cotangent is a constant of the outer
    differentiation -- and is returned as one; for true potentials (Neo-Hooke) the call must fail loudly."""
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
    loss = H.make_loss("neohooke", "quad", mesh, num_gp=2)
    K, u = H.fields("neohooke", mesh, loss, seed=1, batch=2)
    ut = torch.tensor(u, device="cuda", requires_grad=True)
    mean, _ = loss.ComputeBatchLoss(torch.tensor(K, device="cuda"), ut)
    with pytest.raises(NotImplementedError):
        torch.autograd.grad(mean, ut, create_graph=True)
