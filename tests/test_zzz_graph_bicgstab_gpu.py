"""First run of the CUDA-graph replay of the device-scalar BiCGSTAB batch (linalg.bicgstab_device(use_graph=True)):
written without GPU time, so it sits last in the suite.  Must equal the eager device-scalar loop and the host-scalar
loop bit for bit, for converged, maxiter-stopped and loose-tolerance runs."""
import numpy as np
import pytest
import torch

from folax_b200 import linalg
from tests import gpu_helpers as gh

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("precond", [None, "jacobi"])
def test_graph_replay_equals_eager(precond):
    mesh = gh.make_mesh("hexahedron", 4, perturb=0.2, seed=3)
    loss = gh.make_loss("mechanical", "hexahedron", mesh, num_gp=2)
    K, _ = gh.fields("mechanical", mesh, loss, seed=1)
    u0 = loss.ApplyDirichletBCOnDofVector(np.zeros(loss.total_number_of_dofs))
    jac, R = loss.ComputeJacobianMatrixAndResidualVector(K, u0)
    A = linalg.SellOperator(loss, jac)
    diag = A.diagonal() if precond else None
    rhs = -R
    for tol, maxiter in ((1e-10, 3000), (1e-10, 20), (1e-3, 3000)):
        x_h, k_h = linalg.bicgstab(A, rhs, x0=u0, tol=tol, atol=0.0, maxiter=maxiter, M_diagonal=diag)
        x_g, k_g = linalg.bicgstab_device(A, rhs, x0=u0, tol=tol, atol=0.0, maxiter=maxiter, M_diagonal=diag,
                                          check_every=4, use_graph=True)
        assert k_g == k_h, (tol, maxiter, k_g, k_h)
        assert torch.equal(x_g, x_h)
