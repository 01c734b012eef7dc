"""RegressionLoss: the data (mean-squared-error) loss of fol/loss_functions/regression_loss.py:10-109, for completeness
of the loss family.  It is not on the finite-element hot path and has no kernel of its own: three elementwise /
reduction tensor operations on whatever device the arrays live on."""
import numpy as np
import torch

from .loss import Loss


class RegressionLoss(Loss):
    def __init__(self, name: str, loss_settings: dict, fe_mesh) -> None:
        super().__init__(name)
        self.loss_settings = loss_settings
        self.fe_mesh = fe_mesh
        self.dofs = self.loss_settings["nodal_unknows"]           # (sic) regression_loss.py:41

    def Initialize(self, reinitialize=False) -> None:
        if self.initialized and not reinitialize:
            return
        self.non_dirichlet_indices = np.arange(len(self.dofs) * self.fe_mesh.GetNumberOfNodes())
        self.initialized = True

    def GetFullDofVector(self, known_dofs, unknown_dofs):
        return unknown_dofs

    def GetNumberOfUnknowns(self) -> int:
        return None                                               # `pass` in the reference (:65-66)

    def ComputeBatchLoss(self, gt_values, pred_values):
        """(mean, (min, max, mean)) of the squared error over the flattened (batch, -1) arrays (:69-96)."""
        pred = pred_values if isinstance(pred_values, torch.Tensor) else torch.as_tensor(np.asarray(pred_values))
        gt = gt_values if isinstance(gt_values, torch.Tensor) else torch.as_tensor(np.asarray(gt_values))
        gt = gt.to(device=pred.device, dtype=pred.dtype)
        gt = torch.atleast_2d(gt)
        pred = torch.atleast_2d(pred)
        err = (gt.reshape(gt.shape[0], -1) - pred.reshape(pred.shape[0], -1)) ** 2
        return err.mean(), (err.min(), err.max(), err.mean())

    def Finalize(self) -> None:
        pass
