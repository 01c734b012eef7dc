"""Duplicate-keeping COO Jacobian with the surface callers use on jax.experimental.sparse.BCOO
(fe_loss.py:316; consumers touch .data, .indices[:,0/1], .shape, @ vector, .todense():
fe_solver.py:61, 71-72 and the unit tests)."""
import torch


class BCOO:
    def __init__(self, args, shape):
        self.data, self.indices = args
        self.shape = tuple(shape)

    @property
    def dtype(self):
        return self.data.dtype

    @property
    def nse(self):
        return self.data.shape[0]

    def todense(self):
        n, m = self.shape
        out = torch.zeros(n * m, dtype=self.data.dtype, device=self.data.device)
        flat = self.indices[:, 0].to(torch.int64) * m + self.indices[:, 1].to(torch.int64)
        out.index_add_(0, flat, self.data)
        return out.view(n, m)

    def __matmul__(self, v):
        v = torch.as_tensor(v, dtype=self.data.dtype, device=self.data.device)
        rows = self.indices[:, 0].to(torch.int64)
        cols = self.indices[:, 1].to(torch.int64)
        if v.dim() == 1:
            out = torch.zeros(self.shape[0], dtype=self.data.dtype, device=self.data.device)
            return out.index_add_(0, rows, self.data * v[cols])
        out = torch.zeros((self.shape[0], v.shape[1]), dtype=self.data.dtype, device=self.data.device)
        return out.index_add_(0, rows, self.data[:, None] * v[cols])

    @property
    def T(self):
        return BCOO((self.data, self.indices.flip(1)), (self.shape[1], self.shape[0]))

    def to_scipy_csr(self):
        """Host hand-off used by the reference's direct solvers (fe_solver.py:71-72)."""
        import scipy.sparse as sp
        idx = self.indices.cpu().numpy()
        return sp.csr_array((self.data.cpu().numpy(), (idx[:, 0], idx[:, 1])), shape=self.shape)
