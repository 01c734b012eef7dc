"""folax_b200: B200-native (sm_100a) finite-element residual/Jacobian assembly and batched
physics-loss kernels behind the folax `fol.loss_functions` API.  No CPU fallback."""
from . import _lib
from .mesh import (Mesh, create_2D_square_mesh, create_3D_box_mesh, create_3D_tetra_box_mesh,
                   perturb_interior_nodes)
from .sparse import BCOO

__all__ = ["Mesh", "BCOO", "create_2D_square_mesh", "create_3D_box_mesh", "create_3D_tetra_box_mesh",
           "perturb_interior_nodes"]
