"""Host-side detection of the structured Quad4 grid behind fol_energy_and_grads_grid (energy_plan.grid_structure)."""
import numpy as np

import folax_b200
from folax_b200 import energy_plan
from tests.test_zy6_energy_grid_gpu import grid_mesh


def test_grid_detection():
    assert energy_plan.grid_structure(*_cc(grid_mesh(7, 4))) is not None
    g = energy_plan.grid_structure(*_cc(grid_mesh(7, 4, shear=0.2)))
    assert g is not None and g["jinv"][1] != 0.0
    g = energy_plan.grid_structure(*_cc(grid_mesh(7, 4)))
    assert g["jinv"][1] == 0.0 and g["jinv"][2] == 0.0 and abs(g["wdetj"] - 0.1 * 0.07 / 4) < 1e-15
    m = grid_mesh(7, 4)
    folax_b200.perturb_interior_nodes(m, 0.2, 0)
    assert energy_plan.grid_structure(*_cc(m)) is None                 # not one shape
    m = grid_mesh(7, 4)
    m.elements_nodes["quad"] = m.elements_nodes["quad"][::-1].copy()
    assert energy_plan.grid_structure(*_cc(m)) is None                 # another element order
    sq = folax_b200.create_2D_square_mesh(1.0, 257)
    g = energy_plan.grid_structure(*_cc(sq))
    assert g is not None and (g["nx"], g["ny"]) == (256, 256)          # BASELINE.json configs[2]


def _cc(mesh):
    return np.asarray(mesh.GetNodesCoordinates()), mesh.GetElementsNodes("quad")
