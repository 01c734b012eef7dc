"""CPU oracle for the folax finite-element hot path -- TEST INFRASTRUCTURE ONLY.

This package is a NumPy float64 restatement of the reference arithmetic in
``fol/loss_functions`` / ``fol/geometries`` / ``fol/constitutive_material_models``
(Neural-Mechanics-Lab/folax).  It exists to check the CUDA kernels, not to serve users:

* only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
  ``--impl reference`` legs may import it;
* nothing under ``folax_b200/`` imports it, and the product path raises when the CUDA
  library is missing instead of falling back to this code.

Parity pinning: JAX is not installable in this image, so the reference cannot be executed
here; the oracle is instead pinned against the literal known-answer arrays of the
reference's own unit tests (``tests/golden/reference_unit_goldens.json``, extracted by
``tests/golden/make_golden.py``) -- see ``tests/test_oracle_golden.py``.
"""
