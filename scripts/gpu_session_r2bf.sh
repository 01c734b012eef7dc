#!/bin/bash
# A/B of the leaner hand-off (elasticity layout 3 vs 2, J2 layout 2 vs 1) + the Hex8 parity tests under the lean layouts
O=gpurun_out/r2bf; mkdir -p $O
for L in 2 3 2 3; do FOL_HEX_LAYOUT=$L timeout 200 python scripts/hex_layout_ab.py mech >> $O/lean_ab.jsonl 2>> $O/ab.err; done
for L in 1 2; do FOL_J2_LAYOUT=$L timeout 200 python scripts/hex_layout_ab.py j2 >> $O/lean_ab.jsonl 2>> $O/ab.err; done
cat $O/lean_ab.jsonl
FOL_HEX_LAYOUT=3 FOL_J2_LAYOUT=2 timeout 200 python -m pytest -q -m gpu tests/test_assembly_gpu.py tests/test_elastoplasticity_gpu.py tests/test_full_size_gpu.py tests/test_golden_gpu.py tests/test_zy2_config5_slabs_gpu.py tests/test_zy4_hex_j2_tuned_gpu.py tests/test_zy5_fused_halo_gpu.py tests/test_zz9_slab_solve_gpu.py > $O/suite_lean.log 2>&1; echo "lean tests rc=$?"; tail -2 $O/suite_lean.log
