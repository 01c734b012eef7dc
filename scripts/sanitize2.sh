#!/bin/bash
# compute-sanitizer (memcheck, racecheck, synccheck) over the kernels of the second half of round 2 at small sizes:
# structured-grid loss + VJP (cp.async.bulk / mbarrier ring), one-launch BiCGSTAB, tuned Hex8 thermal / Quad4 elasticity.
O=gpurun_out/sanitize2; mkdir -p $O
T="tests/test_zy6_energy_grid_gpu.py::test_grid_kernel_matches_the_oracle tests/test_zy6_energy_grid_gpu.py::test_grid_kernel_float32_sample_pairs tests/test_zy6_energy_grid_gpu.py::test_dirichlet_on_interior_and_top_rows tests/test_zzz2_fused_bicgstab_gpu.py::test_fused_solve_stops_at_maxiter_and_on_a_converged_start tests/test_zzz2_fused_bicgstab_gpu.py::test_fused_solve_float32 tests/test_assembly_gpu.py::test_tuned_hex_thermal_kernel_matches_oracle_and_generic tests/test_assembly_gpu.py::test_tuned_quad_mech_kernel_matches_oracle_and_generic"
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --target-processes all --error-exitcode 7 python -m pytest $T -m gpu -q -x -k "not 16-9-11 and not 300-33 and not 16-130 and not 256-12 and not 257-5" > $O/$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" $O/$tool.log | tail -4
done
