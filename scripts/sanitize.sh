#!/bin/bash
# compute-sanitizer (memcheck, racecheck, synccheck) over the tests of this round's kernels at their small sizes.
O=gpurun_out/sanitize; mkdir -p $O
T="tests/test_zy4_hex_j2_tuned_gpu.py::test_tuned_kernel_matches_oracle_over_two_load_steps tests/test_zy5_fused_halo_gpu.py tests/test_batch_loss_gpu.py tests/test_apply_jacobian_gpu.py::test_batched_products_equal_per_sample_products tests/test_elastoplasticity_gpu.py"
A="tests/test_assembly_gpu.py::test_tuned_hex_f32_kernel_matches_oracle_and_generic tests/test_assembly_gpu.py::test_tuned_hex_kernel_matches_oracle_and_generic"
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --target-processes all --error-exitcode 7 python -m pytest $T $A -m gpu -q -x -k "not 64 and not 16-9-11 and not 16-16-12" > $O/$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" $O/$tool.log | tail -4
done
