#!/bin/bash
# compute-sanitizer memcheck over the WHOLE single-GPU suite except the full-size (128^3 / 1024-sample) tests
O=gpurun_out/sanitize_suite; mkdir -p $O
timeout 2000 compute-sanitizer --tool memcheck --target-processes all --error-exitcode 7 python -m pytest tests -m gpu -q \
  --deselect tests/test_full_size_gpu.py --deselect tests/test_distributed_gpu.py --deselect tests/test_zz9_slab_solve_gpu.py \
  --deselect tests/test_plan_host_gpu.py -k "not 300-33 and not 16-130 and not chunk_height" > $O/memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" $O/memcheck.log | tail -5
