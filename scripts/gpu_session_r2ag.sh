#!/bin/bash
# 2-GPU regression: the two 2-GPU tests + bench.py at N = 2 as the driver launches it
O=gpurun_out/r2scale; mkdir -p $O
timeout 900 python -m pytest tests/test_distributed_gpu.py tests/test_zz9_slab_solve_gpu.py tests/test_zy5_fused_halo_gpu.py tests/test_zy2_config5_slabs_gpu.py -m gpu -q > $O/tests_2gpu.log 2>&1; echo "2gpu tests rc=$?"; tail -3 $O/tests_2gpu.log
NGPUS=2 bash scripts/gpu_session_scale.sh
