#!/bin/bash
# two GPUs: the 2-GPU tests and bench.py --gpus 2 with the dense layouts as defaults (fused in-kernel halo push variants)
O=gpurun_out/r2bc; mkdir -p $O
timeout 600 python -m pytest tests/test_distributed_gpu.py -m gpu -q > $O/gpu_tests_2gpu.log 2>&1; echo "2gpu tests rc=$?"; tail -3 $O/gpu_tests_2gpu.log
NGPUS=2 bash scripts/gpu_session_scale.sh
cp gpurun_out/r2scale/bench_n2.json gpurun_out/r2scale/bench_n2.err $O/ 2>/dev/null
