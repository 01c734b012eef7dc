#!/bin/bash
# Round-2 session j (2 GPUs): the 2-GPU tests (fused + layered halo, J2 slabs, slab solve) and bench.py at N=2.
O=gpurun_out/r2j; mkdir -p $O
nvidia-smi -L > $O/gpus.txt
timeout 900 python -m pytest tests/test_distributed_gpu.py tests/test_zz9_slab_solve_gpu.py -m gpu -q > $O/tests_2gpu.log 2>&1
echo "2gpu tests rc=$?"; tail -12 $O/tests_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err
echo "bench n2 rc=$?"; cut -c1-2500 $O/bench_n2.json; tail -4 $O/bench_n2.err | cut -c1-300
