#!/bin/bash
# Round-2 session l (2 GPUs): symmetric-tile kernels: parity tests on one GPU, 2-GPU tests, fused A/B, bench N=1 and N=2.
O=gpurun_out/r2l; mkdir -p $O
timeout 900 python -m pytest tests/test_assembly_gpu.py tests/test_full_size_gpu.py tests/test_zy4_hex_j2_tuned_gpu.py tests/test_zy5_fused_halo_gpu.py tests/test_elastoplasticity_gpu.py tests/test_golden_gpu.py -m gpu -q > $O/tests.log 2>&1
echo "tests rc=$?"; tail -5 $O/tests.log
timeout 900 python -m pytest tests/test_distributed_gpu.py tests/test_zz9_slab_solve_gpu.py -m gpu -q > $O/tests_2gpu.log 2>&1
echo "2gpu tests rc=$?"; tail -5 $O/tests_2gpu.log
N=128 timeout 300 python scripts/fused_ab.py > $O/fused_ab.json 2> $O/fused_ab.err; echo "rc=$?"; cat $O/fused_ab.json
N=128 timeout 300 python scripts/j2_bench.py > $O/j2_bench_128.txt 2>&1; tail -3 $O/j2_bench_128.txt
timeout 900 python bench.py --no-extras > $O/bench_n1_noextras.json 2> $O/bench_n1.err; echo "bench n1 rc=$?"; cut -c1-2200 $O/bench_n1_noextras.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err
echo "bench n2 rc=$?"; cut -c1-4500 $O/bench_n2.json; tail -2 $O/bench_n2.err | cut -c1-300
