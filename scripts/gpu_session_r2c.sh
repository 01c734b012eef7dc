#!/bin/bash
O=gpurun_out/r2c; mkdir -p $O
timeout 900 python -m pytest tests/test_zy5_fused_halo_gpu.py tests/test_zy4_hex_j2_tuned_gpu.py tests/test_elastoplasticity_gpu.py tests/test_zy2_config5_slabs_gpu.py tests/test_assembly_gpu.py tests/test_full_size_gpu.py -m gpu -q -x > $O/tests.log 2>&1
echo "tests rc=$?"; tail -15 $O/tests.log
N=128 timeout 300 python scripts/j2_bench.py > $O/j2_bench_128.txt 2>&1; echo "j2 rc=$?"; cat $O/j2_bench_128.txt | tail -4
N=128 timeout 600 ncu --set full --clock-control none --import-source on -k regex:assemble_hex_j2 -s 4 -c 1 -o $O/hex_j2_v2 python scripts/j2_bench.py > $O/ncu_j2.log 2>&1; echo "ncu rc=$?"
N=70 DISP=0.02 LOAD_STEPS=5 timeout 900 python scripts/newton_bench.py > $O/newton_70.json 2> $O/newton_70.err; echo "newton70 rc=$?"; cat $O/newton_70.json; grep "bicgstab info\|Newton" $O/newton_70.err | head -40 | cut -c1-220
