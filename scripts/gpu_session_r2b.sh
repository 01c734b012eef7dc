#!/bin/bash
# Round-2 session b: the tuned J2 kernel (tests, timing at 64^3 and 128^3, ncu), Newton bench debug.
O=gpurun_out/r2b; mkdir -p $O
timeout 600 python -m pytest tests/test_zy4_hex_j2_tuned_gpu.py tests/test_elastoplasticity_gpu.py tests/test_zy2_config5_slabs_gpu.py tests/test_batch_loss_gpu.py -m gpu -q > $O/j2_tests.log 2>&1
echo "j2 tests rc=$?"; tail -15 $O/j2_tests.log
N=64 timeout 300 python scripts/j2_bench.py > $O/j2_bench_64.txt 2>&1; echo "j2 rc=$?"; cat $O/j2_bench_64.txt | tail -4
N=128 timeout 300 python scripts/j2_bench.py > $O/j2_bench_128.txt 2>&1; echo "j2 rc=$?"; cat $O/j2_bench_128.txt | tail -4
N=128 timeout 600 ncu --set full --clock-control none --import-source on -k regex:assemble_hex_j2 -s 4 -c 1 -o $O/hex_j2 python scripts/j2_bench.py > $O/ncu_j2.log 2>&1; echo "ncu rc=$?"
N=40 LOAD_STEPS=2 timeout 600 python scripts/newton_bench.py > $O/newton_40.json 2> $O/newton_40.err; echo "newton40 rc=$?"; cat $O/newton_40.json; grep -c bicgstab $O/newton_40.err; grep "bicgstab info" $O/newton_40.err | head -12
N=70 LOAD_STEPS=4 timeout 900 python scripts/newton_bench.py > $O/newton_70.json 2> $O/newton_70.err; echo "newton70 rc=$?"; cat $O/newton_70.json; grep "bicgstab info\|Newton" $O/newton_70.err | head -30 | cut -c1-200
