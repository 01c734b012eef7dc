#!/bin/bash
# Round-2 session d: CSR hand-off test, J2 v3 timing, the new bench.py at N=1 (both arms), ncu captures for the traffic files.
O=gpurun_out/r2d; mkdir -p $O
timeout 600 python -m pytest tests/test_plan_host_gpu.py tests/test_zy4_hex_j2_tuned_gpu.py tests/test_zy5_fused_halo_gpu.py -m gpu -q > $O/tests.log 2>&1
echo "tests rc=$?"; tail -8 $O/tests.log
N=128 timeout 300 python scripts/j2_bench.py > $O/j2_bench_128.txt 2>&1; echo "j2 rc=$?"; cat $O/j2_bench_128.txt | tail -4
timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench.err; echo "bench rc=$?"; cut -c1-3000 $O/bench_n1.json; tail -5 $O/bench.err
timeout 600 python bench.py --impl reference --steps 10 --warmup 1 > $O/bench_reference.json 2> $O/bench_ref.err; echo "ref rc=$?"; cut -c1-600 $O/bench_reference.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:assemble_hex_mech -s 5 -c 1 -o $O/assemble_hex python bench.py --no-extras --steps 3 > $O/ncu_hex.log 2>&1; echo "ncu hex rc=$?"
N=128 timeout 600 ncu --set full --clock-control none --import-source on -k regex:assemble_hex_j2 -s 4 -c 1 -o $O/hex_j2_v3 python scripts/j2_bench.py > $O/ncu_j2.log 2>&1; echo "ncu j2 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --no-extras --steps 2 --warmup 3 > $O/launches.log 2>&1; echo "launch list rc=$?"
