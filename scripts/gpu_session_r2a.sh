#!/bin/bash
# Round-2 first GPU session: the whole -m gpu suite WITHOUT -x (the count VERDICT item 1 asks for), then the timings
# of the (f) rows that never ran on hardware, then the headline bench.  Output in gpurun_out/r2a/.
O=gpurun_out/r2a; mkdir -p $O
nvidia-smi -L > $O/gpus.txt
timeout 900 python -m pytest tests -m gpu -q > $O/all_tests.log 2>&1
echo "all tests rc=$?"; tail -15 $O/all_tests.log
N=${N:-128} timeout 600 python scripts/solver_bench.py > $O/solver_bench.json 2> $O/solver_bench.err
echo "solver bench rc=$?"; cat $O/solver_bench.json; tail -3 $O/solver_bench.err
N=${NT:-70} timeout 600 python scripts/newton_bench.py > $O/newton_bench.json 2> $O/newton_bench.err
echo "newton bench rc=$?"; cat $O/newton_bench.json; tail -3 $O/newton_bench.err
timeout 300 python scripts/j2_bench.py > $O/j2_bench.txt 2>&1; echo "j2 rc=$?"; tail -8 $O/j2_bench.txt
timeout 600 python bench.py > $O/bench_n1.json 2> $O/bench.err
echo "bench rc=$?"; cut -c1-1500 $O/bench_n1.json
