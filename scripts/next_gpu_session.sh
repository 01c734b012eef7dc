#!/bin/bash
# One gpurun call for the first GPU session after round 1's late additions (adjoint responses, solvers / Krylov
# kernels, nested VJP, configs[3]/[4] tests were written after the round's GPU minutes were spent):
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash scripts/next_gpu_session.sh'
# 1. the not-yet-run GPU tests first (short, the news), 2. the whole GPU suite, 3. timings of the new kernels,
# 4. the headline bench, 5. launch list + one full ncu capture of the SELL SpMV.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
python -m pytest tests/test_zz1_responses_gpu.py tests/test_zz2_solvers_gpu.py tests/test_zz7_second_order_gpu.py \
       tests/test_zy1_config3_newton_gpu.py tests/test_zy2_config5_slabs_gpu.py tests/test_zz8_ad_variants_gpu.py tests/test_zy3_kratos_class_gpu.py tests/test_zz5_element_api_gpu.py tests/test_zzz_graph_bicgstab_gpu.py -m gpu -q > gpurun_out/new_tests.log 2>&1
echo "new tests rc=$?"; tail -5 gpurun_out/new_tests.log
python -m pytest tests -m gpu -q -x > gpurun_out/all_tests.log 2>&1
echo "all tests rc=$?"; tail -3 gpurun_out/all_tests.log
python examples/sensitivity_analysis_2D.py 81 > gpurun_out/example_sa.log 2>&1; echo "example rc=$?"; tail -6 gpurun_out/example_sa.log
python examples/neo_hooke_newton_3D.py 12 > gpurun_out/example_newton.log 2>&1; echo "example rc=$?"; tail -7 gpurun_out/example_newton.log | cut -c1-200
N=${N:-128} python scripts/solver_bench.py > gpurun_out/solver_bench.json 2> gpurun_out/solver_bench.err
echo "solver bench rc=$?"; cat gpurun_out/solver_bench.json
N=${NT:-70} python scripts/newton_bench.py > gpurun_out/newton_bench.json 2> gpurun_out/newton_bench.err
echo "newton bench rc=$?"; cat gpurun_out/newton_bench.json
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench.err
echo "bench rc=$?"; cut -c1-600 gpurun_out/bench_n1.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/solver_launches.csv \
    env N=64 MAXITER=20 python scripts/solver_bench.py > gpurun_out/solver_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sell_spmv_kernel -s 3 -c 1 -o gpurun_out/sell_spmv \
    env N=${N:-128} MAXITER=5 python scripts/solver_bench.py > gpurun_out/solver_ncu_full.log 2>&1
echo "ncu rc=$?"
