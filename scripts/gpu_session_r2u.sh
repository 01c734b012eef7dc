#!/bin/bash
O=gpurun_out/r2u; mkdir -p $O
timeout 900 python -m pytest tests/test_zy5_fused_halo_gpu.py tests/test_zy4_hex_j2_tuned_gpu.py tests/test_assembly_gpu.py tests/test_full_size_gpu.py -m gpu -q > $O/tests.log 2>&1; echo "tests rc=$?"; tail -2 $O/tests.log
timeout 900 python -m pytest tests/test_distributed_gpu.py tests/test_zz9_slab_solve_gpu.py -m gpu -q > $O/tests_2gpu.log 2>&1; echo "2gpu tests rc=$?"; tail -2 $O/tests_2gpu.log
N=128 timeout 300 python scripts/fused_ab.py > $O/fused_ab.json 2>$O/err; cat $O/fused_ab.json
N=128 timeout 600 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none -k regex:assemble_hex_mech -s 360 -c 1 python scripts/fused_ab.py 2>&1 | grep -E "inst_executed|duration" | head -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-extras > $O/bench_n2.json 2> $O/bench_n2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2u/bench_n2.json')); print('n2', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['per_rank'], d['halo_check']['bitwise_equal'], d['sustained']['value'])
PY
timeout 600 python bench.py --no-extras > $O/bench_n1.json 2>>$O/err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2u/bench_n1.json')); print('n1', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['sustained']['value'])
PY
