#!/bin/bash
O=gpurun_out/r2af; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/gpu_suite.log 2>&1; echo "suite rc=$?"; tail -3 $O/gpu_suite.log
timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference > $O/bench_n1_reference.json 2> $O/bench_n1_reference.err; echo "bench ref rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2af/bench_n1.json'))
f=d['fol_loss_grad']
print('value', d['value'], 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'])
print('fol', f['value'], 'phys', f['physics_only_samples_per_s'], 'f32', f['physics_only_f32_samples_per_s_per_gpu'], f['kernel_only'], f['roofline_physics']['frac'])
print('newton', d['newton']['per_newton_iteration_ms'], d['newton'].get('krylov_ms_per_iteration'), d['newton']['krylov_iterations'])
print('wall', d['bench_wall_s'])
r=json.load(open('gpurun_out/r2af/bench_n1_reference.json')); print('ref', r['value'], r.get('cpu_baseline'))
PY
DTYPE=float32 timeout 300 ncu --set full --clock-control none --import-source on -k regex:energy_grid -s 3 -c 1 -o $O/energy_grid_f32 python scripts/energy_variants.py > $O/ncu_f32.log 2>&1; echo "ncu f32 rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:energy_grid -s 3 -c 1 -o $O/energy_grid_f64 python scripts/energy_variants.py > $O/ncu_f64.log 2>&1; echo "ncu f64 rc=$?"
