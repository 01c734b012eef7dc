#!/bin/bash
# full GPU suite (grid kernel now serves every structured thermal quad mesh), bench N=1 both arms, launch list of a physics step
O=gpurun_out/r2ae; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/gpu_suite.log 2>&1; echo "suite rc=$?"; tail -4 $O/gpu_suite.log
timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2ae/bench_n1.json'))
f=d['fol_loss_grad']
print('value', d['value'], 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'])
print('fol', f['value'], 'phys', f['physics_only_samples_per_s'], 'f32', f['physics_only_f32_samples_per_s_per_gpu'], f['kernel_only'], f['roofline_physics']['frac'], f['roofline_physics']['executed'])
print('newton', d['newton']['per_newton_iteration_ms'], 'config1', d['config1'])
print('wall', d['bench_wall_s'])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_physics.csv python scripts/energy_variants.py > $O/ncu_launches.log 2>&1
python profiles/launch_summary.py $O/launches_physics.csv 2>/dev/null | head -20
