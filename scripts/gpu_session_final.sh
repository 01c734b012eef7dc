#!/bin/bash
# final validation of the round: whole GPU suite (no -x), smoke(), bench.py both arms, launch list of the bench command
O=gpurun_out/final; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/gpu_suite.log 2>&1; echo "suite rc=$?"; tail -3 $O/gpu_suite.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference > $O/bench_n1_reference.json 2> $O/bench_n1_reference.err; echo "bench ref rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/final/bench_n1.json'))
f=d['fol_loss_grad']
print('value', d['value'], 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], d['clocks'])
print('fol', f['value'], 'phys', f['physics_only_samples_per_s'], 'f32', f['physics_only_f32_samples_per_s_per_gpu'], f['kernel_only']['f64_ms'], f['kernel_only']['f32_ms'], f['roofline_physics']['frac'], f['f32_network_f32_physics']['value'])
print('newton', d['newton']['per_newton_iteration_ms'], d['newton'].get('krylov_ms_per_iteration'), d['newton'].get('krylov_roofline'))
print('j2', d['j2']['value'], d['j2']['config']['global_mesh'], 'config1', d['config1']['assemble_and_solve_ms'], 'wall', d['bench_wall_s'])
r=json.load(open('gpurun_out/final/bench_n1_reference.json')); print('ref', r['value'])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_n1.csv python bench.py --no-extras --steps 2 --warmup 1 > $O/ncu_launch.log 2>&1; echo "launch list rc=$?"
python profiles/launch_summary.py $O/launches_n1.csv 2>/dev/null | head -12
