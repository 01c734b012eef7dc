#!/bin/bash
# Round-2 session m: batched nested VJP tests, Krylov variants at configs[3] size, the full bench at N=1.
O=gpurun_out/r2m; mkdir -p $O
timeout 900 python -m pytest tests/test_apply_jacobian_gpu.py tests/test_zz7_second_order_gpu.py tests/test_batch_loss_gpu.py tests/test_full_size_gpu.py tests/test_zz1_responses_gpu.py -m gpu -q > $O/tests.log 2>&1
echo "tests rc=$?"; tail -5 $O/tests.log
for v in "DEVICE_SCALARS=0" "DEVICE_SCALARS=1 CHECK_EVERY=16" "DEVICE_SCALARS=1 USE_GRAPH=1 CHECK_EVERY=16" "DEVICE_SCALARS=1 USE_GRAPH=1 CHECK_EVERY=64"; do
  env $v N=70 DISP=0.004 LOAD_STEPS=1 timeout 600 python scripts/newton_bench.py 2> $O/newton.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$v', {k:d[k] for k in ('newton_iterations','krylov_iterations','krylov_s','assembly_s','operator_s')}, d['final_residual_norms'])"
done
timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2m/bench_n1.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','bench_wall_s','extra_keys')})
print('roofline', d['roofline']['frac'], d['roofline']['kernel_ms'], 'sustained', d['sustained'])
print('e2e', d['e2e']['value'], d['e2e']['csr'])
print('cpu', d.get('cpu_baseline'))
for k in ('j2','spmv','newton','config1'):
    print(k, json.dumps(d.get(k))[:900])
f=d['fol_loss_grad']; print('fol', {k:f[k] for k in ('value','physics_only_samples_per_s','physics_only_f32_samples_per_s_per_gpu','ms_per_step')}, f['roofline_physics'], f['strong'])
PY
tail -3 $O/bench.err | cut -c1-300
