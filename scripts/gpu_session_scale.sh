#!/bin/bash
# bench.py at N GPUs (N from the environment), as the driver launches it.
N=${NGPUS:-8}; O=gpurun_out/r2scale; mkdir -p $O
nvidia-smi -L | wc -l
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 3 > $O/bench_n$N.json 2> $O/bench_n$N.err
echo "bench n$N rc=$?"; python - <<PY
import json
d=json.load(open('$O/bench_n$N.json'))
print({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches','bench_wall_s')}, d['roofline']['kernel_ms'], d['clocks'])
print('sustained', d['sustained']['value'], 'halo', d.get('halo_check'))
j=d['j2']; print('j2', j.get('value'), j.get('ms_per_step'), j.get('roofline',{}).get('kernel_ms'), j.get('error'))
f=d['fol_loss_grad']; print('fol', f.get('value'), f.get('physics_only_samples_per_s'), f.get('ms_per_step'), f.get('strong'), f.get('error'))
PY
tail -3 $O/bench_n$N.err | cut -c1-300
