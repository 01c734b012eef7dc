#!/bin/bash
O=gpurun_out/r2n; mkdir -p $O
i=0
for v in "DEVICE_SCALARS=0" "DEVICE_SCALARS=1 CHECK_EVERY=16" "DEVICE_SCALARS=1 USE_GRAPH=1 CHECK_EVERY=16" "DEVICE_SCALARS=1 USE_GRAPH=1 CHECK_EVERY=64"; do
  i=$((i+1))
  env $v N=70 DISP=0.004 LOAD_STEPS=1 timeout 600 python scripts/newton_bench.py > $O/newton_$i.out 2> $O/newton_$i.err
  echo "$v rc=$?"; tail -1 $O/newton_$i.out | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('newton_iterations','krylov_iterations','krylov_s','assembly_s','operator_s')}, d['final_residual_norms'])"
done
