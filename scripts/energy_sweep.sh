#!/bin/bash
# GPU sweep of the pipelined batched-loss kernel variants -> gpurun_out/energy_sweep.jsonl
out=gpurun_out/energy_sweep.jsonl; : > $out
run() { env "$@" timeout 120 python scripts/energy_variants.py >> $out 2>>gpurun_out/energy_sweep.err; }
run FOL_ENERGY_V1=1
run FOL_ENERGY_VARIANT=0
run FOL_ENERGY_VARIANT=0
run FOL_ENERGY_VARIANT=1
cut -c1-120 $out
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_throttle_reasons.active --format=csv
