#!/bin/bash
# GPU sweep of the pipelined batched-loss kernel variants -> gpurun_out/energy_sweep.jsonl
out=gpurun_out/energy_sweep.jsonl; : > $out
run() { env "$@" timeout 120 python scripts/energy_variants.py >> $out 2>>gpurun_out/energy_sweep.err; }
run FOL_ENERGY_V1=1
run FOL_ENERGY_VARIANT=0
run FOL_ENERGY_VARIANT=1
run FOL_ENERGY_VARIANT=3 FOL_ENERGY_MAX_ELEMS=160 FOL_ENERGY_TILE_NODES=128
run FOL_ENERGY_VARIANT=0 DTYPE=float32
run FOL_ENERGY_V1=1 DTYPE=float32
cut -c1-200 $out
