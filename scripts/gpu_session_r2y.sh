#!/bin/bash
O=gpurun_out/r2y; mkdir -p $O
timeout 300 ncu --set full --clock-control none --import-source on -k regex:energy_grid -s 3 -c 1 -o $O/energy_grid_f64 python scripts/energy_variants.py > $O/ncu_f64.log 2>&1; echo "ncu f64 rc=$?"
DTYPE=float32 timeout 300 ncu --set full --clock-control none --import-source on -k regex:energy_grid -s 3 -c 1 -o $O/energy_grid_f32 python scripts/energy_variants.py > $O/ncu_f32.log 2>&1; echo "ncu f32 rc=$?"
ls -la $O
