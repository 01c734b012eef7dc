#!/bin/bash
O=gpurun_out/r2z; mkdir -p $O
timeout 900 python -m pytest tests/test_zy6_energy_grid_gpu.py tests/test_batch_loss_gpu.py -m gpu -q -x > $O/tests.log 2>&1; echo "tests rc=$?"; tail -15 $O/tests.log
for dt in float64 float32; do
    echo "== $dt"; DTYPE=$dt timeout 300 python scripts/energy_variants.py 2>>$O/err | tee -a $O/variants.jsonl
done
for rows in 32 43 64 86 128; do
  echo "== f64 rows=$rows"; FOL_ENERGY_GRID_ROWS=$rows timeout 300 python scripts/energy_variants.py 2>>$O/err | head -1 | cut -c1-120 | tee -a $O/rows.jsonl
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:energy_grid -s 3 -c 1 -o $O/energy_grid_f64 python scripts/energy_variants.py > $O/ncu_f64.log 2>&1; echo "ncu f64 rc=$?"
DTYPE=float32 timeout 300 ncu --set full --clock-control none --import-source on -k regex:energy_grid -s 3 -c 1 -o $O/energy_grid_f32 python scripts/energy_variants.py > $O/ncu_f32.log 2>&1; echo "ncu f32 rc=$?"
tail -5 $O/err
