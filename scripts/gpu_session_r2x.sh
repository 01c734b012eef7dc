#!/bin/bash
# structured-grid thermal loss + VJP kernel: parity tests, then A/B against the tile kernels (f64, f32), chunk heights
O=gpurun_out/r2x; mkdir -p $O
timeout 900 python -m pytest tests/test_zy6_energy_grid_gpu.py tests/test_batch_loss_gpu.py -m gpu -q -x > $O/tests.log 2>&1; echo "tests rc=$?"; tail -15 $O/tests.log
for dt in float64 float32; do
  for grid in 0 1; do
    echo "== $dt grid=$grid"; DTYPE=$dt FOL_ENERGY_GRID=$grid timeout 300 python scripts/energy_variants.py 2>>$O/err | tee -a $O/variants.jsonl
  done
done
for rows in 16 24 32 43 64 86 128; do
  echo "== f64 rows=$rows"; FOL_ENERGY_GRID_ROWS=$rows timeout 300 python scripts/energy_variants.py 2>>$O/err | head -1 | tee -a $O/rows.jsonl
done
tail -5 $O/err
