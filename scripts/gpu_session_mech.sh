#!/bin/bash
O=gpurun_out/mech; mkdir -p $O
timeout 900 python -m pytest tests/test_zy6_energy_grid_gpu.py tests/test_batch_loss_gpu.py tests/test_zz7_second_order_gpu.py -m gpu -q -x > $O/tests.log 2>&1; echo "tests rc=$?"; tail -5 $O/tests.log
for dt in float64 float32; do for g in 1 0; do
  echo "== mech $dt grid=$g"; PHYS=mech DTYPE=$dt FOL_ENERGY_GRID=$g timeout 300 python scripts/energy_variants.py 2>>$O/err | cut -c1-120
done; done
tail -3 $O/err
