#!/bin/bash
O=gpurun_out/phys; mkdir -p $O
timeout 900 python -m pytest tests/test_batch_loss_gpu.py tests/test_zz7_second_order_gpu.py tests/test_zy6_energy_grid_gpu.py tests/test_implicit_scalar_gpu.py -m gpu -q > $O/tests.log 2>&1; echo "tests rc=$?"; tail -2 $O/tests.log
for dt in float64 float32; do DTYPE=$dt timeout 300 python scripts/energy_variants.py 2>>$O/err | cut -c1-100; done
B=128 timeout 300 python scripts/physics_step_profile.py 2>>$O/err | head -2
