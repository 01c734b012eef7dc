#!/bin/bash
# grid kernel: parity tests + timings (float64, float32); NCU=1 adds the two full captures
O=gpurun_out/${TAG:-grid}; mkdir -p $O
timeout 900 python -m pytest tests/test_zy6_energy_grid_gpu.py tests/test_batch_loss_gpu.py -m gpu -q -x > $O/tests.log 2>&1; echo "tests rc=$?"; tail -4 $O/tests.log
for dt in float64 float32; do
    echo "== $dt"; DTYPE=$dt timeout 300 python scripts/energy_variants.py 2>>$O/err | cut -c1-150 | tee -a $O/variants.jsonl
done
if [ -n "$NCU" ]; then
timeout 300 ncu --set full --clock-control none --import-source on -k regex:energy_grid -s 3 -c 1 -o $O/energy_grid_f64 python scripts/energy_variants.py > $O/ncu_f64.log 2>&1; echo "ncu f64 rc=$?"
DTYPE=float32 timeout 300 ncu --set full --clock-control none --import-source on -k regex:energy_grid -s 3 -c 1 -o $O/energy_grid_f32 python scripts/energy_variants.py > $O/ncu_f32.log 2>&1; echo "ncu f32 rc=$?"
fi
