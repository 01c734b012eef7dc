#!/bin/bash
# Round-2 session i: final thermal quad kernel (wide tiles + hoisted phase-B loads): parity, timing, ncu.
O=gpurun_out/r2i; mkdir -p $O
timeout 600 python -m pytest tests/test_batch_loss_gpu.py tests/test_full_size_gpu.py tests/test_implicit_scalar_gpu.py tests/test_zz7_second_order_gpu.py -m gpu -q > $O/tests.log 2>&1
echo "tests rc=$?"; tail -5 $O/tests.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke.log
out=$O/energy_ab.jsonl; : > $out
run() { env "$@" timeout 120 python scripts/energy_variants.py >> $out 2>>$O/energy_ab.err; }
run A=default
run FOL_ENERGY_MAX_ELEMS=192 FOL_ENERGY_TILE_NODES=160
run A=default DTYPE=float32
run FOL_ENERGY_MAX_ELEMS=192 FOL_ENERGY_TILE_NODES=160 DTYPE=float32
cut -c1-230 $out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:energy_qt -s 3 -c 1 -o $O/energy_qt_f64 python scripts/energy_variants.py > $O/ncu_f64.log 2>&1; echo "ncu rc=$?"
DTYPE=float32 timeout 300 ncu --set full --clock-control none --import-source on -k regex:energy_qt -s 3 -c 1 -o $O/energy_qt_f32 python scripts/energy_variants.py > $O/ncu_f32.log 2>&1; echo "ncu rc=$?"
