#!/bin/bash
# Round-2 session f: affine variant of the sample-vectorised thermal quad kernel: parity + A/B.
O=gpurun_out/r2f; mkdir -p $O
timeout 600 python -m pytest tests/test_batch_loss_gpu.py tests/test_full_size_gpu.py -m gpu -q > $O/tests.log 2>&1
echo "tests rc=$?"; tail -8 $O/tests.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
out=$O/energy_ab.jsonl; : > $out
run() { env "$@" timeout 120 python scripts/energy_variants.py >> $out 2>>$O/energy_ab.err; }
run FOL_ENERGY_QT=0
run FOL_ENERGY_AFFINE=0
run FOL_ENERGY_QT_MINB=2
run FOL_ENERGY_QT_MINB=3
run FOL_ENERGY_QT=0 DTYPE=float32
run FOL_ENERGY_AFFINE=0 DTYPE=float32
run FOL_ENERGY_QT_MINB=2 DTYPE=float32
run FOL_ENERGY_QT_MINB=3 DTYPE=float32
cut -c1-300 $out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:energy_qt -s 3 -c 1 -o $O/energy_qt_affine_f64 python scripts/energy_variants.py > $O/ncu_f64.log 2>&1; echo "ncu rc=$?"
DTYPE=float32 timeout 300 ncu --set full --clock-control none --import-source on -k regex:energy_qt -s 3 -c 1 -o $O/energy_qt_affine_f32 python scripts/energy_variants.py > $O/ncu_f32.log 2>&1; echo "ncu rc=$?"
