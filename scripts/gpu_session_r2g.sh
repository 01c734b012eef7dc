#!/bin/bash
O=gpurun_out/r2g; mkdir -p $O
out=$O/energy_ab.jsonl; : > $out
run() { env "$@" timeout 120 python scripts/energy_variants.py >> $out 2>>$O/energy_ab.err; }
run FOL_ENERGY_QT=0
run FOL_ENERGY_QT_MINB=2
run FOL_ENERGY_QT_MINB=3
run FOL_ENERGY_QT_MINB=3 FOL_ENERGY_YCHUNKS=4
run FOL_ENERGY_QT_MINB=3 FOL_ENERGY_YCHUNKS=8
run FOL_ENERGY_QT=0 DTYPE=float32
run FOL_ENERGY_QT_MINB=2 DTYPE=float32
run FOL_ENERGY_QT_MINB=3 DTYPE=float32
run FOL_ENERGY_AFFINE=0 FOL_ENERGY_QT_MINB=3 DTYPE=float32
cut -c1-200 $out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:energy_qt -s 3 -c 1 -o $O/energy_qt_affine_f64 python scripts/energy_variants.py > $O/ncu_f64.log 2>&1; echo "ncu rc=$?"
