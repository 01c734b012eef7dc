#!/bin/bash
O=gpurun_out/r2h; mkdir -p $O
out=$O/energy_ab.jsonl; : > $out
run() { env "$@" timeout 120 python scripts/energy_variants.py >> $out 2>>$O/energy_ab.err; }
run FOL_ENERGY_QT_BLOCK=0
run FOL_ENERGY_QT_BLOCK=256
run FOL_ENERGY_MAX_ELEMS=256 FOL_ENERGY_TILE_NODES=216
run FOL_ENERGY_MAX_ELEMS=256 FOL_ENERGY_TILE_NODES=200
run FOL_ENERGY_MAX_ELEMS=256 FOL_ENERGY_TILE_NODES=225
run FOL_ENERGY_QT_BLOCK=256 DTYPE=float32
run FOL_ENERGY_MAX_ELEMS=256 FOL_ENERGY_TILE_NODES=216 DTYPE=float32
cut -c1-260 $out; tail -3 $O/energy_ab.err
FOL_ENERGY_MAX_ELEMS=256 FOL_ENERGY_TILE_NODES=216 timeout 300 ncu --set full --clock-control none --import-source on -k regex:energy_qt -s 3 -c 1 -o $O/energy_qt_wide_f64 python scripts/energy_variants.py > $O/ncu_f64.log 2>&1; echo "ncu rc=$?"
