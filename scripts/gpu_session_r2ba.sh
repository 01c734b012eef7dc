#!/bin/bash
# A/B of the denser layouts of the two tuned Hex8 f64 kernels (elasticity: 16 instead of 12 warps / SM, J2: 12 instead
# of 10), parity suite under the new layouts, ncu capture of the dense elasticity kernel
O=gpurun_out/r2ba; mkdir -p $O
for L in 0 2 0 2; do FOL_HEX_LAYOUT=$L timeout 300 python scripts/hex_layout_ab.py mech >> $O/ab.jsonl 2>> $O/ab.err; done
for L in 0 1 0 1; do FOL_J2_LAYOUT=$L timeout 300 python scripts/hex_layout_ab.py j2 >> $O/ab.jsonl 2>> $O/ab.err; done
FOL_HEX_LAYOUT=2 FOL_HEX_DIAG=nostore timeout 300 python scripts/hex_layout_ab.py mech >> $O/ab.jsonl 2>> $O/ab.err
cat $O/ab.jsonl
FOL_HEX_LAYOUT=2 FOL_J2_LAYOUT=1 timeout 900 python -m pytest tests -m gpu -q > $O/suite_dense.log 2>&1; echo "suite rc=$?"; tail -3 $O/suite_dense.log
FOL_HEX_LAYOUT=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:assemble_hex_mech -s 5 -c 1 -o $O/hex_dense python bench.py --no-extras --steps 3 > $O/ncu_hex.log 2>&1; echo "ncu rc=$?"
FOL_J2_LAYOUT=1 N=128 timeout 600 ncu --set full --clock-control none --import-source on -k regex:assemble_hex_j2 -s 4 -c 1 -o $O/j2_dense python scripts/j2_bench.py > $O/ncu_j2.log 2>&1; echo "ncu j2 rc=$?"
ls -la $O
