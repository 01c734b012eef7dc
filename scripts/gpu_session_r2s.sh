#!/bin/bash
O=gpurun_out/r2s; mkdir -p $O
N=128 timeout 600 ncu --set full --clock-control none --import-source on -k regex:assemble_hex_mech -s 5 -c 1 -o $O/plain python scripts/fused_ab.py > $O/ncu_plain.log 2>&1; echo "rc=$?"
N=128 timeout 600 ncu --set full --clock-control none --import-source on -k regex:assemble_hex_mech -s 360 -c 1 -o $O/fused python scripts/fused_ab.py > $O/ncu_fused.log 2>&1; echo "rc=$?"
