#!/bin/bash
O=gpurun_out/r2t; mkdir -p $O
timeout 600 python -m pytest tests/test_zy5_fused_halo_gpu.py -m gpu -q > $O/tests.log 2>&1; echo "tests rc=$?"; tail -2 $O/tests.log
N=128 timeout 300 python scripts/fused_ab.py > $O/fused_ab.json 2>$O/err; cat $O/fused_ab.json
N=128 timeout 600 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none -k regex:assemble_hex_mech -s 360 -c 1 python scripts/fused_ab.py 2>&1 | grep -E "inst_executed|duration|assemble_hex" | head -4
