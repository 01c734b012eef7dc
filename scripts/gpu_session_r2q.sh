#!/bin/bash
# Round-2 session q: float32 tuned Hex8 kernel: parity, then the sweep of the other configurations.
O=gpurun_out/r2q; mkdir -p $O
timeout 900 python -m pytest tests/test_assembly_gpu.py -m gpu -q > $O/tests.log 2>&1
echo "tests rc=$?"; tail -6 $O/tests.log
timeout 900 python scripts/sweep_bench.py > $O/sweep.jsonl 2> $O/sweep.err; echo "sweep rc=$?"; cut -c1-330 $O/sweep.jsonl; tail -3 $O/sweep.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:assemble_hex_mech_f32 -s 2 -c 1 -o $O/hex_f32 python scripts/sweep_bench.py > $O/ncu.log 2>&1; echo "ncu rc=$?"
