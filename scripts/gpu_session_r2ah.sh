#!/bin/bash
# tuned Hex8 thermal kernel: parity tests, throughput at 128^3 (tuned vs generic), one ncu capture
O=gpurun_out/r2ah; mkdir -p $O
timeout 900 python -m pytest tests/test_assembly_gpu.py tests/test_golden_gpu.py -m gpu -q -x -k "thermal or golden" > $O/tests.log 2>&1; echo "tests rc=$?"; tail -4 $O/tests.log
ONLY=thermal timeout 600 python scripts/sweep_bench.py 2>$O/err | tee $O/sweep_thermal.jsonl | cut -c1-300
timeout 300 ncu --set full --clock-control none --import-source on -k regex:assemble_hex_thermal -s 3 -c 1 -o $O/hex_thermal env ONLY=thermal python scripts/sweep_bench.py > $O/ncu.log 2>&1; echo "ncu rc=$?"
tail -3 $O/err
