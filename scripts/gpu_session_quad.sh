#!/bin/bash
O=gpurun_out/quad; mkdir -p $O
timeout 900 python -m pytest tests/test_assembly_gpu.py -m gpu -q -k "quad" > $O/tests.log 2>&1; echo "tests rc=$?"; tail -2 $O/tests.log
ONLY=quad timeout 600 python scripts/sweep_bench.py 2>$O/err | tee $O/sweep_quad.jsonl | cut -c1-300
