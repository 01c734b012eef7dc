#!/bin/bash
# whole GPU suite without -x (after the revert of the split-loop kernels + the one-launch BiCGSTAB), then configs[3]
# Newton with the multi-launch and the one-launch Krylov loop
O=gpurun_out/r2v; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/gpu_suite.log 2>&1; echo "suite rc=$?"; tail -4 $O/gpu_suite.log
N=70 DISP=0.004 LOAD_STEPS=1 timeout 600 python scripts/newton_bench.py > $O/newton_default.json 2>$O/newton_default.err; cat $O/newton_default.json
N=70 DISP=0.004 LOAD_STEPS=1 FUSED=1 timeout 600 python scripts/newton_bench.py > $O/newton_fused.json 2>$O/newton_fused.err; cat $O/newton_fused.json; tail -3 $O/newton_fused.err
