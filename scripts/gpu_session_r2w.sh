#!/bin/bash
O=gpurun_out/r2w; mkdir -p $O
timeout 600 python -m pytest tests/test_zzz2_fused_bicgstab_gpu.py -m gpu -q 2>&1 | tail -3
N=70 timeout 600 python scripts/krylov_fused_bench.py > $O/krylov_tet70.json 2>$O/err; cat $O/krylov_tet70.json
HEX=128 ITERS=50 timeout 600 python scripts/krylov_fused_bench.py > $O/krylov_hex128.json 2>>$O/err; cat $O/krylov_hex128.json
N=25 ITERS=400 timeout 600 python scripts/krylov_fused_bench.py > $O/krylov_tet25.json 2>>$O/err; cat $O/krylov_tet25.json
tail -5 $O/err
