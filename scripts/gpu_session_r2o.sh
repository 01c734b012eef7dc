#!/bin/bash
O=gpurun_out/r2o; mkdir -p $O
N=128 timeout 300 python scripts/fused_ab.py > $O/ab_store.json 2>$O/err1; cat $O/ab_store.json
FOL_HEX_DIAG=nostore N=128 timeout 300 python scripts/fused_ab.py > $O/ab_nostore.json 2>$O/err2; cat $O/ab_nostore.json
