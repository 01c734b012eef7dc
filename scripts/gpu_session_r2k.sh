#!/bin/bash
O=gpurun_out/r2k; mkdir -p $O
N=128 timeout 300 python scripts/fused_ab.py > $O/fused_ab.json 2> $O/fused_ab.err; echo "rc=$?"; cat $O/fused_ab.json; tail -3 $O/fused_ab.err
