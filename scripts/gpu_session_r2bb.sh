#!/bin/bash
# dense layouts as defaults: A/B of the L2 evict-first hint on the Ke bulk stores, then the round's final validation
O=gpurun_out/r2bb; mkdir -p $O
for H in 0 1 0 1; do FOL_HEX_HINT=$H timeout 300 python scripts/hex_layout_ab.py mech >> $O/hint_ab.jsonl 2>> $O/ab.err; done
for H in 0 1; do FOL_HEX_HINT=$H timeout 300 python scripts/hex_layout_ab.py j2 >> $O/hint_ab.jsonl 2>> $O/ab.err; done
cat $O/hint_ab.jsonl
bash scripts/gpu_session_final.sh
