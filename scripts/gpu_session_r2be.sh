#!/bin/bash
# store-path microbenchmark (second set of modes) + A/B of the LSU store path in the real kernel
O=gpurun_out/r2be; mkdir -p $O
timeout 200 folax_b200/lib/store_path_bench | tee $O/store_path2.jsonl
for S in 0 1 2 0 1; do FOL_HEX_STORE=$S timeout 300 python scripts/hex_layout_ab.py mech | sed "s/^{/{\"store\": $S, /" >> $O/store_ab.jsonl 2>> $O/ab.err; done
cat $O/store_ab.jsonl
