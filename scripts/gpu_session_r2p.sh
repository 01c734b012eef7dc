#!/bin/bash
# Round-2 session p: compact shared-memory layout + second staging slot of the north-star kernel: parity and A/B.
O=gpurun_out/r2p; mkdir -p $O
timeout 900 python -m pytest tests/test_assembly_gpu.py tests/test_full_size_gpu.py tests/test_zy5_fused_halo_gpu.py tests/test_golden_gpu.py tests/test_plan_host_gpu.py -m gpu -q > $O/tests.log 2>&1
echo "tests rc=$?"; tail -4 $O/tests.log
for L in 0 1; do
  FOL_HEX_LAYOUT=$L N=128 timeout 300 python scripts/fused_ab.py > $O/ab_layout$L.json 2>$O/err_$L; echo "layout $L sustained:"; cat $O/ab_layout$L.json
  FOL_HEX_LAYOUT=$L timeout 300 python bench.py --no-extras > $O/bench_layout$L.json 2>>$O/err_$L; python - <<PY
import json
d=json.load(open('$O/bench_layout$L.json')); print('layout $L burst', d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], 'sustained', d['sustained']['ms_per_step'], d['clocks']['sm_mhz'])
PY
done
FOL_HEX_LAYOUT=1 FOL_HEX_DIAG=nostore N=128 timeout 300 python scripts/fused_ab.py > $O/ab_layout1_nostore.json 2>>$O/err_1; cat $O/ab_layout1_nostore.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:assemble_hex_mech -s 5 -c 1 -o $O/assemble_hex_l1 python bench.py --no-extras --steps 3 > $O/ncu_hex.log 2>&1; echo "ncu rc=$?"
