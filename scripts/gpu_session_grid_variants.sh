#!/bin/bash
# compile-time variants of the grid kernel, rebuilt ON the GPU box (nvcc is in the image): ring depth, register caps
O=gpurun_out/${TAG:-gridv}; mkdir -p $O
run() {
  name=$1; shift
  FOL_GRID_DEFS="-DFOL_GRID_ONLY_NL4 $*" python -m folax_b200.build > /dev/null 2>$O/build_$name.err || { echo "build $name failed"; tail -3 $O/build_$name.err; return; }
  grep -E "Used|spill" folax_b200/lib/obj/energy_grid.o.log | paste - - | grep -E "IdLi1ELi4ELb1ELb1|IfLi2ELi4ELb1ELb1" -A0 > /dev/null
  for dt in float64 float32; do
    r=$(DTYPE=$dt timeout 300 python scripts/energy_variants.py 2>>$O/err | head -1 | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print(d['ms'], d['samples_per_s'], d['hash'][1])")
    echo "$name $dt $r" | tee -a $O/variants.txt
  done
}
run base
run prefetch -DFOL_GRID_PREFETCH=1
run prefetch_r80 -DFOL_GRID_PREFETCH=1 -DFOL_GRID_REGS64=80
