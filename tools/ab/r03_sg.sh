#!/bin/bash
# slab grouping of the planned parts: 1,2,rest,1 (default) against 1,1,rest,1
out=gpurun_out; tag=r03sg
for rep in 1 2 3; do
for env in "SDFK_SLAB_GROUPING=0" "SDFK_SLAB_GROUPING=1"; do
  echo "== $env" | tee -a $out/${tag}.txt
  env $env SLABS=0 python tools/time_tomesh.py 1024 readme 2>&1 | tee -a $out/${tag}.txt
  env $env SLABS=0 python tools/time_tomesh.py 1024 csg50 2>&1 | tee -a $out/${tag}.txt
done
done
