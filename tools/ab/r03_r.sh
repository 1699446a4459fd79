#!/bin/bash
out=gpurun_out; tag=r03u
for rep in 1 2 3; do
for env in "SDFK_UNIFORM_SLABS=1" "X=1"; do
  echo "== $env" | tee -a $out/${tag}_tomesh.txt
  env $env SLABS=0 python tools/time_tomesh.py 1024 readme 2>&1 | tee -a $out/${tag}_tomesh.txt
  env $env SLABS=0 python tools/time_tomesh.py 1024 csg50 2>&1 | tee -a $out/${tag}_tomesh.txt
  env $env SLABS=0 python tools/time_tomesh.py 1024 perf 2>&1 | tee -a $out/${tag}_tomesh.txt
done
done
