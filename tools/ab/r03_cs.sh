#!/bin/bash
# second copy stream for the mesh download (normals + triangles), interleaved A/B
out=gpurun_out; tag=r03cs
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for rep in 1 2 3; do
for env in "SDFK_ONE_COPY_STREAM=1" "X=1"; do
  echo "== $env" | tee -a $out/${tag}.txt
  env $env SLABS=0 python tools/time_tomesh.py 1024 readme 2>&1 | tee -a $out/${tag}.txt
  env $env SLABS=0 python tools/time_tomesh.py 1024 perf 2>&1 | tee -a $out/${tag}.txt
done
done
