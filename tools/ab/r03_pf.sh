#!/bin/bash
# mc_emit_verts with the cross-kind L1 prefetch (variant pf) against the default
out=gpurun_out; tag=r03pf
SDFK_LIB=sdfkit_b200/libsdfk_pf.so python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mesh_matches or white_noise" 2>&1 | tail -1
for rep in 1 2; do
  REPS=2 python tools/time_sample.py 1024 readme 2>&1 | grep "signs=1" | tee -a $out/${tag}_stages.txt
  echo "== variant pf" | tee -a $out/${tag}_stages.txt
  SDFK_LIB=sdfkit_b200/libsdfk_pf.so REPS=2 python tools/time_sample.py 1024 readme 2>&1 | grep "signs=1" | tee -a $out/${tag}_stages.txt
done
