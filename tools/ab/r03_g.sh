#!/bin/bash
out=gpurun_out; tag=r03g
for pol in 0 1 2 3; do
  echo "== store policy $pol" | tee -a $out/${tag}_store.txt
  SDFK_JIT_DEFINES="-DSDFK_STORE_POLICY=$pol" REPS=5 python tools/time_sample.py 1024 readme 2>&1 | grep "signs=1" | tee -a $out/${tag}_store.txt
done
