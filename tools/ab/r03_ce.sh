#!/bin/bash
out=gpurun_out; tag=r03ce
REPS=2 python tools/time_sample.py 1024 readme 2>&1 | grep "signs=1" | tee $out/${tag}_stages.txt
for v in "$@"; do
  echo "== variant $v" | tee -a $out/${tag}_stages.txt
  SDFK_LIB=sdfkit_b200/libsdfk_$v.so REPS=2 python tools/time_sample.py 1024 readme 2>&1 | grep "signs=1" | tee -a $out/${tag}_stages.txt
done
REPS=2 python tools/time_sample.py 1024 readme 2>&1 | grep "signs=1" | tee -a $out/${tag}_stages.txt
