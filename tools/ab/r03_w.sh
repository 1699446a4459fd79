#!/bin/bash
out=gpurun_out; tag=r03w
for z in 0 8 16 64; do
  echo "== csg50 zsplit $z" | tee -a $out/${tag}.txt
  SDFK_ZSPLIT=$z REPS=3 python tools/time_sample.py 1024 csg50 2>&1 | grep "signs=1" | cut -c1-40 | tee -a $out/${tag}.txt
done
for z in 0 16 64; do
  echo "== perf zsplit $z" | tee -a $out/${tag}.txt
  SDFK_ZSPLIT=$z REPS=3 python tools/time_sample.py 1024 perf 2>&1 | grep "signs=1" | cut -c1-40 | tee -a $out/${tag}.txt
done
REPS=3 python tools/time_sample.py 1024 readme 2>&1 | grep "signs=1" | tee -a $out/${tag}.txt
