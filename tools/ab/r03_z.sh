#!/bin/bash
# warps per CTA of the sampling kernels (a CTA = that many consecutive work items = x tiles of a row)
out=gpurun_out; tag=r03z
for w in 8 4 2 16; do
  echo "== $w warps per CTA" | tee -a $out/${tag}.txt
  SDFK_SAMPLE_WARPS=$w SDFK_JIT_DEFINES="-DSDFK_SAMPLE_WARPS=$w" REPS=5 python tools/time_sample.py 1024 readme 2>&1 | grep "signs=1" | cut -c1-40 | tee -a $out/${tag}.txt
done
