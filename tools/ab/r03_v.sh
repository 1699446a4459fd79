#!/bin/bash
out=gpurun_out; tag=r03v
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for b in 1 0 3 12; do
  SDFK_RENDER_BANDS=$b python tools/time_toimage.py readme 2>&1 | tee -a $out/${tag}_toimage.txt
  SDFK_RENDER_BANDS=$b python tools/time_toimage.py perf 2>&1 | tee -a $out/${tag}_toimage.txt
done
REPS=3 python tools/time_sample.py 1024 readme 2>&1 | tee $out/${tag}_stages.txt
