#!/bin/bash
out=gpurun_out; tag=r03h
for cfg in "0 0" "2 0" "4 0" "8 0" "2 4" "4 4" "8 4" "16 4" "4 3" "8 3" "4 2"; do
  set -- $cfg
  echo "== zsplit $1 blocks/SM $2" | tee -a $out/${tag}_zsplit.txt
  SDFK_ZSPLIT=$1 SDFK_SAMPLE_BPS=$2 REPS=5 python tools/time_sample.py 1024 readme 2>&1 | grep "signs=1" | cut -c1-40 | tee -a $out/${tag}_zsplit.txt
done
