#!/bin/bash
out=gpurun_out; tag=r03m
for cfg in "0 0" "4 100000" "8 100000" "16 100000" "32 100000"; do
  set -- $cfg
  echo "== zsplit $1 blocks/SM cap $2" | tee -a $out/${tag}_zsplit.txt
  SDFK_ZSPLIT=$1 SDFK_SAMPLE_BPS=$2 REPS=5 python tools/time_sample.py 1024 readme 2>&1 | grep "signs=1" | cut -c1-40 | tee -a $out/${tag}_zsplit.txt
  SDFK_NO_DIST8=1 SDFK_ZSPLIT=$1 SDFK_SAMPLE_BPS=$2 REPS=5 python tools/time_sample.py 1024 readme 2>&1 | grep "colors=0 signs=1" | cut -c1-40 | sed 's/^/no dist8: /' | tee -a $out/${tag}_zsplit.txt
done
echo "== csg50" | tee -a $out/${tag}_zsplit.txt
for cfg in "0 0" "32 100000"; do
  set -- $cfg
  SDFK_ZSPLIT=$1 SDFK_SAMPLE_BPS=$2 REPS=3 python tools/time_sample.py 1024 csg50 2>&1 | grep "signs=1" | cut -c1-40 | tee -a $out/${tag}_zsplit.txt
done
