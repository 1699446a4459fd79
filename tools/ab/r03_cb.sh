#!/bin/bash
# mc_emit_verts from corner blocks: parity, then stage times for the variants given as arguments
out=gpurun_out; tag=${TAG:-r03cb}
python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; tail -2 $out/${tag}_pytest.log
REPS=2 python tools/time_sample.py 1024 readme 2>&1 | grep "signs=1" | tee $out/${tag}_stages.txt
for v in "$@"; do
  echo "== variant $v" | tee -a $out/${tag}_stages.txt
  SDFK_LIB=sdfkit_b200/libsdfk_$v.so REPS=2 python tools/time_sample.py 1024 readme 2>&1 | grep "signs=1" | tee -a $out/${tag}_stages.txt
done
