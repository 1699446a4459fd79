#!/bin/bash
# round-3 session A/B: GPU parity suite, then the mesh stage times at 1024^3 for the library variants given as arguments
out=gpurun_out; tag=${TAG:-r03a}
python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; tail -2 $out/${tag}_pytest.log
REPS=3 python tools/time_sample.py 1024 readme 2>&1 | tee $out/${tag}_stages.txt
for v in "$@"; do
  echo "== variant $v" | tee -a $out/${tag}_stages.txt
  SDFK_LIB=sdfkit_b200/libsdfk_$v.so REPS=3 python tools/time_sample.py 1024 readme 2>&1 | tee -a $out/${tag}_stages.txt
done
