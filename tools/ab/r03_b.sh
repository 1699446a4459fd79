#!/bin/bash
out=gpurun_out; tag=${TAG:-r03b}
python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $out/${tag}_pytest.log 2>&1; tail -2 $out/${tag}_pytest.log
REPS=3 python tools/time_sample.py 1024 readme 2>&1 | tee $out/${tag}_stages.txt
for v in "$@"; do
  echo "== variant $v" | tee -a $out/${tag}_stages.txt
  SDFK_LIB=sdfkit_b200/libsdfk_$v.so REPS=3 python tools/time_sample.py 1024 readme 2>&1 | tee -a $out/${tag}_stages.txt
done
N="ncu --set full --import-source on --clock-control none"
REPS=1 $N -k regex:mc_compact --launch-skip 1 -c 1 -o $out/${tag}_compact -f python tools/time_sample.py 1024 readme > /dev/null 2>&1
REPS=1 $N -k regex:mc_emit_verts --launch-skip 4 -c 4 -o $out/${tag}_emit_verts -f python tools/time_sample.py 1024 readme > /dev/null 2>&1
ls -la $out | grep ${tag}
