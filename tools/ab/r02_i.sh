#!/bin/bash
out=gpurun_out; tag=${1:-r02i}
python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py -x -q -k "mesh or noise or slab or step or fused or pipelined or chunked or census or goldens or multi_to" > $out/${tag}_pytest.log 2>&1; tail -2 $out/${tag}_pytest.log
python tools/time_sample.py 1024 readme 2>&1 | tail -4 > $out/${tag}_time.txt
python tools/time_sample.py 1024 csg50 2>&1 | tail -2 >> $out/${tag}_time.txt
cat $out/${tag}_time.txt
REPS=1 ncu --set full --import-source on --clock-control none -k regex:mc_emit_verts --launch-skip 1 -c 1 -o $out/${tag}_emit_verts -f python tools/time_sample.py 1024 readme > /dev/null 2>&1
