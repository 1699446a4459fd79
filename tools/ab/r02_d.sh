#!/bin/bash
out=gpurun_out; tag=r02d
python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; tail -3 $out/${tag}_pytest.log
SDFK_PLAIN_BODY=1 python tools/time_kernels.py > $out/${tag}_kernels.txt 2>&1
python tools/time_kernels.py >> $out/${tag}_kernels.txt 2>&1
cat $out/${tag}_kernels.txt
N="ncu --set full --import-source on --clock-control none"
REPS=1 $N -k regex:sdfk_k_sample$ --launch-skip 2 -c 1 -o $out/${tag}_k1_csg50 -f python tools/time_sample.py 1024 csg50 > /dev/null 2>&1
REPS=1 $N -k regex:sdfk_k_render$ --launch-skip 2 -c 1 -o $out/${tag}_k5_readme -f python tools/time_render.py readme > /dev/null 2>&1
REPS=1 $N -k regex:sample_dist --launch-skip 5 -c 1 -o $out/${tag}_k1d -f python tools/time_sample.py 1024 readme > /dev/null 2>&1
