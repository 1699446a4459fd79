#!/bin/bash
out=gpurun_out; tag=r02g
N="ncu --set full --import-source on --clock-control none"
REPS=1 $N -k regex:sdfk_k_sample$ --launch-skip 2 -c 1 -o $out/${tag}_k1_csg50 -f python tools/time_sample.py 1024 csg50 > /dev/null 2>&1
REPS=1 $N -k regex:sdfk_k_sample$ --launch-skip 2 -c 1 -o $out/${tag}_k1_readme -f python tools/time_sample.py 1024 readme > /dev/null 2>&1
REPS=1 $N -k regex:sdfk_k_render$ --launch-skip 2 -c 1 -o $out/${tag}_k5_readme -f python tools/time_render.py readme > /dev/null 2>&1
REPS=1 $N -k regex:sdfk_k_render$ --launch-skip 2 -c 1 -o $out/${tag}_k5_perf -f python tools/time_render.py perf > /dev/null 2>&1
REPS=1 $N -k regex:sample_dist --launch-skip 5 -c 1 -o $out/${tag}_k1d -f python tools/time_sample.py 1024 readme > /dev/null 2>&1
ls $out | grep $tag
