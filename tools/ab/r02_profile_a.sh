#!/bin/bash
# round-2 first profiling pass: host facts, FP32 peak, full ncu captures of the kernels this round works on
out=gpurun_out; tag=r02a
( nproc; free -g; lscpu | head -30; nvidia-smi topo -m; numactl -H 2>/dev/null ) > $out/${tag}_host.txt 2>&1
python tools/fp32_peak.py $out/${tag}_fp32_peak.json > /dev/null 2>&1
python tools/time_sample.py 1024 csg50 > $out/${tag}_time_csg50.txt 2>&1
python tools/time_render.py readme > $out/${tag}_time_render.txt 2>&1
python tools/time_render.py perf >> $out/${tag}_time_render.txt 2>&1
N="ncu --set full --import-source on --clock-control none"
REPS=1 $N -k regex:sdfk_k_sample$ --launch-skip 2 -c 1 -o $out/${tag}_k1_csg50 -f python tools/time_sample.py 1024 csg50 > /dev/null 2>&1
REPS=1 $N -k regex:sdfk_k_render$ --launch-skip 2 -c 1 -o $out/${tag}_k5_readme -f python tools/time_render.py readme > /dev/null 2>&1
REPS=1 $N -k regex:sdfk_k_render$ --launch-skip 2 -c 1 -o $out/${tag}_k5_perf -f python tools/time_render.py perf > /dev/null 2>&1
REPS=1 $N -k regex:sample_dist --launch-skip 5 -c 1 -o $out/${tag}_k1d -f python tools/time_sample.py 1024 readme > /dev/null 2>&1
REPS=1 $N -k regex:mc_emit_verts --launch-skip 1 -c 1 -o $out/${tag}_emit_verts -f python tools/time_sample.py 1024 readme > /dev/null 2>&1
REPS=1 $N -k regex:mc_emit_tris --launch-skip 1 -c 1 -o $out/${tag}_emit_tris -f python tools/time_sample.py 1024 readme > /dev/null 2>&1
REPS=1 $N -k regex:mc_compact --launch-skip 1 -c 1 -o $out/${tag}_compact -f python tools/time_sample.py 1024 readme > /dev/null 2>&1
ls -la $out | grep ${tag}_
