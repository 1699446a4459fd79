#!/bin/bash
# One measurement pass on a 1-GPU B200 box (run under gpurun): tests, bench (+ reference arm), ncu launch list of the same
# command, full ncu captures of every kernel of the path.  Outputs under gpurun_out/<tag>_*; tools/summarise_round.sh turns
# them into profiles/<tag>_*.
tag=${1:-r02}
out=gpurun_out
python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; tail -1 $out/${tag}_pytest.log
python bench.py --steps 20 --warmup 5 > $out/${tag}_bench_1gpu.json 2> $out/${tag}_bench_1gpu.err
python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_reference.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_step.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-fused --no-configs --no-strong > /dev/null 2>&1
N="ncu --set full --import-source on --clock-control none"
REPS=1 $N -k regex:sdfk_k_sample$ --launch-skip 2 -c 1 -o $out/${tag}_k1_readme -f python tools/time_sample.py 1024 readme > /dev/null 2>&1
REPS=1 $N -k regex:sdfk_k_sample$ --launch-skip 2 -c 1 -o $out/${tag}_k1_csg50 -f python tools/time_sample.py 1024 csg50 > /dev/null 2>&1
REPS=1 $N -k regex:sample_dist --launch-skip 5 -c 1 -o $out/${tag}_k1d -f python tools/time_sample.py 1024 readme > /dev/null 2>&1
REPS=1 $N -k regex:classify_signs --launch-skip 1 -c 1 -o $out/${tag}_k2s -f python tools/time_sample.py 1024 readme > /dev/null 2>&1
REPS=1 $N -k regex:mc_compact --launch-skip 1 -c 1 -o $out/${tag}_k4a_compact -f python tools/time_sample.py 1024 readme > /dev/null 2>&1
REPS=1 $N -k regex:mc_emit_tris --launch-skip 1 -c 1 -o $out/${tag}_k4b_emit_tris -f python tools/time_sample.py 1024 readme > /dev/null 2>&1
REPS=1 $N -k regex:mc_emit_verts --launch-skip 1 -c 1 -o $out/${tag}_k4b_emit_verts -f python tools/time_sample.py 1024 readme > /dev/null 2>&1
REPS=1 $N -k regex:sdfk_k_render$ --launch-skip 2 -c 1 -o $out/${tag}_k5_readme -f python tools/time_render.py readme > /dev/null 2>&1
REPS=1 $N -k regex:sdfk_k_render$ --launch-skip 2 -c 1 -o $out/${tag}_k5_perf -f python tools/time_render.py perf > /dev/null 2>&1
python tools/time_kernels.py > $out/${tag}_kernels.txt 2>&1
python tools/time_tomesh.py 1024 readme > $out/${tag}_tomesh.txt 2>&1
ls $out | grep ${tag}_ | wc -l
