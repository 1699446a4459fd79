#!/bin/bash
# One measurement pass on a B200 box (run under gpurun): tests, bench (+ reference arm), ncu launch lists of the same
# commands, full ncu captures of the top kernels, the BASELINE config table.  Outputs under gpurun_out/<tag>_*.
tag=${1:-r01s2}
out=gpurun_out
python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; tail -1 $out/${tag}_pytest.log
python bench.py > $out/${tag}_bench_1gpu.json 2> $out/${tag}_bench_1gpu.err
python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_reference.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_step.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-fused > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches_full.csv python bench.py --steps 2 --warmup 1 --no-cpu > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:sdfk_k_sample$ --launch-skip 3 -c 1 -o $out/${tag}_k1_sample -f python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-fused > /dev/null 2>&1
REPS=1 ncu --set full --import-source on --clock-control none -k regex:sample_dist --launch-skip 5 -c 1 -o $out/${tag}_k1d_sample_dist -f python tools/time_sample.py 1024 readme > /dev/null 2>&1
REPS=1 ncu --set full --import-source on --clock-control none -k regex:classify_signs --launch-skip 1 -c 1 -o $out/${tag}_k2s_classify_signs -f python tools/time_sample.py 1024 readme > /dev/null 2>&1
REPS=1 ncu --set full --import-source on --clock-control none -k regex:mc_emit_verts --launch-skip 1 -c 1 -o $out/${tag}_k4b_emit_verts -f python tools/time_sample.py 1024 readme > /dev/null 2>&1
python tools/run_configs.py > $out/${tag}_configs.txt 2> $out/${tag}_configs.err
python tools/time_tomesh.py 1024 readme > $out/${tag}_tomesh.txt 2>&1
ls $out | grep ${tag}_ | wc -l
