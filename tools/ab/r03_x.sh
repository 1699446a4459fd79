#!/bin/bash
# refresh of the 1-GPU bench line + launch list after the store probe moved into the library
tag=r02; out=gpurun_out
python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; tail -1 $out/${tag}_pytest.log
python bench.py --steps 20 --warmup 5 > $out/${tag}_bench_1gpu.json 2> $out/${tag}_bench_1gpu.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_step.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-fused --no-configs --no-strong > /dev/null 2>&1
python tools/time_tomesh.py 1024 readme > $out/${tag}_tomesh.txt 2>&1
SLABS=0 python tools/time_tomesh.py 1024 readme >> $out/${tag}_tomesh.txt 2>&1
tail -c 300 $out/${tag}_bench_1gpu.err
