#!/bin/bash
out=gpurun_out; tag=r02m
python -m pytest tests/test_gpu_multi.py -x -q > $out/${tag}_pytest.log 2>&1; tail -2 $out/${tag}_pytest.log
python tools/time_multi.py 1024 1 2 > $out/${tag}_multi.txt 2>&1; cat $out/${tag}_multi.txt
python tools/time_multi.py 2048 2 2>&1 | tail -3
