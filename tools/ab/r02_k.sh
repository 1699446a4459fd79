#!/bin/bash
out=gpurun_out; tag=r02k
python tools/time_tomesh.py 1024 readme
SDFK_SLABS_AHEAD=3 SLABS=8,16 python tools/time_tomesh.py 1024 readme
SDFK_SLABS_AHEAD=1 SLABS=8,16 python tools/time_tomesh.py 1024 readme
SLABS=0 python tools/time_tomesh.py 2048 readme
SLABS=0 python tools/time_tomesh.py 512 readme
python tools/time_multi.py 1024 1
python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; tail -2 $out/${tag}_pytest.log
