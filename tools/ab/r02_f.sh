#!/bin/bash
out=gpurun_out; tag=r02f
SDFK_CTAB=1 python tools/time_kernels.py > $out/${tag}_kernels.txt 2>&1
cat $out/${tag}_kernels.txt
python -m pytest tests/test_gpu_parity.py -x -q -k "color_table" > $out/${tag}_pytest.log 2>&1; tail -2 $out/${tag}_pytest.log
