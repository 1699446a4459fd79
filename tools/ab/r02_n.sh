#!/bin/bash
out=gpurun_out; tag=r02n
python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py -x -q -k "render or depth or tga or image or default_device" > $out/${tag}_pytest.log 2>&1; tail -3 $out/${tag}_pytest.log
python tools/time_kernels.py > $out/${tag}_kernels.txt 2>&1
SDFK_NO_RENDER4=1 python tools/time_kernels.py >> $out/${tag}_kernels.txt 2>&1
cat $out/${tag}_kernels.txt
