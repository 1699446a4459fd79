#!/bin/bash
# 8-GPU pass: strong scaling through the multi-GPU context + the N=8 bench line (weak, parity at 2048^3)
out=gpurun_out; tag=r02h
nvidia-smi topo -m > $out/${tag}_topo.txt 2>&1; nproc >> $out/${tag}_topo.txt; free -g >> $out/${tag}_topo.txt
python tools/time_multi.py 1024 > $out/${tag}_multi.txt 2>&1
SDFK_TRACE=1 python tools/time_multi.py 1024 8 > $out/${tag}_multi_trace.txt 2>&1
cat $out/${tag}_multi.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 10 --warmup 3 > $out/${tag}_bench_8gpu.json 2> $out/${tag}_bench_8gpu.err; tail -c 600 $out/${tag}_bench_8gpu.err
