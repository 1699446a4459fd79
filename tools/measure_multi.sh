#!/bin/bash
# bench.py on N GPUs of one box under torchrun (what the driver runs), + the multi-GPU context timing table
N=${1:-8}; tag=${2:-r02}; out=gpurun_out
nvidia-smi topo -m > $out/${tag}_topo_${N}gpu.txt 2>&1
extra=""; [ "$N" != "8" ] && extra="--no-strong"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus $N --steps 20 --warmup 5 $extra > $out/${tag}_bench_${N}gpu.json 2> $out/${tag}_bench_${N}gpu.err; tail -c 400 $out/${tag}_bench_${N}gpu.err
if [ "$N" == "8" ]; then python tools/time_multi.py 1024 > $out/${tag}_multi_1024.txt 2>&1; python tools/time_multi.py 2048 2 4 8 > $out/${tag}_multi_2048.txt 2>&1; cat $out/${tag}_multi_1024.txt $out/${tag}_multi_2048.txt; fi
