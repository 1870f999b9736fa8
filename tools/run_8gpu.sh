#!/bin/bash
# Development: the one 8-GPU session of a round (gpurun --gpus 8 charges eight-fold): everything that needs more than two
# GPUs, each step under its own timeout, results under gpurun_out/r2/.
cd "$(dirname "$0")/.."
O=gpurun_out/r2; mkdir -p $O
T() { local to=$1 n=$2; shift 2; timeout $to python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) "$@"; }
for n in 4 8; do T 120 $n tools/pcie_probe.py 2>$O/pcie$n.err | grep probe | tee $O/pcie$n.log | cut -c1-260; done
T 400 8 bench.py --gpus 8 --steps 20 --warmup 5 2>$O/bench8.err | grep metric | tee $O/bench8.json | cut -c1-400
for n in 4 8; do T 200 $n bench.py --config clip --gpus $n 2>$O/clip$n.err | grep config | tee $O/clip$n.json | cut -c1-500; done
T 400 8 bench.py --config train --gpus 8 --steps 3 --warmup 3 2>$O/train8.err | grep config | tee $O/train8.json | cut -c1-1200
