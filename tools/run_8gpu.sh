#!/bin/bash
# Development: the 8-GPU session of a round (gpurun --gpus 8 charges eight-fold): everything that needs more than two
# GPUs, each step under its own timeout, results under gpurun_out/r2/.  STEPS selects what runs (default: all).
cd "$(dirname "$0")/.."
O=gpurun_out/r2; mkdir -p $O
STEPS=${STEPS:-"probe bench clip train"}
T() { local to=$1 n=$2; shift 2; timeout $to python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) "$@"; }
for s in $STEPS; do
  case $s in
    probe) for n in 4 8; do T 120 $n tools/pcie_probe.py 2>$O/pcie$n.err | grep probe | tee $O/pcie$n.log | cut -c1-260; done ;;
    bench) T 400 8 bench.py --gpus 8 --steps 20 --warmup 5 2>$O/bench8.err | grep metric | tee $O/bench8.json | cut -c1-400 ;;
    clip)  for n in ${CLIP_N:-4 8}; do T 200 $n bench.py --config clip --gpus $n 2>$O/clip$n.err | grep config | tee $O/clip${n}${TAG}.json | cut -c1-500; done ;;
    train) for n in ${TRAIN_N:-8}; do T 400 $n bench.py --config train --gpus $n --steps 3 --warmup 3 2>$O/train$n.err | grep config | tee $O/train${n}${TAG}.json | cut -c1-1200; done ;;
  esac
done
