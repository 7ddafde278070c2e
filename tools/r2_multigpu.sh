#!/bin/bash
# Multi-GPU evidence (run with gpurun --gpus 8): c2 weak scaling and c4 strong scaling at 8 GPUs (smaller N: separate calls).
# Outputs: gpurun_out/mg/*
O=gpurun_out/mg; mkdir -p $O
run() { # name gpus args...
  local name=$1 n=$2; shift 2
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
    bench.py --gpus $n --steps 3 --warmup 3 --no-cpu-baseline "$@" > $O/$name.json 2> $O/$name.err
  echo "$name rc=$?"; tail -c 300 $O/$name.json
}
for c in "$@"; do
  case $c in
    c2_n8) run c2_n8 8 --config c2;;
    c4_n8) run c4_n8 8 --config c4;;
    c4_n4) run c4_n4 4 --config c4;;
    c4_n2) run c4_n2 2 --config c4;;
  esac
done
