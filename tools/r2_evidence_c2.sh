#!/bin/bash
# Round-2 ncu evidence for the default-width update kernel (ppo_grad_tc128c_kernel) at config 2, run on the GPU box through
# gpurun.  The command is run once WITHOUT ncu first; numbers printed under ncu are never bench values.
# Outputs: gpurun_out/c2k/*
set -u
O=gpurun_out/c2k; mkdir -p $O
NCU="ncu --clock-control none"
B="python bench.py --steps 1 --warmup 1 --epochs 2 --no-cpu-baseline --no-e2e"
$B > $O/c2_plain.log 2>&1 || exit 1
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/launches_c2_epochs2.csv $B > $O/launches_ncu.log 2>&1
$NCU --set full --import-source on -k regex:ppo_grad_tc128c -s 40 -c 2 -o $O/tc128c $B > $O/tc128c_ncu.log 2>&1
ls -la $O
