#!/bin/bash
# Round-2 ncu evidence (run on the GPU box through gpurun): cold GAE counters, the rollout kernel, the wide-policy
# GEMM kernel and the C3 launch list.  Every command is run once WITHOUT ncu first; numbers printed under ncu are
# never bench values.  Outputs: gpurun_out/ev/*
set -u
O=gpurun_out/ev; mkdir -p $O
NCU="ncu --clock-control none"
python tools/gae_ncu.py 128 65536 > $O/gae_plain.log 2>&1 || exit 1
$NCU --set full --import-source on -k regex:gae_tn -s 2 -c 1 -o $O/gae_c2 python tools/gae_ncu.py 128 65536 > $O/gae_ncu.log 2>&1
$NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -k regex:gae_tn -s 2 -c 1 --csv --log-file $O/gae_1m.csv python tools/gae_ncu.py 128 1048576 > /dev/null 2>&1
B="python bench.py --steps 1 --warmup 1 --epochs 1 --no-cpu-baseline --no-e2e"
$B > $O/c2_plain.log 2>&1 || exit 1
$NCU --set full --import-source on -k regex:rollout_tc -s 1 -c 1 -o $O/rollout_c2 $B > $O/rollout_ncu.log 2>&1
$B --config c3 > $O/c3_plain.log 2>&1 || exit 1
$NCU --metrics gpu__time_duration.sum -k regex:'tc_gemm|gather|loss|reduce|apply|adv_sums|perm|pack' -c 1200 --csv --log-file $O/c3_update_launches.csv $B --config c3 > /dev/null 2>&1
$NCU --metrics gpu__time_duration.sum -k regex:'rollout|steps|gemm|observe|act' -c 60 --csv --log-file $O/c3_rollout_launches.csv $B --config c3 > /dev/null 2>&1
$NCU --set full --import-source on -k regex:tc_gemm -s 40 -c 12 -o $O/tc_gemm_c3 $B --config c3 > $O/tc_gemm_ncu.log 2>&1
ls -la $O
