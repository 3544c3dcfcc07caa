#!/bin/bash
# dev aid (run under gpurun): launch list + full ncu captures of the top kernels of one 2^20 commit
set -x
TAG=${1:-r1c}
B="python bench.py --log-n 20 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-prove"
$B > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $B > gpurun_out/${TAG}_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:hash_rows_kernel -c 1 -f -o gpurun_out/${TAG}_hash $B > gpurun_out/${TAG}_ncu_hash.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"ntt_|lde_last" -c 5 -f -o gpurun_out/${TAG}_ntt $B > gpurun_out/${TAG}_ncu_ntt.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:merkle_level_kernel -c 2 -f -o gpurun_out/${TAG}_merkle $B > gpurun_out/${TAG}_ncu_merkle.log 2>&1
