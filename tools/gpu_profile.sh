#!/bin/bash
# dev aid (run under gpurun): launch list + full ncu captures of the top kernels of one 2^22 commit (the bench configuration)
#   bash tools/gpu_profile.sh [TAG] [LOG_N]
set -x
TAG=${1:-r2f}
LOGN=${2:-22}
B="python bench.py --log-n $LOGN --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-prove --no-parity"
$B > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $B > gpurun_out/${TAG}_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:hash_rows_kernel -c 1 -f -o gpurun_out/${TAG}_hash $B > gpurun_out/${TAG}_ncu_hash.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"ntt_|lde_last" -c 7 -f -o gpurun_out/${TAG}_ntt $B > gpurun_out/${TAG}_ncu_ntt.log 2>&1
