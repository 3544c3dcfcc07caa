#!/bin/bash
# do the multi-threaded prover's stalls coincide with clock / power events?
cd /root/repo/qp-zk-circuits_b200
nvidia-smi --query-gpu=timestamp,clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_power_brake_slowdown,clocks_event_reasons.sw_power_cap,clocks_event_reasons.sw_thermal_slowdown --format=csv,noheader -lms 50 > ../gpurun_out/s3_clocks.csv &
SMI=$!
sleep 0.5
for rep in 1 2 3 4 5 6 7 8 9 10 11 12; do date +%H:%M:%S.%N | cut -c1-12; ./prove_bench --threads 4 --proofs 64 --perm 0 | python3 -c "import sys,json; d=json.loads(sys.stdin.read()); print('proofs/s %.1f'%d['proofs_per_s'], 'maxproof %.1f ms'%d['max_proof_ms'])"; done
kill $SMI
