#!/bin/bash
cd /root/repo/qp-zk-circuits_b200
for rep in 1 2 3 4 5 6 7 8 9 10 11 12; do ./prove_bench --threads 4 --proofs 64 --perm 0 | python3 -c "import sys,json; d=json.loads(sys.stdin.read()); print('single-stream H2D, zeroize 1: proofs/s %.1f'%d['proofs_per_s'], 'maxproof %.1f ms'%d['max_proof_ms'])"; done
./prove_bench --threads 1 --proofs 64 --perm 0 | cut -c1-200
