#!/bin/bash
cd /root/repo/qp-zk-circuits_b200
for rep in 1 2 3 4 5 6 7 8; do for z in 2 1; do ./prove_bench --threads 4 --proofs 64 --perm 1 --zeroize $z | python3 -c "import sys,json; d=json.loads(sys.stdin.read()); print('perm 1 zeroize', $z, 'proofs/s %.1f'%d['proofs_per_s'], 'maxproof %.1f ms'%d['max_proof_ms'])"; done; done
for t in 1 2 6 8; do ./prove_bench --threads $t --proofs 64 --perm 1 --zeroize 2 | python3 -c "import sys,json; d=json.loads(sys.stdin.read()); print('threads', $t, 'zeroize 2 proofs/s %.1f'%d['proofs_per_s'], 'lat %.2f maxproof %.1f ms'%(d['latency_ms'], d['max_proof_ms']))"; done
