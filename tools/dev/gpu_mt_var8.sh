#!/bin/bash
cd /root/repo/qp-zk-circuits_b200
for rep in 1 2 3 4 5 6 7 8; do ./prove_bench --threads 4 --proofs 64 --perm 0 --breakdown 1 2>&1 | grep -E "longest|proofs_per_s" | sed 's/"proofs_identical.*//' ; echo; done
