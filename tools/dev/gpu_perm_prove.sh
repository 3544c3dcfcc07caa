#!/bin/bash
# leaf-shaped proof with the permutation argument computed on the device vs the random stand-in columns
cd qp-zk-circuits_b200
for perm in 0 1; do
  ./prove_bench --threads 1 --proofs 64 --perm $perm
  ./prove_bench --threads 4 --proofs 64 --perm $perm
done
./prove_bench --threads 1 --proofs 32 --perm 1 --breakdown 1
./prove_bench --threads 1 --proofs 16 --perm 1 --log-n 16
