#!/bin/bash
# multi-threaded proof throughput for every library variant, three repetitions each
cd /root/repo/qp-zk-circuits_b200
for lib in ../gpurun_variants/libqpzk_w*.so; do
  cp $lib ./libqpzk_test.so
  for rep in 1 2 3; do
    for t in 1 4; do
      LD_PRELOAD=$lib ./prove_bench --threads $t --proofs 64 --perm 0 | sed "s|^|$(basename $lib) |" | cut -c1-170
    done
  done
done
rm -f ./libqpzk_test.so
