#!/bin/bash
# Build libqpzk variants with different numbers of sparse partial rounds (dev tuning aid).
set -e
cd /root/repo
mkdir -p gpurun_variants
for m in "$@"; do
  QPZK_SPARSE_ROUNDS=$m python tools/gen_poseidon_constants.py >/dev/null
  make -C qp-zk-circuits_b200 -j8 >/dev/null 2>&1
  cp qp-zk-circuits_b200/libqpzk.so gpurun_variants/libqpzk_m$m.so
  echo built m=$m
done
