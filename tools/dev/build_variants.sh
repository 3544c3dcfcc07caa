#!/bin/bash
# Build libqpzk variants (dev tuning aid):  build_variants.sh name1="-DFOO=1" name2="-DBAR=2 -DBAZ" ...
set -e
cd /root/repo
mkdir -p gpurun_variants
for spec in "$@"; do
  name="${spec%%=*}"; flags="${spec#*=}"
  make -C qp-zk-circuits_b200 clean >/dev/null
  make -C qp-zk-circuits_b200 -j8 EXTRA="$flags" >/dev/null 2>&1
  cp qp-zk-circuits_b200/libqpzk.so gpurun_variants/libqpzk_$name.so
  echo "built $name ($flags): $(tools/dev/sass_hist.sh gpurun_variants/libqpzk_$name.so poseidon_bench_kernel | head -1)"
done
make -C qp-zk-circuits_b200 clean >/dev/null; make -C qp-zk-circuits_b200 -j8 >/dev/null 2>&1
