#!/bin/bash
# usage: tools/multi_gpu_session.sh TAG N [extra bench args]   -- N-rank bench, both exchange modes
TAG=$1; N=$2; shift 2
O=gpurun_out; mkdir -p $O
nvidia-smi topo -m > $O/${TAG}_topo.txt 2>&1
for ex in ${EXCHANGES:-p2p nccl}; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --exchange $ex "$@" > $O/${TAG}_n${N}_$ex.json 2> $O/${TAG}_n${N}_$ex.err
  echo "n=$N $ex rc=$?"; tail -c 1500 $O/${TAG}_n${N}_$ex.json; echo; tail -3 $O/${TAG}_n${N}_$ex.err
done
