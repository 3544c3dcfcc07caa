#!/bin/bash
# One gpurun session: GPU tests, bench, launch list, sanitizer subset.  usage: tools/gpu_session.sh TAG [what...]
# what: tests | scale | bench | launches | sanitize | ncu_hash | ncu_ntt   (default: tests bench)
TAG=${1:-r2}; shift
WHAT=${@:-tests bench}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/${TAG}_gpu.txt 2>&1
for w in $WHAT; do
  case $w in
    tests)    timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > $O/${TAG}_tests.log 2>&1; echo "tests rc=$?" | tee -a $O/${TAG}_rc.txt; tail -5 $O/${TAG}_tests.log;;
    scale)    timeout 1500 python -m pytest tests/test_gpu_scale.py -m gpu -x -q --durations=15 > $O/${TAG}_scale.log 2>&1; echo "scale rc=$?" | tee -a $O/${TAG}_rc.txt; tail -5 $O/${TAG}_scale.log;;
    bench)    timeout 1200 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?" | tee -a $O/${TAG}_rc.txt; tail -c 1600 $O/${TAG}_bench.json;;
    benchq)   timeout 900 python bench.py --no-prove --no-cpu-baseline > $O/${TAG}_benchq.json 2> $O/${TAG}_benchq.err; echo "benchq rc=$?" | tee -a $O/${TAG}_rc.txt; tail -c 1600 $O/${TAG}_benchq.json;;
    refarm)   timeout 600 python bench.py --impl reference --steps 2 --warmup 3 > $O/${TAG}_ref.json 2> $O/${TAG}_ref.err; echo "ref rc=$?" | tee -a $O/${TAG}_rc.txt; cat $O/${TAG}_ref.json;;
    launches) B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-prove --no-parity"
              timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv $B > $O/${TAG}_ncu_launch.log 2>&1; echo "launches rc=$?" | tee -a $O/${TAG}_rc.txt;;
    ncu_hash) B="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-prove --no-parity"
              timeout 900 ncu --set full --clock-control none --import-source on -k regex:hash_rows_kernel -c 1 -f -o $O/${TAG}_hash $B > $O/${TAG}_ncu_hash.log 2>&1; echo "ncu_hash rc=$?" | tee -a $O/${TAG}_rc.txt;;
    ncu_ntt)  B="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-prove --no-parity"
              timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ntt_|lde_last" -c 6 -f -o $O/${TAG}_ntt $B > $O/${TAG}_ncu_ntt.log 2>&1; echo "ncu_ntt rc=$?" | tee -a $O/${TAG}_rc.txt;;
    prover)   timeout 900 python -m pytest tests/test_gpu_prover.py -m gpu -x -q --durations=8 > $O/${TAG}_prover.log 2>&1; echo "prover rc=$?" | tee -a $O/${TAG}_rc.txt; tail -15 $O/${TAG}_prover.log;;
    pbench)   PB=qp-zk-circuits_b200/prove_bench; G=qp-zk-circuits_b200/host/poseidon_gate.qpgp
              ( $PB --threads 1 --proofs 64 --perm 1 --breakdown 1
                $PB --threads 1 --proofs 64 --perm 1 --gates $G
                $PB --prover 1 --threads 1 --proofs 128 --graph 0
                $PB --prover 1 --threads 1 --proofs 128
                $PB --prover 1 --threads 1 --proofs 128 --gates $G
                for K in 2 3 4 6 8; do $PB --prover 1 --threads 1 --inflight $K --proofs 256; done
                for K in 2 4 8; do $PB --prover 1 --threads 1 --inflight $K --proofs 256 --gates $G; done
                $PB --prover 1 --threads 2 --inflight 2 --proofs 256 --gates $G
                $PB --prover 1 --threads 4 --inflight 2 --proofs 256 --gates $G
                for LN in 15 16; do $PB --prover 1 --threads 1 --proofs 24 --log-n $LN --routed 60 --gates $G; done
                $PB --prover 1 --threads 1 --inflight 2 --proofs 32 --log-n 15 --routed 60 --gates $G
              ) > $O/${TAG}_pbench.jsonl 2> $O/${TAG}_pbench.err; echo "pbench rc=$?" | tee -a $O/${TAG}_rc.txt; cut -c1-330 $O/${TAG}_pbench.jsonl; tail -5 $O/${TAG}_pbench.err;;
    sanitize) for tool in memcheck racecheck synccheck; do
                timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_subset.py > $O/${TAG}_san_$tool.log 2>&1; echo "sanitize $tool rc=$?" | tee -a $O/${TAG}_rc.txt; tail -4 $O/${TAG}_san_$tool.log
              done;;
  esac
done
