#!/bin/bash
# dev aid (run under gpurun): timeline of concurrent provers.  usage: run_trace.sh TAG [prove_bench args...]
TAG=$1; shift
cd "$(dirname "$0")/../.."
export QPZK_TRACE_OUT=gpurun_out/${TAG}_trace.csv
CUDA_INJECTION64_PATH=$PWD/tools/trace/libqpzktrace.so ./qp-zk-circuits_b200/prove_bench "$@" > gpurun_out/${TAG}_traced.json 2> gpurun_out/${TAG}_traced.err
python tools/trace/summarize.py gpurun_out/${TAG}_trace.csv > gpurun_out/${TAG}_trace_summary.txt 2>&1
