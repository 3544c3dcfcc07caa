#!/usr/bin/env python3
"""Runs only the permutation-argument measurement of bench.py (rows a16/a17) -- the command profiled with ncu:
    python tools/dev/perm_bench.py [log_n ...]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "qp-zk-circuits_b200"))
import torch  # noqa: E402
import bench  # noqa: E402
import qpzk  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
ctx = qpzk.Context(0, stream=torch.cuda.current_stream().cuda_stream)
for k in [int(a) for a in sys.argv[1:]] or [13, 18]:
    print(json.dumps(bench.gpu_perm_argument(ctx, dev, k, reps=2)))
