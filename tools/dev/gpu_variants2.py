"""Run every gpurun_variants/libqpzk_<prefix>*.so: parity of a small Merkle tree against the oracle, the Poseidon
microbenchmark, and the stage times of a 2^18 x 135 commit.   python tools/dev/gpu_variants2.py <prefix>"""
import glob
import os
import subprocess
import sys

code = r'''
import os, sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/qp-zk-circuits_b200')
import qpzk
from oracle import oracle as orc
ctx=qpzk.Context(0)
rng=np.random.default_rng(1)
leaves=rng.integers(0,orc.P,size=(256,135),dtype=np.uint64)
ok=(qpzk.MerkleTree.new(ctx,leaves,2).cap==orc.MerkleTree(leaves,2).cap()).all()
r=[ctx.bench_poseidon(64) for _ in range(3)]
cols=rng.integers(0,orc.P,size=(135,1<<18),dtype=np.uint64)
ctx.set_timing(True)
best=None
for _ in range(4):
    b=qpzk.PolynomialBatch.from_values(ctx,cols,3,None,4); t=ctx.last_commit_timing(); b.free()
    best=t if best is None or t["leaf_hash_ms"]<best["leaf_hash_ms"] else best
print(os.environ["QPZK_LIB"].split("/")[-1], "parity", bool(ok), "perms/s %.3e"%max(r), " ".join("%s=%.2f"%(k,v) for k,v in best.items()), flush=True)
'''
for lib in sorted(glob.glob('/root/repo/gpurun_variants/libqpzk_%s*.so' % (sys.argv[1] if len(sys.argv) > 1 else ""))):
    subprocess.run([sys.executable, '-c', code], env=dict(os.environ, QPZK_LIB=lib))
