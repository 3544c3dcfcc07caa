import sys, os, subprocess, glob
code = r'''
import sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/qp-zk-circuits_b200')
import qpzk
from oracle import oracle as orc
ctx=qpzk.Context(0)
rng=np.random.default_rng(1)
leaves=rng.integers(0,orc.P,size=(256,135),dtype=np.uint64)
ok=(qpzk.MerkleTree.new(ctx,leaves,2).cap==orc.MerkleTree(leaves,2).cap()).all()
r=[ctx.bench_poseidon(64) for _ in range(3)]
print(os.environ["QPZK_LIB"].split("/")[-1], "parity", ok, "perms/s %.3e"%max(r))
'''
code = "import os\n" + code
for lib in sorted(glob.glob('/root/repo/gpurun_variants/libqpzk_m*.so')):
    env = dict(os.environ, QPZK_LIB=lib)
    subprocess.run([sys.executable, '-c', code], env=env)
