# dev aid: parity of the Poseidon kernels vs the oracle + permutations/s of every library given on the command line
import sys, os, subprocess, glob
code = r'''
import os, sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/qp-zk-circuits_b200')
import qpzk
from oracle import oracle as orc
ctx=qpzk.Context(0)
rng=np.random.default_rng(1)
ok=True
for (nl,ll,ch) in [(256,135,2),(64,8,0),(128,33,3),(32,4,1)]:
    leaves=rng.integers(0,orc.P,size=(nl,ll),dtype=np.uint64)
    ok = ok and bool((qpzk.MerkleTree.new(ctx,leaves,ch).cap==orc.MerkleTree(leaves,ch).cap()).all())
edge=np.full((16,24),orc.P-1,dtype=np.uint64)
ok = ok and bool((qpzk.MerkleTree.new(ctx,edge,0).cap==orc.MerkleTree(edge,0).cap()).all())
r=[ctx.bench_poseidon(64) for _ in range(3)]
print(os.environ["QPZK_LIB"].split("/")[-1], "parity", ok, "perms/s %.3e"%max(r), flush=True)
'''
libs = sys.argv[1:] or sorted(glob.glob('/root/repo/gpurun_variants/*.so'))
for lib in libs:
    subprocess.run([sys.executable, '-c', code], env=dict(os.environ, QPZK_LIB=os.path.abspath(lib)))
