# dev aid: leaf-hash time of a real 2^18 x 135 commit for every library variant on the command line / in gpurun_variants
import sys, os, subprocess, glob
code = r'''
import os, sys, numpy as np, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/qp-zk-circuits_b200')
import qpzk
ctx=qpzk.Context(0); ctx.set_timing(True)
x=torch.randint(0,2**63-1,(135,1<<18),device="cuda",dtype=torch.int64)
ts=[]
for i in range(4):
    b=qpzk.PolynomialBatch.from_device(ctx,x.data_ptr(),18,135,3,4,values=True); t=ctx.last_commit_timing(); b.free(); ts.append(t)
t=ts[-1]
print(os.environ["QPZK_LIB"].split("/")[-1], "leaf %.3f levels %.3f lde %.3f intt %.3f ms; microbench %.3e perms/s"%(t["leaf_hash_ms"],t["merkle_levels_ms"],t["lde_ms"],t["intt_ms"],max(ctx.bench_poseidon(64) for _ in range(3))), flush=True)
'''
libs = sys.argv[1:] or sorted(glob.glob('/root/repo/gpurun_variants/*.so'))
for lib in libs:
    subprocess.run([sys.executable, '-c', code], env=dict(os.environ, QPZK_LIB=os.path.abspath(lib)))
