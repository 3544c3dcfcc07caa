import sys, time, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/qp-zk-circuits_b200'); sys.path.insert(0,'/root/repo/tests')
from conftest import splitmix_columns
import qpzk
ctx=qpzk.Context(0)
log_n=6; Cs=[5,9]
cols=[splitmix_columns(77+i,c,1<<log_n) for i,c in enumerate(Cs)]
gb=[qpzk.PolynomialBatch.from_values(ctx,c,3,None,4) for c in cols]
inst=qpzk.FriInstanceInfo([([3,4],[(0,0),(1,1)])])
for pb in [8,12]:
    ch=qpzk.Challenger(ctx); ch.observe_elements([1,2,3])
    p=qpzk.fri_params(log_n,3,4,pb,5,4,2)
    ctx.synchronize(); t0=time.time()
    proof=qpzk.PolynomialBatch.prove_openings(inst,gb,ch,p)
    t1=time.time()
    ch2=qpzk.Challenger(ctx); ch2.observe_elements([1,2,3])
    proof2=qpzk.PolynomialBatch.prove_openings(inst,gb,ch2,p,pow_witness=int(proof[-1]))
    t2=time.time()
    print(pb,'witness',proof[-1],'grind %.4fs forced %.4fs'%(t1-t0,t2-t1),(proof==proof2).all())
print('imad/s %.3e'%ctx.bench_imad(4096)); print('poseidon perms/s %.3e'%ctx.bench_poseidon(32))
