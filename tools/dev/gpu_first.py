import sys, time, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/qp-zk-circuits_b200'); sys.path.insert(0,'/root/repo/tests')
from conftest import splitmix_columns
from oracle import oracle as orc
import qpzk
ctx=qpzk.Context(0)
print(qpzk.load_library().qpzk_version())
# merkle / poseidon
rng=np.random.default_rng(3)
for (nl,ll,ch) in [(1,5,0),(8,3,0),(8,4,3),(16,7,2),(64,8,4),(512,135,4),(1024,32,4),(2,9,1),(4096,20,0)]:
    leaves=rng.integers(0,orc.P,size=(nl,ll),dtype=np.uint64)
    t=qpzk.MerkleTree.new(ctx,leaves,ch); o=orc.MerkleTree(leaves,ch)
    ok=(t.cap==o.cap()).all()
    i=nl//3
    ok2=(t.prove(i)==o.prove(i)).all() and (t.get(i)==leaves[i]).all()
    print('merkle',nl,ll,ch,ok,ok2)
# commit
for (log_n,C_) in [(0,3),(1,2),(3,5),(5,9),(8,135),(10,20),(12,16),(13,135),(14,7),(16,5)]:
    cols=splitmix_columns(1000+log_n,C_,1<<log_n)
    ch=min(4,log_n+3)
    t0=time.time(); b=qpzk.PolynomialBatch.from_values(ctx,cols,3,None,ch); cap=b.cap; t1=time.time()
    o=orc.PolynomialBatch.from_values(cols,3,ch); t2=time.time()
    okc=(b.polynomials==o.coeffs()).all()
    okl=(b.merkle_tree.rows(0,1<<(log_n+3))==o.leaves()).all()
    okcap=(cap==o.cap()).all()
    print('commit',log_n,C_,okc,okl,okcap,'gpu %.3fs cpu %.3fs'%(t1-t0,t2-t1))
    b2=qpzk.PolynomialBatch.from_coeffs(ctx,o.coeffs(),3,None,ch)
    print('  from_coeffs', (b2.cap==o.cap()).all())
