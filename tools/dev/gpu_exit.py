import sys, faulthandler, numpy as np
faulthandler.enable()
sys.path.insert(0,'/root/repo/qp-zk-circuits_b200')
import qpzk
mode=sys.argv[1]
ctx=qpzk.Context(0)
if mode=='ctx': pass
elif mode=='batch':
    b=qpzk.PolynomialBatch.from_values(ctx,np.arange(64,dtype=np.uint64).reshape(2,32),3,None,4); print(b.cap[0])
elif mode=='chal':
    ch=qpzk.Challenger(ctx); ch.observe_elements([1,2,3]); print(ch.get_challenge())
elif mode=='bench':
    print(ctx.bench_imad(64))
print('done',mode, flush=True)
