# dev aid: stage timings of commits of different shapes (device-resident input)
import sys, numpy as np, torch, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/qp-zk-circuits_b200')
import qpzk
ctx=qpzk.Context(0); ctx.set_timing(True)
for log_n, C_ in [(18,135),(18,20),(18,16),(16,20),(20,20),(13,20)]:
    x=torch.randint(0,2**63-1,(C_,1<<log_n),device="cuda",dtype=torch.int64)
    for i in range(3):
        torch.cuda.synchronize(); t0=time.perf_counter()
        b=qpzk.PolynomialBatch.from_device(ctx,x.data_ptr(),log_n,C_,3,4,values=True); t=ctx.last_commit_timing(); dt=time.perf_counter()-t0; b.free()
    print(log_n, C_, {k:round(v,3) for k,v in t.items()}, "wall %.2f ms"%(1e3*dt), flush=True)
