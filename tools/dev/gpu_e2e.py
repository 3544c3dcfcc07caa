# dev aid: where does the host-buffer commit spend its time?
import sys, time, numpy as np, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/qp-zk-circuits_b200')
import qpzk
log_n, C_ = 20, 135
n = 1 << log_n
dev = torch.device("cuda", 0)
for own in (False, True):
    ctx = qpzk.Context(0) if own else qpzk.Context(0, stream=torch.cuda.current_stream().cuda_stream)
    full = torch.randint(0, 2**63 - 1, (C_, n), device=dev, dtype=torch.int64)
    host = torch.empty((C_, n), dtype=torch.int64).pin_memory()
    host.copy_(full); torch.cuda.synchronize()
    hnp = host.numpy().view(np.uint64)
    def sync():
        ctx.synchronize(); torch.cuda.synchronize()
    for rep in range(3):
        sync(); t0 = time.perf_counter()
        b = qpzk.PolynomialBatch.from_values(ctx, hnp, 3, None, 4)
        t1 = time.perf_counter(); sync(); t2 = time.perf_counter()
        cap = b.cap; t3 = time.perf_counter()
        b.free(); t4 = time.perf_counter(); sync(); t5 = time.perf_counter()
        print("own_stream=%s host path: call %.1f ms, sync %.1f, cap %.1f, free %.1f, sync %.1f" % (own, 1e3*(t1-t0), 1e3*(t2-t1), 1e3*(t3-t2), 1e3*(t4-t3), 1e3*(t5-t4)))
    for rep in range(3):
        sync(); t0 = time.perf_counter()
        full.copy_(host, non_blocking=True); torch.cuda.synchronize()
        t1 = time.perf_counter()
        b = qpzk.PolynomialBatch.from_device(ctx, full.data_ptr(), log_n, C_, 3, 4, values=True)
        sync(); t2 = time.perf_counter()
        cap = b.cap; b.free(); sync(); t3 = time.perf_counter()
        print("own_stream=%s dev path: h2d %.1f ms, commit %.1f, cap+free %.1f" % (own, 1e3*(t1-t0), 1e3*(t2-t1), 1e3*(t3-t2)))
    del ctx
