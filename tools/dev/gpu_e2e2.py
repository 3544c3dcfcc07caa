# dev aid: back-to-back host-buffer commits without extra synchronisation -- which call blocks?
import sys, time, numpy as np, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/qp-zk-circuits_b200')
import qpzk
log_n, C_ = 20, 135
n = 1 << log_n
dev = torch.device("cuda", 0)
ctx = qpzk.Context(0, stream=torch.cuda.current_stream().cuda_stream)
full = torch.randint(0, 2**63 - 1, (C_, n), device=dev, dtype=torch.int64)
host = torch.empty((C_, n), dtype=torch.int64).pin_memory()
host.copy_(full); torch.cuda.synchronize()
hnp = host.numpy().view(np.uint64)
for zero in (True, False):
    ctx.set_zeroize(zero)
    for rep in range(5):
        t0 = time.perf_counter()
        b = qpzk.PolynomialBatch.from_values(ctx, hnp, 3, None, 4)
        t1 = time.perf_counter()
        cap = b.cap; t2 = time.perf_counter()
        b.free(); t3 = time.perf_counter()
        print("zeroize=%s: from_values %.1f ms, cap %.1f, free %.1f" % (zero, 1e3*(t1-t0), 1e3*(t2-t1), 1e3*(t3-t2)))
    torch.cuda.synchronize()
