import torch, time
x=torch.empty(1<<27,dtype=torch.int64).pin_memory()
d=torch.empty(1<<27,dtype=torch.int64,device='cuda')
for _ in range(3):
    torch.cuda.synchronize(); t=time.perf_counter(); d.copy_(x,non_blocking=True); torch.cuda.synchronize(); dt=time.perf_counter()-t
    print('H2D pinned 1GiB: %.1f GB/s'%(x.numel()*8/dt/1e9))
    torch.cuda.synchronize(); t=time.perf_counter(); x.copy_(d,non_blocking=True); torch.cuda.synchronize(); dt=time.perf_counter()-t
    print('D2H pinned 1GiB: %.1f GB/s'%(x.numel()*8/dt/1e9))
