# dev aid: raw pinned host->device bandwidth on this box (context for the e2e number)
import torch, time
n = 1 << 27  # 1 GiB of int64
h = torch.empty(n, dtype=torch.int64).pin_memory()
h.fill_(1)
d = torch.empty(n, dtype=torch.int64, device="cuda")
for _ in range(2):
    d.copy_(h, non_blocking=True); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    d.copy_(h, non_blocking=True)
e1.record(); torch.cuda.synchronize()
print("pinned H2D GB/s: %.1f" % (3 * n * 8 / (e0.elapsed_time(e1) * 1e-3) / 1e9))
p = torch.empty(n, dtype=torch.int64); p.fill_(1)
t0 = time.perf_counter(); d.copy_(p); torch.cuda.synchronize(); t1 = time.perf_counter()
print("pageable H2D GB/s: %.1f" % (n * 8 / (t1 - t0) / 1e9))
e0.record(); h2 = d.to("cpu"); e1.record(); torch.cuda.synchronize()
import subprocess
print(subprocess.run(["nvidia-smi", "--query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max", "--format=csv"], capture_output=True, text=True).stdout)
