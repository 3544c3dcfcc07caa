cd qp-zk-circuits_b200
for i in 1 2 3; do ./prove_bench --threads 4 --proofs 64; done
for i in 1 2; do ./prove_bench --threads 3 --proofs 64; done
python - <<'PY'
import torch, subprocess
x = torch.zeros(1<<20, device="cuda"); torch.cuda.synchronize()
print("under a parent with a CUDA context:")
for i in range(2):
    print(subprocess.run(["./prove_bench","--threads","4","--proofs","64"],capture_output=True,text=True).stdout.strip())
PY
nproc
