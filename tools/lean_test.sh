mkdir -p gpurun_out
python -m pytest tests/test_poseidon2.py tests/test_reference_vectors.py -m gpu -x -q 2>&1 | tail -3
python -m pytest tests/test_gpu_scale.py -m gpu -x -q -k "pass_plans or full_compare" 2>&1 | tail -2
for lean in 1 0; do
  QPZK_NTT_LEAN=$lean python bench.py --steps 3 --warmup 3 --no-prove --no-cpu-baseline --no-e2e --no-parity > gpurun_out/r2j_lean$lean.json 2>/dev/null
  python -c "
import json;d=json.load(open('gpurun_out/r2j_lean$lean.json'));print('lean=$lean', d['ms_per_step'], d['stages_ms'])"
done
for lean in 1 0; do
  QPZK_NTT_LEAN=$lean python bench.py --log-n 20 --steps 3 --warmup 3 --no-prove --no-cpu-baseline --no-e2e --no-parity > gpurun_out/r2j_lean${lean}_20.json 2>/dev/null
  python -c "
import json;d=json.load(open('gpurun_out/r2j_lean${lean}_20.json'));print('2^20 lean=$lean', d['ms_per_step'], d['stages_ms'])"
done
