"""CPU tests of bench.py's bookkeeping (the numbers the roofline is computed from)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_algorithmic_bytes_and_permutations_match_survey_table():
    sys.path.insert(0, ROOT)
    import bench
    for k, gb, mperm in [(16, 0.671, 9.4), (18, 2.682, 37.7), (20, 10.73, 151), (22, 42.92, 604)]:
        assert abs(bench.bytes_alg(k) / 1e9 - gb) / gb < 2e-3
        assert abs(bench.perms(k) / 1e6 - mperm) / mperm < 6e-3


def test_reference_arm_prints_contract_line():
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                                   "--warmup", "0"], text=True, timeout=600)
    line = json.loads(out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "lde_merkle_commit_throughput"
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["value"] > 0
