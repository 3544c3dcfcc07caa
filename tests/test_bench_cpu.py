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


def test_contract_line_fits_the_driver_record():
    """The driver keeps about 1.5 kB of the line: with every optional section present and 8 GPUs it must stay < 1.4 kB
    and carry the contract keys."""
    sys.path.insert(0, ROOT)
    import bench
    stage = {"intt_ms": 22.54412, "lde_ms": 129.50123, "leaf_hash_ms": 452.91345, "merkle_levels_ms": 26.37612, "allgather_ms": 6.12345}
    native = {"latency_ms": 4.72345, "latency_ms_gates": 5.82312, "proofs_per_s": 389.7312, "proofs_per_s_gates": 308.0512,
              "ms_2p15_gates": 13.00412, "ms_2p16_gates": 22.89112, "batch_1024_s": 3.32112}
    out = bench.assemble_line(world=8, steps=5, warmup=3, ms_per_step=637.6212, value=497.30712, log_n=22, leaf_gbs=82.38412,
                              lde_gbs=268.1412, hbm_peak=6539.9, traffic=41436000000.0, int_issue=0.70112, imad_peak=7.2681e12,
                              perm_rate=1.2595e9, stage=stage, launches=110,
                              clocks={"sm_mhz": 1965.0, "sm_max_mhz": 1965.0, "reasons": ["sw_power_cap"], "samples": 17},
                              e2e={"value": 389.012, "unit": "GB/s", "ms_per_step": 110.3, "h2d_bytes_per_step": 4529848320, "d2h_bytes_per_step": 512},
                              parity={"ok": True, "rows_checked": 1024, "columns": 135, "cap_timed_path_equal": True, "cap_e2e_path_equal": True},
                              native=native, proofs_all=(2464.4012, 0.41512),
                              cpu={"value": 0.227912, "cores": 32, "sample": "from_values 2^16x135, 2.9 s", "us_perm": 3.33161, "prove_ms": 852.8612})
    line = json.dumps(bench.compact(out), separators=(",", ":"))
    assert len(line) < 1400, len(line)
    back = json.loads(line)
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks", "parity_check",
              "roofline_lde", "int_pipe", "stages_ms", "proofs_per_s_all_gpus"):
        assert k in back, k
    assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(back["roofline"])
    assert set(("value", "unit", "cores", "kind", "sample")) <= set(back["cpu_baseline"])
