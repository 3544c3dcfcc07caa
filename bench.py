#!/usr/bin/env python3
"""bench.py -- LDE + Poseidon-Merkle commit throughput (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--log-n 22] [--impl ours|reference]

A step = ONE `PolynomialBatch::from_values` of the raw-commit shape of BASELINE.json config 3:
2^log_n trace rows x 135 Goldilocks columns, rate_bits 3, cap_height 4, no blinding, uniform random canonical
inputs (synthetic).  Default log_n = 22, the largest size of the sweep; it fits one GPU.

  value     whole-job commit throughput, algorithmic GB/s:
            bytes_alg = 8*C*n*(1+2^rate_bits) + 32*(2*(2^rate_bits*n - 2^cap_height) + 2^cap_height)
            divided by the device time of the step, inputs already resident in HBM.
  e2e       the same metric through the host-buffer C-ABI call qpzk_commit_from_values (pinned host input,
            host->device copy and the device->host read of the cap inside the timed region).
  N > 1     strong scaling of the SAME commitment: columns are split for the iNTT, the coefficient
            matrix is all-gathered over NCCL, each rank extends and hashes 2^rate_bits/N cosets (whole Merkle
            subtrees), and the cap slices are all-gathered.  Time = max over ranks.
  roofline  the dominant kernel (Poseidon leaf hashing) against the measured HBM peak, as the contract asks; the
            LDE stage (HBM-bound by design) and the integer pipe (what actually bounds Poseidon) are reported
            beside it as roofline_lde / int_pipe.
  --impl reference   the CPU restatement of the reference algorithm (oracle/, OpenMP on every host core, set
            explicitly: torchrun exports OMP_NUM_THREADS=1) on a bounded sample (2^k x 135, k <= 16 sized to the host)
            of the same workload.  The Rust reference itself cannot be built here (no cargo, un-vendored qp-plonky2),
            see DESIGN.md; `cpu_baseline.calibration` states the port's us per permutation per core and the
            reference's published whole-prove() time so that no ratio is read as "x times the reference".
  parity_check       outside the timed region the benchmarked commitment is verified against the oracle
            (oracle/scale_check.py) and the caps of the timed path, the e2e path and a fresh commit must agree; a
            mismatch fails the run.
  output    ONE compact JSON line (<= 1.4 kB); the long-form record goes to gpurun_out/bench_detail_n<N>.json.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "qp-zk-circuits_b200"))

NUM_COLS, RATE_BITS, CAP_HEIGHT = 135, 3, 4
METRIC, UNIT = "lde_merkle_commit_throughput", "GB/s"
P = 0xFFFFFFFF00000001


def bytes_alg(log_n, num_cols=NUM_COLS, rate_bits=RATE_BITS, cap_height=CAP_HEIGHT):
    n, B, H = 1 << log_n, 1 << rate_bits, 1 << cap_height
    return 8 * num_cols * n * (1 + B) + 32 * (2 * (B * n - H) + H)


def perms(log_n, num_cols=NUM_COLS, rate_bits=RATE_BITS, cap_height=CAP_HEIGHT):
    L = (1 << log_n) << rate_bits
    return L * ((num_cols + 7) // 8) + (L - (1 << cap_height))


def splitmix_columns_device(seed, num_cols, n, device):
    """SURVEY.md 8d input recipe on the device: SplitMix64 stream seeded `seed`, rejection-sampled < p, filled
    column-major; returns a (num_cols, n) int64 tensor holding the canonical u64 bit patterns (identical to
    tests/conftest.py splitmix_columns and to host/prove_bench.cpp's SplitMix)."""
    import torch
    total = num_cols * n
    m = total + 64 + total // (1 << 24)                 # rejection probability 2^-32 per draw
    s64 = lambda v: v - (1 << 64) if v >= (1 << 63) else v
    lsr = lambda z, k: (z >> k) & ((1 << (64 - k)) - 1)
    z = torch.arange(1, m + 1, dtype=torch.int64, device=device) * s64(0x9E3779B97F4A7C15) + s64(seed & ((1 << 64) - 1))
    z = (z ^ lsr(z, 30)) * s64(0xBF58476D1CE4E5B9)
    z = (z ^ lsr(z, 27)) * s64(0x94D049BB133111EB)
    z = z ^ lsr(z, 31)
    bad = (z < 0) & (z >= -4294967295)                  # unsigned value >= p = 2^64 - 2^32 + 1
    if bool(bad.any()):
        z = z[~bad]
    return z[:total].reshape(num_cols, n).contiguous()


def ncu_record():
    """Numbers taken from the committed `ncu --set full` capture of hash_rows_kernel at the benchmarked configuration
    (profiles/r2_traffic.json, else r1): DRAM traffic per launch by log_n, warp-level instructions per permutation."""
    for name in ("r2_traffic.json", "r1_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                e = json.load(f)
            return {"source": "profiles/" + name,
                    "traffic": {k: v["dram_read_bytes"] + v["dram_write_bytes"] for k, v in e["per_log_n"].items()},
                    "inst_per_permutation": e.get("inst_executed_per_permutation")}
        except Exception:
            continue
    return {}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "MEASURED_PEAKS.json"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def cpu_commit_sample(log_n_sample=16, repeats=1):
    """The CPU restatement (oracle, OpenMP) on a bounded sample of the workload; returns (GB/s, cores, seconds)."""
    import numpy as np
    from oracle import oracle as orc
    orc.build()
    rng = np.random.default_rng(16)
    cols = rng.integers(0, P, size=(NUM_COLS, 1 << log_n_sample), dtype=np.uint64)
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        b = orc.PolynomialBatch.from_values(cols, RATE_BITS, CAP_HEIGHT)
        dt = time.perf_counter() - t0
        del b
        best = dt if best is None else min(best, dt)
    return bytes_alg(log_n_sample) / best / 1e9, orc.num_threads(), best


def gpu_perm_argument(ctx, dev, log_n, reps=5, routed=80, num_wires=135, num_consts=5, nch=2, chunk=8):
    """SURVEY.md 8a rows a16/a17 at the leaf-circuit column counts: partial products + Z from device-resident routed
    wires and sigmas, then the quotient of the permutation argument from the three committed batches.  Device times
    (CUDA events on the context stream) and algorithmic bytes: every operand read once, every result written once."""
    import numpy as np
    import torch
    import qpzk
    n = 1 << log_n
    gen = torch.Generator(device=dev).manual_seed(99 + log_n)
    rnd = lambda r: torch.randint(0, 2**63 - 1, (r, n), generator=gen, device=dev, dtype=torch.int64)
    wires, cs = rnd(num_wires), rnd(num_consts + routed)
    nchunks = (routed + chunk - 1) // chunk
    zs = torch.empty((nch * nchunks, n), dtype=torch.int64, device=dev)
    ch = np.random.default_rng(log_n).integers(0, P, size=(3, nch), dtype=np.uint64)
    sig_ptr = cs.data_ptr() + num_consts * n * 8
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        a, b = ev(), ev()
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    pp_ms = timed(lambda: ctx.partial_products_device(wires.data_ptr(), sig_ptr, log_n, routed, ch[0], ch[1], zs.data_ptr(), chunk))
    batches = [qpzk.PolynomialBatch.from_device(ctx, t.data_ptr(), log_n, t.shape[0], RATE_BITS, CAP_HEIGHT) for t in (cs, wires, zs)]
    ctx.set_timing(True)
    commit_ms = [0.0]

    def quot():
        qb = qpzk.PolynomialBatch.commit_quotient(batches[0], batches[1], batches[2], num_consts, routed, ch[0], ch[1], ch[2],
                                                  CAP_HEIGHT, chunk)
        commit_ms[0] = sum(ctx.last_commit_timing().values())
        qb.free()

    total_ms = timed(quot)

    # gate constraints as a program: upstream's PoseidonGate -- the gate behind the reference's 135-wire configuration,
    # 123 constraints of degree 7 -- recorded by qpzk/gates.py (about 2900 instructions after multiply-add fusion, 24 registers) and interpreted at
    # every quotient point (qpzk_commit_quotient_prog minus the plain quotient above)
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_poseidon_constants as gen
    from qpzk import gates
    rc = gen.round_constants()
    rec = gates.Recorder()
    gates.poseidon_gate(rec, rc, gen.derive_fast_partial(rc), selector=rec.const(0))
    code, imm, regs = rec.finish()
    prog = qpzk.GateProgram.make(code, imm, regs)
    ops = int(code.shape[0])

    def quot_prog():
        qb = qpzk.PolynomialBatch.commit_quotient_prog(batches[0], batches[1], batches[2], num_consts, routed, ch[0], ch[1], ch[2],
                                                       prog, CAP_HEIGHT, chunk)
        qb.free()

    prog_total_ms = timed(quot_prog)
    for b in batches:
        b.free()
    Lq = n * chunk
    pp_bytes = 8 * n * (2 * routed + nch * nchunks)
    q_bytes = 8 * Lq * (2 * routed + nch * nchunks + nch) + 8 * 3 * nch * Lq   # operands once; values out, iNTT in/out
    q_ms = total_ms - commit_ms[0]
    return {"log_n": log_n, "routed": routed, "challenges": nch,
            "partial_products_ms": pp_ms, "partial_products_GBps": pp_bytes / pp_ms / 1e6,
            "quotient_ms": q_ms, "quotient_GBps": q_bytes / q_ms / 1e6 if q_ms > 0 else None,
            "quotient_commit_ms": commit_ms[0],
            "gate_program": "PoseidonGate (qpzk/gates.py)", "gate_program_ms": prog_total_ms - total_ms, "gate_program_ops": ops,
            "gate_program_terms": int((code[:, 0] == 6).sum()), "gate_program_registers": int(regs),
            "gate_program_ns_per_op_and_point": 1e6 * (prog_total_ms - total_ms) / (ops * n * chunk),
            "what": "a16: qpzk_partial_products_dev; a17: qpzk_commit_quotient minus the time of its 16-column commit"}


def prove_bench(device, *opts, repeats=1):
    """qp-zk-circuits_b200/prove_bench (host/prove_bench.cpp: the native driver of the C ABI) with the given options;
    returns the parsed JSON line of the median run (by proofs/s) with all runs' proofs/s, or None."""
    exe = os.path.join(ROOT, "qp-zk-circuits_b200", "prove_bench")
    if not os.path.exists(exe):
        return None
    try:
        runs = []
        for _ in range(repeats):
            r = subprocess.run([exe, "--device", str(device)] + [str(o) for o in opts], capture_output=True, text=True, timeout=300)
            runs.append(json.loads(r.stdout.strip().splitlines()[-1]))
        runs.sort(key=lambda d: d["proofs_per_s"])
        out = runs[len(runs) // 2]
        out["proofs_per_s_runs"] = [d["proofs_per_s"] for d in runs]
        return out
    except Exception:
        return None


def native_prove_suite(device, world):
    """Leaf-shaped synthetic prove (BASELINE configs 1/2/5 stand-in) through the one-call prover
    (qpzk_prover_prove: device-resident challenges, CUDA-graph replay), from native code:
      latency      one proof at a time (2^13 rows, 135 wires / 80 routed, 85 constants+sigmas, 21 public inputs, FRI
                   rate 3 / cap 4 / 16 PoW bits / 28 queries / arity 4), host buffers in, proof out, witness generation
                   excluded; with and without upstream's PoseidonGate program on every row
      throughput   one host thread keeping 8 proofs in flight on 8 contexts (their graphs overlap on the GPU)
      config 5     1024 proofs sharded over the N GPUs of the job (1024 / N per rank), wall time of the slowest rank
      private batch shapes (common/src/circuit.rs:393-402): 2^15 and 2^16 rows, 135 wires / 60 routed."""
    gp = os.path.join(ROOT, "qp-zk-circuits_b200", "host", "poseidon_gate.qpgp")   # written by __graft_entry__.build()
    g = ["--gates", gp] if os.path.exists(gp) else []
    runs = {"latency": prove_bench(device, "--prover", 1, "--proofs", 128),
            "latency_gates": prove_bench(device, "--prover", 1, "--proofs", 128, *g),
            "throughput": prove_bench(device, "--prover", 1, "--inflight", 8, "--proofs", 512, repeats=3),
            "throughput_gates": prove_bench(device, "--prover", 1, "--inflight", 8, "--proofs", 512, *g, repeats=3),
            "batch_1024": prove_bench(device, "--prover", 1, "--inflight", 8, "--proofs", 1024 // world, *g),
            "private_batch_2p15": prove_bench(device, "--prover", 1, "--proofs", 24, "--log-n", 15, "--routed", 60, *g),
            "private_batch_2p16": prove_bench(device, "--prover", 1, "--proofs", 16, "--log-n", 16, "--routed", 60, *g)}
    get = lambda k, f: (runs.get(k) or {}).get(f)
    return {"runs": runs,
            "latency_ms": get("latency", "latency_ms"), "latency_ms_gates": get("latency_gates", "latency_ms"),
            "proofs_per_s": get("throughput", "proofs_per_s"), "proofs_per_s_gates": get("throughput_gates", "proofs_per_s"),
            "proofs_per_s_gates_runs": get("throughput_gates", "proofs_per_s_runs"),
            "batch_1024_s": get("batch_1024", "slowest_thread_s"),
            "ms_2p15_gates": get("private_batch_2p15", "latency_ms"), "ms_2p16_gates": get("private_batch_2p16", "latency_ms"),
            "launches_per_proof": get("latency_gates", "launches_per_proof"), "graph": get("latency_gates", "graph"),
            "what": native_prove_suite.__doc__}


def cpu_prove_synthetic(proofs=2):
    """The same synthetic prove on the CPU restatement (all host cores); ms per proof."""
    import numpy as np
    from oracle import oracle as orc
    orc.build()
    sys.path.insert(0, os.path.join(ROOT, "qp-zk-circuits_b200"))
    from qpzk import prover
    shape = prover.LEAF_SHAPE
    rng = np.random.default_rng(1)
    cs = orc.PolynomialBatch.from_values(rng.integers(0, P, size=(shape["constants_sigmas"], 1 << shape["log_n"]), dtype=np.uint64), RATE_BITS, CAP_HEIGHT)
    best = None
    for s in range(proofs):
        inp = prover.synthetic_inputs(s, shape)
        t0 = time.perf_counter()
        orc.prove_synthetic(cs, inp["wires"], inp["zs_partial_products"], inp["quotient_coeffs"], inp["public_inputs_hash"], shape["log_n"])
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return 1e3 * best, orc.num_threads()


WORKLOAD = "raw commit (BASELINE config 3): 2^%d rows x %d cols, rate_bits %d, cap_height %d, Poseidon-12 Merkle, blinding off"
# published whole-prove() time of the reference's own CPU prover (other hardware; context for every ratio below)
PUBLISHED_LEAF_PROVE = {"ms": 20.0, "hardware": "Apple M2 Max, 12 cores", "source": "paper/main.tex:449,455"}


def cpu_sample_log_n(cores, budget_s=6.0):
    """Largest size of the sweep's shape (2^k x 135, k <= 16) one step of which fits `budget_s` on `cores` threads,
    assuming 7 us per permutation per core (the port's measured rate; the permutations dominate)."""
    k = 16
    while k > 10 and 7e-6 * perms(k) / max(cores, 1) > budget_s:
        k -= 1
    return k


def port_calibration(orc):
    """us per Poseidon permutation on ONE core of this host (the port's unit of work)."""
    import numpy as np
    rows = np.arange(256 * 128, dtype=np.uint64).reshape(256, 128)
    orc.set_num_threads(1)
    t0 = time.perf_counter()
    orc.MerkleTree(rows, 0)
    dt = time.perf_counter() - t0
    orc.set_num_threads()
    return 1e6 * dt / (256 * 16 + 255)


def run_reference(args):
    """The reference arm: the CPU restatement of the same algorithm (oracle/, C + OpenMP) on every host core, each step
    one from_values of a bounded sample of the workload.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    from oracle import oracle as orc
    orc.build()
    cores = orc.set_num_threads()            # torchrun exports OMP_NUM_THREADS=1: use the host's cores explicitly
    us_perm = port_calibration(orc)
    sample = cpu_sample_log_n(cores)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import splitmix_columns
    cols = splitmix_columns(0x51575A4B00000000 + sample, NUM_COLS, 1 << sample)
    times = []
    t_start = time.perf_counter()
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        b = orc.PolynomialBatch.from_values(cols, RATE_BITS, CAP_HEIGHT)
        dt = time.perf_counter() - t0
        del b
        if i >= args.warmup:
            times.append(dt)
        if time.perf_counter() - t_start > 150 and times:      # bound the arm's wall time whatever the host is
            break
    ms = 1e3 * sum(times) / len(times)
    v = bytes_alg(sample) / (ms * 1e-3) / 1e9
    desc = "one from_values of 2^%d x %d per step (same shape, bounded size), %d steps timed" % (sample, NUM_COLS, len(times))
    print(json.dumps(compact({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic",
        "config": {"workload": WORKLOAD % (args.log_n, NUM_COLS, RATE_BITS, CAP_HEIGHT)},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc,
                         "calibration": {"port_us_per_permutation_per_core": us_perm, "published_leaf_prove": PUBLISHED_LEAF_PROVE},
                         "note": "restated CPU baseline (oracle/), not the reference binary: no cargo here"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })))


def compact(x, sig=4):
    """Round floats to `sig` significant digits (the driver keeps ~1.5 kB of the line)."""
    if isinstance(x, float):
        return float("%.*g" % (sig, x))
    if isinstance(x, dict):
        return {k: compact(v, sig) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [compact(v, sig) for v in x]
    return x


def assemble_line(world, steps, warmup, ms_per_step, value, log_n, leaf_gbs, lde_gbs, hbm_peak, traffic, int_issue, imad_peak,
                  perm_rate, stage, launches, clocks, e2e, parity, native, proofs_all, cpu, exchange="NVLink P2P pull in LDE pass 1"):
    """The contract line (kept under 1.4 kB: the driver stores about 1.5 kB; long-form strings go to the detail file)."""
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic",
        "config": {"workload": "raw commit 2^%dx%d, rate_bits %d, cap_height %d (config 3)" % (log_n, NUM_COLS, RATE_BITS, CAP_HEIGHT),
                   "parallelism": ("coset split x%d, %s" % (world, exchange)) if world > 1 else "1 GPU", "l2": "inputs > L2"},
        "roofline": {"bound": "hbm", "achieved": leaf_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": leaf_gbs / hbm_peak if leaf_gbs else None,
                     "traffic": traffic, "kernel": "hash_rows", "int_issue": int_issue},
        "roofline_lde": {"bound": "hbm", "achieved": lde_gbs, "frac": lde_gbs / hbm_peak if lde_gbs else None},
        "int_pipe": {"imad_per_s": imad_peak, "perms_per_s": perm_rate},
        "stages_ms": {"intt": stage["intt_ms"], "lde": stage["lde_ms"], "leaf_hash": stage["leaf_hash_ms"],
                      "levels": stage["merkle_levels_ms"], "allgather": stage["allgather_ms"]},
        "gpu_launches": int(launches), "clocks": {k: clocks.get(k) for k in ("sm_mhz", "sm_max_mhz", "reasons")},
    }
    if e2e:
        out["e2e"] = {k: e2e[k] for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")}
    if parity is not None:
        out["parity_check"] = {"ok": parity.get("ok"), "rows": parity.get("rows_checked"), "cols": parity.get("columns"),
                               "caps_equal": bool(parity.get("cap_timed_path_equal") and parity.get("cap_e2e_path_equal") is not False)}
        if parity.get("error"):
            out["parity_check"]["error"] = parity["error"][:80]
    if native is not None:
        out["prove"] = {k.replace("_gates", "_g"): native.get(k) for k in ("latency_ms", "latency_ms_gates", "proofs_per_s", "proofs_per_s_gates",
                                                                           "ms_2p15_gates", "ms_2p16_gates") if native.get(k) is not None}
        # config 5: 1024 independent proofs sharded over the job's GPUs (replicas, no data-path collective)
        out["proofs_per_s_all_gpus"] = proofs_all[0] if proofs_all else native.get("proofs_per_s_gates")
        out["batch_1024_proofs_s"] = proofs_all[1] if proofs_all else native.get("batch_1024_s")
    if cpu is not None:
        out["cpu_baseline"] = {"value": cpu["value"], "unit": UNIT, "cores": cpu["cores"], "kind": "port", "sample": cpu["sample"],
                               "calibration": {"port_us_per_perm": cpu["us_perm"], "published_leaf_ms": PUBLISHED_LEAF_PROVE["ms"]}}
        if native is not None and cpu.get("prove_ms") is not None:
            out["prove"]["cpu_port_ms"] = cpu["prove_ms"]
    return out


def parity_check_commit(qpzk, ctx, dev, log_n, values_dev, cap_timed, cap_e2e, seed=22):
    """Outside the timed region: the commitment the benchmark times, checked against the oracle.
    values_dev: the full (C, n) input on this device.  A fresh single-GPU device-resident commit of it must
      (1) pass oracle/scale_check.check_commit: coefficients vs the committed VALUES at a random point for every
          column, 1024 leaf rows x all columns on subgroup cosets in all 8 leaf blocks, their Merkle paths, tree top;
      (2) have the cap the timed loop produced (N > 1: the all-gathered cap of the coset split) and
      (3) the cap of the end-to-end host-buffer path.
    Returns the dict that goes into the JSON line; any mismatch raises."""
    import numpy as np
    from oracle import scale_check
    from oracle import oracle as orc
    orc.set_num_threads()
    t0 = time.perf_counter()
    b = qpzk.PolynomialBatch.from_device(ctx, values_dev.data_ptr(), log_n, values_dev.shape[0], RATE_BITS, CAP_HEIGHT, values=True)
    cap = b.cap
    host = values_dev.cpu().numpy().view(np.uint64)
    r = scale_check.check_commit(b, host, RATE_BITS, num_rows=1024, seed=seed)
    b.free()
    r["cap_timed_path_equal"] = bool((np.asarray(cap_timed).reshape(-1) == cap.reshape(-1)).all())
    r["cap_e2e_path_equal"] = None if cap_e2e is None else bool((np.asarray(cap_e2e).reshape(-1) == cap.reshape(-1)).all())
    r["ok"] = bool(r["ok"] and r["cap_timed_path_equal"] and r["cap_e2e_path_equal"] is not False)
    r["seconds"] = time.perf_counter() - t0
    r["cap0"] = "%016x" % int(cap.reshape(-1)[0])
    return r


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--log-n", type=int, default=22)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-prove", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="N > 1: how the coefficient columns reach the other ranks -- p2p: the first LDE pass reads the peers' "
                         "iNTT outputs over NVLink (fused compute + collective); nccl: all-gather, then compute")
    ap.add_argument("--detail", default=None, help="side file for the long-form record (default gpurun_out/bench_detail_n<N>.json)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    import qpzk

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # Native multi-threaded prover runs FIRST, while this process holds no CUDA context: two processes with contexts
    # on one GPU are time-sliced, which halves the child's throughput.  Every rank does it at once (N > 1: one
    # prover per GPU = BASELINE configs 4/5, replicas without a data-path collective).
    native = {}
    if not args.no_prove:
        native = native_prove_suite(local_rank, world)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    if world != args.gpus and rank == 0:
        print("warning: --gpus %d but WORLD_SIZE %d; using WORLD_SIZE" % (args.gpus, world), file=sys.stderr)
    B = 1 << RATE_BITS
    if world > B or B % world:
        raise SystemExit("the coset split supports 1, 2, 4 or 8 ranks")

    log_n, C_ = args.log_n, NUM_COLS
    n = 1 << log_n
    stream = torch.cuda.current_stream()
    ctx = qpzk.Context(local_rank, stream=stream.cuda_stream)
    ctx.set_timing(True)

    # ---- synthetic input: SplitMix64 seeded 0x51575a4b00000000 + k, rejection-sampled < p, column-major;
    # this rank's column slice resident in HBM (rank 0 keeps the whole matrix for the parity check)
    q = (C_ + world - 1) // world                      # columns per rank (last rank is zero padded)
    c0, c1 = rank * q, min((rank + 1) * q, C_)
    full = splitmix_columns_device(0x51575A4B00000000 + log_n, C_, n, dev)
    if world == 1:
        mine = full
    else:
        mine = torch.zeros((q, n), dtype=torch.int64, device=dev)
        mine[: c1 - c0] = full[c0:c1]
        if rank != 0 or args.no_parity:
            del full
            full = None
    p2p = world > 1 and args.exchange == "p2p"
    cs = None
    if world > 1:
        from qpzk.dist import CosetSplitCommitter        # the multi-GPU commitment lives in the library, not here
        cs = CosetSplitCommitter(ctx, log_n, C_, RATE_BITS, CAP_HEIGHT, rank, world, dev, exchange=args.exchange)
    torch.cuda.synchronize()

    stage = {"intt_ms": 0.0, "lde_ms": 0.0, "leaf_hash_ms": 0.0, "merkle_levels_ms": 0.0, "allgather_ms": 0.0}
    last_cap = [None]

    def step(record):
        if world == 1:
            b = qpzk.PolynomialBatch.from_device(ctx, mine.data_ptr(), log_n, C_, RATE_BITS, CAP_HEIGHT, values=True)
        else:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            cs.intt(mine.data_ptr())
            e0.record()
            b = cs.commit()
            e1.record() if not p2p else None
        if record:
            t = ctx.last_commit_timing()   # synchronises the stream
            for k in ("intt_ms", "lde_ms", "leaf_hash_ms", "merkle_levels_ms"):
                stage[k] += t[k]
            if world > 1 and not p2p:
                stage["allgather_ms"] += e0.elapsed_time(e1) - sum(t.values())   # e0..e1 brackets all-gather + commit
        return b

    def finish(b):
        if world > 1:
            last_cap[0] = cs.gather_cap(b)            # this rank's slice (a tiny D2H read), then the 64-byte all-gather
        else:
            last_cap[0] = b.cap
        b.free()

    # ---- device-resident timing: W warm-up, then exactly K steps between barriers + synchronize
    for _ in range(args.warmup):
        finish(step(False))
    sampler = ClockSampler(local_rank)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    launches0 = ctx.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        finish(step(False))
    ev1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches = ctx.launch_count() - launches0
    clocks = sampler.stop()
    ms_total = ev0.elapsed_time(ev1)
    cap_timed = last_cap[0].cpu().numpy().view(np.uint64).copy() if world > 1 else last_cap[0].copy()
    # per-stage durations from a separate pass (stage events synchronise the stream, so they stay out of the timed loop)
    for _ in range(args.steps):
        finish(step(True))
    for k in stage:
        stage[k] /= args.steps
    t_ms = torch.tensor([ms_total], device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(t_ms.item()) / args.steps
    value = bytes_alg(log_n) / (ms_per_step * 1e-3) / 1e9

    # ---- end to end through the host-buffer C-ABI call (pinned host input, H2D inside, cap D2H inside)
    e2e, cap_e2e = None, None
    if not args.no_e2e:
        host = torch.empty((c1 - c0 if world > 1 else C_, n), dtype=torch.int64).pin_memory()
        host.copy_(mine[: host.shape[0]])
        torch.cuda.synchronize()
        hnp = host.numpy().view(np.uint64)
        copy_stream = torch.cuda.Stream(device=dev)
        pieces = max(1, min(8, host.shape[0]))
        bounds = [host.shape[0] * i // pieces for i in range(pieces + 1)]

        # N > 1: the steps are a stream of commitments (the sweep, an aggregator's proofs), so the input of step k + 1
        # is copied -- in column ranges, on a copy stream -- into a second device buffer while step k computes; every
        # step's host->device copy is inside the timed region, but only the first one is exposed.  (On one GPU
        # qpzk_commit_from_values overlaps the copy with the per-column stages of the same call.)
        bufs2 = [mine, torch.empty_like(mine)] if world > 1 else None
        if world > 1 and q > host.shape[0]:
            bufs2[1][host.shape[0]:].zero_()                          # zero padding columns of the last rank
        copied, consumed = [None, None], [None, None]

        def e2e_prefetch(k):
            dst = bufs2[k & 1]
            if consumed[k & 1] is not None:
                copy_stream.wait_event(consumed[k & 1])                # the iNTT that read this buffer two steps ago
            else:
                copy_stream.wait_stream(stream)
            evs = []
            with torch.cuda.stream(copy_stream):
                for a, z in zip(bounds[:-1], bounds[1:]):
                    dst[a:z].copy_(host[a:z], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(copy_stream)
                    evs.append(ev)
            copied[k & 1] = evs

        def e2e_step(k=0, last=True):
            if world == 1:
                bb = qpzk.PolynomialBatch.from_values(ctx, hnp, RATE_BITS, None, CAP_HEIGHT)
                finish(bb)
                return
            src = bufs2[k & 1]
            if copied[k & 1] is None:
                e2e_prefetch(k)
            evs, copied[k & 1] = copied[k & 1], None
            if not last:
                e2e_prefetch(k + 1)
            for (a, z), ev in zip(zip(bounds[:-1], bounds[1:]), evs):
                stream.wait_event(ev)
                if z > a:
                    cs.intt(src[a:z].data_ptr(), a, z - a)
            if q > host.shape[0]:
                cs.intt(src[host.shape[0]:].data_ptr(), host.shape[0], q - host.shape[0])
            done = torch.cuda.Event()
            done.record(stream)
            consumed[k & 1] = done
            finish(cs.commit())

        for _ in range(2):
            e2e_step()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k in range(args.steps):
            e2e_step(k, last=k + 1 == args.steps)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e_ms = 1e3 * float(dt.item()) / args.steps
        cap_e2e = last_cap[0].cpu().numpy().view(np.uint64).copy() if world > 1 else last_cap[0].copy()
        e2e = {"value": bytes_alg(log_n) / (e2e_ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": e2e_ms,
               "h2d_bytes_per_step": int(host.numel() * 8), "d2h_bytes_per_step": int((1 << CAP_HEIGHT) // world * 32)}
        del host, hnp

    # ---- independent proofs, one prover per GPU (BASELINE configs 4/5: replicas, no data-path collective) ----
    proofs_all = None
    if world > 1 and not args.no_prove:
        v = torch.tensor([native.get("proofs_per_s_gates") or 0.0, native.get("batch_1024_s") or 0.0], device=dev, dtype=torch.float64)
        w = v.clone()
        dist.all_reduce(v)                                  # sum of the per-GPU rates
        dist.all_reduce(w, op=dist.ReduceOp.MAX)            # slowest rank of the sharded batch
        proofs_all = (float(v[0].item()), float(w[1].item()))

    if rank != 0:
        if world > 1:
            dist.barrier()                            # rank 0 finishes its parity check before NCCL goes away
            cs.close()
            dist.destroy_process_group()
        return

    # ---- parity of what was timed (outside the timed region; a mismatch fails the run)
    parity = None
    if not args.no_parity:
        del mine
        ctx.trim()
        try:
            parity = parity_check_commit(qpzk, ctx, dev, log_n, full, cap_timed, cap_e2e, seed=log_n)
        except AssertionError as ex:
            parity = {"ok": False, "error": str(ex)[:200]}

    # ---- rooflines
    hbm_peak, peak_src = measured_peaks()
    L = n << RATE_BITS
    rows_here = L // world
    leaf_bytes = rows_here * C_ * 8 + rows_here * 32                    # read each leaf row once, write its digest
    leaf_gbs = leaf_bytes / (stage["leaf_hash_ms"] * 1e-3) / 1e9 if stage["leaf_hash_ms"] else None
    lde_bytes = 8 * C_ * n * (1 + B) / world                            # read the trace once, write the LDE once
    lde_ms = stage["intt_ms"] + stage["lde_ms"]
    lde_gbs = lde_bytes / (lde_ms * 1e-3) / 1e9 if lde_ms else None
    imad_peak = ctx.bench_imad(8192)
    perms_launch = rows_here * ((C_ + 7) // 8)
    perm_rate_kernel = perms_launch / (stage["leaf_hash_ms"] * 1e-3) if stage["leaf_hash_ms"] else None
    ncu = ncu_record()
    sm_hz = 1e6 * (clocks.get("sm_mhz") or 1965.0)
    int_issue = None
    if ncu.get("inst_per_permutation") and stage["leaf_hash_ms"]:
        # warp instructions issued by one launch (ncu: smsp__inst_executed.sum per permutation x permutations / 32 lanes)
        # over the issue slots of the launch: SMs x 4 schedulers x cycles
        int_issue = ncu["inst_per_permutation"] * perms_launch / 32 / (148 * 4 * sm_hz * stage["leaf_hash_ms"] * 1e-3)
    detail = {
        "config_long": {"l2": "inputs (%.2f GB per rank) and the LDE are larger than the 126 MB L2; no flush needed" % (C_ * n * 8 / world / 1e9),
                        "inputs": "SplitMix64 seed 0x51575a4b00000000 + log_n, rejection-sampled < p, column-major (SURVEY.md 8d)",
                        "bytes_alg": bytes_alg(log_n), "permutations": perms(log_n)},
        "roofline_long": {"kernel": "hash_rows_kernel (Poseidon leaf sponge, %d permutations per row)" % ((C_ + 7) // 8),
                          "traffic_unit": "DRAM bytes read+written per launch (ncu --set full, %s)" % ncu.get("source"),
                          "algorithmic_bytes_per_launch": leaf_bytes, "peak_source": peak_src + " (of measured)",
                          "note": "integer-pipe bound, not HBM bound: int_issue = issued warp instructions / issue slots; "
                                  "HBM fraction reported because the contract asks for it"},
        "roofline_commit": {"bound": "hbm", "achieved": value, "peak": hbm_peak * world, "unit": "GB/s", "frac": value / (hbm_peak * world)},
        "native": native, "stages_ms": stage, "e2e": e2e,
    }
    cpu = None
    if not args.no_cpu_baseline:
        from oracle import oracle as orc
        cores = orc.set_num_threads()
        us_perm = port_calibration(orc)
        k = cpu_sample_log_n(cores, 12.0)
        v, cores, secs = cpu_commit_sample(k)
        cpu = {"value": v, "cores": cores, "sample": "from_values 2^%dx%d, %.1f s" % (k, C_, secs), "us_perm": us_perm,
               "prove_ms": cpu_prove_synthetic()[0] if not args.no_prove else None}
    if not args.no_prove:
        detail["perm_argument"] = [gpu_perm_argument(ctx, dev, k) for k in (13, 18)]
    out = assemble_line(world=world, steps=args.steps, warmup=args.warmup, ms_per_step=ms_per_step, value=value, log_n=log_n,
                        leaf_gbs=leaf_gbs, lde_gbs=lde_gbs, hbm_peak=hbm_peak, traffic=ncu.get("traffic", {}).get(str(log_n)) if world == 1 else None,
                        int_issue=int_issue, imad_peak=imad_peak, perm_rate=perm_rate_kernel, stage=stage, launches=launches,
                        clocks=clocks, e2e=e2e, parity=parity, native=None if args.no_prove else native, proofs_all=proofs_all, cpu=cpu,
                        exchange="NVLink P2P pull in LDE pass 1" if p2p else "NCCL all-gather")
    if parity is not None:
        detail["parity_check"] = parity
    path = args.detail or os.path.join(ROOT, "gpurun_out", "bench_detail_n%d.json" % world)
    try:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "w") as f:
            json.dump({"line": out, "detail": detail}, f, indent=1, default=str)
    except OSError:
        pass
    print(json.dumps(compact(out), separators=(",", ":")))
    if world > 1:
        dist.barrier()
        cs.close()
        dist.destroy_process_group()
    if parity is not None and not parity.get("ok"):
        raise SystemExit("parity check FAILED: %s" % parity)


if __name__ == "__main__":
    main()
