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
  --impl reference   the CPU restatement of the reference algorithm (oracle/, OpenMP on all host cores) on a bounded
            sample (2^16 x 135) of the same workload.  The Rust reference itself cannot be built here (no cargo,
            un-vendored qp-plonky2), see DESIGN.md.
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


def measured_traffic(log_n, world):
    """DRAM bytes (read + write) of one hash_rows_kernel launch from the committed `ncu --set full` capture of the same
    configuration (profiles/r1_traffic.json); None when that configuration was not captured."""
    if world != 1:
        return None
    try:
        with open(os.path.join(ROOT, "profiles", "r1_traffic.json")) as f:
            e = json.load(f)["per_log_n"][str(log_n)]
        return e["dram_read_bytes"] + e["dram_write_bytes"]
    except Exception:
        return None


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


def gpu_prove_synthetic(ctx, proofs=16):
    """Leaf-shaped synthetic prove (BASELINE.json configs 1/2): 3 commits + openings + FRI per proof
    through the host-buffer API, sequential on one context.  Returns (ms per proof, launches per proof)."""
    import numpy as np
    import qpzk
    from qpzk import prover
    shape = prover.LEAF_SHAPE
    rng = np.random.default_rng(1)
    cs = rng.integers(0, P, size=(shape["constants_sigmas"], 1 << shape["log_n"]), dtype=np.uint64)
    circuit = prover.SyntheticCircuit(ctx, cs, shape)
    inputs = [prover.synthetic_inputs(s, shape) for s in range(4)]
    for i in range(3):
        prover.prove_synthetic(circuit, **inputs[i % 4])
    ctx.synchronize()
    l0 = ctx.launch_count()
    t0 = time.perf_counter()
    for i in range(proofs):
        prover.prove_synthetic(circuit, **inputs[i % 4])
    ctx.synchronize()
    dt = (time.perf_counter() - t0) / proofs
    return 1e3 * dt, (ctx.launch_count() - l0) // proofs


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


def native_prove_bench(device, threads, proofs=32, perm=0, zeroize=1, gates=None):
    """The same synthetic prove driven from native code (qp-zk-circuits_b200/host/prove_bench.cpp: C++ threads on the
    C ABI, one context per thread, pinned host buffers).  Returns the parsed JSON line or None."""
    exe = os.path.join(ROOT, "qp-zk-circuits_b200", "prove_bench")
    if not os.path.exists(exe):
        return None
    try:
        runs = []
        for _ in range(3 if threads > 1 else 1):   # multi-threaded throughput varies from run to run: report the median
            r = subprocess.run([exe, "--device", str(device), "--threads", str(threads), "--proofs", str(proofs), "--perm", str(perm),
                                "--zeroize", str(zeroize)] + (["--gates", gates] if gates else []),
                               capture_output=True, text=True, timeout=300)
            runs.append(json.loads(r.stdout.strip().splitlines()[-1]))
        runs.sort(key=lambda d: d["proofs_per_s"])
        out = runs[len(runs) // 2]
        out["proofs_per_s_runs"] = [d["proofs_per_s"] for d in runs]
        return out
    except Exception:
        return None


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


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    from oracle import oracle as orc
    orc.build()
    sample = 16
    rng = np.random.default_rng(16)
    cols = rng.integers(0, P, size=(NUM_COLS, 1 << sample), dtype=np.uint64)
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        b = orc.PolynomialBatch.from_values(cols, RATE_BITS, CAP_HEIGHT)
        dt = time.perf_counter() - t0
        del b
        if i >= args.warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    v = bytes_alg(sample) / (ms * 1e-3) / 1e9
    desc = "one from_values of 2^%d x %d (rate_bits 3, cap_height 4) per step" % (sample, NUM_COLS)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u64 (Goldilocks)", "data": "synthetic",
        "config": {"workload": "raw commit 2^%d x %d, rate_bits 3, cap_height 4 (bounded CPU sample of config 3)" % (sample, NUM_COLS)},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": orc.num_threads(), "kind": "port", "sample": desc,
                         "note": "restated CPU baseline (oracle/, C + OpenMP) -- not the reference binary: no cargo/rustc here"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "prove_synthetic": (lambda r: {"latency_ms": r[0], "proofs_per_s": 1e3 / r[0], "cores": r[1],
                                       "what": "leaf-shaped synthetic prove (2^13; 135/20/16-column commits, openings, FRI) on the CPU restatement"})(cpu_prove_synthetic()),
    }))


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
        tmax = max(1, min(4, (os.cpu_count() or 4) // world))
        native[1] = native_prove_bench(local_rank, 1, 64)          # also warms the clocks
        native[tmax] = native_prove_bench(local_rank, tmax, 64)
        native["perm1"] = native_prove_bench(local_rank, 1, 64, perm=1)
        native["perm%d" % tmax] = native_prove_bench(local_rank, tmax, 64, perm=1)
        native["perm%d_deferred" % tmax] = native_prove_bench(local_rank, tmax, 64, perm=1, zeroize=2)
        gp = os.path.join(ROOT, "qp-zk-circuits_b200", "host", "poseidon_gate.qpgp")   # written by __graft_entry__.build()
        if os.path.exists(gp):
            native["gates1"] = native_prove_bench(local_rank, 1, 64, perm=1, gates=gp)
            native["gates%d_deferred" % tmax] = native_prove_bench(local_rank, tmax, 64, perm=1, zeroize=2, gates=gp)
        native["tmax"] = tmax
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

    # ---- synthetic input: this rank's column slice, canonical uniform values, resident in HBM
    q = (C_ + world - 1) // world                      # columns per rank (last rank is zero padded)
    c0, c1 = rank * q, min((rank + 1) * q, C_)
    gen = torch.Generator(device=dev).manual_seed(0x5157 + log_n)
    full = torch.randint(0, 2**63 - 1, (C_, n), generator=gen, device=dev, dtype=torch.int64)  # < 2^63 < p: canonical
    mine = torch.zeros((q, n), dtype=torch.int64, device=dev)
    mine[: c1 - c0] = full[c0:c1]
    del full
    coeff_mine = torch.empty((q, n), dtype=torch.int64, device=dev)
    coeff_all = torch.empty((world * q, n), dtype=torch.int64, device=dev) if world > 1 else None
    cosets_per = B // world
    cap_all = torch.empty((world, (1 << CAP_HEIGHT) // world * 4), dtype=torch.int64, device=dev) if world > 1 else None
    torch.cuda.synchronize()

    stage = {"intt_ms": 0.0, "lde_ms": 0.0, "leaf_hash_ms": 0.0, "merkle_levels_ms": 0.0, "allgather_ms": 0.0}

    def step(record):
        if world == 1:
            b = qpzk.PolynomialBatch.from_device(ctx, mine.data_ptr(), log_n, C_, RATE_BITS, CAP_HEIGHT, values=True)
        else:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            qpzk._check(ctx.L.qpzk_intt_columns_dev(ctx.h, mine.data_ptr(), coeff_mine.data_ptr(), log_n, q))
            e0.record()
            dist.all_gather_into_tensor(coeff_all, coeff_mine)
            e1.record()
            b = qpzk.PolynomialBatch.cosets_from_coeffs_device(ctx, coeff_all.data_ptr(), log_n, C_, RATE_BITS, CAP_HEIGHT,
                                                               rank * cosets_per, cosets_per)
        if record:
            t = ctx.last_commit_timing()   # synchronises the stream
            for k in ("intt_ms", "lde_ms", "leaf_hash_ms", "merkle_levels_ms"):
                stage[k] += t[k]
            if world > 1:
                stage["allgather_ms"] += e0.elapsed_time(e1)
        return b

    def finish(b):
        cap = torch.from_numpy(b.cap.view(np.int64).reshape(-1))  # tiny D2H read of this rank's cap slice
        if world > 1:
            dist.all_gather_into_tensor(cap_all, cap.to(dev).reshape(1, -1))
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
    e2e = None
    if not args.no_e2e:
        host = torch.empty((c1 - c0 if world > 1 else C_, n), dtype=torch.int64).pin_memory()
        host.copy_(mine[: host.shape[0]])
        torch.cuda.synchronize()
        hnp = host.numpy().view(np.uint64)

        def e2e_step():
            if world == 1:
                bb = qpzk.PolynomialBatch.from_values(ctx, hnp, RATE_BITS, None, CAP_HEIGHT)
                cap = bb.cap
                bb.free()
                return cap
            mine[: host.shape[0]].copy_(host, non_blocking=True)
            bb = step(False)
            finish(bb)

        for _ in range(2):
            e2e_step()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            ts = time.perf_counter()
            e2e_step()
            if os.environ.get("QPZK_BENCH_DEBUG"):
                torch.cuda.synchronize()
                print("e2e step %.1f ms" % (1e3 * (time.perf_counter() - ts)), file=sys.stderr)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e_ms = 1e3 * float(dt.item()) / args.steps
        e2e = {"value": bytes_alg(log_n) / (e2e_ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": e2e_ms,
               "h2d_bytes_per_step": int(host.numel() * 8), "d2h_bytes_per_step": int((1 << CAP_HEIGHT) // world * 32)}

    # ---- independent proofs, one prover per GPU (BASELINE configs 4/5: replicas, no data-path collective) ----
    native_all = None
    if world > 1 and not args.no_prove:
        r = native.get("perm%d_deferred" % native.get("tmax", 0))
        v = torch.tensor([r["proofs_per_s"] if r else 0.0], device=dev, dtype=torch.float64)
        dist.all_reduce(v)
        native_all = float(v.item())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

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
    perm_rate_kernel = rows_here * ((C_ + 7) // 8) / (stage["leaf_hash_ms"] * 1e-3) if stage["leaf_hash_ms"] else None
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u64 (Goldilocks)", "data": "synthetic",
        "config": {"workload": "raw commit (BASELINE config 3): 2^%d rows x %d cols, rate_bits %d, cap_height %d, "
                               "Poseidon-12 Merkle, blinding off" % (log_n, C_, RATE_BITS, CAP_HEIGHT),
                   "l2": "inputs (%.2f GB per rank) and the LDE are larger than the 126 MB L2; no flush needed" % (mine.numel() * 8 / 1e9),
                   "parallelism": "coset split x%d (iNTT column split + NCCL all-gather of coefficients)" % world if world > 1 else "single GPU",
                   "bytes_alg": bytes_alg(log_n), "permutations": perms(log_n)},
        "stages_ms": stage,
        "roofline": {"kernel": "hash_rows_kernel (Poseidon leaf sponge, %d permutations per row)" % ((C_ + 7) // 8),
                     "bound": "hbm", "achieved": leaf_gbs, "peak": hbm_peak, "unit": "GB/s",
                     "frac": leaf_gbs / hbm_peak if leaf_gbs else None, "traffic": measured_traffic(log_n, world),
                     "traffic_unit": "DRAM bytes read+written per launch (ncu --set full, profiles/r1_traffic.json)",
                     "algorithmic_bytes_per_launch": leaf_bytes, "peak_source": peak_src + " (of measured)",
                     "pipes": "ncu: issue slots 71% busy, ALU pipe 80% of peak (profiles/r1_summary.md F)",
                     "note": "integer-pipe bound, not HBM bound: see int_pipe; HBM fraction reported because the contract asks for it"},
        "roofline_lde": {"kernels": "ntt_pass_kernel (iNTT) + ntt_fwd_kernel x2 + lde_last32_tma_kernel", "bound": "hbm", "achieved": lde_gbs, "peak": hbm_peak,
                         "unit": "GB/s", "frac": lde_gbs / hbm_peak if lde_gbs else None, "bytes": lde_bytes},
        "roofline_commit": {"bound": "hbm", "achieved": value, "peak": hbm_peak * world, "unit": "GB/s",
                            "frac": value / (hbm_peak * world)},
        "int_pipe": {"imad_wide_per_s_measured": imad_peak, "poseidon_perms_per_s_in_kernel": perm_rate_kernel,
                     "poseidon_perms_per_s_microbench": ctx.bench_poseidon(32)},
        "gpu_launches": int(launches), "clocks": clocks,
    }
    if e2e:
        out["e2e"] = e2e
    if not args.no_prove:
        ms, nl = gpu_prove_synthetic(ctx)
        out["prove_synthetic"] = {"latency_ms": ms, "proofs_per_s": 1e3 / ms, "gpu_launches_per_proof": int(nl),
                                  "what": "leaf-shaped synthetic prove (BASELINE configs 1/2 stand-in): 2^13 rows; commits of 135 / 20 / 16 "
                                          "random columns, openings at zeta and g*zeta over 256 polynomials, FRI (rate 3, cap 4, 16 PoW bits, "
                                          "28 queries, arity 4); host buffers in, proof out; sequential on one GPU; bit-exact with the CPU "
                                          "restatement (tests/test_gpu_parity.py); witness generation, partial products and quotient evaluation are not part of THIS figure (see native_with_permutation_argument*)"}
        if native_all is not None:
            # min(4, cores/N) host threads per GPU, every rank at once; permutation argument on the device, deferred zeroize
            out["prove_synthetic"]["proofs_per_s_all_gpus"] = native_all
        n1, n4 = native.get(1), native.get(native.get("tmax"))
        if n1 and n4:
            out["prove_synthetic"]["native"] = {
                "latency_ms_1_thread": n1["latency_ms"], "proofs_per_s_1_thread": n1["proofs_per_s"],
                "proofs_per_s_%d_threads" % native["tmax"]: n4["proofs_per_s"], "gpu_launches_per_proof": n1["launches_per_proof"],
                "what": "same proof from C++ host threads on the C ABI (host/prove_bench.cpp), one context per thread, pinned buffers"}
        p1, p4 = native.get("perm1"), native.get("perm%d" % native.get("tmax", 0))
        if p1 and p4:
            out["prove_synthetic"]["native_with_permutation_argument"] = {
                "latency_ms_1_thread": p1["latency_ms"], "proofs_per_s_1_thread": p1["proofs_per_s"],
                "proofs_per_s_%d_threads" % native["tmax"]: p4["proofs_per_s"], "gpu_launches_per_proof": p1["launches_per_proof"],
                "proofs_per_s_%d_threads_runs" % native["tmax"]: p4.get("proofs_per_s_runs"),
                "what": "as above, but the zs/partial-product columns and the quotient chunks are computed on the device from the "
                        "routed wires, the circuit's sigmas and the transcript's challenges (rows a16/a17, permutation argument only)"}
            g1, g4 = native.get("gates1"), native.get("gates%d_deferred" % native["tmax"])
            if g1 and g4:
                out["prove_synthetic"]["native_with_permutation_argument_and_poseidon_gates"] = {
                    "latency_ms_1_thread": g1["latency_ms"], "proofs_per_s_1_thread": g1["proofs_per_s"],
                    "proofs_per_s_%d_threads_deferred_zeroize" % native["tmax"]: g4["proofs_per_s"],
                    "gate_program_ops": g1.get("gate_program_ops"), "gpu_launches_per_proof": g1["launches_per_proof"],
                    "what": "as above, plus the gate constraints of a circuit with upstream's PoseidonGate on every row (123 constraints of "
                            "degree 7), handed over as a qpzk_gate_program and interpreted on the device: everything prove() does after "
                            "witness generation"}
            pd = native.get("perm%d_deferred" % native["tmax"])
            if pd:
                out["prove_synthetic"]["native_with_permutation_argument"].update({
                    "proofs_per_s_%d_threads_deferred_zeroize" % native["tmax"]: pd["proofs_per_s"],
                    "proofs_per_s_%d_threads_deferred_zeroize_runs" % native["tmax"]: pd.get("proofs_per_s_runs"),
                    "zeroize_note": "default: secret-derived buffers are zeroed when released; with several provers on one GPU those "
                                    "fills trigger stalls of 0.2-2 s in about half of the runs (median of 3 reported); "
                                    "qpzk_ctx_set_zeroize(ctx, 2) zeroes them before they leave the context's block cache instead"})
        if not args.no_cpu_baseline:
            cms, cores = cpu_prove_synthetic()
            out["prove_synthetic"]["cpu_port_latency_ms"] = cms
            out["prove_synthetic"]["cpu_cores"] = cores
    if not args.no_prove:
        out["perm_argument"] = [gpu_perm_argument(ctx, dev, k) for k in (13, 18)]
    if not args.no_cpu_baseline:
        v, cores, secs = cpu_commit_sample(16)
        out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                               "sample": "one from_values of 2^16 x %d (same shape, smallest size of the sweep), %.1f s" % (C_, secs),
                               "note": "restated CPU baseline (oracle/, C + OpenMP) -- not the reference binary"}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
