#!/usr/bin/env python
"""bench.py — BN254 G1 MSM throughput (Mpoints/s) on B200, the metric BASELINE.json names.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--logn 20] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A step is one MSM over the rank's shard: 2^logn random BN254 G1 points (k_i * G, generated on the
device) with uniform Fr scalars.  N > 1 shards a (N * 2^logn)-point MSM evenly over the ranks
("scaling": "weak"): every GPU returns one partial point, there is no data-path collective
(SURVEY.md §8e); the N partial points are added on rank 0 after the timed region and checked.

`value`       bases and scalars already resident in HBM; device time from the engine's own CUDA
              events (kgr_last_timing: recorded on the stream the kernels run on), summed over the
              K steps; an L2 flush (write of a 512 MiB buffer) separates the steps.
`e2e`         the same MSM through the reference-facing call (kgr_msm_oneshot = msm_curve_addition
              with host slices): points AND scalars are uploaded from pinned host memory inside
              every timed call and the result is read back; wall clock around the calls.
`roofline`    integer-multiply roofline of the whole pipeline: algorithmic IMADs of
              the reference parameterisation / device time / measured IMAD peak.
`cpu_baseline` the C++ restatement of the reference algorithm (oracle/) on the box's host cores.
`--impl reference` times that CPU restatement alone (the Rust reference cannot be built here).
"""
import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FR_TOP = 0x30644E72E131A029  # top limb of r (bn254/src/fr.rs:11-16)
METRIC = "bn254_g1_msm_throughput"
UNIT = "Mpoints/s"


def ref_window_bits(n):  # groth16/src/msm.rs:7-14
    if n < 4:
        return 1
    if n < 32:
        return 3
    return (n.bit_length() * 69) // 100 + 2


CURVE_IDS = {"bn254_g1": 0, "grumpkin": 1, "bn254_g2": 2}
METRICS = {"bn254_g1": METRIC, "grumpkin": "grumpkin_msm_throughput", "bn254_g2": "bn254_g2_msm_throughput"}


def algorithmic_imads(n, curve_name="bn254_g1"):
    """A(n) = n*W + 2*(2^c - 1)*W point adds with the reference's c and W = ceil(254/c);
    11 field multiplications per add (reference mixed add 9M+2S); 264 32-bit IMADs per multiplication.
    G2: the same adds over Fq2 — 9 Fq2 products at 3 Fq multiplications + 2 Fq2 squares at 2 = 31 per add."""
    c = ref_window_bits(n)
    W = math.ceil(254 / c)
    adds = n * W + 2 * ((1 << c) - 1) * W
    return adds * (31 if curve_name == "bn254_g2" else 11) * 264


def random_scalars(n, seed):
    """Uniform-ish Fr elements as Montgomery limbs (any value < r is a valid Montgomery residue)."""
    rng = np.random.default_rng(seed)
    s = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=(n, 4), dtype=np.uint64)
    s[:, 3] = rng.integers(0, FR_TOP, size=n, dtype=np.uint64)  # top limb < top limb of r  =>  value < r
    return np.ascontiguousarray(s)


class ClockSampler:
    """SM clock / throttle reasons sampled through NVML in a thread while the timed region runs
    (the same counters `nvidia-smi --query-gpu=clocks.sm,clocks_event_reasons.*` prints)."""

    def __init__(self, gpu_index, period_s=0.02):
        self.gpu, self.period, self.rows, self.stop_flag, self.thread, self.h = gpu_index, period_s, [], False, None, None
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((sm, reasons))
            except Exception:
                pass
            time.sleep(self.period)

    def start(self):
        if self.nv is None or self.h is None:
            return
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def stop(self):
        self.stop_flag = True
        if self.thread:
            self.thread.join(timeout=1.0)
        if self.nv is None or not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        nv = self.nv
        smax = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "hw_power_brake_slowdown": 0x80}
        seen = set()
        for _, r in self.rows:
            for name, bit in names.items():
                if r & bit:
                    seen.add(name)
        sm = [x for x, _ in self.rows]
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": smax, "reasons": sorted(seen), "samples": len(sm)}


def cpu_msm_rate(pts, sc, threads, repeats=1, curve=0):
    """Mpoints/s of the restated reference algorithm (oracle) on `threads` host threads."""
    from oracle import oracle as A
    best = None
    out = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        out = A.msm(curve, pts, sc, threads=threads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return pts.shape[0] / best / 1e6, best, out


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU algorithm (C++ restatement; rustc/cargo are absent) on the host cores."""
    if rank != 0:
        return
    from oracle import oracle as A
    cores = os.cpu_count() or 1
    n_full = 1 << args.logn
    # calibrate, then bound the per-step sample so that the whole run stays within a few minutes
    cal_n = 1 << 13
    cid = CURVE_IDS[args.curve]
    pool = A.random_points(cid, cal_n, seed=bytes(range(16)))
    sc_cal = random_scalars(cal_n, 1)
    rate, _, _ = cpu_msm_rate(pool, sc_cal, cores, curve=cid)
    budget_s = 150.0
    n_s = n_full
    while n_s > cal_n and (args.steps + args.warmup) * n_s / (rate * 1e6) > budget_s:
        n_s //= 2
    pts = np.tile(pool, (n_s // cal_n, 1))
    sc = random_scalars(n_s, 2)
    for _ in range(args.warmup):
        cpu_msm_rate(pts, sc, cores, curve=cid)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_msm_rate(pts, sc, cores, curve=cid)
    dt = time.perf_counter() - t0
    value = n_s * args.steps / dt / 1e6
    sample = f"{args.steps} x MSM of 2^{int(math.log2(n_s))} {args.curve} points (workload 2^{args.logn}); {cal_n} distinct points tiled, uniform scalars"
    line = {"impl": "reference", "metric": METRICS[args.curve], "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64 limbs (Montgomery Fq/Fr)",
            "data": "synthetic", "config": {"workload": f"{args.curve} MSM, 2^{args.logn} points per GPU", "sample_points_per_step": n_s},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "C++ restatement of groth16::msm::msm_curve_addition + zkstd arithmetic (oracle/), std::thread over windows like rayon; the Rust reference cannot be compiled in this image"}
    print(json.dumps(line), flush=True)


def run_groth16(args, rank):
    """Groth16 prove latency (BASELINE metric, second half; config #4 scaled as SURVEY H7 suggests): create_proof after witness generation —
    7 FFTs, six G1 and two G2 MSMs, assembly of A, B, C — on the chained x^3 + x + 5 circuit with 2^logm constraints, host buffers in, three
    affine points out.  `--impl reference` times the restated reference code on the same inputs (FFTs on one thread, the eight MSMs on all cores).  The proof is checked against its discrete logs computed from the toxic waste."""
    if rank != 0:
        return
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_next_rows", os.path.join(ROOT, "tools", "bench_next_rows.py"))
    rows = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(rows)
    import kogarashi_b200 as k
    k.init([int(os.environ.get("LOCAL_RANK", "0"))])
    out = []
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        # Measured twice.  In a process that has not run any multi-threaded CPU work yet the call is ~25 % slower however many warm-up proofs
        # precede it (the per-lane host threads run on cores at idle clocks; tools/probe_prover_host.py).  The second pass runs after the CPU
        # baseline of the first has used all host cores, which is the state of a prover process that has generated a witness; both figures are reported.
        rows.bench_groth16(args.logm, out)
        rows.bench_groth16(args.logm, out)
    rec, first = out[1], out[0]
    cpu_ms = rec["cpu_baseline"]["total_seconds"] * 1e3
    ours = args.impl != "reference"
    value = rec["gpu_wall_ms"]["normal"] if ours else cpu_ms
    line = {"metric": "groth16_prove_latency", "value": value, "unit": "ms", "n_gpus": 1, "steps": 5, "warmup": 5, "ms_per_step": value, "higher_is_better": False,
            "scaling": "weak", "vs_baseline": None, "dtype": "u32 limbs (Montgomery Fq/Fr/Fq2)" if ours else "u64 limbs (Montgomery)", "data": "synthetic",
            "config": {"workload": f"Groth16 create_proof after witness generation, chained x^3 + x + 5 circuit, {rec['constraints']} constraints (2^{rec['log_n']} domain)",
                       "timing": "wall clock around Groth16Prover.prove_from_evaluations, best of 5 after 5 warm-up proofs; CRS registered on the GPU"},
            "e2e": {"value": value, "unit": "ms", "h2d_bytes_per_step": (3 * rec["constraints"] + 4 * rec["constraints"]) * 32 if ours else 0,
                    "d2h_bytes_per_step": (1 << rec["log_n"]) * 32 if ours else 0},
            "cpu_baseline": {"value": cpu_ms, "unit": "ms", "cores": rec["cpu_baseline"]["cores"], "kind": "port",
                             "sample": "the whole workload once: " + rec["cpu_baseline"]["note"]},
            "roofline": None, "proof_checked_against_discrete_logs": rec["checked_against_discrete_logs"], "h_bit_exact_with_oracle": rec["h_bit_exact_with_oracle"]}
    if ours:
        line["gpu_launches"] = int(k.launch_count(0))
        line["precomputed_crs_tables_ms"] = rec["gpu_wall_ms"]["precomputed"]
        line["idle_host_ms"] = first["gpu_wall_ms"]["normal"]   # first pass: no multi-threaded CPU work in this process before it
    else:
        line["impl"] = "reference"
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--logn", type=int, default=20, help="log2 of the points per GPU (BASELINE configs[1]: 2^20 on 1 B200)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--curve", default="bn254_g1", choices=["bn254_g1", "grumpkin", "bn254_g2"],
                    help="grumpkin = BASELINE configs[2] (Nova secondary-curve commitment shape); bn254_g2 = next row N3 (the b_g2 MSMs of the Groth16 prover)")
    ap.add_argument("--cpu-sample-logn", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-precompute", action="store_true", help="skip the secondary measurement of the window-collapsed (precomputed table) mode")
    ap.add_argument("--workload", default="msm", choices=["msm", "groth16"],
                    help="groth16 = the other half of BASELINE's metric: prove latency (create_proof after witness generation) on the chained example circuit")
    ap.add_argument("--logm", type=int, default=16, help="log2 of the constraint count for --workload groth16")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.workload == "groth16":
        run_groth16(args, rank)
        return
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    import kogarashi_b200 as k

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    k.init([local_rank])

    n = 1 << args.logn
    curve = CURVE_IDS[args.curve]
    metric = METRICS[args.curve]
    pt_bytes = 128 if curve == k.BN254_G2 else 64
    # this rank's shard of the (world * n)-point vector: bases k_i*G for global indices [rank*n, (rank+1)*n)
    bases, ks = k.Bases.generate(curve, n, seed=1000 + rank, return_scalars=True)
    sc = random_scalars(n, 77 + rank)
    sc_pinned = torch.from_numpy(sc.view(np.int64)).pin_memory()
    d_sc = sc_pinned.cuda(non_blocking=False)
    pts_host = bases.download()
    pts_pinned = torch.from_numpy(pts_host.view(np.int64)).pin_memory()
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: inputs resident in HBM -------------------------------------------------------------
    for _ in range(args.warmup):
        out = k.msm_device(bases, d_sc.data_ptr(), n)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    launches0 = k.launch_count(0)
    dev_ms, phases = 0.0, {}
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        out = k.msm_device(bases, d_sc.data_ptr(), n)
        ms, shape = k.last_timing(0)
        dev_ms += ms["total"]
        for key, v in ms.items():
            phases[key] = phases.get(key, 0.0) + v / args.steps
    barrier()
    wall_ms = (time.perf_counter() - t_wall0) * 1e3
    launches = k.launch_count(0) - launches0
    clocks = sampler.stop()

    # ---- e2e: host buffers through the reference-facing call ----------------------------------------
    for _ in range(2):
        k.msm_oneshot_ptr(curve, pts_pinned.data_ptr(), n, sc_pinned.data_ptr(), n)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out_e2e = k.msm_oneshot_ptr(curve, pts_pinned.data_ptr(), n, sc_pinned.data_ptr(), n)
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    # registered bases, host scalars (what a prover holding the CRS on the GPU pays per MSM)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out_reg = k.msm_host_ptr(bases, sc_pinned.data_ptr(), n)
    barrier()
    e2e_reg_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    assert (k.to_affine(curve, out_e2e) == k.to_affine(curve, out)).all() and (k.to_affine(curve, out_reg) == k.to_affine(curve, out)).all()

    # ---- optional mode for reused vectors (CRS / Pedersen key): window-collapsing table built once at registration --------
    pre = None
    if not args.no_precompute:
        t0 = time.perf_counter()
        bases.precompute(0)
        pre_build_s = time.perf_counter() - t0
        for _ in range(args.warmup):
            out_pre = k.msm_device(bases, d_sc.data_ptr(), n)
        barrier()
        pre_ms = 0.0
        for _ in range(args.steps):
            flush.zero_()
            torch.cuda.synchronize()
            out_pre = k.msm_device(bases, d_sc.data_ptr(), n)
            ms_p, shape_p = k.last_timing(0)
            pre_ms += ms_p["total"]
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            out_pre = k.msm_host_ptr(bases, sc_pinned.data_ptr(), n)
        barrier()
        pre_e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
        assert (k.to_affine(curve, out_pre) == k.to_affine(curve, out)).all()
        pre = (pre_ms / args.steps, pre_e2e_ms, pre_build_s, shape_p)

    # ---- max over ranks, partial sums to rank 0 -----------------------------------------------------
    from kogarashi_b200 import sharding
    stats = torch.tensor([dev_ms, e2e_ms, e2e_reg_ms, wall_ms, pre[0] if pre else 0.0, pre[1] if pre else 0.0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
    parts = sharding.gather_partials(out, device="cuda")  # one 96-byte point per rank, after the timed region
    dev_ms, e2e_ms, e2e_reg_ms, wall_ms, pre_ms_step, pre_e2e_ms = [float(x) for x in stats.cpu()]
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    total_aff = k.to_affine(curve, sharding.combine_partials(curve, parts))

    ms_per_step = dev_ms / args.steps
    value = world * n / ms_per_step / 1e3
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    imad_peak_t = 148 * 64 * 1.965e9 / 1e12  # nominal = measured by kgr_microbench on this pool (profiles/r01_microbench.md): 18.5 T IMAD/s
    alg = algorithmic_imads(n, args.curve)
    achieved_t = alg / (ms_per_step * 1e-3) / 1e12
    hbm_bytes = (pt_bytes + 32) * n  # 64 B point (128 B on G2) + 32 B scalar per pair
    # DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture of this workload (profiles/ncu_traffic.json)
    traffic, pipe_util = None, None
    try:
        ncu = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(f"{args.curve}/{args.logn}", {})
        traffic, pipe_util = ncu.get("dram_bytes_per_launch"), ncu.get("fmaheavy_pipe_active_frac")
    except Exception:
        pass
    # work this implementation actually issues (SURVEY §8d "executed-work fraction"): n*W mixed additions in XYZZ (8M + 2S = 10 products; 28 over Fq2)
    # and about 2*W*B general additions in the bucket reduction (12M + 2S = 14; 40 over Fq2), 264 IMAD per 254-bit product
    g2 = args.curve == "bn254_g2"
    exec_imads = (n * shape["W"] * (28 if g2 else 10) + 2 * shape["W"] * shape["B"] * (40 if g2 else 14)) * 264
    roofline = {"bound": "imad", "achieved": achieved_t, "peak": imad_peak_t, "unit": "T IMAD/s", "frac": achieved_t / imad_peak_t, "traffic": traffic,
                "peak_source": "148 SM x 64 IMAD/clk x 1.965 GHz; kgr_microbench measured 18.5 T mad.lo.u32/s on this pool (MEASURED_PEAKS.json has no integer figure)",
                "kernel": "whole pipeline; k_accumulate is the dominant kernel (see phases_ms)",
                "algorithmic_imads_per_launch": alg,
                "executed": {"imads_per_launch": exec_imads, "achieved": exec_imads / (ms_per_step * 1e-3) / 1e12, "frac": exec_imads / (ms_per_step * 1e-3) / 1e12 / imad_peak_t,
                             "note": "additions this implementation issues (signed digits, its own window size, XYZZ formulas) x products per addition x 264"},
                "ncu_fmaheavy_pipe_active_frac": pipe_util,
                # the dominant kernel on its own: the n*W bucket additions of the reference's inner loop (msm.rs:25-33) over k_accumulate's
                # live CUDA-event duration (phases_ms.accumulate); the 2*(2^c - 1)*W running-sum additions belong to the reduce kernels
                "dominant_kernel": (lambda adds: {"name": "k_accumulate", "ms": phases.get("accumulate"), "algorithmic_imads": adds,
                                                  "achieved": adds / (phases["accumulate"] * 1e-3) / 1e12, "frac": adds / (phases["accumulate"] * 1e-3) / 1e12 / imad_peak_t})(
                    n * math.ceil(254 / ref_window_bits(n)) * (31 if args.curve == "bn254_g2" else 11) * 264),
                "hbm": {"achieved_gbs": hbm_bytes / (ms_per_step * 1e-3) / 1e9, "peak_gbs": peaks.get("hbm_gbs"), "note": f"{pt_bytes + 32} B/point algorithmic traffic; the path is multiply-bound, not HBM-bound"}}
    line = {"metric": metric, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32 limbs (Montgomery Fq/Fr, 8x32-bit)", "data": "synthetic",
            "config": {"workload": f"{args.curve} MSM, 2^{args.logn} points per GPU ({world * n} total), uniform scalars, bases k_i*G", "l2": "flushed between steps (512 MiB write)",
                       "shape": shape, "timing": "sum of per-step CUDA-event durations on the engine stream; wall_ms_per_step includes the flush and host gaps"},
            "wall_ms_per_step": wall_ms / args.steps, "phases_ms": phases, "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": world * n / e2e_ms / 1e3, "unit": UNIT, "h2d_bytes_per_step": (pt_bytes + 32) * n,
                    # the call cuts itself into min(4, n >> 19) pipelined pieces; every piece returns its W window sums (one XYZZ point each)
                    "d2h_bytes_per_step": max(1, min(4, n >> 19)) * shape["W"] * 2 * pt_bytes, "ms_per_step": e2e_ms,
                    "call": "kgr_msm_oneshot (points + scalars uploaded from pinned host memory every call)"},
            "e2e_registered": {"value": world * n / e2e_reg_ms / 1e3, "unit": UNIT, "h2d_bytes_per_step": 32 * n, "d2h_bytes_per_step": shape["W"] * 2 * pt_bytes, "ms_per_step": e2e_reg_ms,
                               "call": "kgr_msm (bases registered once, scalars uploaded every call)"},
            "roofline": roofline, "result_is_identity": bool(int(total_aff[-1]))}
    if pre:
        line["precomputed_bases"] = {
            "note": "secondary numbers, NOT the headline: bases registered with kgr_bases_precompute (table 2^(c*w)*P_i built once, W x the memory); "
                    "all windows share one bucket set, no final doublings; same inputs, same result",
            "value": world * n / pre_ms_step / 1e3, "unit": UNIT, "ms_per_step": pre_ms_step,
            "e2e_registered": {"value": world * n / pre_e2e_ms / 1e3, "ms_per_step": pre_e2e_ms, "h2d_bytes_per_step": 32 * n},
            "table_build_s": pre[2], "shape": pre[3]}

    # ---- checksum of the result: bases are k_i*G, so the MSM must equal (sum k_i s_i) * G --------------
    if world == 1:
        from oracle import oracle as A  # checker only
        from oracle import pyref as B
        r = B.FQ if curve == A.GRUMPKIN else B.FR
        acc = 0
        for a, b in zip(ks, sc):
            acc += B.from_mont(B.limbs_to_int(a), r) * B.from_mont(B.limbs_to_int(b), r)
        g = A.generator(curve)
        one = A.field_op(A.FIELD_FR if curve == A.GRUMPKIN else A.FIELD_FQ, "to_mont", np.array([1, 0, 0, 0], dtype=np.uint64))
        z = np.concatenate([one, np.zeros(4, dtype=np.uint64)]) if curve == A.BN254_G2 else one  # Z = 1 (in Fq2: 1 + 0u)
        exp = A.to_affine(curve, A.scalar_point(curve, np.concatenate([g, z]), np.array(B.int_to_limbs(B.to_mont(acc % r, r)), dtype=np.uint64)))
        line["checksum_ok"] = bool((exp == total_aff).all())

    # ---- cpu_baseline: the restated reference algorithm on the host cores, bounded sample ------------
    if not args.no_cpu_baseline and world == 1:
        cores = os.cpu_count() or 1
        s_logn = args.cpu_sample_logn or min(args.logn, 20)
        ns = 1 << s_logn
        rate, secs, cpu_out = cpu_msm_rate(pts_host[:ns], sc[:ns], cores, curve=curve)
        from oracle import oracle as A
        line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"one MSM over the first 2^{s_logn} pairs of the same inputs, {secs:.2f} s, C++ restatement of the reference algorithm (c={ref_window_bits(ns)})"}
        if ns == n:
            line["cpu_baseline"]["bit_exact_with_gpu"] = bool((A.to_affine(curve, cpu_out) == total_aff).all())
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
