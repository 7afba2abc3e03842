#!/usr/bin/env python
"""bench.py - headline benchmark: batched rnea_grad (iiwa14, FP64, 2^20 knot points per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--robot iiwa14|hyq|atlas] [--op rnea_grad|minv|rnea] [--dtype f64|f32] [--batch B]

One "step" = one pass of the hot path over one batch of synthetic (q, qd, qdd) knot points.
Prints ONE JSON line (rank 0).  Metric = evals/s of BASELINE.json ("rnea_grad & minv evals/sec
at batch 1M"); default workload = configs[1] (iiwa14 rnea_grad FP64, 1M points on one B200).
Multi-GPU (torchrun, one rank per GPU): the batch axis is sharded with no data-path collective;
`value` is the weak-scaling figure (every rank owns `--batch` knot points, time = max over ranks);
the line also carries a `strong` record (`--batch` knot points in total, split over the ranks) and
`gather_ms` (NCCL gather of the strong-scaling result to rank 0 / to all ranks).  The default
invocation adds `quadrants`: the other three cells of the metric (minv iiwa14, rnea_grad Atlas,
minv Atlas at the same batch), each with its own time, roofline fraction and clocks.

`--impl reference` times the reference's own CPU implementation (oracle/_ref = byte-compiled
unmodified RBDReference.py when staged, else the oracle port) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


# ------------------------------------------------------------------------------------------------
# CPU reference arm / cpu_baseline (the only place bench.py executes oracle/)
# ------------------------------------------------------------------------------------------------
_W = {}


def _cpu_worker_init(robot_name, op):
    try:
        from threadpoolctl import threadpool_limits
        _W["limit"] = threadpool_limits(1)
    except Exception:
        pass
    from rbdreference_b200 import robots
    from oracle import build_ref
    from oracle.rbd_oracle import ScalarOracle
    rb = robots.by_name(robot_name)
    Ref = build_ref.load_reference()
    _W["impl"] = Ref(rb) if Ref is not None else ScalarOracle(rb)
    _W["op"] = op


def _cpu_worker_run(args):
    q, qd, qdd = args
    impl, op = _W["impl"], _W["op"]
    acc = 0.0
    for k in range(q.shape[0]):
        if op == "rnea_grad":
            acc += float(impl.rnea_grad(q[k], qd[k], qdd[k])[0, 0])
        elif op == "minv":
            acc += float(impl.minv(q[k])[0, 0])
        elif op == "crba":
            acc += float(impl.crba(q[k])[0, 0])
        elif op == "fd":
            acc += float(impl.forward_dynamics(q[k], qd[k], qdd[k])[0])
        elif op == "fd_grad":
            acc += float(impl.forward_dynamics_grad(q[k], qd[k], qdd[k])[0][0, 0])
        elif op == "ee_grad":
            acc += float(np.asarray(impl.end_effector_pose_gradient(q[k])[0])[0, 0])
        else:
            acc += float(impl.rnea(q[k], qd[k], qdd[k])[0][0])
    return acc


class CpuReference:
    """Runs the reference (or the oracle port) over samples with one process per host core."""

    def __init__(self, robot_name, op):
        from oracle import build_ref
        self.kind = "reference" if build_ref.load_reference() is not None else "port"
        try:
            self.cores = len(os.sched_getaffinity(0))
        except AttributeError:
            self.cores = os.cpu_count() or 1
        self.cores = max(1, min(self.cores, 128))
        ctx = mp.get_context("fork")
        self.pool = ctx.Pool(self.cores, initializer=_cpu_worker_init, initargs=(robot_name, op))

    def run(self, q, qd, qdd):
        """Evaluate every row once; returns elapsed seconds."""
        chunks = [c for c in np.array_split(np.arange(q.shape[0]), self.cores) if len(c)]
        t0 = time.perf_counter()
        self.pool.map(_cpu_worker_run, [(q[c], qd[c], qdd[c]) for c in chunks])
        return time.perf_counter() - t0

    def close(self):
        self.pool.close()
        self.pool.join()


def per_eval_cpu_seconds(robot_name, op):
    """Rough single-core cost used only to size the bounded sample."""
    base = {"iiwa14": 6e-3, "hyq": 9e-3, "atlas": 65e-3, "iiwa14_fb": 20e-3, "hyq_fb": 30e-3,
            "atlas_fb": 120e-3}.get(robot_name, 20e-3)
    return base * {"rnea_grad": 1.0, "minv": 0.15, "rnea": 0.09, "crba": 0.1, "fd": 0.25, "fd_grad": 1.6,
                   "ee_grad": 0.25}[op]


# ------------------------------------------------------------------------------------------------
# clocks sampling (nvidia-smi during the timed region)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock / power / throttle reasons of one GPU, polled every ~2 ms through NVML on a thread
    (nvidia-smi -lms 20 if NVML is unavailable); `window(t0, t1)` summarises one timed region."""
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, gpu_index):
        self.samples = []          # (t, sm_mhz, power_w, set(reasons))
        self.smax = None
        self.proc = None
        self._stop = False
        self.source = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[gpu_index]) if vis and vis.split(",")[gpu_index].isdigit() else gpu_index
            h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
            get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                pynvml.nvmlDeviceGetCurrentClocksThrottleReasons

            def poll():
                while not self._stop:
                    try:
                        clk = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                        r = int(get_reasons(h))
                        try:
                            pw = pynvml.nvmlDeviceGetPowerUsage(h) / 1e3
                        except Exception:
                            pw = None
                        self.samples.append((time.perf_counter(), clk, pw, {k for k, v in bits.items() if r & v}))
                    except Exception:
                        pass
                    time.sleep(0.001)

            self.thread = threading.Thread(target=poll, daemon=True)
            self.thread.start()
            self.source = "nvml"
            return
        except Exception:
            pass
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.QUERY, "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
            self.source = "nvidia-smi"
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.strip().split(",")]
            if len(parts) < 7:
                continue
            try:
                clk, mx = float(parts[0]), float(parts[1])
            except ValueError:
                continue
            self.smax = mx
            try:
                pw = float(parts[2])
            except ValueError:
                pw = None
            self.samples.append((time.perf_counter(), clk, pw,
                                 {nm for nm, val in zip(self.NAMES, parts[3:7]) if val.lower().startswith("active")}))

    def window(self, t0, t1):
        if self.source is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.source == "nvidia-smi":
            time.sleep(0.05)
        sel = [x for x in list(self.samples) if t0 - 0.002 <= x[0] <= t1 + 0.002]
        if not sel:                 # region shorter than the sampling period: the nearest samples
            allx = sorted(list(self.samples), key=lambda x: abs(x[0] - 0.5 * (t0 + t1)))
            sel = allx[:2]
        sm = [x[1] for x in sel]
        power = [x[2] for x in sel if x[2] is not None]
        reasons = set()
        for x in sel:
            reasons |= x[3]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.smax, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power) if power else None, "source": self.source}

    def stop(self, t0=None, t1=None):
        out = self.window(t0, t1) if t0 is not None else None
        self._stop = True
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        return out


# ------------------------------------------------------------------------------------------------
def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            d = json.load(open(path))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def traffic_from_profile(robot, op, dtype, batch=None):
    """dram bytes per launch of the dominant kernel from the committed ncu summary, if any (the Atlas summaries were
    taken on 2^18 knot points; the stored figure is scaled to 2^20 and rescaled here to the batch of the run)."""
    path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    try:
        d = json.load(open(path))
        t = d.get("%s/%s/%s" % (robot, op, dtype))
        if t is not None and batch is not None:
            t = int(t * (batch / float(1 << 20)))
        return t
    except Exception:
        return None


def executed_from_profile(robot, op, dtype):
    """What the dominant kernel's pipes actually did (ncu, committed under profiles/): the equivalent-work fraction can
    exceed 1 because the kernels execute fewer flops than the reference's recursion."""
    path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    try:
        return json.load(open(path)).get("executed", {}).get("%s/%s/%s" % (robot, op, dtype))
    except Exception:
        return None


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--robot", default="iiwa14")
    ap.add_argument("--op", default="rnea_grad", choices=["rnea_grad", "minv", "rnea", "crba", "fd", "fd_grad", "ee_grad"])
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--batch", type=int, default=1 << 20, help="knot points per GPU (weak scaling) / in total (strong record)")
    ap.add_argument("--variant", type=int, default=0,
                    help="kernel family: 0 auto, 1 generic, 2 world/thread, 3 cooperative, 4 hybrid minv, 5 lane minv, 7 chain rnea_grad")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-quadrants", action="store_true", help="skip the other three cells of the metric")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling and gather records (N > 1)")
    return ap.parse_args()


L2_BYTES = 126e6


def workload_name(args):
    return "%s %s %s, %d knot points per GPU" % (args.robot, args.op, "FP64" if args.dtype == "f64" else "FP32", args.batch)


def config_for(args, n, io_bytes_per_eval=None):
    """`config` of the JSON line - the same dict in both arms (ours / reference)."""
    if io_bytes_per_eval is None:
        from rbdreference_b200 import robots
        from rbdreference_b200.model import compile_model
        rb = robots.by_name(args.robot)
        try:
            io_bytes_per_eval = compile_model(rb).io_bytes(args.op, 8 if args.dtype == "f64" else 4)
        except Exception:
            io_bytes_per_eval = 0
    foot = io_bytes_per_eval * args.batch
    return {"workload": workload_name(args), "robot": args.robot, "op": args.op, "n_dof": n,
            "batch_per_gpu": args.batch, "sharding": "batch axis, contiguous slices, no collective",
            "l2": ("inputs + outputs of one step (%.0f MB) are larger than the 126 MB L2" % (foot / 1e6)) if foot > 2 * L2_BYTES
            else ("inputs + outputs of one step are %.0f MB: a 256 MB buffer is rewritten between timed steps to flush the L2" % (foot / 1e6))}


def synth_host(n, B, seed, robot_name=None):
    rng = np.random.default_rng(seed)
    if robot_name and robot_name.endswith("_fb"):           # floating base: q has n + 1 entries, unit quaternion
        from rbdreference_b200 import robots
        return robots.by_name(robot_name).random_state(rng, B)
    return (rng.uniform(-np.pi, np.pi, (B, n)), rng.uniform(-1, 1, (B, n)), rng.uniform(-1, 1, (B, n)))


# ------------------------------------------------------------------------------------------------
# reference arm
# ------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from rbdreference_b200 import robots
    n = robots.by_name(args.robot).get_num_vel()
    cpu = CpuReference(args.robot, args.op)
    # bounded sample per step: ~0.4 s of wall time on this host
    per = per_eval_cpu_seconds(args.robot, args.op)
    sample = int(max(cpu.cores, min(4096, round(0.4 * cpu.cores / per))))
    q, qd, qdd = synth_host(n, sample, 0xB200, args.robot)
    for _ in range(max(args.warmup, 1)):
        cpu.run(q, qd, qdd)
    t = 0.0
    for _ in range(args.steps):
        t += cpu.run(q, qd, qdd)
    cpu.close()
    value = sample * args.steps / t
    line = {
        "impl": "reference", "metric": "%s evals/sec" % args.op, "value": value, "unit": "evals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.dtype, "data": "synthetic",
        "config": config_for(args, n),
        "note": "CPU numpy reference (FP64 arithmetic); each step evaluates a bounded sample of the workload",
        "cpu_baseline": {"value": value, "unit": "evals/s", "cores": cpu.cores, "kind": cpu.kind,
                         "sample": "%d knot points per step (same seeded distribution), one process per core" % sample},
        "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
class Workload:
    """Synthetic device-resident inputs + the step closure of one (robot, op, dtype, batch)."""

    def __init__(self, torch, robot_name, op, dtype, B, dev, seed):
        from rbdreference_b200 import RBDReference, robots
        self.torch, self.op, self.B, self.dev = torch, op, B, dev
        tdtype = torch.float64 if dtype == "f64" else torch.float32
        self.itemsize = 8 if dtype == "f64" else 4
        self.eng = eng = RBDReference(robots.by_name(robot_name), dtype=tdtype)
        n = self.n = eng.n
        # synthetic inputs, resident in HBM: same fp64 draws for both precisions
        gen = torch.Generator(device=dev).manual_seed(seed)
        q = (torch.rand(B, eng.nq, generator=gen, device=dev, dtype=torch.float64) * 2 - 1) * np.pi
        if eng.floating_base:                          # q[0:7] = base position + unit quaternion
            if op not in ("rnea", "rnea_grad", "minv"):
                raise SystemExit("bench.py: floating-base robots support --op rnea | rnea_grad | minv")
            q[:, 0:3] /= np.pi
            q[:, 3:7] /= q[:, 3:7].norm(dim=1, keepdim=True)
        self.q = q = q.to(tdtype)
        self.qd = qd = (torch.rand(B, n, generator=gen, device=dev, dtype=torch.float64) * 2 - 1).to(tdtype)
        self.qdd = qdd = (torch.rand(B, n, generator=gen, device=dev, dtype=torch.float64) * 2 - 1).to(tdtype)
        self.out = None
        if op == "rnea_grad":
            self.out = out = torch.empty(B, n, 2 * n, dtype=tdtype, device=dev)
            self.step = lambda: eng.rnea_grad(q, qd, qdd, out=out)
        elif op == "minv":
            self.out = out = torch.empty(B, n, n, dtype=tdtype, device=dev)
            self.step = lambda: eng.minv(q, out=out)
        elif op == "crba":
            self.out = out = torch.empty(B, n, n, dtype=tdtype, device=dev)
            self.step = lambda: eng.crba(q, out=out)
        elif op == "fd":
            self.step = lambda: eng.forward_dynamics(q, qd, qdd)          # qdd plays the torque u
        elif op == "fd_grad":
            self.step = lambda: eng.forward_dynamics_grad(q, qd, qdd)
        elif op == "ee_grad":
            self.step = lambda: eng.end_effector_pose_gradient(q)          # every leaf joint, default offset
        else:
            self.step = lambda: eng.rnea(q, qd, qdd, outputs="c")
        self.flops = eng.model.flops(op)
        self.io_bytes = eng.model.io_bytes(op, self.itemsize)
        self.flush = self.io_bytes * B <= 2 * L2_BYTES      # footprint not clearly larger than the L2: flush between steps
        self._flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if self.flush else None

    def timed(self, steps, warmup, barrier):
        """W warm-up steps, then exactly `steps` timed steps bracketed by barrier + synchronize.
        -> (total_ms on this rank, per-step ms list, launches, wall t0, wall t1)"""
        torch = self.torch
        for _ in range(warmup):
            self.step()
        barrier()
        launches0 = self.eng.launch_count()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * steps)]
        t0 = time.perf_counter()
        for k in range(steps):
            if self.flush:
                self._flush_buf.fill_(k & 1)            # outside the event pair of the step
            ev[2 * k].record()
            self.step()
            ev[2 * k + 1].record()
        barrier()
        t1 = time.perf_counter()
        launches = self.eng.launch_count() - launches0
        per_step = [ev[2 * k].elapsed_time(ev[2 * k + 1]) for k in range(steps)]
        total = float(sum(per_step)) if self.flush else ev[0].elapsed_time(ev[-1])
        return total, per_step, launches, t0, t1


def rooflines(w, kernel_ms, fma_peak_tflops, robot, op, dtype):
    hbm_peak, hbm_src = measured_peaks()
    ach_tflops = w.flops * w.B / (kernel_ms * 1e-3) / 1e12
    ach_gbs = w.io_bytes * w.B / (kernel_ms * 1e-3) / 1e9
    traffic = traffic_from_profile(robot, op, dtype, w.B)
    roofline = {
        "bound": "fp64_fma" if dtype == "f64" else "fp32_fma",
        "achieved": ach_tflops, "peak": fma_peak_tflops, "unit": "TFLOP/s",
        "frac": (ach_tflops / fma_peak_tflops) if fma_peak_tflops else None,
        "traffic": traffic,
        "peak_source": "FMA micro-benchmark (rbd_measure_fma_peak) on this GPU just before the timed region",
        "flops_per_eval": w.flops, "kernel_ms": kernel_ms, "executed": executed_from_profile(robot, op, dtype),
        "note": "achieved = SURVEY.md 8d algorithmic flops x evals / kernel time (equivalent work: the kernels execute "
                "fewer flops than the reference's recursion; executed-pipe utilisation is in profiles/); tensor cores are not applicable",
    }
    roofline_hbm = {"bound": "hbm", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak,
                    "traffic": traffic, "bytes_per_eval": w.io_bytes, "peak_source": hbm_src}
    return roofline, roofline_hbm


def run_ours(args):
    import torch
    import torch.distributed as dist

    from rbdreference_b200 import RBDReference
    from rbdreference_b200 import dist as rdist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    if args.variant:
        RBDReference.set_kernel_variant(args.variant)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    B = args.batch
    w = Workload(torch, args.robot, args.op, args.dtype, B, dev, 0xB200 + rank)
    eng, n = w.eng, w.n

    # measured FMA peak of this GPU (denominator of the compute roofline), before the timed region
    import ctypes
    peak = ctypes.c_double(0.0)
    ms = ctypes.c_double(0.0)
    with torch.cuda.device(dev):
        rc = eng._lib.rbd_measure_fma_peak(1 if args.dtype == "f64" else 0, ctypes.byref(peak), ctypes.byref(ms),
                                           torch.cuda.current_stream(dev).cuda_stream)
    fma_peak_tflops = peak.value / 1e12 if rc == 0 else None

    # started before the warm-up so that the sampler is already running when the (short) timed region begins
    sampler = ClockSampler(local_rank) if rank == 0 else None
    warm = max(args.warmup, 3)
    total_ms, per_step, launches, t_wall0, t_wall1 = w.timed(args.steps, warm, barrier)
    clocks = sampler.window(t_wall0, t_wall1) if sampler else None
    total_ms_max = max_over_ranks(total_ms)
    ms_per_step = total_ms_max / args.steps
    value = world * B * args.steps / (total_ms_max * 1e-3)
    kernel_ms = float(np.mean(per_step))           # one launch per step; events on the launch stream
    roofline, roofline_hbm = rooflines(w, kernel_ms, fma_peak_tflops, args.robot, args.op, args.dtype)

    # ---- e2e: the public API with HOST buffers, H2D + kernel + D2H inside the timed region ----
    e2e = None
    if not args.no_e2e and args.op in ("rnea_grad", "minv", "crba", "rnea"):
        e2e = measure_e2e(args, eng, torch, dist, dev, rank, world, n, B)

    # ---- strong scaling + gather (N > 1): --batch knot points in total, split over the ranks ----
    strong = weak = gather = None
    if world > 1 and not args.no_strong:
        weak = {"value": value, "unit": "evals/s", "batch_per_gpu": B, "ms_per_step": ms_per_step}
        lo, hi = rdist.shard_bounds(B, rank, world)
        ws = Workload(torch, args.robot, args.op, args.dtype, hi - lo, dev, 0xB200 + rank)
        t_ms, ps, _, _, _ = ws.timed(args.steps, warm, barrier)
        t_ms = max_over_ranks(t_ms)
        strong = {"value": B * args.steps / (t_ms * 1e-3), "unit": "evals/s", "batch_total": B,
                  "batch_per_gpu": hi - lo, "ms_per_step": t_ms / args.steps,
                  "l2": "flushed between steps" if ws.flush else "footprint larger than L2"}
        if ws.out is not None:
            gather = {}
            for name, fn in (("to_rank0", lambda: rdist.gather_to_rank(ws.out, B, 0)), ("to_all", lambda: rdist.gather_to_all(ws.out, B))):
                fn()
                barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                reps = 5
                e0.record()
                for _ in range(reps):
                    r = fn()
                    del r
                e1.record()
                barrier()
                gather[name + "_ms"] = max_over_ranks(e0.elapsed_time(e1) / reps)
            gather["bytes_total"] = int(B * int(np.prod(ws.out.shape[1:])) * ws.itemsize)
            gather["note"] = "NCCL over NVLink: result slices of the strong-scaling batch gathered onto rank 0 (send/recv) and onto every rank (all-gather); compute-only time is strong.ms_per_step"
        del ws

    # ---- the other three cells of the metric (default invocation, one GPU) ----
    quadrants = None
    default_cell = (args.robot, args.op) == ("iiwa14", "rnea_grad")
    if world == 1 and default_cell and not args.no_quadrants:
        quadrants = []
        for robot_name, op in (("iiwa14", "minv"), ("atlas", "rnea_grad"), ("atlas", "minv")):
            del w.out
            torch.cuda.empty_cache()
            wq = Workload(torch, robot_name, op, args.dtype, B, dev, 0xB200 + rank)
            qsteps = max(3, min(args.steps, 10))
            t_ms, ps, ln, q0, q1 = wq.timed(qsteps, 3, barrier)
            rf, rfh = rooflines(wq, float(np.mean(ps)), fma_peak_tflops, robot_name, op, args.dtype)
            quadrants.append({"robot": robot_name, "op": op, "dtype": args.dtype, "batch": B, "value": B * qsteps / (t_ms * 1e-3),
                              "unit": "evals/s", "steps": qsteps, "ms_per_step": t_ms / qsteps, "gpu_launches": int(ln),
                              "roofline": rf, "roofline_hbm": rfh, "clocks": sampler.window(q0, q1) if sampler else None})
            w.out = None
            del wq
            torch.cuda.empty_cache()
    if sampler:
        sampler.stop()

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline = measure_cpu_baseline(args, n)

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    line = {
        "metric": "%s evals/sec" % args.op, "value": value, "unit": "evals/s", "n_gpus": world,
        "steps": args.steps, "warmup": warm, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.dtype, "data": "synthetic",
        "config": config_for(args, n, w.io_bytes),
        "roofline": roofline, "roofline_hbm": roofline_hbm,
        "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
    }
    if quadrants is not None:
        line["quadrants"] = quadrants
    if strong is not None:
        line["weak"], line["strong"], line["gather"] = weak, strong, gather
    print(json.dumps(line), flush=True)


def measure_e2e(args, eng, torch, dist, dev, rank, world, n, B):
    """Same metric through the public API with HOST numpy buffers: one call
    `eng.<op>(q, qd, qdd, out=...)` per step.  The arrays live in pinned host memory
    (`RBDReference.pinned_empty`); the chunked H2D -> kernel -> D2H pipeline is the engine's own
    (rbdreference_b200/hostpipe.py) - nothing is staged here."""
    np_dtype = np.float64 if args.dtype == "f64" else np.float32
    itemsize = np.dtype(np_dtype).itemsize
    host = synth_host(n, B, 0xE2E + rank)
    hq, hqd, hqdd = (eng.pinned_empty(x.shape, np_dtype) for x in host)
    for dst, src in zip((hq, hqd, hqdd), host):
        np.copyto(dst, src)
    out_tail = {"rnea_grad": (n, 2 * n), "minv": (n, n), "crba": (n, n), "rnea": (n,)}[args.op]
    hout = eng.pinned_empty((B,) + out_tail, np_dtype)
    if args.op == "rnea_grad":
        one_step = lambda: eng.rnea_grad(hq, hqd, hqdd, out=hout)
    elif args.op == "minv":
        one_step = lambda: eng.minv(hq, out=hout)
    elif args.op == "crba":
        one_step = lambda: eng.crba(hq, out=hout)
    else:
        one_step = lambda: eng.rnea(hq, hqd, hqdd, outputs="c")

    steps = max(3, min(args.steps, 10))
    for _ in range(2):
        one_step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(steps):
        one_step()                                  # returns when the result is in host memory
    torch.cuda.synchronize(dev)
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    n_in = 1 if args.op in ("minv", "crba") else 3
    h2d = n_in * n * itemsize * B
    d2h = int(np.prod(out_tail)) * itemsize * B
    secs = float(t.item()) / steps
    # what the link gives a lone copy in the dominant direction on this box, measured right here (all ranks at once when
    # N > 1, as in the e2e step): the denominator of `frac_of_pcie`
    big_dir = "d2h" if d2h >= h2d else "h2d"
    probe_h = torch.empty(256 << 20, dtype=torch.uint8, pin_memory=True)
    probe_d = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    copy = (lambda: probe_h.copy_(probe_d, non_blocking=True)) if big_dir == "d2h" else (lambda: probe_d.copy_(probe_h, non_blocking=True))
    copy()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    p0 = time.perf_counter()
    for _ in range(4):
        copy()
    torch.cuda.synchronize(dev)
    tp = torch.tensor([time.perf_counter() - p0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tp, op=dist.ReduceOp.MAX)
    link = 4 * (256 << 20) / float(tp.item()) / 1e9
    return {"value": world * B / secs, "unit": "evals/s", "h2d_bytes_per_step": int(h2d),
            "d2h_bytes_per_step": int(d2h), "steps": steps, "ms_per_step": 1e3 * secs,
            "pcie_gbs": {"h2d": h2d / secs / 1e9, "d2h": d2h / secs / 1e9},
            "pcie_probe": {"direction": big_dir, "gbs_per_gpu": link,
                           "note": "plain 256 MB pinned copy, every rank at once, measured after the e2e steps"},
            "frac_of_pcie": (max(h2d, d2h) / secs / 1e9) / link,
            "api": "RBDReference.%s(numpy, ..., out=numpy) on pinned host arrays; chunked 3-stream pipeline inside the engine" % args.op}


def measure_cpu_baseline(args, n):
    """The reference on the host cores over a bounded sample sized for 10-20 s of wall time."""
    cpu = CpuReference(args.robot, args.op)
    ncal = 16 * cpu.cores
    q, qd, qdd = synth_host(n, ncal, 0xB201, args.robot)
    cpu.run(q[: cpu.cores], qd[: cpu.cores], qdd[: cpu.cores])                # warm the workers
    rate = ncal / cpu.run(q, qd, qdd)                                         # calibration: evals/s on all cores
    sample = int(max(cpu.cores, min(1 << 20, round(30.0 * rate))))
    q, qd, qdd = synth_host(n, sample, 0xB200, args.robot)
    t = cpu.run(q, qd, qdd)
    cpu.close()
    return {"value": sample / t, "unit": "evals/s", "cores": cpu.cores, "kind": cpu.kind,
            "sample": "%d knot points of the same seeded distribution, one process per core, %.1f s" % (sample, t)}


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
