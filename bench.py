#!/usr/bin/env python
"""bench.py - headline benchmark: batched rnea_grad (iiwa14, FP64, 2^20 knot points per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--robot iiwa14|hyq|atlas] [--op rnea_grad|minv|rnea] [--dtype f64|f32] [--batch B]

One "step" = one pass of the hot path over one batch of synthetic (q, qd, qdd) knot points.
Prints ONE JSON line (rank 0).  Metric = evals/s of BASELINE.json ("rnea_grad & minv evals/sec
at batch 1M"); default workload = configs[1] (iiwa14 rnea_grad FP64, 1M points on one B200).
Multi-GPU (torchrun, one rank per GPU): the batch axis is sharded, every rank owns `--batch`
knot points (weak scaling), no data-path collective; time = max over ranks.

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
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.lines = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.QUERY, "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], None, set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                clk, mx = float(parts[0]), float(parts[1])
            except ValueError:
                continue
            smax = mx
            if t0 - 0.05 <= ts <= t1 + 0.05:
                sm.append(clk)
                try:
                    power.append(float(parts[2]))
                except ValueError:
                    pass
                for nm, val in zip(names, parts[3:7]):
                    if val.lower().startswith("active"):
                        reasons.add(nm)
        if not sm:  # timed region shorter than the sampling period: use everything we saw
            for ts, line in self.lines:
                parts = [p.strip() for p in line.split(",")]
                try:
                    sm.append(float(parts[0]))
                except (ValueError, IndexError):
                    pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power) if power else None}


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


def traffic_from_profile(robot, op, dtype):
    """dram bytes per launch of the dominant kernel from the committed ncu summary, if any."""
    path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    try:
        d = json.load(open(path))
        return d.get("%s/%s/%s" % (robot, op, dtype))
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
    ap.add_argument("--batch", type=int, default=1 << 20, help="knot points per GPU")
    ap.add_argument("--variant", type=int, default=0, help="kernel family: 0 auto, 1 generic, 2 world/thread, 3 cooperative, 4 hybrid minv, 5 lane minv, 6 lane2 minv")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def workload_name(args):
    return "%s %s %s, %d knot points per GPU" % (args.robot, args.op, "FP64" if args.dtype == "f64" else "FP32", args.batch)


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
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args), "robot": args.robot, "op": args.op,
                   "note": "CPU numpy reference; each step evaluates a bounded sample of the workload"},
        "cpu_baseline": {"value": value, "unit": "evals/s", "cores": cpu.cores, "kind": cpu.kind,
                         "sample": "%d knot points per step (same seeded distribution), one process per core" % sample},
        "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    from rbdreference_b200 import RBDReference, robots

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

    tdtype = torch.float64 if args.dtype == "f64" else torch.float32
    itemsize = 8 if args.dtype == "f64" else 4
    robot = robots.by_name(args.robot)
    eng = RBDReference(robot, dtype=tdtype)
    if args.variant:
        RBDReference.set_kernel_variant(args.variant)
    n, B = eng.n, args.batch
    model = eng.model

    # synthetic inputs, resident in HBM: same fp64 draws for both precisions
    gen = torch.Generator(device=dev).manual_seed(0xB200 + rank)
    q = (torch.rand(B, eng.nq, generator=gen, device=dev, dtype=torch.float64) * 2 - 1) * np.pi
    if eng.floating_base:                          # q[0:7] = base position + unit quaternion
        if args.op not in ("rnea", "rnea_grad", "minv"):
            raise SystemExit("bench.py: floating-base robots support --op rnea | rnea_grad | minv")
        q[:, 0:3] /= np.pi
        q[:, 3:7] /= q[:, 3:7].norm(dim=1, keepdim=True)
    q = q.to(tdtype)
    qd = (torch.rand(B, n, generator=gen, device=dev, dtype=torch.float64) * 2 - 1).to(tdtype)
    qdd = (torch.rand(B, n, generator=gen, device=dev, dtype=torch.float64) * 2 - 1).to(tdtype)
    if args.op == "rnea_grad":
        out = torch.empty(B, n, 2 * n, dtype=tdtype, device=dev)
        step = lambda: eng.rnea_grad(q, qd, qdd, out=out)
    elif args.op == "minv":
        out = torch.empty(B, n, n, dtype=tdtype, device=dev)
        step = lambda: eng.minv(q, out=out)
    elif args.op == "crba":
        out = torch.empty(B, n, n, dtype=tdtype, device=dev)
        step = lambda: eng.crba(q, out=out)
    elif args.op == "fd":
        out = None
        step = lambda: eng.forward_dynamics(q, qd, qdd)          # qdd plays the torque u
    elif args.op == "fd_grad":
        out = None
        step = lambda: eng.forward_dynamics_grad(q, qd, qdd)
    elif args.op == "ee_grad":
        out = None
        step = lambda: eng.end_effector_pose_gradient(q)          # every leaf joint, default offset
    else:
        out = None
        step = lambda: eng.rnea(q, qd, qdd, outputs="c")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # measured FMA peak of this GPU (denominator of the compute roofline), before the timed region
    import ctypes
    peak = ctypes.c_double(0.0)
    ms = ctypes.c_double(0.0)
    with torch.cuda.device(dev):
        rc = eng._lib.rbd_measure_fma_peak(1 if args.dtype == "f64" else 0, ctypes.byref(peak), ctypes.byref(ms),
                                           torch.cuda.current_stream(dev).cuda_stream)
    fma_peak_tflops = peak.value / 1e12 if rc == 0 else None

    # started before the warm-up so that nvidia-smi is already streaming when the (short) timed region begins
    sampler = ClockSampler(local_rank) if rank == 0 else None
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    launches0 = eng.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    t_wall0 = time.perf_counter()
    ev[0].record()
    for k in range(args.steps):
        step()
        ev[k + 1].record()
    barrier()
    t_wall1 = time.perf_counter()
    launches = eng.launch_count() - launches0
    total_ms = ev[0].elapsed_time(ev[-1])
    per_step = [ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps)]
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None

    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    ms_per_step = total_ms_max / args.steps
    value = world * B * args.steps / (total_ms_max * 1e-3)

    # ---- e2e: public API with HOST buffers, H2D + kernel + D2H inside the timed region ----
    e2e = None
    if not args.no_e2e and args.op not in ("fd", "fd_grad", "ee_grad") and not eng.floating_base:
        e2e = measure_e2e(args, eng, torch, dist, dev, rank, world, n, B, tdtype, itemsize)

    # ---- roofline of the dominant (only) kernel: algorithmic flops / bytes per launch ----
    flops = model.flops(args.op)
    io_bytes = model.io_bytes(args.op, itemsize)
    kernel_ms = float(np.mean(per_step))           # one launch per step; events on the launch stream
    hbm_peak, hbm_src = measured_peaks()
    ach_tflops = flops * B / (kernel_ms * 1e-3) / 1e12
    ach_gbs = io_bytes * B / (kernel_ms * 1e-3) / 1e9
    traffic = traffic_from_profile(args.robot, args.op, args.dtype)
    roofline = {
        "bound": "fp64_fma" if args.dtype == "f64" else "fp32_fma",
        "achieved": ach_tflops, "peak": fma_peak_tflops, "unit": "TFLOP/s",
        "frac": (ach_tflops / fma_peak_tflops) if fma_peak_tflops else None,
        "traffic": traffic,
        "peak_source": "FMA micro-benchmark (rbd_measure_fma_peak) on this GPU just before the timed region",
        "flops_per_eval": flops, "kernel_ms": kernel_ms,
        "note": "achieved = SURVEY.md 8d algorithmic flops x evals / kernel time; tensor cores are not applicable",
    }
    roofline_hbm = {"bound": "hbm", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak,
                    "traffic": traffic, "bytes_per_eval": io_bytes, "peak_source": hbm_src}

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
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": workload_name(args), "robot": args.robot, "op": args.op, "n_dof": n,
                   "batch_per_gpu": B, "sharding": "batch axis, contiguous slices, no collective",
                   "l2": "inputs+outputs per step (%.0f MB) exceed the 126 MB L2" % ((io_bytes * B) / 1e6)},
        "roofline": roofline, "roofline_hbm": roofline_hbm,
        "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
    }
    print(json.dumps(line), flush=True)


def measure_e2e(args, eng, torch, dist, dev, rank, world, n, B, tdtype, itemsize):
    """Same metric through RBDReference.<op>() with pinned HOST inputs and a HOST result buffer.

    The batch is cut into chunks that flow through three streams' worth of work (H2D, kernel,
    D2H) so copies overlap compute; everything is inside the timed region.
    """
    nchunk = 16 if B >= (1 << 16) else 1
    bounds = np.linspace(0, B, nchunk + 1).astype(np.int64)
    np_dtype = np.float64 if itemsize == 8 else np.float32
    hq, hqd, hqdd = (torch.from_numpy(x.astype(np_dtype)).pin_memory() for x in synth_host(n, B, 0xE2E + rank))
    if args.op == "rnea_grad":
        out_tail = (n, 2 * n)
    elif args.op in ("minv", "crba"):
        out_tail = (n, n)
    else:
        out_tail = (n,)
    hout = torch.empty((B,) + out_tail, dtype=tdtype).pin_memory()
    cmax = int(np.max(np.diff(bounds)))
    nbuf = 3
    dq = [torch.empty(cmax, n, dtype=tdtype, device=dev) for _ in range(nbuf)]
    dqd = [torch.empty(cmax, n, dtype=tdtype, device=dev) for _ in range(nbuf)]
    dqdd = [torch.empty(cmax, n, dtype=tdtype, device=dev) for _ in range(nbuf)]
    dout = [torch.empty((cmax,) + out_tail, dtype=tdtype, device=dev) for _ in range(nbuf)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(nbuf)]

    def one_step():
        for ci in range(nchunk):
            lo, hi = int(bounds[ci]), int(bounds[ci + 1])
            m = hi - lo
            k = ci % nbuf
            with torch.cuda.stream(streams[k]):
                dq[k][:m].copy_(hq[lo:hi], non_blocking=True)
                if args.op not in ("minv", "crba"):
                    dqd[k][:m].copy_(hqd[lo:hi], non_blocking=True)
                    dqdd[k][:m].copy_(hqdd[lo:hi], non_blocking=True)
                if args.op == "rnea_grad":
                    eng.rnea_grad(dq[k][:m], dqd[k][:m], dqdd[k][:m], out=dout[k][:m])
                elif args.op == "minv":
                    eng.minv(dq[k][:m], out=dout[k][:m])
                elif args.op == "crba":
                    eng.crba(dq[k][:m], out=dout[k][:m])
                else:
                    dout[k][:m].copy_(eng.rnea(dq[k][:m], dqd[k][:m], dqdd[k][:m], outputs="c"))
                hout[lo:hi].copy_(dout[k][:m], non_blocking=True)
        for s in streams:
            s.synchronize()

    steps = max(3, min(args.steps, 10))
    for _ in range(2):
        one_step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(steps):
        one_step()
    torch.cuda.synchronize(dev)
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    n_in = 1 if args.op in ("minv", "crba") else 3
    h2d = n_in * n * itemsize * B
    d2h = int(np.prod(out_tail)) * itemsize * B
    return {"value": world * B * steps / float(t.item()), "unit": "evals/s", "h2d_bytes_per_step": int(h2d),
            "d2h_bytes_per_step": int(d2h), "steps": steps, "chunks": nchunk,
            "api": "RBDReference.%s on %d-way chunked pinned host buffers, 3 streams" % (args.op, nchunk)}


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
