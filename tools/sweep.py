#!/usr/bin/env python
"""Batch-size sweep of BASELINE.json configs[4]: rnea_grad (and optionally other ops) for iiwa14
and Atlas, B = 2^10 .. 2^24 knot points on ONE GPU (the multi-GPU points come from
`bench.py --gpus N`, which shards the same per-GPU batch).

    python tools/sweep.py [--ops rnea_grad,minv] [--robots iiwa14,atlas] [--dtype f64] [--max-gb 120]

One JSON line per (robot, op, B): ms per launch (median of `--reps`, CUDA events, inputs resident),
evals/s, algorithmic GB/s and TFLOP/s, and the SM clock / throttle reasons sampled (NVML, bench.ClockSampler) while
that point ran.  A 256 MB buffer is rewritten between launches whenever one launch's inputs + outputs are not
clearly larger than the 126 MB L2.  Batches whose dense inputs + outputs exceed --max-gb are
skipped (16M Atlas rnea_grad results are 242 GB - more than one B200 holds).
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ops", default="rnea_grad")
    ap.add_argument("--robots", default="iiwa14,atlas")
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--reps", type=int, default=7)
    ap.add_argument("--max-gb", type=float, default=120.0)
    ap.add_argument("--min-log2", type=int, default=10)
    ap.add_argument("--max-log2", type=int, default=24)
    args = ap.parse_args()
    import time
    import torch
    from rbdreference_b200 import RBDReference, robots
    from bench import ClockSampler

    dev = torch.device("cuda", 0)
    sampler = ClockSampler(0)
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    torch.cuda.set_device(dev)
    td = torch.float64 if args.dtype == "f64" else torch.float32
    isz = 8 if args.dtype == "f64" else 4
    for rname in args.robots.split(","):
        eng = RBDReference(robots.by_name(rname), dtype=td)
        n = eng.n
        for op in args.ops.split(","):
            flops, nbytes = eng.model.flops(op), eng.model.io_bytes(op, isz)
            for lg in range(args.min_log2, args.max_log2 + 1):
                B = 1 << lg
                if B * nbytes / 1e9 > args.max_gb:
                    print(json.dumps({"robot": rname, "op": op, "dtype": args.dtype, "batch": B,
                                      "skipped": "%.0f GB of inputs + outputs" % (B * nbytes / 1e9)}), flush=True)
                    continue
                gen = torch.Generator(device=dev).manual_seed(0xB200)
                q = ((torch.rand(B, n, generator=gen, device=dev, dtype=torch.float64) * 2 - 1) * np.pi).to(td)
                qd = (torch.rand(B, n, generator=gen, device=dev, dtype=torch.float64) * 2 - 1).to(td)
                qdd = (torch.rand(B, n, generator=gen, device=dev, dtype=torch.float64) * 2 - 1).to(td)
                if op == "rnea_grad":
                    out = torch.empty(B, n, 2 * n, dtype=td, device=dev)
                    fn = lambda: eng.rnea_grad(q, qd, qdd, out=out)
                elif op == "minv":
                    out = torch.empty(B, n, n, dtype=td, device=dev)
                    fn = lambda: eng.minv(q, out=out)
                elif op == "crba":
                    out = torch.empty(B, n, n, dtype=td, device=dev)
                    fn = lambda: eng.crba(q, out=out)
                else:
                    out = None
                    fn = lambda: eng.rnea(q, qd, qdd, outputs="c")
                for _ in range(3):
                    fn()
                torch.cuda.synchronize()
                ts = []
                flush = B * nbytes <= 2 * 126e6
                reps = args.reps if B * nbytes > 32e6 else max(args.reps, 40)      # short launches: more samples under the clock sampler
                t0 = time.perf_counter()
                for k in range(reps):
                    if flush:
                        flush_buf.fill_(k & 1)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    fn()
                    e1.record()
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1))
                t1 = time.perf_counter()
                ms = float(np.median(ts))
                print(json.dumps({"robot": rname, "op": op, "dtype": args.dtype, "batch": B, "ms": ms,
                                  "evals_per_s": B / (ms * 1e-3), "alg_gb_per_s": nbytes * B / (ms * 1e-3) / 1e9,
                                  "alg_tflop_per_s": flops * B / (ms * 1e-3) / 1e12, "l2_flushed": bool(flush),
                                  "clocks": sampler.window(t0, t1)}), flush=True)
                del q, qd, qdd, out
                torch.cuda.empty_cache()
    sampler.stop()


if __name__ == "__main__":
    main()
