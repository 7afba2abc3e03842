#!/usr/bin/env python
"""Times the eight per-pass helper entry points on one GPU and reports the
achieved HBM bandwidth of each against MEASURED_PEAKS.json - they are HBM-bound by construction
(their (6,n,NB) / (n,6,n) tensors are inputs and outputs).

    python tools/bench_passes.py [--robot iiwa14] [--batch 262144] [--dtype f64] [--reps 10]

One JSON line per pass: algorithmic bytes per knot point (every tensor the pass must read plus
every tensor it must write, dense), ms per launch (CUDA events, median), GB/s, fraction of peak.
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
    ap.add_argument("--robot", default="iiwa14")
    ap.add_argument("--batch", type=int, default=1 << 18)
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--reps", type=int, default=10)
    args = ap.parse_args()
    import torch
    from rbdreference_b200 import RBDReference, robots

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    td = torch.float64 if args.dtype == "f64" else torch.float32
    isz = 8 if args.dtype == "f64" else 4
    rb = robots.by_name(args.robot)
    eng = RBDReference(rb, dtype=td)
    n, B = eng.n, args.batch
    NB = eng.NB                                             # bodies (= n for a fixed base, n - 5 with a floating base)
    gen = torch.Generator(device=dev).manual_seed(0xB200)
    if eng.floating_base:                                   # q carries a unit quaternion: 4096 states from the robot, tiled
        hq, hqd, hqdd = rb.random_state(np.random.default_rng(0xB200), 4096)
        rep = (B + 4095) // 4096
        q, qd, qdd = (torch.as_tensor(x, device=dev, dtype=td).repeat(rep, 1)[:B].contiguous() for x in (hq, hqd, hqdd))
    else:
        q = ((torch.rand(B, n, generator=gen, device=dev, dtype=torch.float64) * 2 - 1) * np.pi).to(td)
        qd = (torch.rand(B, n, generator=gen, device=dev, dtype=torch.float64) * 2 - 1).to(td)
        qdd = (torch.rand(B, n, generator=gen, device=dev, dtype=torch.float64) * 2 - 1).to(td)
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
        src = "MEASURED_PEAKS.json"
    except Exception:
        peak, src = 6650.0, "fallback (B200_PROFILING.md)"

    v, a, f = eng.rnea_fpass(q, qd, qdd)
    dvq, daq, dfq = eng.rnea_grad_fpass_dq(q, qd, v, a)
    dvd, dad, dfd = eng.rnea_grad_fpass_dqd(q, qd, v)
    fa = f.clone()
    eng.rnea_bpass(q, fa)
    Mb, Fb, U, D = eng.minv_bpass(q)
    v6, t6, nn, f6 = 6 * NB, 6 * n * NB, n * n, 6 * n * NB
    # (name, callable, values read + written per knot point)
    passes = [
        ("rnea_fpass", lambda: eng.rnea_fpass(q, qd, qdd), eng.nq + 2 * n + 3 * v6),
        ("rnea_bpass", lambda: eng.rnea_bpass(q, fa), n + 2 * v6 + n),
        ("rnea_grad_fpass_dq", lambda: eng.rnea_grad_fpass_dq(q, qd, v, a), 2 * n + 2 * v6 + 3 * t6),
        ("rnea_grad_fpass_dqd", lambda: eng.rnea_grad_fpass_dqd(q, qd, v), 2 * n + v6 + 3 * t6),
        ("rnea_grad_bpass_dq", lambda: eng.rnea_grad_bpass_dq(q, fa, dfq), n + v6 + 2 * t6 + nn),
        ("rnea_grad_bpass_dqd", lambda: eng.rnea_grad_bpass_dqd(q, dfd), n + 2 * t6 + nn),
        ("minv_bpass", lambda: eng.minv_bpass(q), n + nn + f6 + v6 + n),
        ("minv_fpass", lambda: eng.minv_fpass(q, Mb, Fb, U, D), n + 2 * nn + 2 * f6 + v6 + n),
    ]
    for name, fn, vals in passes:
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(args.reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts))
        gbs = vals * isz * B / (ms * 1e-3) / 1e9
        print(json.dumps({"pass": name, "robot": args.robot, "dtype": args.dtype, "batch": B, "bytes_per_knot": vals * isz,
                          "ms": ms, "gb_per_s": gbs, "hbm_peak_gb_per_s": peak, "frac": gbs / peak, "peak_source": src,
                          "note": "timed through RBDReference.<pass>() on CUDA tensors: includes output allocation"}), flush=True)


if __name__ == "__main__":
    main()
