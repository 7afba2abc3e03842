#!/usr/bin/env python
"""Latency of one MPC-sized step (rnea_grad + minv + forward_dynamics_grad on B knot points), eager
calls against a replayed CUDA graph of the same calls.

    python tools/small_batch_latency.py [--robot iiwa14] [--batch 1024] [--reps 200]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--robot", default="iiwa14")
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=200)
    args = ap.parse_args()
    import torch
    from rbdreference_b200 import RBDReference, robots
    eng = RBDReference(robots.by_name(args.robot))
    n, B = eng.n, args.batch
    g = torch.Generator(device="cuda").manual_seed(1)
    q, qd, u = (torch.rand(B, n, generator=g, device="cuda", dtype=torch.float64) * 2 - 1 for _ in range(3))
    dc = torch.empty(B, n, 2 * n, dtype=torch.float64, device="cuda")
    M = torch.empty(B, n, n, dtype=torch.float64, device="cuda")

    def step():
        eng.rnea_grad(q, qd, u, out=dc)
        eng.minv(q, out=M)
        return eng.forward_dynamics_grad(q, qd, u)

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.reps):
        step()
    torch.cuda.synchronize()
    eager = (time.perf_counter() - t0) / args.reps
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        step()
    for _ in range(5):
        graph.replay()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.reps):
        graph.replay()
    torch.cuda.synchronize()
    replay = (time.perf_counter() - t0) / args.reps
    print(json.dumps({"robot": args.robot, "batch": B, "step": "rnea_grad + minv + forward_dynamics_grad (7 launches)",
                      "eager_us": eager * 1e6, "graph_replay_us": replay * 1e6}))


if __name__ == "__main__":
    main()
