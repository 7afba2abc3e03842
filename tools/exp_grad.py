"""Quick GPU experiment: rnea_grad kernel variants on one robot - parity against the CPU oracle on a
small batch, then CUDA-event timing at the bench size.  python tools/exp_grad.py [robot] [log2B]"""
import os
import sys
import json

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from rbdreference_b200 import RBDReference, robots
from oracle.rbd_oracle import BatchOracle


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "iiwa14"
    lb = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    variants = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [3, 7]
    rb = robots.by_name(name)
    bo = BatchOracle(rb)
    rng = np.random.default_rng(1)
    for dtype, tol in ((torch.float64, 1e-10), (torch.float32, 1e-4)):
        eng = RBDReference(rb, dtype=dtype)
        n = eng.n
        Bs = 20000 + 13
        q, qd, qdd = rng.uniform(-np.pi, np.pi, (Bs, n)), rng.uniform(-1, 1, (Bs, n)), rng.uniform(-1, 1, (Bs, n))
        ref = bo.rnea_grad(q, qd, qdd)
        refc = bo.rnea(q, qd, qdd)[0]
        refd = bo.rnea_grad(q, qd, None, GRAVITY=-3.0, USE_VELOCITY_DAMPING=True)
        tq, tqd, tqdd = (torch.as_tensor(x, device="cuda", dtype=dtype) for x in (q, qd, qdd))
        B = 1 << lb
        g = torch.Generator(device="cuda").manual_seed(7)
        bq = (torch.rand((B, n), generator=g, device="cuda", dtype=torch.float64) * 2 - 1) * np.pi
        bqd = torch.rand((B, n), generator=g, device="cuda", dtype=torch.float64) * 2 - 1
        bqdd = torch.rand((B, n), generator=g, device="cuda", dtype=torch.float64) * 2 - 1
        bq, bqd, bqdd = bq.to(dtype), bqd.to(dtype), bqdd.to(dtype)
        out = torch.empty((B, n, 2 * n), device="cuda", dtype=dtype)
        for var in variants:
            RBDReference.set_kernel_variant(var)
            c_out = torch.empty((Bs, n), device="cuda", dtype=dtype)
            got = eng.rnea_grad(tq, tqd, tqdd, c_out=c_out).cpu().numpy()
            err = float(np.max(np.abs(got - ref)) / np.max(np.abs(ref)))
            errc = float(np.max(np.abs(c_out.cpu().numpy() - refc)) / np.max(np.abs(refc)))
            gotd = eng.rnea_grad(tq, tqd, None, GRAVITY=-3.0, USE_VELOCITY_DAMPING=True).cpu().numpy()
            errd = float(np.max(np.abs(gotd - refd)) / np.max(np.abs(refd)))
            for _ in range(3):
                eng.rnea_grad(bq, bqd, bqdd, out=out)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 20
            e0.record()
            for _ in range(reps):
                eng.rnea_grad(bq, bqd, bqdd, out=out)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            print(json.dumps({"robot": name, "dtype": str(dtype), "variant": var, "rel_err": err, "rel_err_c": errc,
                              "rel_err_damp": errd, "ok": bool(max(err, errc, errd) < tol), "B": B, "ms": ms,
                              "evals_per_s": B / ms * 1e3}), flush=True)
        RBDReference.set_kernel_variant(0)


if __name__ == "__main__":
    main()
