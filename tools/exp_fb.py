"""Quick GPU experiment: floating-base rnea_grad / minv kernel families (process-wide variant 1 = knot point per thread,
0 = automatic) - parity against the scalar CPU oracle on a few knot points, then CUDA-event timing.
python tools/exp_fb.py [robot] [log2B] [variants] [ops]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from rbdreference_b200 import RBDReference, robots
from oracle.rbd_oracle_fb import FloatingScalarOracle


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "hyq"
    lb = int(sys.argv[2]) if len(sys.argv) > 2 else 18
    variants = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [1, 0]
    ops = sys.argv[4].split(",") if len(sys.argv) > 4 else ["rnea_grad", "minv"]
    rb = robots.by_name(name + "_fb")
    so = FloatingScalarOracle(rb)
    Bs = 67
    q, qd, qdd = rb.random_state(np.random.default_rng(1), Bs)
    for dtype, tol in ((torch.float64, 1e-10), (torch.float32, 1e-4)):
        eng = RBDReference(rb, dtype=dtype)
        t = lambda x: torch.as_tensor(x, device="cuda", dtype=dtype)
        B = 1 << lb
        bq, bqd, bqdd = (t(x) for x in rb.random_state(np.random.default_rng(2), 4096))
        bq, bqd, bqdd = (x.repeat(B // 4096, 1) for x in (bq, bqd, bqdd))
        for op in ops:
            for var in variants:
                RBDReference.set_kernel_variant(var)
                err = 0.0
                for damp in (False, True):
                    if op == "rnea_grad":
                        got = eng.rnea_grad(t(q), t(qd), t(qdd), USE_VELOCITY_DAMPING=damp).cpu().numpy()
                        for k in range(0, Bs, 11):
                            ref = np.asarray(so.rnea_grad(q[k], qd[k], qdd[k], USE_VELOCITY_DAMPING=damp))
                            err = max(err, float(np.max(np.abs(got[k] - ref)) / np.max(np.abs(ref))))
                    else:
                        got = eng.minv(t(q), output_dense=damp).cpu().numpy()
                        for k in range(0, Bs, 11):
                            ref = np.asarray(so.minv(q[k], output_dense=damp))
                            err = max(err, float(np.max(np.abs(got[k] - ref)) / np.max(np.abs(ref))))
                call = (lambda: eng.rnea_grad(bq, bqd, bqdd, out=out)) if op == "rnea_grad" else (lambda: eng.minv(bq, out=out))
                nv = eng.n
                out = torch.empty((B, nv, 2 * nv) if op == "rnea_grad" else (B, nv, nv), device="cuda", dtype=dtype)
                for _ in range(3):
                    call()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                reps = 5
                e0.record()
                for _ in range(reps):
                    call()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / reps
                print(json.dumps({"robot": name + "_fb", "op": op, "dtype": str(dtype), "variant": var, "rel_err": err,
                                  "ok": bool(err < tol), "B": B, "ms": ms, "evals_per_s": B / ms * 1e3}), flush=True)
        RBDReference.set_kernel_variant(0)


if __name__ == "__main__":
    main()
