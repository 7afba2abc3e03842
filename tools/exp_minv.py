"""Quick GPU experiment: minv kernel variants on one robot - parity against the CPU oracle on a small
batch, then CUDA-event timing.  python tools/exp_minv.py [robot] [log2B] [variants]"""
import os
import sys
import json

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from rbdreference_b200 import RBDReference, robots
from oracle.rbd_oracle import BatchOracle


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "atlas"
    lb = int(sys.argv[2]) if len(sys.argv) > 2 else 18
    variants = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [4, 8]
    rb = robots.by_name(name) if not name.startswith("tree") else robots.random_tree(int(name[4:]), seed=3, branching=0.5, prismatic=0.3)
    bo = BatchOracle(rb)
    rng = np.random.default_rng(1)
    n = rb.get_num_vel()
    Bs = 777
    q = rng.uniform(-np.pi, np.pi, (Bs, n))
    ref = bo.minv(q)
    for dtype, tol in ((torch.float64, 1e-10), (torch.float32, 1e-4)):
        eng = RBDReference(rb, dtype=dtype)
        tq = torch.as_tensor(q, device="cuda", dtype=dtype)
        B = 1 << lb
        g = torch.Generator(device="cuda").manual_seed(7)
        bq = ((torch.rand((B, n), generator=g, device="cuda", dtype=torch.float64) * 2 - 1) * np.pi).to(dtype)
        out = torch.empty((B, n, n), device="cuda", dtype=dtype)
        for var in variants:
            eng.set_variant(var)
            got = eng.minv(tq).cpu().numpy()
            err = float(np.max(np.abs(got - ref)) / np.max(np.abs(ref)))
            asym = float(np.max(np.abs(got - got.transpose(0, 2, 1))))
            for _ in range(3):
                eng.minv(bq, out=out)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 10
            e0.record()
            for _ in range(reps):
                eng.minv(bq, out=out)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            print(json.dumps({"robot": name, "dtype": str(dtype), "variant": var, "rel_err": err, "asym": asym,
                              "ok": bool(err < tol), "B": B, "ms": ms, "evals_per_s": B / ms * 1e3}), flush=True)


if __name__ == "__main__":
    main()
