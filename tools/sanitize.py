"""Memory-safety and race evidence without compute-sanitizer (the tool is closed on this GPU pool: its wrapper
refuses to start; see profiles/README.md).  Every kernel family runs at B = 33 and B = 1000 (partial warps, partial
tiles, tail CTAs) with

  * canaries: every output lives inside a larger allocation whose guard bands (4 KB before and after) are filled
    with a NaN pattern; a write outside the result's extent changes a guard word;
  * determinism: each call runs twice into fresh buffers; the warp-cooperative kernels exchange data through
    shared-memory transposes guarded by __syncwarp / __syncthreads - a missing barrier shows up as run-to-run
    differences (the arithmetic itself is order-fixed, so results must be bit-identical);
  * poisoned inputs beyond the batch: the inputs are carved out of larger NaN-filled allocations, so a read past
    the batch end poisons the result (checked with isfinite).

    python tools/sanitize.py            # prints one JSON line"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from rbdreference_b200 import RBDReference, robots

GUARD = 512          # elements on each side


class Arena:
    def __init__(self, dtype):
        self.dtype = dtype
        self.live = []

    def out(self, *shape):
        n = int(np.prod(shape))
        raw = torch.full((n + 2 * GUARD,), float("nan"), dtype=self.dtype, device="cuda")
        view = raw[GUARD:GUARD + n].view(*shape)
        self.live.append((raw, n))
        return view

    def inp(self, arr):
        n = arr.size
        raw = torch.full((n + 2 * GUARD,), float("nan"), dtype=self.dtype, device="cuda")
        raw[GUARD:GUARD + n] = torch.as_tensor(arr.reshape(-1), dtype=self.dtype, device="cuda")
        return raw[GUARD:GUARD + n].view(*arr.shape)

    def check(self, what):
        torch.cuda.synchronize()
        for raw, n in self.live:
            assert bool(torch.isnan(raw[:GUARD]).all()) and bool(torch.isnan(raw[GUARD + n:]).all()), "guard band written: " + what
            assert bool(torch.isfinite(raw[GUARD:GUARD + n]).all()), "non-finite result (read past the inputs?): " + what
        self.live = []


def main():
    rng = np.random.default_rng(0)
    stats = {"calls": 0, "guard_checks": 0, "determinism_checks": 0}

    def twice(fn, what):
        a = fn()
        b = fn()
        a = a if isinstance(a, (tuple, list)) else (a,)
        b = b if isinstance(b, (tuple, list)) else (b,)
        for x, y in zip(a, b):
            assert torch.equal(x, y), "run-to-run difference: " + what
        stats["calls"] += 2
        stats["determinism_checks"] += 1

    for name in ("iiwa14", "hyq", "atlas"):
        rb = robots.by_name(name)
        for dtype in (torch.float64, torch.float32):
            eng = RBDReference(rb, dtype=dtype)
            n = eng.n
            ar = Arena(dtype)
            for B in (33, 1000):
                q, qd, qdd = (ar.inp(rng.uniform(-1, 1, (B, n))) for _ in range(3))
                tag = "%s %s B=%d" % (name, dtype, B)
                for variant in (0, 1, 2, 3, 4, 5, 7, 8, 9):
                    eng.set_variant(variant)
                    twice(lambda: eng.rnea_grad(q, qd, qdd, out=ar.out(B, n, 2 * n), c_out=ar.out(B, n)), tag + " rnea_grad v%d" % variant)
                    twice(lambda: eng.minv(q, out=ar.out(B, n, n)), tag + " minv v%d" % variant)
                    twice(lambda: eng.rnea(q, qd, qdd, outputs="c"), tag + " rnea v%d" % variant)
                    ar.check(tag + " variant %d" % variant)
                    stats["guard_checks"] += 1
                eng.set_variant(-1)
                c, v, a, f = eng.rnea(q, qd, qdd)
                twice(lambda: eng.crba(q, out=ar.out(B, n, n)), tag + " crba")
                twice(lambda: eng.aba(q, qd, qdd), tag + " aba")
                twice(lambda: eng.forward_dynamics(q, qd, qdd), tag + " fd")
                twice(lambda: eng.forward_dynamics_grad(q, qd, qdd), tag + " fd_grad")
                twice(lambda: eng.end_effector_pose_gradient(q), tag + " ee_grad")
                twice(lambda: eng.rnea_fpass(q, qd, qdd), tag + " rnea_fpass")
                twice(lambda: eng.rnea_bpass(q, f.clone()), tag + " rnea_bpass")
                twice(lambda: eng.rnea_grad_fpass_dq(q, qd, v, a), tag + " fpass_dq")
                twice(lambda: eng.rnea_grad_fpass_dqd(q, qd, v), tag + " fpass_dqd")
                _, _, dfq = eng.rnea_grad_fpass_dq(q, qd, v, a)
                _, _, dfd = eng.rnea_grad_fpass_dqd(q, qd, v)
                twice(lambda: eng.rnea_grad_bpass_dq(q, f, dfq.clone()), tag + " bpass_dq")
                twice(lambda: eng.rnea_grad_bpass_dqd(q, dfd.clone(), True), tag + " bpass_dqd")
                twice(lambda: eng.minv_bpass(q), tag + " minv_bpass")
                M, F, U, D = eng.minv_bpass(q)
                twice(lambda: eng.minv_fpass(q, M.clone(), F.clone(), U, D), tag + " minv_fpass")
                ar.check(tag + " helpers")
                stats["guard_checks"] += 1
    # floating base: the cooperative kernels (family 0) and the thread-per-knot-point kernels (family 1)
    for name in ("hyq_fb", "iiwa14_fb", "atlas_fb"):
        rb = robots.by_name(name)
        for dtype in (torch.float64, torch.float32):
            eng = RBDReference(rb, dtype=dtype)
            nv = eng.n
            ar = Arena(dtype)
            for B in (33, 1000):
                q, qd, qdd = (ar.inp(x) for x in rb.random_state(rng, B))
                tag = "%s %s B=%d" % (name, dtype, B)
                for family in (0, 1):
                    RBDReference.set_kernel_variant(family)
                    twice(lambda: eng.rnea_grad(q, qd, qdd, USE_VELOCITY_DAMPING=True, out=ar.out(B, nv, 2 * nv), c_out=ar.out(B, nv)),
                          tag + " rnea_grad family %d" % family)
                    twice(lambda: eng.minv(q, out=ar.out(B, nv, nv)), tag + " minv family %d" % family)
                    twice(lambda: eng.rnea(q, qd, qdd, outputs="c"), tag + " rnea family %d" % family)
                    ar.check(tag + " family %d" % family)
                    stats["guard_checks"] += 1
                RBDReference.set_kernel_variant(0)
                if dtype == torch.float64 and B == 33:
                    twice(lambda: eng.rnea_grad_passes(q, qd, qdd), tag + " passes")
                    twice(lambda: eng.minv_passes(q), tag + " minv passes")
    torch.cuda.synchronize()
    print(json.dumps(dict(stats, ok=True)))


if __name__ == "__main__":
    main()
