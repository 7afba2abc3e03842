"""GPU parity tests (run with `-m gpu` on the B200 box): every C-ABI entry point, called
through the `RBDReference` host API, against the reference's golden outputs and the oracle.

Bars (BASELINE.json north_star): rel = max|x - ref| / max|ref| per tensor,
FP64 <= 1e-10, FP32 <= 1e-4.  Nothing here reads /root/reference.
"""
import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR, row_scaled_err, FB_CASES, GOLDEN_CASES, TOL_F32, TOL_F64, load_ee_golden, load_fb_golden, load_fbpass_golden, make_fb_robot, make_robot, random_states, rel_err
from oracle.rbd_oracle import BatchOracle

pytestmark = pytest.mark.gpu

requires_cuda = pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a CUDA device")


def _engine(rb, dtype=torch.float64):
    from rbdreference_b200 import RBDReference
    return RBDReference(rb, dtype=dtype)


def _t(x, dtype=torch.float64):
    return torch.as_tensor(np.ascontiguousarray(x), dtype=dtype, device="cuda")


# ---------------------------------------------------------------------------------------------
# golden vectors produced by the unmodified reference
# ---------------------------------------------------------------------------------------------
@requires_cuda
def test_fused_drivers_vs_reference_golden(golden):
    name, rb, g = golden
    eng = _engine(rb)
    q, qd, qdd = g["q"], g["qd"], g["qdd"]
    c, v, a, f = eng.rnea(q, qd, qdd)
    for got, key in ((c, "c"), (v, "v"), (a, "a"), (f, "f")):
        assert rel_err(got, g[key]) < TOL_F64, key
    c0, _, a0, f0 = eng.rnea(q, qd)
    assert rel_err(c0, g["c_noqdd"]) < TOL_F64 and rel_err(a0, g["a_noqdd"]) < TOL_F64
    assert rel_err(f0, g["f_noqdd"]) < TOL_F64
    assert rel_err(eng.rnea(q, qd, qdd, GRAVITY=-3.7)[0], g["c_galt"]) < TOL_F64
    assert rel_err(eng.rnea(q, qd, qdd, outputs="c"), g["c"]) < TOL_F64
    assert rel_err(eng.rnea_grad(q, qd, qdd), g["dc_du"]) < TOL_F64
    assert rel_err(eng.rnea_grad(q, qd, qdd, USE_VELOCITY_DAMPING=True), g["dc_du_damped"]) < TOL_F64
    assert rel_err(eng.rnea_grad(q, qd), g["dc_du_noqdd"]) < TOL_F64
    assert rel_err(eng.minv(q), g["Minv"]) < TOL_F64
    assert rel_err(eng.minv(q, output_dense=False), g["Minv_sparse"]) < TOL_F64


@requires_cuda
def test_generic_body_frame_kernels_vs_reference_golden(golden):
    """The generic (reference-recursion) fused kernels stay reachable and correct."""
    from rbdreference_b200 import RBDReference
    name, rb, g = golden
    eng = _engine(rb)
    q, qd, qdd = g["q"], g["qd"], g["qdd"]
    RBDReference.set_kernel_variant(1)
    try:
        assert rel_err(eng.rnea_grad(q, qd, qdd), g["dc_du"]) < TOL_F64
        assert rel_err(eng.rnea_grad(q, qd, qdd, USE_VELOCITY_DAMPING=True), g["dc_du_damped"]) < TOL_F64
        assert rel_err(eng.rnea_grad(q, qd), g["dc_du_noqdd"]) < TOL_F64
        assert rel_err(eng.minv(q), g["Minv"]) < TOL_F64
        assert rel_err(eng.minv(q, output_dense=False), g["Minv_sparse"]) < TOL_F64
        e32 = _engine(rb, torch.float32)
        assert rel_err(e32.rnea_grad(q, qd, qdd), g["dc_du"]) < TOL_F32
        assert rel_err(e32.minv(q), g["Minv"]) < TOL_F32
    finally:
        RBDReference.set_kernel_variant(0)


@requires_cuda
@pytest.mark.parametrize("variant", [2, 3, 4, 5, 7, 8, 9])
def test_world_kernel_variants_vs_reference_golden(golden, variant):
    """Both world-frame mappings (2: knot point per thread, 3: body per lane) against the goldens."""
    from rbdreference_b200 import RBDReference
    name, rb, g = golden
    q, qd, qdd = g["q"], g["qd"], g["qdd"]
    RBDReference.set_kernel_variant(variant)
    try:
        for dtype, tol in ((torch.float64, TOL_F64), (torch.float32, TOL_F32)):
            eng = _engine(rb, dtype)
            assert rel_err(eng.rnea_grad(q, qd, qdd), g["dc_du"]) < tol
            assert rel_err(eng.rnea_grad(q, qd, qdd, USE_VELOCITY_DAMPING=True), g["dc_du_damped"]) < tol
            assert rel_err(eng.rnea_grad(q, qd), g["dc_du_noqdd"]) < tol
            assert rel_err(eng.minv(q), g["Minv"]) < tol
        eng = _engine(rb)
        B = 333
        bo = BatchOracle(rb)
        qq, qqd, qqdd = random_states(eng.n, B, seed=variant)
        cbuf = torch.empty(B, eng.n, dtype=torch.float64, device="cuda")
        dc = eng.rnea_grad(_t(qq), _t(qqd), _t(qqdd), c_out=cbuf)
        assert rel_err(dc.cpu().numpy(), bo.rnea_grad(qq, qqd, qqdd)) < TOL_F64
        assert rel_err(cbuf.cpu().numpy(), bo.rnea(qq, qqd, qqdd)[0]) < TOL_F64
        Mref = bo.minv(qq)
        assert rel_err(eng.minv(_t(qq)).cpu().numpy(), Mref) < TOL_F64
        eng32 = _engine(rb, torch.float32)
        assert rel_err(eng32.minv(_t(qq).float()).cpu().numpy(), Mref) < TOL_F32
    finally:
        RBDReference.set_kernel_variant(0)


@requires_cuda
def test_world_kernels_selected_for_rigid_robots(golden):
    name, rb, g = golden
    assert _engine(rb).uses_world_kernels()


@requires_cuda
def test_non_rigid_inertia_falls_back_to_generic_kernels():
    """A spatial inertia without rigid-body structure is legal for the reference; the engine must
    detect it and run the body-frame kernels."""
    from rbdreference_b200 import robots

    class Odd(robots.Robot):
        def get_Imat_by_id(self, i):
            I = super().get_Imat_by_id(i)
            I[0, 4] += 0.01 * (i + 1); I[4, 0] += 0.01 * (i + 1)     # symmetric, but not [[Ibar, hx],[hx^T, m]]
            return I

        def get_Imats_dict_by_id(self):
            return {i: self.get_Imat_by_id(i) for i in range(self.get_num_bodies())}

    rb = Odd("odd", robots.hyq().joints)
    eng, bo = _engine(rb), BatchOracle(rb)
    assert not eng.uses_world_kernels()
    q, qd, qdd = random_states(12, 64, seed=2)
    assert rel_err(eng.rnea_grad(q, qd, qdd), bo.rnea_grad(q, qd, qdd)) < TOL_F64
    assert rel_err(eng.minv(q), bo.minv(q)) < TOL_F64


@requires_cuda
@pytest.mark.parametrize("variant", [0, 1, 3])
def test_pass_helpers_vs_reference_golden(golden, variant):
    """variant 0: lane / body-per-lane (gradient fpass: one ancestor distance per round) pass kernels; variant 1: the
    generic knot-point-per-thread ones; variant 3: gradient fpass with one column per lane and shared-memory tiles."""
    from rbdreference_b200 import RBDReference
    RBDReference.set_kernel_variant(variant)
    try:
        _check_pass_helpers(golden)
    finally:
        RBDReference.set_kernel_variant(0)


def _check_pass_helpers(golden):
    name, rb, g = golden
    eng = _engine(rb)
    q, qd, qdd = g["q"], g["qd"], g["qdd"]
    v, a, f = eng.rnea_fpass(q, qd, qdd)
    assert rel_err(v, g["v"]) < TOL_F64 and rel_err(a, g["a"]) < TOL_F64
    assert rel_err(f, g["f_fpass"]) < TOL_F64
    c, f_ret = eng.rnea_bpass(q, f)
    assert f_ret is f                                            # in-place contract (:619,:621)
    assert rel_err(c, g["c"]) < TOL_F64 and rel_err(f, g["f"]) < TOL_F64
    dv, da, df = eng.rnea_grad_fpass_dq(q, qd, g["v"], g["a"])
    for got, key in ((dv, "dv_dq"), (da, "da_dq"), (df, "df_dq")):
        assert rel_err(got, g[key]) < TOL_F64, key
    dv2, da2, df2 = eng.rnea_grad_fpass_dqd(q, qd, g["v"])
    for got, key in ((dv2, "dv_dqd"), (da2, "da_dqd"), (df2, "df_dqd")):
        assert rel_err(got, g[key]) < TOL_F64, key
    dc_dq = eng.rnea_grad_bpass_dq(q, g["f"], df)
    assert rel_err(dc_dq, g["dc_dq"]) < TOL_F64
    assert rel_err(df, g["df_dq_acc"]) < TOL_F64                 # df_dq mutated in place (:1291)
    dc_dqd = eng.rnea_grad_bpass_dqd(q, df2)
    assert rel_err(dc_dqd, g["dc_dqd"]) < TOL_F64
    assert rel_err(df2, g["df_dqd_acc"]) < TOL_F64
    Mb, Fb, U, D = eng.minv_bpass(q)
    for got, key in ((Mb, "Minv_b"), (Fb, "F_b"), (U, "U"), (D, "Dinv")):
        assert rel_err(got, g[key]) < TOL_F64, key
    M = eng.minv_fpass(q, Mb, Fb, U, D)
    assert M is Mb                                               # returned object is the argument (:783)
    assert rel_err(M, g["Minv_sparse"]) < TOL_F64
    assert rel_err(Fb, g["F_f"]) < TOL_F64
    assert rel_err(eng.rnea_grad_passes(q, qd, qdd), g["dc_du"]) < TOL_F64
    assert rel_err(eng.minv_passes(q), g["Minv"]) < TOL_F64


@requires_cuda
def test_single_state_reference_shapes(golden):
    """(n,) numpy in -> reference-shaped numpy out, one knot point (README.md:9-12)."""
    name, rb, g = golden
    eng = _engine(rb)
    n = eng.n
    q, qd, qdd = g["q"][0], g["qd"][0], g["qdd"][0]
    c, v, a, f = eng.rnea(q, qd, qdd)
    assert c.shape == (n,) and v.shape == (6, n) and a.shape == (6, n) and f.shape == (6, n)
    assert isinstance(c, np.ndarray) and rel_err(c, g["c"][0]) < TOL_F64
    dc = eng.rnea_grad(q, qd, qdd)
    assert dc.shape == (n, 2 * n) and rel_err(dc, g["dc_du"][0]) < TOL_F64
    M = eng.minv(q)
    assert M.shape == (n, n) and rel_err(M, g["Minv"][0]) < TOL_F64
    f_in = g["f_fpass"][0].copy()
    c2, f_out = eng.rnea_bpass(q, f_in)
    assert f_out is f_in and rel_err(f_in, g["f"][0]) < TOL_F64


@requires_cuda
def test_fp32_vs_reference_golden(golden):
    name, rb, g = golden
    eng = _engine(rb, torch.float32)
    q, qd, qdd = g["q"], g["qd"], g["qdd"]
    c, v, a, f = eng.rnea(q, qd, qdd)
    assert c.dtype == np.float32
    for got, key in ((c, "c"), (v, "v"), (a, "a"), (f, "f")):
        assert rel_err(got, g[key]) < TOL_F32, key
    assert rel_err(eng.rnea_grad(q, qd, qdd), g["dc_du"]) < TOL_F32
    assert rel_err(eng.minv(q), g["Minv"]) < TOL_F32
    assert rel_err(eng.rnea_grad_passes(q, qd, qdd), g["dc_du"]) < TOL_F32
    assert rel_err(eng.minv_passes(q), g["Minv"]) < TOL_F32


# ---------------------------------------------------------------------------------------------
# seeded batches against the oracle
# ---------------------------------------------------------------------------------------------
@requires_cuda
@pytest.mark.parametrize("name,B", [("iiwa14", 4096), ("hyq", 2048), ("atlas", 1024), ("tree13", 777)])
def test_batched_cuda_tensors_vs_oracle(name, B):
    rb = make_robot(name)
    eng, bo = _engine(rb), BatchOracle(rb)
    n = eng.n
    q, qd, qdd = random_states(n, B, seed=0xB200)
    tq, tqd, tqdd = _t(q), _t(qd), _t(qdd)
    c, v, a, f = eng.rnea(tq, tqd, tqdd)
    assert c.is_cuda and c.shape == (B, n) and v.shape == (B, 6, n)
    rc, rv, ra, rf = bo.rnea(q, qd, qdd)
    for got, ref in ((c, rc), (v, rv), (a, ra), (f, rf)):
        assert rel_err(got.cpu().numpy(), ref) < TOL_F64
    dc = eng.rnea_grad(tq, tqd, tqdd)
    assert dc.shape == (B, n, 2 * n)
    assert rel_err(dc.cpu().numpy(), bo.rnea_grad(q, qd, qdd)) < TOL_F64
    M = eng.minv(tq)
    assert rel_err(M.cpu().numpy(), bo.minv(q)) < TOL_F64
    out = torch.empty(B, n, 2 * n, dtype=torch.float64, device="cuda")
    cbuf = torch.empty(B, n, dtype=torch.float64, device="cuda")
    ret = eng.rnea_grad(tq, tqd, tqdd, out=out, c_out=cbuf)
    assert ret.data_ptr() == out.data_ptr() and torch.equal(out, dc)
    assert rel_err(cbuf.cpu().numpy(), rc) < TOL_F64


@requires_cuda
@pytest.mark.parametrize("name,B", [("iiwa14", 1000), ("hyq", 777), ("atlas", 333), ("tree13", 257), ("tree9", 65)])
def test_rnea_kernel_paths_vs_oracle(name, B):
    """rnea: lane kernel (shared-memory / local-memory f rows, c only / all outputs, FP64 / FP32)
    and the generic body-frame kernel (variant 1) against the oracle."""
    from rbdreference_b200 import RBDReference
    rb = make_robot(name)
    bo = BatchOracle(rb)
    q, qd, qdd = random_states(rb.get_num_vel(), B, seed=7 + B)
    rc, rv, ra, rf = bo.rnea(q, qd, qdd)
    rc0 = bo.rnea(q, qd)[0]
    for variant in (0, 1, 2):      # 0: cooperative c-only kernel (small batch) / lane kernel; 1: generic; 2: lane kernel
        RBDReference.set_kernel_variant(variant)
        try:
            for dtype, tol in ((torch.float64, TOL_F64), (torch.float32, TOL_F32)):
                eng = _engine(rb, dtype)
                tq, tqd, tqdd = _t(q, dtype), _t(qd, dtype), _t(qdd, dtype)
                assert rel_err(eng.rnea(tq, tqd, tqdd, outputs="c").cpu().numpy(), rc) < tol
                assert rel_err(eng.rnea(tq, tqd, outputs="c").cpu().numpy(), rc0) < tol
                c, v, a, f = eng.rnea(tq, tqd, tqdd)
                for got, ref in ((c, rc), (v, rv), (a, ra), (f, rf)):
                    assert rel_err(got.cpu().numpy(), ref) < tol
                # the two passes on their own (lane kernel modes 1 / 2, or the generic pass kernels)
                v1, a1, f1 = eng.rnea_fpass(tq, tqd, tqdd)
                assert rel_err(v1.cpu().numpy(), rv) < tol and rel_err(a1.cpu().numpy(), ra) < tol
                c2, f2 = eng.rnea_bpass(tq, f1)
                assert f2.data_ptr() == f1.data_ptr()                       # in place (:619)
                assert rel_err(c2.cpu().numpy(), rc) < tol and rel_err(f2.cpu().numpy(), rf) < tol
        finally:
            RBDReference.set_kernel_variant(0)


@requires_cuda
@pytest.mark.parametrize("name,B", [("iiwa14", 131), ("hyq", 67), ("atlas", 9), ("tree13", 33)])
def test_gradient_passes_batched_vs_oracle(name, B):
    """The four gradient passes on ragged batches (partial warp groups), FP64 and FP32."""
    rb = make_robot(name)
    bo = BatchOracle(rb)
    n = rb.get_num_vel()
    q, qd, qdd = random_states(n, B, seed=3 * B)
    c, v, a, f = bo.rnea(q, qd, qdd)
    ref = bo.rnea_grad(q, qd, qdd)
    for dtype, tol in ((torch.float64, TOL_F64), (torch.float32, TOL_F32)):
        eng = _engine(rb, dtype)
        tq, tqd = _t(q, dtype), _t(qd, dtype)
        dv, da, df = eng.rnea_grad_fpass_dq(tq, tqd, _t(v, dtype), _t(a, dtype))
        dv2, da2, df2 = eng.rnea_grad_fpass_dqd(tq, tqd, _t(v, dtype))
        dc_dq = eng.rnea_grad_bpass_dq(tq, _t(f, dtype), df)
        dc_dqd = eng.rnea_grad_bpass_dqd(tq, df2)
        got = torch.cat((dc_dq, dc_dqd), dim=2).cpu().numpy()
        assert rel_err(got, ref) < tol


# ---------------------------------------------------------------------------------------------
# concurrency: one immutable handle, many host threads and streams
# ---------------------------------------------------------------------------------------------
@requires_cuda
@pytest.mark.parametrize("name", ["iiwa14", "atlas"])
def test_concurrent_threads_and_streams_are_deterministic(name):
    """Four host threads share one engine, each on its own stream (scratch buffers and temporaries
    are stream-ordered pool allocations): results must equal the single-threaded ones bit for bit."""
    import threading
    rb = make_robot(name)
    eng = _engine(rb)
    n = eng.n
    B = 2048 if name == "iiwa14" else 512
    data = []
    for t in range(4):
        q, qd, qdd = random_states(n, B, seed=100 + t)
        tq, tqd, tqdd = _t(q), _t(qd), _t(qdd)
        ref = (eng.rnea_grad(tq, tqd, tqdd).clone(), eng.minv(tq).clone(),
               torch.cat(eng.forward_dynamics_grad(tq, tqd, tqdd), dim=2).clone())
        data.append((tq, tqd, tqdd, ref))
    torch.cuda.synchronize()
    errors = []

    def work(t):
        try:
            tq, tqd, tqdd, ref = data[t]
            stream = torch.cuda.Stream()
            with torch.cuda.stream(stream):
                for _ in range(10):
                    got = (eng.rnea_grad(tq, tqd, tqdd), eng.minv(tq), torch.cat(eng.forward_dynamics_grad(tq, tqd, tqdd), dim=2))
                    stream.synchronize()
                    for g, r in zip(got, ref):
                        if not torch.equal(g, r):
                            errors.append("thread %d: result differs from the single-threaded run" % t)
                            return
        except Exception as exc:  # pragma: no cover
            errors.append("thread %d: %r" % (t, exc))

    threads = [threading.Thread(target=work, args=(t,)) for t in range(4)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors


@requires_cuda
@pytest.mark.parametrize("name", ["iiwa14", "atlas"])
def test_calls_are_cuda_graph_capturable(name):
    """Small-batch MPC loops are launch-bound: the fused drivers only enqueue stream-ordered work
    (kernels, pool allocations), so a whole step can be captured in a CUDA graph and replayed."""
    rb = make_robot(name)
    eng = _engine(rb)
    n, B = eng.n, 256
    sq, sqd, sqdd = (torch.zeros(B, n, dtype=torch.float64, device="cuda") for _ in range(3))
    out_dc = torch.empty(B, n, 2 * n, dtype=torch.float64, device="cuda")
    out_M = torch.empty(B, n, n, dtype=torch.float64, device="cuda")
    q, qd, qdd = random_states(n, B, seed=5)
    sq.copy_(_t(q)); sqd.copy_(_t(qd)); sqdd.copy_(_t(qdd))
    # warm up outside the capture (the library creates its scratch memory pool on first use)
    eng.rnea_grad(sq, sqd, sqdd, out=out_dc); eng.minv(sq, out=out_M); eng.forward_dynamics_grad(sq, sqd, sqdd)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        eng.rnea_grad(sq, sqd, sqdd, out=out_dc)
        eng.minv(sq, out=out_M)
        fd1, fd2 = eng.forward_dynamics_grad(sq, sqd, sqdd)
    for seed in (6, 7):
        q, qd, qdd = random_states(n, B, seed=seed)
        sq.copy_(_t(q)); sqd.copy_(_t(qd)); sqdd.copy_(_t(qdd))
        graph.replay()
        torch.cuda.synchronize()
        got = (out_dc.clone(), out_M.clone(), fd1.clone(), fd2.clone())
        e1, e2 = eng.forward_dynamics_grad(sq, sqd, sqdd)
        ref = (eng.rnea_grad(sq, sqd, sqdd), eng.minv(sq), e1, e2)
        for g, r in zip(got, ref):
            assert torch.equal(g, r)


# ---------------------------------------------------------------------------------------------
# sizes at the edges of the kernels' design space
# ---------------------------------------------------------------------------------------------
def _edge_robot(kind):
    from rbdreference_b200 import robots
    if kind == "one":
        return robots.random_tree(1, seed=3)
    if kind == "two":
        return robots.random_tree(2, seed=4, branching=0.0)
    if kind == "chain32":           # RBD_MAX_DOF bodies in one chain: depth 31, five pointer-jumping rounds
        return robots.random_tree(32, seed=5, branching=0.0, prismatic=0.1)
    if kind == "bush32":            # RBD_MAX_DOF bodies, many roots and branch points
        return robots.random_tree(32, seed=6, branching=0.6, prismatic=0.2)
    if kind == "stars":             # every body hangs off the base: n root components of one body
        return _stars(9)
    raise KeyError(kind)


def _stars(n):
    from rbdreference_b200 import robots
    rb = robots.random_tree(n, seed=8, branching=0.0)
    for j in rb.joints:
        j.parent = -1
    return robots.Robot("stars%d" % n, rb.joints)


@requires_cuda
@pytest.mark.parametrize("kind", ["one", "two", "chain32", "bush32", "stars"])
def test_edge_topologies_all_drivers_vs_oracle(kind):
    """n = 1, 2 and RBD_MAX_DOF, the deepest and the flattest trees: every fused driver, every kernel
    variant that serves the size, both precisions, ragged batch."""
    from rbdreference_b200 import RBDReference
    rb = _edge_robot(kind)
    bo = BatchOracle(rb)
    n = rb.get_num_vel()
    B = 77
    q, qd, qdd = random_states(n, B, seed=31)
    u = np.random.default_rng(9).uniform(-5, 5, (B, n))
    rc, rv, ra, rf = bo.rnea(q, qd, qdd)
    rdc, rM, rH, raba = bo.rnea_grad(q, qd, qdd), bo.minv(q), bo.crba(q), bo.aba(q, qd, u)
    c0 = bo.rnea(q, qd)[0]
    rqdd = np.einsum("bij,bj->bi", rM, u - c0)
    # rnea / rnea_grad / crba keep the plain bar on every robot.  Minv (and what is built on it) of a deep
    # chain of random links is ill-conditioned: its bar scales with cond(M) taken from the oracle.
    cond = float(np.max(np.linalg.cond(rH)))
    scale_m = max(1.0, cond * 2.3e-16 / TOL_F64 * 50.0)
    scale = 1.0
    for variant in (0, 1, 2, 3, 4, 5, 7, 8, 9):
        RBDReference.set_kernel_variant(variant)
        try:
            eng = _engine(rb)
            tq, tqd, tqdd, tu = _t(q), _t(qd), _t(qdd), _t(u)
            c, v, a, f = eng.rnea(tq, tqd, tqdd)
            for got, ref in ((c, rc), (v, rv), (a, ra), (f, rf)):
                assert rel_err(got.cpu().numpy(), ref) < TOL_F64 * scale, (variant, "rnea")
            assert rel_err(eng.rnea_grad(tq, tqd, tqdd).cpu().numpy(), rdc) < TOL_F64 * scale, (variant, "rnea_grad")
            assert rel_err(eng.minv(tq).cpu().numpy(), rM) < TOL_F64 * scale_m, (variant, "minv")
            assert rel_err(eng.crba(tq).cpu().numpy(), rH) < TOL_F64 * scale, (variant, "crba")
            if variant in (0, 1):
                assert rel_err(eng.aba(tq, tqd, tu).cpu().numpy(), raba) < TOL_F64 * scale_m * 10, (variant, "aba")
                assert rel_err(eng.forward_dynamics(tq, tqd, tu).cpu().numpy(), rqdd) < TOL_F64 * scale_m * 10, (variant, "fd")
                d1, d2 = eng.forward_dynamics_grad(tq, tqd, tu)
                dc = bo.rnea_grad(q, qd, rqdd)
                # one bar for [qdd_dq | qdd_dqd]: dc_dqd is exactly zero for some of these robots (a single
                # body on a fixed axis has no velocity-dependent torque), so its own scale would be rounding noise
                got = torch.cat((d1, d2), dim=2).cpu().numpy()
                assert rel_err(got, -np.einsum("bij,bjk->bik", rM, dc)) < TOL_F64 * scale_m * 10, (variant, "fd_grad")
                e32 = _engine(rb, torch.float32)
                f32 = torch.float32
                # FP32: n <= 16 at 10x the bar (prismatic joints far from the base), n = 32 at 30x for rnea / rnea_grad
                # (world-frame quantities of a 32-link chain reach |p| ~ 10 m: m |p|^2 terms cost ~1.5 digits);
                # FP32 Minv of the n = 32 robots is ill-conditioned beyond single precision and is not asserted
                f32bar = TOL_F32 * (10 if n <= 16 else 30)
                assert rel_err(e32.rnea_grad(_t(q, f32), _t(qd, f32), _t(qdd, f32)).cpu().numpy(), rdc) < f32bar, (variant, "rnea_grad f32")
                assert rel_err(e32.rnea(_t(q, f32), _t(qd, f32), _t(qdd, f32), outputs="c").cpu().numpy(), rc) < f32bar, (variant, "rnea f32")
                if n <= 16:
                    assert rel_err(e32.minv(_t(q, f32)).cpu().numpy(), rM) < TOL_F32 * 10
        finally:
            RBDReference.set_kernel_variant(0)
    # the pass helpers (column-per-lane tiles at G = 8 / 32) on the same robots
    eng = _engine(rb)
    dv, da, df = eng.rnea_grad_fpass_dq(tq, tqd, _t(rv), _t(ra))
    dv2, da2, df2 = eng.rnea_grad_fpass_dqd(tq, tqd, _t(rv))
    got = torch.cat((eng.rnea_grad_bpass_dq(tq, _t(rf), df), eng.rnea_grad_bpass_dqd(tq, df2)), dim=2)
    assert rel_err(got.cpu().numpy(), rdc) < TOL_F64 * scale


@requires_cuda
def test_aba_vs_reference_golden_and_oracle(golden):
    """aba (RBDReference.py:817): goldens from the live reference (default and alternate gravity),
    batched oracle, FP32."""
    name, rb, g = golden
    bo = BatchOracle(rb)
    n = rb.get_num_vel()
    eng = _engine(rb)
    assert rel_err(eng.aba(g["q"], g["qd"], g["u"]), g["aba_qdd"]) < TOL_F64
    assert rel_err(eng.aba(g["q"], g["qd"], g["u"], GRAVITY=-3.7), g["aba_qdd_galt"]) < TOL_F64
    one = eng.aba(g["q"][0], g["qd"][0], g["u"][0])
    assert one.shape == (n,) and rel_err(one, g["aba_qdd"][0]) < TOL_F64
    B = 513
    q, qd, _ = random_states(n, B, seed=21)
    tau = np.random.default_rng(5).uniform(-10, 10, (B, n))
    ref = bo.aba(q, qd, tau)
    got = eng.aba(_t(q), _t(qd), _t(tau))
    assert got.shape == (B, n) and rel_err(got.cpu().numpy(), ref) < TOL_F64
    e32 = _engine(rb, torch.float32)
    got32 = e32.aba(_t(q, torch.float32), _t(qd, torch.float32), _t(tau, torch.float32))
    assert rel_err(got32.cpu().numpy(), ref) < 20 * TOL_F32      # qdd = (u - U.a)/d amplifies FP32 rounding on light distal links


@requires_cuda
def test_crba_vs_reference_golden_and_oracle(golden):
    """crba (RBDReference.py:1090-1124): goldens from the live reference, batched oracle, both kernel
    families, both precisions, and H @ Minv = I."""
    from rbdreference_b200 import RBDReference
    name, rb, g = golden
    bo = BatchOracle(rb)
    n = rb.get_num_vel()
    B = 301
    q, _, _ = random_states(n, B, seed=11)
    Href = bo.crba(q)
    for variant in (0, 1):
        RBDReference.set_kernel_variant(variant)
        try:
            eng = _engine(rb)
            assert rel_err(eng.crba(g["q"]), g["H"]) < TOL_F64
            H1 = eng.crba(g["q"][0])
            assert H1.shape == (n, n) and rel_err(H1, g["H"][0]) < TOL_F64
            H = eng.crba(_t(q))
            assert H.shape == (B, n, n) and rel_err(H.cpu().numpy(), Href) < TOL_F64
            assert torch.equal(H, H.transpose(1, 2))
            e32 = _engine(rb, torch.float32)
            assert rel_err(e32.crba(_t(q, torch.float32)).cpu().numpy(), Href) < TOL_F32
            if variant == 0:
                eye = torch.bmm(H, eng.minv(_t(q)))
                assert float((eye - torch.eye(n, dtype=torch.float64, device="cuda")).abs().max()) < 1e-8
        finally:
            RBDReference.set_kernel_variant(0)


@requires_cuda
@pytest.mark.parametrize("B", [1, 2, 31, 32, 33, 127, 129, 1000])
def test_ragged_batch_sizes(B):
    rb = make_robot("hyq")
    eng, bo = _engine(rb), BatchOracle(rb)
    q, qd, qdd = random_states(eng.n, B, seed=B)
    assert rel_err(eng.rnea_grad(q, qd, qdd), bo.rnea_grad(q, qd, qdd)) < TOL_F64
    assert rel_err(eng.minv(q), bo.minv(q)) < TOL_F64
    assert rel_err(eng.rnea(q, qd, qdd)[0], bo.rnea(q, qd, qdd)[0]) < TOL_F64


@requires_cuda
def test_empty_batch_and_bad_shapes():
    eng = _engine(make_robot("iiwa14"))
    z = torch.empty(0, 7, dtype=torch.float64, device="cuda")
    assert eng.rnea_grad(z, z, z).shape == (0, 7, 14)
    assert eng.minv(z).shape == (0, 7, 7)
    with pytest.raises(ValueError):
        eng.rnea(np.zeros(6), np.zeros(6))
    with pytest.raises(ValueError):
        eng.rnea_grad(np.zeros((3, 7)), np.zeros((4, 7)))


@requires_cuda
def test_fp32_batch_vs_fp64_oracle():
    """FP32 kernels fed the FP64 draws cast down stay within 1e-4 on Atlas."""
    rb = make_robot("atlas")
    eng, bo = _engine(rb, torch.float32), BatchOracle(rb)
    q, qd, qdd = random_states(30, 512, seed=5)
    assert rel_err(eng.minv(q), bo.minv(q)) < TOL_F32
    assert rel_err(eng.rnea_grad(q, qd, qdd), bo.rnea_grad(q, qd, qdd)) < TOL_F32


# ---------------------------------------------------------------------------------------------
# full-size, size-independent properties (no oracle at 2^20 points)
# ---------------------------------------------------------------------------------------------
@requires_cuda
def test_full_size_properties_iiwa_1m():
    rb = make_robot("iiwa14")
    eng, bo = _engine(rb), BatchOracle(rb)
    n, B = 7, 1 << 20
    gen = torch.Generator(device="cuda").manual_seed(0xB200)
    q = (torch.rand(B, n, generator=gen, device="cuda", dtype=torch.float64) * 2 - 1) * np.pi
    qd = torch.rand(B, n, generator=gen, device="cuda", dtype=torch.float64) * 2 - 1
    qdd = torch.rand(B, n, generator=gen, device="cuda", dtype=torch.float64) * 2 - 1
    dc = eng.rnea_grad(q, qd, qdd)
    M = eng.minv(q)
    assert torch.isfinite(dc).all() and torch.isfinite(M).all()
    assert torch.equal(M, M.transpose(1, 2))                      # mirror is exact (:799-804)
    # linearity in qdd: c(q,qd,qdd) - c(q,qd,0) = H qdd  =>  Minv (c - c0) = qdd
    c = eng.rnea(q, qd, qdd, outputs="c")
    c0 = eng.rnea(q, qd, outputs="c")
    back = torch.matmul(M, (c - c0).unsqueeze(-1)).squeeze(-1)
    assert float((back - qdd).abs().max()) < 1e-9
    # gravity-free, motion-free torque is exactly zero
    zero = torch.zeros_like(q)
    assert float(eng.rnea(q, zero, zero, GRAVITY=0.0, outputs="c").abs().max()) == 0.0
    # dc/dqd does not depend on qdd; damping only touches its diagonal
    dc2 = eng.rnea_grad(q, qd, None, USE_VELOCITY_DAMPING=True)
    diff = dc2[:, :, n:] - dc[:, :, n:]
    damp = torch.as_tensor(eng.model.damping, device="cuda")
    assert float((diff - torch.diag(damp)).abs().max()) < 1e-12
    # sharded evaluation (disjoint slices, same kernel) is bit-identical to the unsharded one
    from rbdreference_b200.dist import shard_bounds
    lo, hi = shard_bounds(B, 3, 8)
    assert torch.equal(eng.rnea_grad(q[lo:hi], qd[lo:hi], qdd[lo:hi]), dc[lo:hi])
    # spot-check 4096 strided points of the big batch against the oracle
    idx = torch.arange(0, B, B // 4096, device="cuda")[:4096]
    ref = bo.rnea_grad(q[idx].cpu().numpy(), qd[idx].cpu().numpy(), qdd[idx].cpu().numpy())
    assert rel_err(dc[idx].cpu().numpy(), ref) < TOL_F64
    assert rel_err(M[idx].cpu().numpy(), bo.minv(q[idx].cpu().numpy())) < TOL_F64


@requires_cuda
def test_full_size_properties_atlas_256k():
    rb = make_robot("atlas")
    n, B = 30, 1 << 18
    bo = BatchOracle(rb)
    gen = torch.Generator(device="cuda").manual_seed(0xA71A5)
    q64 = (torch.rand(B, n, generator=gen, device="cuda", dtype=torch.float64) * 2 - 1) * np.pi
    for dtype, tol in ((torch.float64, TOL_F64), (torch.float32, TOL_F32)):
        eng = _engine(rb, dtype)
        M = eng.minv(q64.to(dtype))
        assert torch.isfinite(M).all()
        # tile kernel: both triangles come from the forward recursion itself (:771 updates whole rows; it is what
        # output_dense=False returns) instead of the copy of :799-804 - symmetric to rounding
        asym = float((M - M.transpose(1, 2)).abs().max() / M.abs().max())
        assert asym < (1e-13 if dtype == torch.float64 else 1e-5)
        # block structure: pelvis-rooted components (torso+arms | l_leg | r_leg) do not couple
        assert float(M[:, :18, 18:].abs().max()) == 0.0 and float(M[:, 18:24, 24:].abs().max()) == 0.0
        idx = torch.arange(0, B, B // 512, device="cuda")[:512]
        H = bo.crba(q64[idx].cpu().numpy())
        MH = M[idx].double().cpu().numpy() @ H
        assert np.max(np.abs(MH - np.eye(n))) < (1e-9 if dtype == torch.float64 else 2e-2)
        assert rel_err(M[idx].cpu().numpy(), bo.minv(q64[idx].cpu().numpy())) < tol


@requires_cuda
def test_forward_dynamics_vs_reference_golden(golden):
    """rbd_forward_dynamics / rbd_forward_dynamics_grad (rnea + minv + rnea_grad plus the
    hand-written per-knot-point product kernel) against outputs of the unmodified reference."""
    name, rb, g = golden
    q, qd, u = g["q"], g["qd"], g["u"]
    # Bars: the north-star bars apply to c, Minv and dc_du; what they allow in those factors is propagated
    # through the products  Minv (u - c)  and  -Minv dc_du  (row sums of |Minv| reach 1e3-1e4 for light links).
    Minv, n = g["Minv"], rb.get_num_vel()
    rows = np.max(np.sum(np.abs(Minv), axis=-1))
    for dtype, tol in ((torch.float64, TOL_F64), (torch.float32, TOL_F32)):
        eng = _engine(rb, dtype)
        ref_qdd = g["fd_qdd"]
        bar_qdd = tol * (rows * np.max(np.abs(g["c_noqdd"])) + np.max(np.abs(Minv)) * n * np.max(np.abs(u))) / np.max(np.abs(ref_qdd))
        assert rel_err(eng.forward_dynamics(q, qd, u), ref_qdd) < max(bar_qdd, tol)
        d1, d2 = eng.forward_dynamics_grad(q, qd, u)
        for got, key in ((d1, "fd_dq"), (d2, "fd_dqd")):
            scale = np.max(np.abs(g["dc_du"]))          # magnitude of the rnea_grad factor
            bar = tol * 2 * rows * scale / np.max(np.abs(g[key]))
            assert rel_err(got, g[key]) < max(bar, tol), (key, rel_err(got, g[key]), bar)
    eng = _engine(rb)
    one = eng.forward_dynamics(q[0], qd[0], u[0])               # single knot point, reference shapes
    assert one.shape == (eng.n,) and rel_err(one, g["fd_qdd"][0]) < 100 * TOL_F64
    e1, e2 = eng.forward_dynamics_grad(q[0], qd[0], u[0])
    assert e1.shape == (eng.n, eng.n) and rel_err(e2, g["fd_dqd"][0]) < 100 * TOL_F64


@requires_cuda
@pytest.mark.parametrize("name,B", [("iiwa14", 1000), ("hyq", 333), ("atlas", 97)])
def test_forward_dynamics_batches_vs_oracle(name, B):
    """Ragged batches (not multiples of the kernels' group sizes) against the CPU oracle."""
    rb = make_robot(name)
    eng, bo = _engine(rb), BatchOracle(rb)
    n = eng.n
    q, qd, _ = random_states(n, B, seed=77)
    u = np.random.default_rng(78).uniform(-10, 10, (B, n))
    Minv = bo.minv(q)
    c = bo.rnea(q, qd)[0]
    qdd_ref = np.einsum("bij,bj->bi", Minv, u - c)
    qdd = eng.forward_dynamics(_t(q), _t(qd), _t(u)).cpu().numpy()
    assert rel_err(qdd, qdd_ref) < 10 * TOL_F64
    dc = bo.rnea_grad(q, qd, qdd_ref)
    from rbdreference_b200 import RBDReference
    for variant in (0, 1):      # 0: FP64 tensor-core product (mma.sync m8n8k4); 1: register-tiled product + generic drivers
        RBDReference.set_kernel_variant(variant)
        try:
            d1, d2 = eng.forward_dynamics_grad(_t(q), _t(qd), _t(u))
            assert rel_err(d1.cpu().numpy(), -np.einsum("bij,bjk->bik", Minv, dc[:, :, :n])) < 10 * TOL_F64
            assert rel_err(d2.cpu().numpy(), -np.einsum("bij,bjk->bik", Minv, dc[:, :, n:])) < 10 * TOL_F64
        finally:
            RBDReference.set_kernel_variant(0)
    e32 = _engine(rb, torch.float32)                       # FP32: register-tiled product
    f1, f2 = e32.forward_dynamics_grad(_t(q, torch.float32), _t(qd, torch.float32), _t(u, torch.float32))
    assert rel_err(f1.cpu().numpy(), -np.einsum("bij,bjk->bik", Minv, dc[:, :, :n])) < 50 * TOL_F32
    assert rel_err(f2.cpu().numpy(), -np.einsum("bij,bjk->bik", Minv, dc[:, :, n:])) < 50 * TOL_F32
    assert eng.forward_dynamics(_t(q[:0]), _t(qd[:0]), _t(u[:0])).shape == (0, n)     # empty batch


@requires_cuda
def test_forward_dynamics_compositions():
    """forward_dynamics / forward_dynamics_grad as compositions."""
    from oracle.rbd_oracle import ScalarOracle
    rb = make_robot("iiwa14")
    eng, so = _engine(rb), ScalarOracle(rb)
    q, qd, u = random_states(7, 4, seed=21)
    qdd = eng.forward_dynamics(q, qd, u)
    dq, dqd = eng.forward_dynamics_grad(q, qd, u)
    for k in range(4):
        assert rel_err(qdd[k], so.forward_dynamics(q[k], qd[k], u[k])) < 1e-9
        r1, r2 = so.forward_dynamics_grad(q[k], qd[k], u[k])
        assert rel_err(dq[k], r1) < 1e-9 and rel_err(dqd[k], r2) < 1e-9


# ---------------------------------------------------------------------------------------------
# end-effector kinematics: RBDReference.py:220-386
# ---------------------------------------------------------------------------------------------
@requires_cuda
@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_end_effector_pose_and_gradient_vs_reference_golden(name):
    """Default leaves, named moving + fixed joints, zero / non-zero offset; single knot point
    (reference shapes: lists of (6,1) / (6,n)) and batched; FP64 and FP32."""
    rb = make_robot(name)
    eng, e32 = _engine(rb), _engine(rb, torch.float32)
    q, cases = load_ee_golden(name)
    n = eng.n
    for names, off, pose, grad in cases:
        P = eng.end_effector_pose(q[0], names, off)
        G = eng.end_effector_pose_gradient(q[0], names, off)
        assert isinstance(P, list) and len(P) == pose.shape[1] and P[0].shape == (6, 1) and G[0].shape == (6, n)
        assert rel_err(np.stack(P)[:, :, 0], pose[0]) < TOL_F64 and rel_err(np.stack(G), grad[0]) < TOL_F64
        Pb = eng.end_effector_pose(q, names, off)
        Gb, Pb2 = eng.end_effector_pose_gradient(q, names, off, return_pose=True)
        assert Pb.shape == pose.shape and Gb.shape == grad.shape
        assert rel_err(Pb, pose) < TOL_F64 and rel_err(Pb2, pose) < TOL_F64 and rel_err(Gb, grad) < TOL_F64
        assert rel_err(e32.end_effector_pose(q, names, off), pose) < TOL_F32
        assert rel_err(e32.end_effector_pose_gradient(q, names, off), grad) < 20 * TOL_F32   # 1/(x^2+y^2) in the rpy columns
    with pytest.raises(ValueError, match="Could not find joint or fixed joint named"):
        eng.end_effector_pose(q[0], ["no_such_joint"])


@requires_cuda
@pytest.mark.parametrize("name,B", [("iiwa14", 4099), ("hyq", 1001), ("atlas", 1025), ("tree13", 333)])
def test_end_effector_batched_vs_oracle(name, B):
    """Ragged batches on the device against the vectorised oracle; zero columns off the chain."""
    rb = make_robot(name)
    eng, bo = _engine(rb), BatchOracle(rb)
    n = eng.n
    q = np.random.default_rng(B).uniform(-np.pi, np.pi, (B, n))
    Pref, Gref = bo.end_effector_pose(q, gradient=True)
    tq = _t(q)
    P = eng.end_effector_pose(tq)
    G, P2 = eng.end_effector_pose_gradient(tq, return_pose=True)
    assert P.is_cuda and G.shape == (B, len(rb.get_leaf_nodes()), 6, n)
    assert rel_err(P.cpu().numpy(), Pref) < TOL_F64 and torch.equal(P, P2)
    # per knot point scaling: a pose close to the pitch singularity has huge rpy derivatives
    Gn, scale = G.cpu().numpy(), np.maximum(1.0, np.abs(Gref).max(axis=(1, 2, 3), keepdims=True))
    assert np.max(np.abs(Gn - Gref) / scale) < TOL_F64
    for e, leaf in enumerate(rb.get_leaf_nodes()):
        off_chain = [j for j in range(n) if j != leaf and j not in rb.get_ancestors_by_id(leaf)]
        assert np.all(Gn[:, e][:, :, off_chain] == 0.0)
    assert eng.end_effector_pose(tq[:0]).shape == (0, len(rb.get_leaf_nodes()), 6)


@requires_cuda
def test_end_effector_full_size_properties():
    """2^20 iiwa14 knot points: the flange position keeps its distance to the last joint's origin,
    rpy stay in range, the gradient matches central differences of the pose on a sample, and
    sharded == unsharded bit for bit."""
    rb = make_robot("iiwa14")
    eng = _engine(rb)
    B, n = 1 << 20, eng.n
    g = torch.Generator(device="cuda").manual_seed(0xEE)
    q = (torch.rand(B, n, dtype=torch.float64, device="cuda", generator=g) * 2 - 1) * np.pi
    names = ["iiwa_joint_7", "iiwa_joint_ee"]
    P = eng.end_effector_pose(q, names)
    assert torch.isfinite(P).all()
    d = (P[:, 0, :3] - P[:, 1, :3]).norm(dim=1)
    assert float((d - 0.045).abs().max()) < 1e-12            # fixed flange offset (robots.iiwa14)
    assert float(P[..., 3].abs().max()) <= np.pi and float(P[..., 4].abs().max()) <= np.pi / 2 + 1e-12
    G = eng.end_effector_pose_gradient(q, names)
    assert float(G[:, :, :3, :].abs().max()) < 2.0            # |dp/dq_j| <= reach of the arm
    h = 1e-6
    idx = torch.arange(0, B, 4099, device="cuda")
    for j in range(n):
        dq = torch.zeros(n, dtype=torch.float64, device="cuda"); dq[j] = h
        num = (eng.end_effector_pose(q[idx] + dq, names) - eng.end_effector_pose(q[idx] - dq, names))[..., :3] / (2 * h)
        assert float((num - G[idx][:, :, :3, j]).abs().max()) < 1e-7
    half = B // 2 + 17
    G2 = torch.cat((eng.end_effector_pose_gradient(q[:half], names), eng.end_effector_pose_gradient(q[half:], names)))
    assert torch.equal(G, G2)


# ---------------------------------------------------------------------------------------------
# floating base: the reference's `floating_base` branches of rnea / rnea_grad / minv
# ---------------------------------------------------------------------------------------------
@pytest.fixture(params=[0, 1], ids=["coop", "thread"])
def fb_family(request):
    """Both kernel families of the fused floating-base drivers: 0 = automatic (warp-cooperative kernels in base
    coordinates), 1 = one knot point per thread (the body-frame recursion of round 1)."""
    from rbdreference_b200 import RBDReference
    RBDReference.set_kernel_variant(request.param)
    yield request.param
    RBDReference.set_kernel_variant(0)


@requires_cuda
@pytest.mark.parametrize("name", FB_CASES)
def test_floating_base_vs_reference_golden(name, fb_family):
    rb = make_fb_robot(name)
    g = load_fb_golden(name)
    eng, e32 = _engine(rb), _engine(rb, torch.float32)
    assert eng.floating_base and eng.n == rb.get_num_vel() and eng.nq == rb.get_num_pos()
    q, qd, qdd = g["q"], g["qd"], g["qdd"]
    c, v, a, f = eng.rnea(q, qd, qdd)
    for got, key in ((c, "c"), (v, "v"), (a, "a"), (f, "f")):
        assert rel_err(got, g[key]) < TOL_F64, key
    assert rel_err(eng.rnea(q, qd)[0], g["c_noqdd"]) < TOL_F64
    assert rel_err(eng.rnea(q, qd, qdd, GRAVITY=-3.7)[0], g["c_galt"]) < TOL_F64
    assert rel_err(eng.rnea(q, qd, qdd, outputs="c"), g["c"]) < TOL_F64
    assert rel_err(eng.rnea_grad(q, qd, qdd), g["dc_du"]) < TOL_F64
    assert rel_err(eng.rnea_grad(q, qd, qdd, USE_VELOCITY_DAMPING=True), g["dc_du_damped"]) < TOL_F64
    assert rel_err(eng.rnea_grad(q, qd), g["dc_du_noqdd"]) < TOL_F64
    assert rel_err(eng.minv(q), g["Minv"]) < TOL_F64
    assert rel_err(eng.minv(q, output_dense=False), g["Minv_sparse"]) < TOL_F64
    # one knot point, reference shapes
    c1, v1, _, _ = eng.rnea(q[0], qd[0], qdd[0])
    assert c1.shape == (eng.n,) and v1.shape == (6, eng.NB) and rel_err(c1, g["c"][0]) < TOL_F64
    assert rel_err(eng.rnea_grad(q[0], qd[0], qdd[0]), g["dc_du"][0]) < TOL_F64
    assert rel_err(eng.minv(q[0]), g["Minv"][0]) < TOL_F64
    # FP32
    assert rel_err(e32.rnea(q, qd, qdd)[0], g["c"]) < TOL_F32
    assert rel_err(e32.rnea_grad(q, qd, qdd), g["dc_du"]) < TOL_F32
    # the cooperative kernel meets the north-star FP32 bar; the thread-per-knot-point recursion in body frames does not
    assert rel_err(e32.minv(q), g["Minv"]) < (TOL_F32 if fb_family == 0 else 5 * TOL_F32)
    assert row_scaled_err(eng.minv(q), g["Minv"]) < 1e3 * TOL_F64
    # compositions (RBDReference.py:1369-1384); conditioning of Minv enters, as for the fixed base
    assert rel_err(eng.forward_dynamics(q, qd, g["u"]), g["fd_qdd"]) < 10 * TOL_F64
    d1, d2 = eng.forward_dynamics_grad(q, qd, g["u"])
    assert rel_err(d1, g["fd_dq"]) < 10 * TOL_F64 and rel_err(d2, g["fd_dqd"]) < 10 * TOL_F64
    from rbdreference_b200 import RBDReference
    RBDReference.set_kernel_variant(1)            # register-tiled product instead of the FP64 tensor-core one
    try:
        t1, t2 = eng.forward_dynamics_grad(q, qd, g["u"])
        assert rel_err(t1, g["fd_dq"]) < 10 * TOL_F64 and rel_err(t2, g["fd_dqd"]) < 10 * TOL_F64
    finally:
        RBDReference.set_kernel_variant(fb_family)
    with pytest.raises(NotImplementedError):
        eng.crba(q)


@requires_cuda
@pytest.mark.parametrize("name", FB_CASES)
def test_floating_base_pass_helpers_vs_reference_golden(name, fb_family):
    """The eight per-pass entry points of a floating-base robot against arrays of the unmodified reference: batched
    (numpy and CUDA tensors), one knot point with reference shapes, in-place contracts, both precisions, and the
    reference's own call sequence composed from the helpers."""
    rb = make_fb_robot(name)
    g = load_fbpass_golden(name)
    q, qd, qdd = g["q"], g["qd"], g["qdd"]
    for dtype, tol in ((torch.float64, TOL_F64), (torch.float32, TOL_F32)):
        eng = _engine(rb, dtype)
        v, a, f = eng.rnea_fpass(q, qd, qdd)
        for got, key in ((v, "v"), (a, "a"), (f, "f")):
            assert rel_err(got, g[key]) < tol, key
        fa = _t(g["f"], dtype)
        c, fr = eng.rnea_bpass(_t(q, dtype), fa)
        assert fr is fa and rel_err(c.cpu().numpy(), g["c"]) < tol and rel_err(fa.cpu().numpy(), g["f_acc"]) < tol
        Mb, Fb, U, D = eng.minv_bpass(q)
        for got, key in ((Mb, "Minv_b"), (Fb, "F_b"), (U, "U"), (D, "Dinv")):
            assert rel_err(got, g[key]) < tol * (5 if dtype == torch.float32 else 1), key
        Mf, Ff = _t(g["Minv_b"], dtype), _t(g["F_b"], dtype)
        out = eng.minv_fpass(_t(q, dtype), Mf, Ff, _t(g["U"], dtype), _t(g["Dinv"], dtype))
        assert out is Mf
        assert rel_err(Mf.cpu().numpy(), g["Minv_f"]) < tol * (5 if dtype == torch.float32 else 1)
        assert rel_err(Ff.cpu().numpy(), g["F_f"]) < tol * (5 if dtype == torch.float32 else 1)
        dv, da, df = eng.rnea_grad_fpass_dq(q, qd, g["v"], g["a"])
        for got, key in ((dv, "dv_dq"), (da, "da_dq"), (df, "df_dq")):
            assert rel_err(got, g[key]) < tol, key
        dv, da, df = eng.rnea_grad_fpass_dqd(q, qd, g["v"])
        for got, key in ((dv, "dv_dqd"), (da, "da_dqd"), (df, "df_dqd")):
            assert rel_err(got, g[key]) < tol, key
        dfa = _t(g["df_dq"], dtype)
        dc = eng.rnea_grad_bpass_dq(_t(q, dtype), _t(g["f_acc"], dtype), dfa)
        assert rel_err(dc.cpu().numpy(), g["dc_dq"]) < tol and rel_err(dfa.cpu().numpy(), g["df_dq_acc"]) < tol
        dfa = _t(g["df_dqd"], dtype)
        dc = eng.rnea_grad_bpass_dqd(_t(q, dtype), dfa)
        assert rel_err(dc.cpu().numpy(), g["dc_dqd"]) < tol and rel_err(dfa.cpu().numpy(), g["df_dqd_acc"]) < tol
        assert rel_err(eng.rnea_grad_bpass_dqd(q, g["df_dqd"].copy(), USE_VELOCITY_DAMPING=True), g["dc_dqd_damped"]) < tol
    eng = _engine(rb)
    # one knot point, reference shapes, numpy in place
    f0 = g["f"][0].copy()
    c0, fr0 = eng.rnea_bpass(q[0], f0)
    assert fr0 is f0 and c0.shape == (eng.n,) and rel_err(f0, g["f_acc"][0]) < TOL_F64
    M0, F0, U0, D0 = eng.minv_bpass(q[0])
    assert M0.shape == (eng.n, eng.n) and F0.shape == (eng.n, 6, eng.n) and U0.shape == (eng.n, 6) and D0.shape == (eng.n,)
    assert eng.minv_fpass(q[0], M0, F0, U0, D0) is M0 and rel_err(M0, g["Minv_f"][0]) < TOL_F64
    # the reference's own call sequences (:1353-1367, :793-804) composed from the helpers equal the fused drivers
    fb = load_fb_golden(name)
    assert rel_err(eng.rnea_grad_passes(fb["q"], fb["qd"], fb["qdd"]), fb["dc_du"]) < TOL_F64
    assert rel_err(eng.rnea_grad_passes(fb["q"], fb["qd"], fb["qdd"], USE_VELOCITY_DAMPING=True), fb["dc_du_damped"]) < TOL_F64
    assert rel_err(eng.minv_passes(fb["q"]), fb["Minv"]) < TOL_F64
    assert rel_err(eng.minv_passes(fb["q"], output_dense=False), fb["Minv_sparse"]) < TOL_F64


@requires_cuda
def test_floating_base_grad_fpass_dq_needs_six_bodies():
    """RBDReference.py:1168 indexes bodies 0..5: the reference raises IndexError for smaller robots, the library refuses."""
    from rbdreference_b200 import robots
    from rbdreference_b200._capi import RbdError
    rb = robots.FloatingBaseRobot(robots.random_tree(3, seed=4), name="tiny_fb")
    eng = _engine(rb)
    q, qd, qdd = rb.random_state(np.random.default_rng(0), 4)
    v, a, f = eng.rnea_fpass(q, qd, qdd)
    with pytest.raises(RbdError):
        eng.rnea_grad_fpass_dq(q, qd, v, a)
    eng.rnea_grad_fpass_dqd(q, qd, v)


@requires_cuda
@pytest.mark.parametrize("name,B", [("hyq", 257), ("atlas", 65)])
def test_floating_base_batched_vs_oracle_and_identities(name, B, fb_family):
    """Ragged device batches against the scalar oracle; Minv inverts the mass matrix assembled
    from the engine's own rnea columns; dc_dqd matches central differences of rnea."""
    from oracle.rbd_oracle_fb import FloatingScalarOracle
    rb = make_fb_robot(name)
    eng, so = _engine(rb), FloatingScalarOracle(rb)
    nv = eng.n
    q, qd, qdd = rb.random_state(np.random.default_rng(B), B)
    tq, tqd, tqdd = _t(q), _t(qd), _t(qdd)
    cbuf = torch.empty(B, nv, dtype=torch.float64, device="cuda")
    dc = eng.rnea_grad(tq, tqd, tqdd, c_out=cbuf).cpu().numpy()
    M = eng.minv(tq).cpu().numpy()
    c = eng.rnea(tq, tqd, tqdd, outputs="c").cpu().numpy()
    if fb_family == 1:
        assert np.array_equal(c, cbuf.cpu().numpy())              # the same body-frame recursion in both kernels
    else:
        assert rel_err(cbuf.cpu().numpy(), c) < 1e-12              # composite forces in base coordinates vs the recursion
    for k in range(0, B, 8):
        assert rel_err(c[k], so.rnea(q[k], qd[k], qdd[k])[0]) < TOL_F64
        assert rel_err(cbuf[k].cpu().numpy(), so.rnea(q[k], qd[k], qdd[k])[0]) < TOL_F64
        assert rel_err(dc[k], so.rnea_grad(q[k], qd[k], qdd[k])) < TOL_F64
        assert rel_err(M[k], so.minv(q[k])) < TOL_F64
    zero = torch.zeros_like(tqd)
    c0 = eng.rnea(tq, zero, zero, GRAVITY=0.0, outputs="c")
    H = torch.stack([eng.rnea(tq, zero, torch.eye(nv, dtype=torch.float64, device="cuda")[j].expand(B, nv).contiguous(),
                              GRAVITY=0.0, outputs="c") - c0 for j in range(nv)], dim=2)
    eye = torch.eye(nv, dtype=torch.float64, device="cuda")
    assert float((torch.as_tensor(M, device="cuda") @ H - eye).abs().max()) < 1e-9
    h = 1e-6
    scale = np.abs(dc).max(axis=(1, 2), keepdims=True)
    for j in (0, 4, 5, 6, nv - 1):
        dv = torch.zeros(nv, dtype=torch.float64, device="cuda"); dv[j] = h
        num = ((eng.rnea(tq, tqd + dv, tqdd, outputs="c") - eng.rnea(tq, tqd - dv, tqdd, outputs="c")) / (2 * h)).cpu().numpy()
        assert np.max(np.abs(num - dc[:, :, nv + j]) / scale[:, :, 0]) < 1e-7
    assert eng.minv(tq[:0]).shape == (0, nv, nv)


@requires_cuda
def test_floating_base_with_prismatic_joints_vs_oracle(fb_family):
    """A branched tree with prismatic joints on a floating base: the reference's prismatic quirk of rnea_grad (:1292)
    and the prismatic branches of minv, in base coordinates (cooperative family) and in body frames, both precisions,
    damping on, ragged batch."""
    from oracle.rbd_oracle_fb import FloatingScalarOracle
    from rbdreference_b200 import robots
    tree = robots.random_tree(13, seed=2, branching=0.5, prismatic=0.3)
    assert any(j.kind == "prismatic" for j in tree.joints)
    fb = robots.FloatingBaseRobot(tree, name="tree13_fb")
    eng, e32, so = _engine(fb), _engine(fb, torch.float32), FloatingScalarOracle(fb)
    B = 75
    q, qd, qdd = fb.random_state(np.random.default_rng(21), B)
    dc = eng.rnea_grad(_t(q), _t(qd), _t(qdd), USE_VELOCITY_DAMPING=True).cpu().numpy()
    M = eng.minv(_t(q)).cpu().numpy()
    dc32 = e32.rnea_grad(_t(q, torch.float32), _t(qd, torch.float32), _t(qdd, torch.float32), USE_VELOCITY_DAMPING=True).cpu().numpy()
    M32 = e32.minv(_t(q, torch.float32)).cpu().numpy()
    for k in (0, 1, 31, 32, 74):
        ref = np.asarray(so.rnea_grad(q[k], qd[k], qdd[k], USE_VELOCITY_DAMPING=True))
        Mref = np.asarray(so.minv(q[k]))
        cond = np.linalg.cond(Mref)
        assert rel_err(dc[k], ref) < TOL_F64 and rel_err(dc32[k], ref) < TOL_F32
        assert rel_err(M[k], Mref) < TOL_F64 * max(1.0, cond / 1e4)
        assert rel_err(M32[k], Mref) < (TOL_F32 if fb_family == 0 else 5 * TOL_F32) * max(1.0, cond / 1e4)


@requires_cuda
def test_end_effector_and_floating_base_calls_are_cuda_graph_capturable_and_stream_safe():
    """The new entry points only enqueue stream-ordered work too: captured in a CUDA graph, replayed on
    new inputs, and issued concurrently on two streams they return what the eager calls return."""
    rb, fb = make_robot("hyq"), make_fb_robot("hyq")
    eng, feng = _engine(rb), _engine(fb)
    B = 192
    sq = torch.zeros(B, eng.n, dtype=torch.float64, device="cuda")
    fq, fqd, fqdd = (torch.zeros(B, k, dtype=torch.float64, device="cuda") for k in (feng.nq, feng.n, feng.n))

    def load(seed):
        sq.copy_(_t(np.random.default_rng(seed).uniform(-np.pi, np.pi, (B, eng.n))))
        a, b, c = fb.random_state(np.random.default_rng(seed + 100), B)
        fq.copy_(_t(a)); fqd.copy_(_t(b)); fqdd.copy_(_t(c))

    load(1)
    eng.end_effector_pose_gradient(sq, return_pose=True); feng.rnea_grad(fq, fqd, fqdd); feng.forward_dynamics_grad(fq, fqd, fqdd)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        g_grad, g_pose = eng.end_effector_pose_gradient(sq, return_pose=True)
        g_dc = feng.rnea_grad(fq, fqd, fqdd)
        g_M = feng.minv(fq)
        g_f1, g_f2 = feng.forward_dynamics_grad(fq, fqd, fqdd)
    for seed in (2, 3):
        load(seed)
        graph.replay()
        torch.cuda.synchronize()
        e_grad, e_pose = eng.end_effector_pose_gradient(sq, return_pose=True)
        e_f1, e_f2 = feng.forward_dynamics_grad(fq, fqd, fqdd)
        for g, r in ((g_grad, e_grad), (g_pose, e_pose), (g_dc, feng.rnea_grad(fq, fqd, fqdd)), (g_M, feng.minv(fq)),
                     (g_f1, e_f1), (g_f2, e_f2)):
            assert torch.equal(g, r)
    # two streams at once
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    with torch.cuda.stream(s1):
        a1 = [eng.end_effector_pose_gradient(sq) for _ in range(4)]
    with torch.cuda.stream(s2):
        a2 = [feng.rnea_grad(fq, fqd, fqdd) for _ in range(4)]
    torch.cuda.synchronize()
    assert all(torch.equal(x, e_grad) for x in a1) and all(torch.equal(x, g_dc) for x in a2)


@requires_cuda
@pytest.mark.parametrize("kind", ["one", "two", "chain32", "bush32", "stars"])
def test_edge_topologies_end_effector_and_floating_base(kind):
    """n = 1, 2 and RBD_MAX_DOF bodies, the deepest chain (one end effector, 32 columns: dense tile)
    and the flattest trees (many end effectors with one-joint chains: compact tile); the same trees
    on a floating base (up to 37 velocity coordinates), both precisions."""
    from oracle.rbd_oracle_fb import FloatingScalarOracle
    from rbdreference_b200 import robots
    rb = _edge_robot(kind)
    eng, e32, bo = _engine(rb), _engine(rb, torch.float32), BatchOracle(rb)
    B = 97
    q = np.random.default_rng(9).uniform(-np.pi, np.pi, (B, eng.n))
    Pref, Gref = bo.end_effector_pose(q, gradient=True)
    G, P = eng.end_effector_pose_gradient(_t(q), return_pose=True)
    scale = np.maximum(1.0, np.abs(Gref).max(axis=(1, 2, 3), keepdims=True))     # per knot point (rpy singularities)
    assert rel_err(P.cpu().numpy(), Pref) < TOL_F64
    assert np.max(np.abs(G.cpu().numpy() - Gref) / scale) < TOL_F64
    G32 = e32.end_effector_pose_gradient(_t(q, torch.float32)).cpu().numpy()
    assert np.max(np.abs(G32 - Gref) / scale) < 20 * TOL_F32
    if rb.get_num_bodies() < 2:
        return
    fb = robots.FloatingBaseRobot(_edge_robot(kind) if rb.get_num_bodies() < 32 else robots.random_tree(31, seed=5, branching=0.3))
    feng, so = _engine(fb), FloatingScalarOracle(fb)
    fq, fqd, fqdd = fb.random_state(np.random.default_rng(10), 33)
    c = feng.rnea(_t(fq), _t(fqd), _t(fqdd), outputs="c").cpu().numpy()
    M = feng.minv(_t(fq)).cpu().numpy()
    for k in (0, 32):
        assert rel_err(c[k], so.rnea(fq[k], fqd[k], fqdd[k])[0]) < TOL_F64
        assert rel_err(M[k], so.minv(fq[k])) < TOL_F64
    if fb.get_num_bodies() >= 6:                     # the reference's gradient needs NB >= 6 (:1168)
        dc = feng.rnea_grad(_t(fq), _t(fqd), _t(fqdd), USE_VELOCITY_DAMPING=True).cpu().numpy()
        for k in (0, 32):
            assert rel_err(dc[k], so.rnea_grad(fq[k], fqd[k], fqdd[k], USE_VELOCITY_DAMPING=True)) < TOL_F64


# ---------------------------------------------------------------------------------------------
# round 2: serial-chain rnea_grad kernel, host-buffer pipeline, result-buffer validation, per-handle variants
# ---------------------------------------------------------------------------------------------
def _chain_robot(n, prismatic_at=()):
    import copy
    from rbdreference_b200 import robots
    joints = copy.deepcopy(robots.random_tree(n, seed=100 + n, branching=0.0, prismatic=0.0).joints)
    for i, j in enumerate(joints):
        j.parent = i - 1
        if i in prismatic_at:
            j.kind = "prismatic"
    return robots.Robot("chain%d" % n, joints)


@requires_cuda
@pytest.mark.parametrize("case", ["iiwa14", "chain4", "chain5", "chain6", "chain8", "chain7_prismatic"])
def test_chain_kernel_vs_oracle(case):
    """variant 7 (one knot point per lane, serial chains): ragged batch, damping, qdd=None, alternate gravity, c_out,
    both precisions; prismatic joints exercise the reference's :1292 quirk path of the kernel."""
    from rbdreference_b200 import RBDReference, robots
    rb = robots.iiwa14() if case == "iiwa14" else (_chain_robot(7, prismatic_at=(0, 3, 6)) if case == "chain7_prismatic"
                                                   else _chain_robot(int(case[5:])))
    bo = BatchOracle(rb)
    n = rb.get_num_vel()
    B = 32 * 5 + 13
    q, qd, qdd = random_states(n, B, seed=77)
    for dtype, tol in ((torch.float64, TOL_F64), (torch.float32, TOL_F32 * (1 if case == "iiwa14" else 10))):
        eng = _engine(rb, dtype)
        eng.set_variant(7)
        before = eng.launch_count()
        cbuf = torch.empty(B, n, dtype=dtype, device="cuda")
        got = eng.rnea_grad(_t(q, dtype), _t(qd, dtype), _t(qdd, dtype), c_out=cbuf)
        assert eng.launch_count() == before + 1
        assert rel_err(got.cpu().numpy(), bo.rnea_grad(q, qd, qdd)) < tol
        assert rel_err(cbuf.cpu().numpy(), bo.rnea(q, qd, qdd)[0]) < tol
        got = eng.rnea_grad(_t(q, dtype), _t(qd, dtype), None, GRAVITY=-3.3, USE_VELOCITY_DAMPING=True)
        assert rel_err(got.cpu().numpy(), bo.rnea_grad(q, qd, None, GRAVITY=-3.3, USE_VELOCITY_DAMPING=True)) < tol
        # agrees with the cooperative kernel to rounding
        e3 = _engine(rb, dtype)
        e3.set_variant(3)
        ref3 = e3.rnea_grad(_t(q, dtype), _t(qd, dtype), _t(qdd, dtype))
        assert rel_err(eng.rnea_grad(_t(q, dtype), _t(qd, dtype), _t(qdd, dtype)).cpu().numpy(), ref3.cpu().numpy()) < tol


@requires_cuda
def test_chain_kernel_far_angles_and_nan():
    """|q| beyond the fast range of the kernel's own sincos takes the library path; NaN inputs stay NaN-local."""
    from rbdreference_b200 import robots
    rb = robots.iiwa14()
    bo = BatchOracle(rb)
    q, qd, qdd = random_states(7, 64, seed=5)
    q[3] *= 1e6
    q[10, 2] = 2.0e5
    eng = _engine(rb)
    eng.set_variant(7)
    got = eng.rnea_grad(_t(q), _t(qd), _t(qdd)).cpu().numpy()
    assert rel_err(got, bo.rnea_grad(q, qd, qdd)) < 1e-9       # sin/cos of 1e6-sized angles carry ~1e-10 of argument rounding
    q[7, 1] = np.nan
    got = eng.rnea_grad(_t(q), _t(qd), _t(qdd)).cpu().numpy()
    assert np.isnan(got[7]).any() and not np.isnan(np.delete(got, 7, axis=0)).any()


@requires_cuda
def test_per_handle_variant_is_independent():
    from rbdreference_b200 import RBDReference, robots
    rb = robots.iiwa14()
    e_a, e_b = _engine(rb), _engine(rb)
    e_a.set_variant(1)                       # generic body-frame kernels on this handle only
    q, qd, qdd = random_states(7, 50, seed=3)
    ra = e_a.rnea_grad(_t(q), _t(qd), _t(qdd)).cpu().numpy()
    rb_ = e_b.rnea_grad(_t(q), _t(qd), _t(qdd)).cpu().numpy()
    assert rel_err(ra, rb_) < TOL_F64 and not np.array_equal(ra, rb_)     # different kernels, same result to rounding
    e_a.set_variant(-1)
    assert np.array_equal(e_a.rnea_grad(_t(q), _t(qd), _t(qdd)).cpu().numpy(), rb_)
    with pytest.raises(Exception):
        e_a.set_variant(6)
    # floating-base handles: 1 = one knot point per thread, 0 / -1 = cooperative kernels in base coordinates
    fb = make_fb_robot("hyq")
    f_a, f_b = _engine(fb), _engine(fb)
    f_a.set_variant(1)
    fq, fqd, fqdd = fb.random_state(np.random.default_rng(5), 70)
    for call in (lambda e: e.rnea_grad(_t(fq), _t(fqd), _t(fqdd)), lambda e: e.minv(_t(fq))):
        xa, xb = call(f_a).cpu().numpy(), call(f_b).cpu().numpy()
        assert rel_err(xa, xb) < TOL_F64 and not np.array_equal(xa, xb)
    f_a.set_variant(-1)
    assert np.array_equal(f_a.minv(_t(fq)).cpu().numpy(), f_b.minv(_t(fq)).cpu().numpy())
    with pytest.raises(Exception):
        f_a.set_variant(7)


@requires_cuda
@pytest.mark.parametrize("pinned", [True, False])
def test_host_numpy_pipeline_bit_identical_to_device_path(pinned):
    """eng.rnea_grad / minv / rnea on (B, n) numpy arrays run the chunked pinned pipeline inside the engine; the result is
    bit-identical to the CUDA-tensor call (same kernels on the same values), for pinned and pageable arrays, with and
    without out=."""
    from rbdreference_b200 import RBDReference, robots
    rb = robots.iiwa14()
    eng = _engine(rb)
    B = (1 << 18) + 37
    q, qd, qdd = random_states(7, B, seed=11)
    if pinned:
        hq, hqd, hqdd = (RBDReference.pinned_empty(x.shape) for x in (q, qd, qdd))
        for d, s_ in zip((hq, hqd, hqdd), (q, qd, qdd)):
            np.copyto(d, s_)
        out = RBDReference.pinned_empty((B, 7, 14))
    else:
        hq, hqd, hqdd = q, qd, qdd
        out = np.empty((B, 7, 14))
    ref = eng.rnea_grad(_t(q), _t(qd), _t(qdd)).cpu().numpy()
    got = eng.rnea_grad(hq, hqd, hqdd, out=out)
    assert got is out and np.array_equal(got, ref)
    got2 = eng.rnea_grad(hq, hqd, hqdd)
    assert isinstance(got2, np.ndarray) and np.array_equal(got2, ref)
    assert np.array_equal(eng.minv(hq), eng.minv(_t(q)).cpu().numpy())
    c, v, a, f = eng.rnea(hq, hqd, hqdd)
    rc, rv, ra, rf = eng.rnea(_t(q), _t(qd), _t(qdd))
    for x, y in ((c, rc), (v, rv), (a, ra), (f, rf)):
        assert np.array_equal(x, y.cpu().numpy())
    assert np.array_equal(eng.rnea(hq, hqd, None, outputs="c"), eng.rnea(_t(q), _t(qd), None, outputs="c").cpu().numpy())
    # float32 engine fed float64 numpy: converted on the host, same result as converting first
    e32 = _engine(rb, torch.float32)
    assert np.array_equal(e32.minv(hq[:5000]), e32.minv(_t(q[:5000], torch.float32)).cpu().numpy())
    # tiny and empty batches go through the same path
    assert np.array_equal(eng.rnea_grad(hq[:3], hqd[:3], hqdd[:3]), eng.rnea_grad(_t(q[:3]), _t(qd[:3]), _t(qdd[:3])).cpu().numpy())
    assert eng.rnea_grad(hq[:0], hqd[:0], hqdd[:0]).shape == (0, 7, 14)


@requires_cuda
def test_result_buffers_are_validated():
    from rbdreference_b200 import robots
    rb = robots.iiwa14()
    eng = _engine(rb)
    q, qd, qdd = (_t(x) for x in random_states(7, 16, seed=1))
    good = torch.empty(16, 7, 14, dtype=torch.float64, device="cuda")
    eng.rnea_grad(q, qd, qdd, out=good)
    bad = [torch.empty(16, 7, 14, dtype=torch.float64),                          # CPU tensor
           torch.empty(16, 7, 14, dtype=torch.float32, device="cuda"),            # wrong dtype
           torch.empty(15, 7, 14, dtype=torch.float64, device="cuda"),            # wrong shape
           torch.empty(16, 14, 7, dtype=torch.float64, device="cuda").transpose(1, 2)]   # not contiguous
    for b in bad:
        with pytest.raises(ValueError):
            eng.rnea_grad(q, qd, qdd, out=b)
    with pytest.raises(ValueError):
        eng.rnea_grad(q, qd, qdd, c_out=torch.empty(16, 8, dtype=torch.float64, device="cuda"))
    with pytest.raises(ValueError):
        eng.minv(q, out=torch.empty(16, 7, 7, dtype=torch.float64))
    with pytest.raises(ValueError):
        eng.crba(q, out=torch.empty(16, 7, 8, dtype=torch.float64, device="cuda"))
    with pytest.raises(ValueError):
        eng.rnea_grad(q[0], qd[0], qdd[0], out=good)                               # unbatched call with out=
    with pytest.raises(ValueError):
        eng.rnea_grad(q.cpu().numpy(), qd.cpu().numpy(), qdd.cpu().numpy(), out=good)   # numpy call with a torch out
    with pytest.raises(ValueError):
        eng.minv(q.cpu().numpy(), out=np.empty((16, 7, 7), dtype=np.float32))      # numpy out of the wrong dtype


@requires_cuda
@pytest.mark.parametrize("name", ["iiwa14", "atlas"])
def test_fused_drivers_vs_wide_reference_golden(name):
    """256 iiwa14 / 24 Atlas states of the unmodified reference (tests/golden/wide_*.npz) through every kernel family
    that serves the robot, with the per-tensor bar AND a per-row bar (rows of Minv / dc_du of light distal links are
    orders of magnitude below the tensor's largest entry; the row bar is 100x the tensor bar: a row's own scale can
    sit 1e-2 below the entries it is computed from)."""
    import os
    g = np.load(os.path.join(GOLDEN_DIR, "wide_" + name + ".npz"))
    rb = make_robot(name)
    q, qd, qdd = g["q"], g["qd"], g["qdd"]
    for dtype, tol in ((torch.float64, TOL_F64), (torch.float32, TOL_F32)):
        eng = _engine(rb, dtype)
        tq, tqd, tqdd = _t(q, dtype), _t(qd, dtype), _t(qdd, dtype)
        for variant in (0, 1, 2, 3, 4, 5, 7, 8, 9):
            eng.set_variant(variant)
            dc = eng.rnea_grad(tq, tqd, tqdd).cpu().numpy()
            M = eng.minv(tq).cpu().numpy()
            c = eng.rnea(tq, tqd, tqdd, outputs="c").cpu().numpy()
            for got, key in ((c, "c"), (dc, "dc_du"), (M, "Minv")):
                assert rel_err(got, g[key]) < tol, (variant, key)
                assert row_scaled_err(got, g[key]) < 100 * tol, (variant, key)
