"""CPU tests: the oracle restatement against the reference's golden outputs and identities."""
import numpy as np
import pytest

from conftest import GOLDEN_DIR, row_scaled_err, FB_CASES, GOLDEN_CASES, load_ee_golden, load_fb_golden, load_fbpass_golden, make_fb_robot, rel_err, random_states, make_robot
from oracle.rbd_oracle import BatchOracle, ScalarOracle
from oracle import build_ref

PIN = 1e-11   # oracle vs reference golden vectors (both float64, different summation order)


def test_golden_matches_robot_tables(golden):
    name, rb, g = golden
    n = rb.get_num_vel()
    assert np.array_equal(g["parent"], [rb.get_parent_id(i) for i in range(n)])
    assert np.allclose(g["S"], np.stack([rb.get_S_by_id(i) for i in range(n)]), atol=0)
    assert np.allclose(g["I"], np.stack([rb.get_Imat_by_id(i) for i in range(n)]), rtol=0, atol=1e-15)
    assert np.allclose(g["X_at_0p3"], np.stack([rb.get_Xmat_Func_by_id(i)(0.3) for i in range(n)]), rtol=0, atol=1e-15)


def test_scalar_oracle_vs_golden(golden):
    name, rb, g = golden
    so = ScalarOracle(rb)
    q, qd, qdd = g["q"], g["qd"], g["qdd"]
    for k in range(q.shape[0]):
        v, a, f = so.rnea_fpass(q[k], qd[k], qdd[k])
        assert rel_err(f, g["f_fpass"][k]) < PIN
        c, f2 = so.rnea_bpass(q[k], f)
        assert f2 is f                                   # in-place contract
        for got, key in ((c, "c"), (v, "v"), (a, "a"), (f, "f")):
            assert rel_err(got, g[key][k]) < PIN, key
        assert rel_err(so.rnea(q[k], qd[k])[0], g["c_noqdd"][k]) < PIN
        assert rel_err(so.rnea(q[k], qd[k], qdd[k], GRAVITY=-3.7)[0], g["c_galt"][k]) < PIN
        assert rel_err(so.rnea_grad(q[k], qd[k], qdd[k]), g["dc_du"][k]) < PIN
        assert rel_err(so.rnea_grad(q[k], qd[k], qdd[k], USE_VELOCITY_DAMPING=True), g["dc_du_damped"][k]) < PIN
        assert rel_err(so.rnea_grad(q[k], qd[k]), g["dc_du_noqdd"][k]) < PIN
        dv, da, df = so.rnea_grad_fpass_dq(q[k], qd[k], g["v"][k], g["a"][k])
        for got, key in ((dv, "dv_dq"), (da, "da_dq"), (df, "df_dq")):
            assert rel_err(got, g[key][k]) < PIN, key
        dv2, da2, df2 = so.rnea_grad_fpass_dqd(q[k], qd[k], g["v"][k])
        for got, key in ((dv2, "dv_dqd"), (da2, "da_dqd"), (df2, "df_dqd")):
            assert rel_err(got, g[key][k]) < PIN, key
        assert rel_err(so.rnea_grad_bpass_dq(q[k], g["f"][k], df), g["dc_dq"][k]) < PIN
        assert rel_err(df, g["df_dq_acc"][k]) < PIN
        assert rel_err(so.rnea_grad_bpass_dqd(q[k], df2), g["dc_dqd"][k]) < PIN
        assert rel_err(df2, g["df_dqd_acc"][k]) < PIN
        Mb, Fb, U, D = so.minv_bpass(q[k])
        for got, key in ((Mb, "Minv_b"), (Fb, "F_b"), (U, "U"), (D, "Dinv")):
            assert rel_err(got, g[key][k]) < PIN, key
        M = so.minv_fpass(q[k], Mb, Fb, U, D)
        assert M is Mb
        assert rel_err(Fb, g["F_f"][k]) < PIN
        assert rel_err(so.minv(q[k]), g["Minv"][k]) < PIN
        assert rel_err(so.minv(q[k], output_dense=False), g["Minv_sparse"][k]) < PIN
        assert rel_err(so.crba(q[k]), g["H"][k]) < PIN
        assert rel_err(so.forward_dynamics(q[k], qd[k], g["u"][k]), g["fd_qdd"][k]) < PIN
        r1, r2 = so.forward_dynamics_grad(q[k], qd[k], g["u"][k])
        assert rel_err(r1, g["fd_dq"][k]) < PIN and rel_err(r2, g["fd_dqd"][k]) < PIN
        assert rel_err(so.aba(q[k], qd[k], g["u"][k]), g["aba_qdd"][k]) < PIN
        assert rel_err(so.aba(q[k], qd[k], g["u"][k], GRAVITY=-3.7), g["aba_qdd_galt"][k]) < PIN


def test_batch_oracle_vs_golden(golden):
    name, rb, g = golden
    bo = BatchOracle(rb)
    q, qd, qdd = g["q"], g["qd"], g["qdd"]
    c, v, a, f = bo.rnea(q, qd, qdd)
    for got, key in ((c, "c"), (v, "v"), (a, "a"), (f, "f")):
        assert rel_err(got, g[key]) < PIN, key
    assert rel_err(bo.rnea(q, qd)[0], g["c_noqdd"]) < PIN
    dc, parts = bo.rnea_grad(q, qd, qdd, return_parts=True)
    assert rel_err(dc, g["dc_du"]) < PIN
    for key in ("dv_dq", "da_dq", "df_dq", "dv_dqd", "da_dqd", "df_dqd", "df_dq_acc", "df_dqd_acc"):
        assert rel_err(parts[key], g[key]) < PIN, key
    assert rel_err(bo.rnea_grad(q, qd, qdd, USE_VELOCITY_DAMPING=True), g["dc_du_damped"]) < PIN
    M, mp = bo.minv(q, return_parts=True)
    assert rel_err(M, g["Minv"]) < PIN
    for key, gk in (("Minv_b", "Minv_b"), ("F_b", "F_b"), ("U", "U"), ("D", "Dinv"), ("F_f", "F_f")):
        assert rel_err(mp[key], g[gk]) < PIN, key
    assert rel_err(bo.minv(q, output_dense=False), g["Minv_sparse"]) < PIN
    assert rel_err(bo.crba(q), g["H"]) < PIN
    assert rel_err(bo.aba(q, qd, g["u"]), g["aba_qdd"]) < PIN
    assert rel_err(bo.aba(q, qd, g["u"], GRAVITY=-3.7), g["aba_qdd_galt"]) < PIN


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_end_effector_oracle_vs_golden(name):
    """end_effector_pose / _gradient (RBDReference.py:220-386) of both oracle layers against the
    reference's outputs: default leaves, named moving + fixed joints, zero and non-zero offset."""
    rb = make_robot(name)
    so, bo = ScalarOracle(rb), BatchOracle(rb)
    q, cases = load_ee_golden(name)
    n = rb.get_num_vel()
    for names, off, pose, grad in cases:
        for k in range(q.shape[0]):
            P = so.end_effector_pose(q[k], names, off)
            G = so.end_effector_pose_gradient(q[k], names, off)
            assert len(P) == pose.shape[1] and P[0].shape == (6, 1) and G[0].shape == (6, n)
            assert rel_err(np.stack(P)[:, :, 0], pose[k]) < PIN
            assert rel_err(np.stack(G), grad[k]) < PIN
        Pb, Gb = bo.end_effector_pose(q, names, off, gradient=True)
        assert rel_err(Pb, pose) < PIN and rel_err(Gb, grad) < PIN
    with pytest.raises(ValueError, match="Could not find joint or fixed joint named"):
        so.end_effector_pose(q[0], ["no_such_joint"])


@pytest.mark.parametrize("name", ["iiwa14", "atlas", "tree13"])
def test_end_effector_gradient_is_the_derivative_of_the_pose(name):
    rb = make_robot(name)
    bo = BatchOracle(rb)
    n = rb.get_num_vel()
    q = np.random.default_rng(5).uniform(-1.2, 1.2, (6, n))
    G = bo.end_effector_pose_gradient(q)
    h = 1e-6
    for j in range(n):
        dq = np.zeros(n); dq[j] = h
        num = (bo.end_effector_pose(q + dq) - bo.end_effector_pose(q - dq)) / (2 * h)
        num[..., 3:] = (num[..., 3:] + np.pi / (2 * h)) % (np.pi / h) - np.pi / (2 * h)     # angle wrap
        assert np.max(np.abs(num - G[..., j])) < 1e-6 * max(1.0, np.max(np.abs(G)))


@pytest.mark.parametrize("name", FB_CASES)
def test_floating_base_oracle_vs_golden(name):
    """The floating-base branches against the unmodified reference's outputs."""
    from oracle.rbd_oracle_fb import FloatingScalarOracle
    rb = make_fb_robot(name)
    g = load_fb_golden(name)
    so = FloatingScalarOracle(rb)
    assert np.allclose(rb.get_Xmat_Func_by_id(0)(g["q"][0, 0:7]), g["X0_first"], rtol=0, atol=1e-15)
    for k in range(g["q"].shape[0]):
        q, qd, qdd = g["q"][k], g["qd"][k], g["qdd"][k]
        c, v, a, f = so.rnea(q, qd, qdd)
        for got, key in ((c, "c"), (v, "v"), (a, "a"), (f, "f")):
            assert rel_err(got, g[key][k]) < PIN, key
        assert rel_err(so.rnea(q, qd)[0], g["c_noqdd"][k]) < PIN
        assert rel_err(so.rnea(q, qd, qdd, GRAVITY=-3.7)[0], g["c_galt"][k]) < PIN
        assert rel_err(so.rnea_grad(q, qd, qdd), g["dc_du"][k]) < PIN
        assert rel_err(so.rnea_grad(q, qd, qdd, USE_VELOCITY_DAMPING=True), g["dc_du_damped"][k]) < PIN
        assert rel_err(so.rnea_grad(q, qd), g["dc_du_noqdd"][k]) < PIN
        assert rel_err(so.minv(q), g["Minv"][k]) < PIN
        assert rel_err(so.minv(q, output_dense=False), g["Minv_sparse"][k]) < PIN
        assert rel_err(so.forward_dynamics(q, qd, g["u"][k]), g["fd_qdd"][k]) < PIN
        r1, r2 = so.forward_dynamics_grad(q, qd, g["u"][k])
        assert rel_err(r1, g["fd_dq"][k]) < 10 * PIN and rel_err(r2, g["fd_dqd"][k]) < 10 * PIN


@pytest.mark.parametrize("name", FB_CASES)
def test_floating_base_pass_oracle_vs_reference_golden(name):
    """The eight floating-base per-pass restatements against arrays of the unmodified reference, including the
    in-place behaviour (f, Minv and F, df are updated through the caller's arrays)."""
    from oracle.rbd_oracle_fb import FloatingScalarOracle
    rb = make_fb_robot(name)
    so = FloatingScalarOracle(rb)
    g = load_fbpass_golden(name)
    tol = 1e-11
    for k in range(g["q"].shape[0]):
        q, qd, qdd = g["q"][k], g["qd"][k], g["qdd"][k]
        v, a, f = so.rnea_fpass(q, qd, qdd)
        for got, key in ((v, "v"), (a, "a"), (f, "f")):
            assert rel_err(got, g[key][k]) < tol, key
        fa = g["f"][k].copy()
        c, fr = so.rnea_bpass(q, fa)
        assert fr is fa and rel_err(c, g["c"][k]) < tol and rel_err(fa, g["f_acc"][k]) < tol
        Mb, Fb, U, D = so.minv_bpass(q)
        for got, key in ((Mb, "Minv_b"), (Fb, "F_b"), (U, "U"), (D, "Dinv")):
            assert rel_err(got, g[key][k]) < tol, key
        Mf, Ff = g["Minv_b"][k].copy(), g["F_b"][k].copy()
        assert so.minv_fpass(q, Mf, Ff, g["U"][k], g["Dinv"][k]) is Mf
        assert rel_err(Mf, g["Minv_f"][k]) < tol and rel_err(Ff, g["F_f"][k]) < tol
        dv, da, df = so.rnea_grad_fpass_dq(q, qd, g["v"][k], g["a"][k])
        for got, key in ((dv, "dv_dq"), (da, "da_dq"), (df, "df_dq")):
            assert rel_err(got, g[key][k]) < tol, key
        dv, da, df = so.rnea_grad_fpass_dqd(q, qd, g["v"][k])
        for got, key in ((dv, "dv_dqd"), (da, "da_dqd"), (df, "df_dqd")):
            assert rel_err(got, g[key][k]) < tol, key
        dfa = g["df_dq"][k].copy()
        assert rel_err(so.rnea_grad_bpass_dq(q, g["f_acc"][k], dfa), g["dc_dq"][k]) < tol
        assert rel_err(dfa, g["df_dq_acc"][k]) < tol
        dfa = g["df_dqd"][k].copy()
        assert rel_err(so.rnea_grad_bpass_dqd(q, dfa), g["dc_dqd"][k]) < tol and rel_err(dfa, g["df_dqd_acc"][k]) < tol
        assert rel_err(so.rnea_grad_bpass_dqd(q, g["df_dqd"][k].copy(), USE_VELOCITY_DAMPING=True), g["dc_dqd_damped"][k]) < tol


def test_floating_base_identities():
    """Minv inverts the mass matrix assembled from rnea columns; dc_dqd and the joint columns of
    dc_dq match central differences of rnea (the base columns of dc_dq are derivatives along base
    twists, not along q[0:7])."""
    from oracle.rbd_oracle_fb import FloatingScalarOracle
    rb = make_fb_robot("hyq")
    so = FloatingScalarOracle(rb)
    nv = so.n
    q, qd, qdd = rb.random_state(np.random.default_rng(3))
    c0 = so.rnea(q, 0 * qd, 0 * qdd, GRAVITY=0.0)[0]
    M = np.stack([so.rnea(q, 0 * qd, np.eye(nv)[j], GRAVITY=0.0)[0] - c0 for j in range(nv)], axis=1)
    assert np.max(np.abs(M - M.T)) < 1e-12
    assert np.max(np.abs(so.minv(q) @ M - np.eye(nv))) < 1e-11
    dc = so.rnea_grad(q, qd, qdd)
    h, scale = 1e-6, np.max(np.abs(dc))
    for j in range(nv):
        dv = np.zeros(nv); dv[j] = h
        num = (so.rnea(q, qd + dv, qdd)[0] - so.rnea(q, qd - dv, qdd)[0]) / (2 * h)
        assert np.max(np.abs(num - dc[:, nv + j])) / scale < 1e-8
        if j >= 6:
            dq = np.zeros(nv + 1); dq[j + 1] = h
            num = (so.rnea(q + dq, qd, qdd)[0] - so.rnea(q - dq, qd, qdd)[0]) / (2 * h)
            assert np.max(np.abs(num - dc[:, j])) / scale < 1e-8


@pytest.mark.parametrize("name", ["iiwa14", "hyq", "atlas"])
def test_identities(name):
    """SURVEY.md section 4 identity table, on the oracle alone."""
    rb = make_robot(name)
    bo = BatchOracle(rb)
    n = rb.get_num_vel()
    q, qd, qdd = random_states(n, 16, seed=7)
    M, H = bo.minv(q), bo.crba(q)
    assert np.max(np.abs(M @ H - np.eye(n))) < 1e-11
    assert np.max(np.abs(M - np.swapaxes(M, 1, 2))) == 0.0
    c = bo.rnea(q, qd, qdd)[0]
    c0 = bo.rnea(q, qd, np.zeros_like(qdd))[0]
    assert rel_err(np.einsum("bij,bj->bi", H, qdd) + c0, c) < 1e-12
    assert np.array_equal(bo.rnea(q, qd)[0], c0)
    assert np.max(np.abs(bo.rnea(q, 0 * qd, 0 * qdd, GRAVITY=0.0)[0])) == 0.0
    # gradient vs central differences of rnea
    dc = bo.rnea_grad(q[:4], qd[:4], qdd[:4])
    eps = 1e-6
    for j in range(n):
        dq = np.zeros(n); dq[j] = eps
        num_q = (bo.rnea(q[:4] + dq, qd[:4], qdd[:4])[0] - bo.rnea(q[:4] - dq, qd[:4], qdd[:4])[0]) / (2 * eps)
        num_d = (bo.rnea(q[:4], qd[:4] + dq, qdd[:4])[0] - bo.rnea(q[:4], qd[:4] - dq, qdd[:4])[0]) / (2 * eps)
        scale = max(1.0, np.max(np.abs(dc)))
        assert np.max(np.abs(num_q - dc[:, :, j])) / scale < 1e-7
        assert np.max(np.abs(num_d - dc[:, :, n + j])) / scale < 1e-7
    # structural zeros: dc entries vanish unless i, j lie on one root-to-leaf branch
    for i in range(n):
        for j in range(n):
            on_branch = (i in rb.get_subtree_by_id(j)) or (j in rb.get_subtree_by_id(i))
            if not on_branch:
                assert np.all(dc[:, i, j] == 0) and np.all(dc[:, i, n + j] == 0)


def test_staged_reference_agrees_when_present():
    """If oracle/_ref holds the byte-compiled reference, check the oracle against it live."""
    Ref = build_ref.load_reference()
    if Ref is None:
        pytest.skip("oracle/_ref not staged (run `python oracle/build_ref.py` in the build container)")
    rb = make_robot("hyq")
    ref, so = Ref(rb), ScalarOracle(rb)
    q, qd, qdd = random_states(12, 3, seed=11)
    for k in range(3):
        assert rel_err(so.rnea_grad(q[k], qd[k], qdd[k]), ref.rnea_grad(q[k], qd[k], qdd[k])) < PIN
        assert rel_err(so.minv(q[k]), ref.minv(q[k])) < PIN
        assert rel_err(so.rnea(q[k], qd[k], qdd[k])[0], ref.rnea(q[k], qd[k], qdd[k])[0]) < PIN


@pytest.mark.parametrize("name", ["iiwa14", "atlas"])
def test_batch_oracle_vs_wide_reference_golden(name):
    """256 (iiwa14) / 24 (Atlas) states of the unmodified reference's fused drivers: per-tensor and per-row bars."""
    import os
    from oracle.rbd_oracle import BatchOracle
    g = np.load(os.path.join(GOLDEN_DIR, "wide_" + name + ".npz"))
    bo = BatchOracle(make_robot(name))
    q, qd, qdd = g["q"], g["qd"], g["qdd"]
    for got, key in ((bo.rnea(q, qd, qdd)[0], "c"), (bo.rnea_grad(q, qd, qdd), "dc_du"), (bo.minv(q), "Minv")):
        assert rel_err(got, g[key]) < 1e-11, key
        assert row_scaled_err(got, g[key]) < 1e-9, key
