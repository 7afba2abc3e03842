"""Generate tests/golden/*.npz from the UNMODIFIED reference (run in the build container).

    python oracle/make_golden.py            # needs /root/reference/RBDReference.py

The reference ships no golden vectors, so parity is pinned on the
outputs of the reference itself: for each robot a handful of seeded states is pushed
through every hot-path function of /root/reference/RBDReference.py and the results are
stored.  The GPU box has no /root/reference; tests there read only these files.
TEST INFRASTRUCTURE - not imported by the product package.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root")

from reference.RBDReference import RBDReference  # noqa: E402  (the live, unmodified reference)
from rbdreference_b200 import robots  # noqa: E402

CASES = [
    ("iiwa14", lambda: robots.iiwa14(), 8),
    ("hyq", lambda: robots.hyq(), 6),
    ("atlas", lambda: robots.atlas(), 3),
    ("tree9", lambda: robots.random_tree(9, seed=1), 6),
    ("tree13", lambda: robots.random_tree(13, seed=2, branching=0.5, prismatic=0.3), 4),
]


def run_case(name, make, B, seed):
    rb = make()
    ref = RBDReference(rb)
    n = rb.get_num_vel()
    rng = np.random.default_rng(seed)
    q = rng.uniform(-np.pi, np.pi, (B, n))
    qd = rng.uniform(-1.0, 1.0, (B, n))
    qdd = rng.uniform(-1.0, 1.0, (B, n))
    u = rng.uniform(-10.0, 10.0, (B, n))
    out = dict(q=q, qd=qd, qdd=qdd, u=u, gravity_alt=np.array(-3.7))
    keys = ["c", "v", "a", "f", "c_noqdd", "a_noqdd", "f_noqdd", "c_galt", "f_fpass", "dc_du", "dc_du_damped",
            "dc_du_noqdd", "dv_dq", "da_dq", "df_dq", "dv_dqd", "da_dqd", "df_dqd", "dc_dq", "dc_dqd",
            "df_dq_acc", "df_dqd_acc", "Minv", "Minv_sparse", "Minv_b", "F_b", "U", "Dinv", "F_f", "H",
            "fd_qdd", "fd_dq", "fd_dqd", "aba_qdd", "aba_qdd_galt"]
    acc = {k: [] for k in keys}
    for k in range(B):
        v, a, f = ref.rnea_fpass(q[k], qd[k], qdd[k])
        acc["f_fpass"].append(f.copy())
        c, v, a, f = ref.rnea(q[k], qd[k], qdd[k])
        for key, val in (("c", c), ("v", v), ("a", a), ("f", f)):
            acc[key].append(val.copy())
        c0, _, a0, f0 = ref.rnea(q[k], qd[k])
        acc["c_noqdd"].append(c0); acc["a_noqdd"].append(a0); acc["f_noqdd"].append(f0)
        acc["c_galt"].append(ref.rnea(q[k], qd[k], qdd[k], GRAVITY=-3.7)[0])
        acc["dc_du"].append(ref.rnea_grad(q[k], qd[k], qdd[k]))
        acc["dc_du_damped"].append(ref.rnea_grad(q[k], qd[k], qdd[k], USE_VELOCITY_DAMPING=True))
        acc["dc_du_noqdd"].append(ref.rnea_grad(q[k], qd[k]))
        dv, da, df = ref.rnea_grad_fpass_dq(q[k], qd[k], v, a)
        acc["dv_dq"].append(dv.copy()); acc["da_dq"].append(da.copy()); acc["df_dq"].append(df.copy())
        dv2, da2, df2 = ref.rnea_grad_fpass_dqd(q[k], qd[k], v)
        acc["dv_dqd"].append(dv2.copy()); acc["da_dqd"].append(da2.copy()); acc["df_dqd"].append(df2.copy())
        acc["dc_dq"].append(ref.rnea_grad_bpass_dq(q[k], f, df))      # mutates df
        acc["dc_dqd"].append(ref.rnea_grad_bpass_dqd(q[k], df2))      # mutates df2
        acc["df_dq_acc"].append(df.copy()); acc["df_dqd_acc"].append(df2.copy())
        acc["Minv"].append(ref.minv(q[k]))
        acc["Minv_sparse"].append(ref.minv(q[k], output_dense=False))
        Mb, Fb, U, D = ref.minv_bpass(q[k])
        acc["Minv_b"].append(Mb.copy()); acc["F_b"].append(Fb.copy()); acc["U"].append(U.copy()); acc["Dinv"].append(D.copy())
        ref.minv_fpass(q[k], Mb, Fb, U, D)
        acc["F_f"].append(Fb.copy())
        acc["H"].append(ref.crba(q[k]))
        acc["fd_qdd"].append(ref.forward_dynamics(q[k], qd[k], u[k]))
        fdq, fdqd = ref.forward_dynamics_grad(q[k], qd[k], u[k])
        acc["fd_dq"].append(fdq); acc["fd_dqd"].append(fdqd)
        acc["aba_qdd"].append(ref.aba(q[k], qd[k], u[k]))
        acc["aba_qdd_galt"].append(ref.aba(q[k], qd[k], u[k], GRAVITY=-3.7))
    for key in keys:
        out[key] = np.stack(acc[key])
    # model tables, to detect drift of rbdreference_b200/robots.py against the fixture
    out["parent"] = np.array([rb.get_parent_id(i) for i in range(n)])
    out["S"] = np.stack([np.asarray(rb.get_S_by_id(i), float).reshape(-1) for i in range(n)])
    out["I"] = np.stack([rb.get_Imat_by_id(i) for i in range(n)])
    out["X_at_0p3"] = np.stack([rb.get_Xmat_Func_by_id(i)(0.3) for i in range(n)])
    path = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(path, **out)
    print("wrote %s (%d states, %.1f KB)" % (path, B, os.path.getsize(path) / 1024))


# end-effector kinematics (RBDReference.py:190-386): which end effectors each golden case requests
EE_SPECS = {
    "iiwa14": [None, ["iiwa_joint_ee", "iiwa_joint_4", "iiwa_tool_tip"]],
    "hyq": [None, ["rh_foot_joint", "lf_hfe_joint"]],
    "atlas": [None, ["l_hand_mount", "back_bkx", "head_camera", "r_leg_akx"]],
    "tree9": [None],
    "tree13": [None, ["j5", "j12"]],
}
EE_OFFSETS = [(0.0, 0.0, 0.0, 1.0), (0.1, -0.2, 0.3, 1.0)]


def run_ee_case(name, make, B, seed):
    """tests/golden/ee_<name>.npz: end_effector_pose / _gradient of the unmodified reference."""
    import json
    import warnings
    warnings.simplefilter("ignore")          # the reference builds np.matrix objects
    rb = make()
    ref = RBDReference(rb)
    n = rb.get_num_vel()
    q = np.random.default_rng(seed).uniform(-np.pi, np.pi, (B, n))
    out = dict(q=q, specs=np.array(json.dumps(EE_SPECS[name])), offsets=np.array(EE_OFFSETS))
    for si, names in enumerate(EE_SPECS[name]):
        for oi, off in enumerate(EE_OFFSETS):
            offs = [np.matrix([list(off)])]
            pose = np.stack([np.stack([np.asarray(p)[:, 0] for p in ref.end_effector_pose(q[k], names, offs)])
                             for k in range(B)])
            grad = np.stack([np.stack([np.asarray(g) for g in ref.end_effector_pose_gradient(q[k], names, offs)])
                             for k in range(B)])
            out["pose_s%d_o%d" % (si, oi)] = pose
            out["grad_s%d_o%d" % (si, oi)] = grad
    path = os.path.join(ROOT, "tests", "golden", "ee_" + name + ".npz")
    np.savez_compressed(path, **out)
    print("wrote %s (%d states, %.1f KB)" % (path, B, os.path.getsize(path) / 1024))


def run_wide_case(name, make, B, seed):
    """tests/golden/wide_<name>.npz: the fused drivers' outputs (c, dc_du, Minv) of the unmodified reference on MANY
    states - breadth for the headline operations, next to the depth (every intermediate) of <name>.npz."""
    rb = make()
    ref = RBDReference(rb)
    n = rb.get_num_vel()
    rng = np.random.default_rng(seed)
    q = rng.uniform(-np.pi, np.pi, (B, n))
    qd = rng.uniform(-1.0, 1.0, (B, n))
    qdd = rng.uniform(-1.0, 1.0, (B, n))
    c = np.stack([ref.rnea(q[k], qd[k], qdd[k])[0] for k in range(B)])
    dc = np.stack([ref.rnea_grad(q[k], qd[k], qdd[k]) for k in range(B)])
    M = np.stack([ref.minv(q[k]) for k in range(B)])
    path = os.path.join(ROOT, "tests", "golden", "wide_" + name + ".npz")
    np.savez_compressed(path, q=q, qd=qd, qdd=qdd, c=c, dc_du=dc, Minv=M)
    print("wrote %s (%d states, %.1f KB)" % (path, B, os.path.getsize(path) / 1024))


# floating-base branches: the wrapped robots of rbdreference_b200.robots
FB_CASES = [("hyq_fb", 5), ("atlas_fb", 3), ("iiwa14_fb", 5), ("tree9_fb", 4)]


def make_fb_robot(name):
    if name == "tree9_fb":
        return robots.FloatingBaseRobot(robots.random_tree(9, seed=1), name="tree9_fb")
    return robots.by_name(name)


def run_fb_case(name, B, seed):
    """tests/golden/fb_<name>.npz: rnea / rnea_grad / minv of the unmodified reference, floating base."""
    import warnings
    warnings.simplefilter("ignore")
    rb = make_fb_robot(name)
    ref = RBDReference(rb)
    q, qd, qdd = rb.random_state(np.random.default_rng(seed), B)
    keys = ["c", "v", "a", "f", "c_noqdd", "c_galt", "dc_du", "dc_du_damped", "dc_du_noqdd", "Minv", "Minv_sparse",
            "fd_qdd", "fd_dq", "fd_dqd"]
    acc = {k: [] for k in keys}
    u = np.random.default_rng(seed + 500).uniform(-10.0, 10.0, qd.shape)
    for k in range(B):
        acc["fd_qdd"].append(ref.forward_dynamics(q[k], qd[k], u[k]))
        fdq, fdqd = ref.forward_dynamics_grad(q[k], qd[k], u[k])
        acc["fd_dq"].append(fdq); acc["fd_dqd"].append(fdqd)
        c, v, a, f = ref.rnea(q[k], qd[k], qdd[k])
        for key, val in (("c", c), ("v", v), ("a", a), ("f", f)):
            acc[key].append(np.array(val))
        acc["c_noqdd"].append(ref.rnea(q[k], qd[k])[0])
        acc["c_galt"].append(ref.rnea(q[k], qd[k], qdd[k], GRAVITY=-3.7)[0])
        acc["dc_du"].append(ref.rnea_grad(q[k], qd[k], qdd[k]))
        acc["dc_du_damped"].append(ref.rnea_grad(q[k], qd[k], qdd[k], USE_VELOCITY_DAMPING=True))
        acc["dc_du_noqdd"].append(ref.rnea_grad(q[k], qd[k]))
        acc["Minv"].append(ref.minv(q[k]))
        acc["Minv_sparse"].append(ref.minv(q[k], output_dense=False))
    out = dict(q=q, qd=qd, qdd=qdd, u=u, X0_first=rb.get_Xmat_Func_by_id(0)(q[0, 0:7]))
    for key in keys:
        out[key] = np.stack(acc[key])
    path = os.path.join(ROOT, "tests", "golden", "fb_" + name[:-3] + ".npz")
    np.savez_compressed(path, **out)
    print("wrote %s (%d states, %.1f KB)" % (path, B, os.path.getsize(path) / 1024))


# per-pass intermediates of the floating-base branches (the eight helpers)
FBPASS_CASES = [("hyq_fb", 3), ("atlas_fb", 1), ("iiwa14_fb", 3), ("tree9_fb", 2)]


def run_fbpass_case(name, B, seed):
    """tests/golden/fbpass_<name>.npz: every array the unmodified reference's eight per-pass helpers take and return
    for floating-base robots (copies are taken around the calls that mutate their arguments)."""
    import warnings
    warnings.simplefilter("ignore")
    rb = make_fb_robot(name)
    ref = RBDReference(rb)
    q, qd, qdd = rb.random_state(np.random.default_rng(seed), B)
    keys = ["v", "a", "f", "c", "f_acc", "Minv_b", "F_b", "U", "Dinv", "Minv_f", "F_f", "dv_dq", "da_dq", "df_dq", "dv_dqd",
            "da_dqd", "df_dqd", "dc_dq", "df_dq_acc", "dc_dqd", "dc_dqd_damped", "df_dqd_acc"]
    acc = {k: [] for k in keys}
    for k in range(B):
        v, a, f = ref.rnea_fpass(q[k], qd[k], qdd[k])
        acc["v"].append(v.copy()); acc["a"].append(a.copy()); acc["f"].append(f.copy())
        c, f_acc = ref.rnea_bpass(q[k], f.copy())
        acc["c"].append(c); acc["f_acc"].append(f_acc.copy())
        Mb, Fb, U, D = ref.minv_bpass(q[k])
        acc["Minv_b"].append(Mb.copy()); acc["F_b"].append(Fb.copy()); acc["U"].append(U.copy()); acc["Dinv"].append(D.copy())
        Mf, Ff = Mb.copy(), Fb.copy()
        ref.minv_fpass(q[k], Mf, Ff, U, D)
        acc["Minv_f"].append(Mf); acc["F_f"].append(Ff)
        dv, da, df = ref.rnea_grad_fpass_dq(q[k], qd[k], v, a)
        acc["dv_dq"].append(dv.copy()); acc["da_dq"].append(da.copy()); acc["df_dq"].append(df.copy())
        dv2, da2, df2 = ref.rnea_grad_fpass_dqd(q[k], qd[k], v)
        acc["dv_dqd"].append(dv2.copy()); acc["da_dqd"].append(da2.copy()); acc["df_dqd"].append(df2.copy())
        dfa = df.copy()
        acc["dc_dq"].append(ref.rnea_grad_bpass_dq(q[k], f_acc, dfa)); acc["df_dq_acc"].append(dfa)
        dfa2 = df2.copy()
        acc["dc_dqd"].append(ref.rnea_grad_bpass_dqd(q[k], dfa2)); acc["df_dqd_acc"].append(dfa2)
        acc["dc_dqd_damped"].append(ref.rnea_grad_bpass_dqd(q[k], df2.copy(), USE_VELOCITY_DAMPING=True))
    out = dict(q=q, qd=qd, qdd=qdd)
    for key in keys:
        out[key] = np.stack([np.asarray(x, dtype=float) for x in acc[key]])
    path = os.path.join(ROOT, "tests", "golden", "fbpass_" + name[:-3] + ".npz")
    np.savez_compressed(path, **out)
    print("wrote %s (%d states, %.1f KB)" % (path, B, os.path.getsize(path) / 1024))


if __name__ == "__main__":
    only_ee = "--ee" in sys.argv          # regenerate only the end-effector fixtures
    only_fb = "--fb" in sys.argv          # regenerate only the floating-base fixtures
    if "--wide" in sys.argv:              # only the many-state fixtures of the fused drivers
        run_wide_case("iiwa14", lambda: robots.iiwa14(), 256, seed=5000)
        run_wide_case("atlas", lambda: robots.atlas(), 24, seed=5001)
        sys.exit(0)
    if "--fbpass" in sys.argv:            # only the floating-base per-pass fixtures
        for idx, (name, B) in enumerate(FBPASS_CASES):
            run_fbpass_case(name, B, seed=4000 + idx)
        sys.exit(0)
    for idx, (name, make, B) in enumerate(CASES):
        if not (only_ee or only_fb):
            run_case(name, make, B, seed=1000 + idx)
        if not only_fb:
            run_ee_case(name, make, B, seed=2000 + idx)
    if not only_ee:
        for idx, (name, B) in enumerate(FB_CASES):
            run_fb_case(name, B, seed=3000 + idx)
        for idx, (name, B) in enumerate(FBPASS_CASES):
            run_fbpass_case(name, B, seed=4000 + idx)
