"""Fixed-base robot objects with the URDFParser getter surface RBDReference consumes.

The reference takes "an instance of Robot Object class created by URDFparser"
(/root/reference/RBDReference.py:7, README.md:8).  URDFParser is an external,
un-vendored package and there are no URDF files in this environment, so this module
provides a small stand-in `Robot` that answers the getters the hot path calls
(SURVEY.md section 8b lists every call site) plus builders for the three benchmark
topologies named in BASELINE.json:

* `iiwa14()`  - 7-DoF serial arm,
* `hyq()`     - 12-DoF quadruped, four 3-DoF legs rooted at the fixed trunk,
* `atlas()`   - 30-DoF humanoid tree (3 roots at the fixed pelvis, depth 10),
* `random_tree(n, seed)` - random topology / axes / inertias for property tests.

Kinematic and inertial numbers are written from public knowledge of the URDFs and are
labelled synthetic: parity never depends on them (engine and checker consume the same
object) and throughput depends only on the topology.

Conventions (Featherstone, spatial_v2): spatial vectors are [angular; linear];
`Xmat(q)` maps parent-frame motion vectors to child-frame coordinates;
X(q) = XJ(q) @ Xtree with Xtree = xrot(E) @ xlt(r).
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Sequence

import numpy as np

__all__ = ["Robot", "FloatingBaseRobot", "JointSpec", "FixedJoint", "JointHandle", "quat_rotation", "iiwa14", "hyq", "atlas", "random_tree", "by_name"]


# ----------------------------------------------------------------------------------------
# spatial algebra building blocks (host side, float64)
# ----------------------------------------------------------------------------------------
def skew(r: Sequence[float]) -> np.ndarray:
    x, y, z = (float(t) for t in r)
    return np.array([[0.0, -z, y], [z, 0.0, -x], [-y, x, 0.0]])


def rot_axis(axis: Sequence[float], theta: float) -> np.ndarray:
    """Coordinate transform E for a rotation of `theta` about unit `axis` (= R^T)."""
    a = np.asarray(axis, dtype=float)
    a = a / np.linalg.norm(a)
    K = skew(a)
    R = np.eye(3) + math.sin(theta) * K + (1.0 - math.cos(theta)) * (K @ K)
    return R.T


def rpy_transform(rpy: Sequence[float]) -> np.ndarray:
    """E = rx(roll) @ ry(pitch) @ rz(yaw) = (Rz Ry Rx)^T, the URDF origin rotation."""
    r, p, y = (float(t) for t in rpy)
    return rot_axis([1, 0, 0], r) @ rot_axis([0, 1, 0], p) @ rot_axis([0, 0, 1], y)


def xrot(E: np.ndarray) -> np.ndarray:
    X = np.zeros((6, 6))
    X[:3, :3] = E
    X[3:, 3:] = E
    return X


def xlt(r: Sequence[float]) -> np.ndarray:
    X = np.eye(6)
    X[3:, :3] = -skew(r)
    return X


def mcI(m: float, c: Sequence[float], Ic: np.ndarray) -> np.ndarray:
    """Spatial inertia from mass, centre of mass and rotational inertia about the COM."""
    C = skew(c)
    I = np.zeros((6, 6))
    I[:3, :3] = np.asarray(Ic, dtype=float) + m * (C @ C.T)
    I[:3, 3:] = m * C
    I[3:, :3] = m * C.T
    I[3:, 3:] = m * np.eye(3)
    return I


# ----------------------------------------------------------------------------------------
# robot object
# ----------------------------------------------------------------------------------------
class JointSpec:
    """One 1-DoF joint and the body it moves."""

    def __init__(self, name, parent, kind, axis, xyz, rpy, mass, com, inertia_diag,
                 inertia_offdiag=(0.0, 0.0, 0.0), damping=0.0):
        self.name = name
        self.parent = int(parent)
        self.kind = kind  # "revolute" | "prismatic"
        self.axis = np.asarray(axis, dtype=float) / np.linalg.norm(np.asarray(axis, dtype=float))
        self.xyz = np.asarray(xyz, dtype=float)
        self.rpy = np.asarray(rpy, dtype=float)
        self.mass = float(mass)
        self.com = np.asarray(com, dtype=float)
        ixx, iyy, izz = inertia_diag
        ixy, ixz, iyz = inertia_offdiag
        self.Ic = np.array([[ixx, ixy, ixz], [ixy, iyy, iyz], [ixz, iyz, izz]], dtype=float)
        self.damping = float(damping)


class JointHandle:
    """What `robot.get_joint_by_name` hands back (RBDReference.py:204-206 only calls get_id)."""

    def __init__(self, jid: int, name: str):
        self._id = int(jid)
        self.name = name

    def get_id(self) -> int:
        return self._id


class FixedJoint:
    """A fixed joint hanging off the child link of a moving joint (e.g. a tool flange):
    `parent_name` names that moving joint, the transform maps flange-frame points to its frame
    (RBDReference.py:277-280, :374-386)."""

    def __init__(self, name, parent_name, xyz=(0.0, 0.0, 0.0), rpy=(0.0, 0.0, 0.0)):
        self.name = name
        self.parent_name = parent_name
        self.xyz = np.asarray(xyz, dtype=float)
        self.rpy = np.asarray(rpy, dtype=float)
        self._id = -1

    def get_id(self) -> int:
        return self._id

    def get_transformation_matrix_hom(self) -> np.ndarray:
        T = np.eye(4)
        T[:3, :3] = rpy_transform(self.rpy).T
        T[:3, 3] = self.xyz
        return T


class Robot:
    """Duck-typed stand-in for a URDFParser robot (fixed base, 1-DoF joints).

    Getter names and return conventions follow the call sites in
    /root/reference/RBDReference.py (e.g. :570-574, :595, :662, :666, :1339).
    """

    floating_base = False

    def __init__(self, name: str, joints: List[JointSpec], fixed_joints=None):
        self.name = name
        self.joints = joints
        n = len(joints)
        self._n = n
        self._parent = [j.parent for j in joints]
        for i, p in enumerate(self._parent):
            if not (-1 <= p < i):
                raise ValueError("bodies must be topologically ordered (parent id < child id)")
        self._S = []
        self._Xtree = []
        self._I = []
        for j in joints:
            S = np.zeros(6)
            if j.kind == "revolute":
                S[:3] = j.axis
            elif j.kind == "prismatic":
                S[3:] = j.axis
            else:
                raise ValueError("joint kind must be revolute or prismatic")
            self._S.append(S)
            self._Xtree.append(xrot(rpy_transform(j.rpy)) @ xlt(j.xyz))
            self._I.append(mcI(j.mass, j.com, j.Ic))
        self._children = [[] for _ in range(n)]
        for i, p in enumerate(self._parent):
            if p >= 0:
                self._children[p].append(i)
        self._subtree = [self._collect_subtree(i) for i in range(n)]
        self._Xfuncs = [self._make_xfunc(i) for i in range(n)]
        self._hom_funcs = [[self._make_hom_func(i, order) for i in range(n)] for order in range(3)]
        self._joint_handles = [JointHandle(i, j.name) for i, j in enumerate(joints)]
        self._fixed_joints = []
        for fid, fj in enumerate(fixed_joints or []):
            if fj.parent_name not in [j.name for j in joints]:
                raise ValueError("fixed joint %r: unknown parent joint %r" % (fj.name, fj.parent_name))
            fj._id = fid
            self._fixed_joints.append(fj)

    # -- construction helpers -------------------------------------------------------------
    def _collect_subtree(self, i: int) -> List[int]:
        out = [i]
        for c in self._children[i]:
            out.extend(self._collect_subtree(c))
        return sorted(out)

    def _make_xfunc(self, i: int) -> Callable[[float], np.ndarray]:
        j = self.joints[i]
        Xtree = self._Xtree[i]
        axis = j.axis
        if j.kind == "revolute":
            def xfunc(q, _axis=axis, _Xtree=Xtree):
                return xrot(rot_axis(_axis, float(q))) @ _Xtree
        else:
            def xfunc(q, _axis=axis, _Xtree=Xtree):
                return xlt(_axis * float(q)) @ _Xtree
        return xfunc

    def _make_hom_func(self, i: int, order: int) -> Callable[[float], np.ndarray]:
        """4x4 homogeneous transform of joint i (child-frame point -> parent-frame point) or its
        first / second derivative in q: the point-transform counterpart of `Xmat(q)`."""
        j = self.joints[i]
        Rt = rpy_transform(j.rpy).T            # Etree^T
        K = skew(j.axis)
        K2 = K @ K
        r = j.xyz
        if j.kind == "revolute":
            def hom(q, _Rt=Rt, _K=K, _K2=K2, _r=r, _o=order):
                s, c = math.sin(float(q)), math.cos(float(q))
                T = np.zeros((4, 4))
                if _o == 0:
                    T[:3, :3] = _Rt @ (np.eye(3) + s * _K + (1.0 - c) * _K2)
                    T[:3, 3] = _r
                    T[3, 3] = 1.0
                elif _o == 1:
                    T[:3, :3] = _Rt @ (c * _K + s * _K2)
                else:
                    T[:3, :3] = _Rt @ (-s * _K + c * _K2)
                return T
        else:
            d = Rt @ j.axis
            def hom(q, _Rt=Rt, _d=d, _r=r, _o=order):
                T = np.zeros((4, 4))
                if _o == 0:
                    T[:3, :3] = _Rt
                    T[:3, 3] = _r + _d * float(q)
                    T[3, 3] = 1.0
                elif _o == 1:
                    T[:3, 3] = _d
                return T
        return hom

    # -- URDFParser getter surface -----------------------------------------
    def get_num_bodies(self) -> int:
        return self._n

    def get_num_vel(self) -> int:
        return self._n

    def get_num_pos(self) -> int:
        return self._n

    def get_num_joints(self) -> int:
        return self._n

    def get_parent_id(self, i: int) -> int:
        return self._parent[i]

    def get_parent_id_array(self) -> List[int]:
        return list(self._parent)

    def get_S_by_id(self, i: int) -> np.ndarray:
        return self._S[i].copy()

    def get_joint_index_q(self, i: int) -> int:
        return i

    def get_joint_index_v(self, i: int) -> int:
        return i

    def get_joint_index_f(self, i: int) -> int:
        return i

    def get_Xmat_Func_by_id(self, i: int) -> Callable[[float], np.ndarray]:
        return self._Xfuncs[i]

    def get_Imat_by_id(self, i: int) -> np.ndarray:
        return self._I[i].copy()

    def get_Imats_dict_by_id(self) -> Dict[int, np.ndarray]:
        return {i: self._I[i].copy() for i in range(self._n)}

    def get_subtree_by_id(self, i: int) -> List[int]:
        return list(self._subtree[i])

    def get_ancestors_by_id(self, i: int) -> List[int]:
        out = []
        p = self._parent[i]
        while p != -1:
            out.append(p)
            p = self._parent[p]
        return out

    def get_damping_by_id(self, i: int) -> float:
        return self.joints[i].damping

    def get_joint_names(self) -> List[str]:
        return [j.name for j in self.joints]

    # -- getters used by the end-effector kinematics (RBDReference.py:190-386) -------------
    def get_Xmat_hom_Func_by_id(self, i: int) -> Callable[[float], np.ndarray]:
        return self._hom_funcs[0][i]

    def get_dXmat_hom_Func_by_id(self, i: int) -> Callable[[float], np.ndarray]:
        return self._hom_funcs[1][i]

    def get_d2Xmat_hom_Func_by_id(self, i: int) -> Callable[[float], np.ndarray]:
        return self._hom_funcs[2][i]

    def get_leaf_nodes(self) -> List[int]:
        return [i for i in range(self._n) if not self._children[i]]

    def get_joint_by_name(self, name: str):
        for h in self._joint_handles:
            if h.name == name:
                return h
        return None

    def get_fixed_joint_by_name(self, name: str):
        for fj in self._fixed_joints:
            if fj.name == name:
                return fj
        return None

    def get_fixed_joint_by_id(self, fid: int):
        return self._fixed_joints[fid]


# ----------------------------------------------------------------------------------------
# benchmark robots
# ----------------------------------------------------------------------------------------
_PI = math.pi


def iiwa14() -> Robot:
    """KUKA LBR iiwa14: 7 revolute-z joints in series (synthetic-from-memory numbers)."""
    rows = [
        # xyz,                 rpy,                 mass, com,                     Ixx,Iyy,Izz
        ((0, 0, 0.1575),      (0, 0, 0),           5.76, (0, -0.03, 0.12),        (0.033, 0.0333, 0.0123)),
        ((0, 0, 0.2025),      (_PI / 2, 0, _PI),   6.35, (0.0003, 0.059, 0.042),  (0.0305, 0.0304, 0.011)),
        ((0, 0.2045, 0),      (_PI / 2, 0, _PI),   3.5,  (0, 0.03, 0.13),         (0.025, 0.0238, 0.0076)),
        ((0, 0, 0.2155),      (_PI / 2, 0, 0),     3.5,  (0, 0.067, 0.034),       (0.017, 0.0164, 0.006)),
        ((0, 0.1845, 0),      (-_PI / 2, _PI, 0),  3.5,  (0.0001, 0.021, 0.076),  (0.01, 0.0087, 0.00449)),
        ((0, 0, 0.2155),      (_PI / 2, 0, 0),     1.8,  (0, 0.0006, 0.0004),     (0.0049, 0.0047, 0.0036)),
        ((0, 0.081, 0),       (-_PI / 2, _PI, 0),  1.2,  (0, 0, 0.02),            (0.001, 0.001, 0.001)),
    ]
    joints = []
    for i, (xyz, rpy, m, com, idiag) in enumerate(rows):
        joints.append(JointSpec("iiwa_joint_%d" % (i + 1), i - 1, "revolute", (0, 0, 1), xyz, rpy,
                                m, com, idiag, damping=0.5))
    fixed = [FixedJoint("iiwa_joint_ee", "iiwa_joint_7", xyz=(0, 0, 0.045)),
             FixedJoint("iiwa_tool_tip", "iiwa_joint_7", xyz=(0.02, -0.01, 0.16), rpy=(0.1, -0.2, 0.3))]
    return Robot("iiwa14", joints, fixed)


def hyq() -> Robot:
    """HyQ with the trunk fixed: 4 legs x (HAA -> HFE -> KFE), 12 DoF, 4 roots."""
    joints = []
    legs = [("lf", +1, +1), ("rf", +1, -1), ("lh", -1, +1), ("rh", -1, -1)]
    for k, (leg, fx, fy) in enumerate(legs):
        base = 3 * k
        # hip abduction/adduction about the trunk x axis
        joints.append(JointSpec(leg + "_haa_joint", -1, "revolute", (1, 0, 0),
                                (fx * 0.3735, fy * 0.207, 0.0), (0, 0, 0),
                                2.93, (0.04263, 0.0, fy * 0.16931), (0.05071, 0.054530, 0.005571),
                                inertia_offdiag=(0.000036 * fy, -0.000216, 0.000513 * fy), damping=0.1))
        # hip flexion/extension about y
        joints.append(JointSpec(leg + "_hfe_joint", base, "revolute", (0, 1, 0),
                                (0.08, 0.0, 0.0), (0, 0, 0),
                                2.638, (0.15074, -0.02625 * fy, 0.0), (0.003357, 0.027879, 0.028079),
                                inertia_offdiag=(0.000241 * fy, -0.001448, 0.000130 * fy), damping=0.1))
        # knee flexion/extension about y
        joints.append(JointSpec(leg + "_kfe_joint", base + 1, "revolute", (0, 1, 0),
                                (0.35, 0.0, 0.0), (0, 0, 0),
                                0.881, (0.1254, 0.0005 * fy, -0.0001), (0.000468, 0.012237, 0.012029),
                                inertia_offdiag=(0.0, 0.000035, 0.0), damping=0.1))
    fixed = [FixedJoint(leg + "_foot_joint", leg + "_kfe_joint", xyz=(0.33, 0.0, 0.0)) for leg, _, _ in legs]
    return Robot("hyq", joints, fixed)


def atlas() -> Robot:
    """Atlas v5 with the pelvis fixed: 30 DoF, roots {back_bkz, l_leg_hpz, r_leg_hpz}.

    Level widths 3,3,3,5,4,4,2,2,2,2.  Numbers are synthetic,
    shaped after the public atlas_v5 URDF (mixed x/y/z axes).
    """
    J: List[JointSpec] = []

    def add(name, parent, axis, xyz, mass, com, idiag, rpy=(0, 0, 0), ioff=(0.0, 0.0, 0.0)):
        J.append(JointSpec(name, parent, "revolute", axis, xyz, rpy, mass, com, idiag,
                           inertia_offdiag=ioff, damping=0.1))
        return len(J) - 1

    bkz = add("back_bkz", -1, (0, 0, 1), (-0.0125, 0, 0), 2.27, (-0.0113, 0, 0.0757), (0.0039, 0.0034, 0.0017))
    bky = add("back_bky", bkz, (0, 1, 0), (0, 0, 0.162), 0.8, (-0.0082, -0.0131, 0.0306), (0.00045, 0.00069, 0.00083))
    bkx = add("back_bkx", bky, (1, 0, 0), (0, 0, 0.05), 84.4, (-0.0622, 0.0023, 0.3158), (1.577, 1.602, 0.565),
              ioff=(-0.032, 0.102, 0.047))
    add("neck_ry", bkx, (0, 1, 0), (0.2546, 0, 0.6215), 1.42, (-0.075, 0.0003, 0.0262), (0.0039, 0.0042, 0.0047))
    for side, sy in (("l", 1.0), ("r", -1.0)):
        shz = add(side + "_arm_shz", bkx, (0, 0, 1), (0.1406, sy * 0.2256, 0.4776), 3.45,
                  (0.0, sy * -0.0014, 0.0147), (0.0049, 0.0037, 0.0067), ioff=(0.0, 0.0, sy * 0.0003))
        shx = add(side + "_arm_shx", shz, (1, 0, 0), (0, sy * 0.11, -0.245), 3.012,
                  (0.0, sy * -0.0014, 0.0147), (0.0032, 0.0033, 0.0044), ioff=(0.0, 0.0, sy * -0.0002))
        ely = add(side + "_arm_ely", shx, (0, 1, 0), (0, sy * 0.187, -0.016), 3.388,
                  (0.0, sy * -0.0014, 0.0147), (0.0029, 0.0031, 0.0035), ioff=(sy * 0.0001, 0.0, 0.0))
        elx = add(side + "_arm_elx", ely, (1, 0, 0), (0, sy * 0.119, 0.0092), 2.509,
                  (0.0, sy * -0.0014, 0.0147), (0.0035, 0.0018, 0.0031))
        wry = add(side + "_arm_wry", elx, (0, 1, 0), (0, sy * 0.29955, -0.00921), 0.8,
                  (0.0, sy * 0.01, 0.0), (0.00079, 0.00067, 0.00093))
        wrx = add(side + "_arm_wrx", wry, (1, 0, 0), (0, 0, 0), 0.5, (0.0, sy * 0.005, 0.0),
                  (0.0005, 0.0004, 0.0006))
        add(side + "_arm_wry2", wrx, (0, 1, 0), (0, 0, 0), 0.6, (0.0, sy * 0.02, 0.0),
            (0.0006, 0.0005, 0.0007), ioff=(0.0, sy * 0.00002, 0.0))
    for side, sy in (("l", 1.0), ("r", -1.0)):
        hpz = add(side + "_leg_hpz", -1, (0, 0, 1), (0, sy * 0.089, 0), 2.409,
                  (0.0, 0.0, 0.0), (0.00254, 0.00254, 0.00131), ioff=(0.0, 0.0, sy * 0.0001))
        hpx = add(side + "_leg_hpx", hpz, (1, 0, 0), (0, 0, 0), 0.648, (0.0195, 0.0, 0.0264),
                  (0.00074, 0.00074, 0.00054))
        hpy = add(side + "_leg_hpy", hpx, (0, 1, 0), (0.05, sy * 0.0225, -0.066), 9.209,
                  (0.0, 0.0, -0.21), (0.09, 0.09, 0.02), ioff=(0.0, 0.001, sy * 0.001))
        kny = add(side + "_leg_kny", hpy, (0, 1, 0), (-0.05, 0, -0.374), 5.479,
                  (0.001, 0.0, -0.187), (0.077, 0.076, 0.01), ioff=(0.0, -0.003, 0.0))
        aky = add(side + "_leg_aky", kny, (0, 1, 0), (0, 0, -0.422), 0.125, (0.0, 0.0, 0.0),
                  (0.00001, 0.000012, 0.000011))
        add(side + "_leg_akx", aky, (1, 0, 0), (0, 0, 0), 2.05, (0.027, 0.0, -0.067),
            (0.002, 0.007, 0.008), ioff=(0.0, 0.002, 0.0))
    fixed = [FixedJoint("l_hand_mount", "l_arm_wry2", xyz=(0, 0.1, 0), rpy=(0, 0, _PI / 2)),
             FixedJoint("r_hand_mount", "r_arm_wry2", xyz=(0, -0.1, 0), rpy=(0, 0, -_PI / 2)),
             FixedJoint("head_camera", "neck_ry", xyz=(0.1, 0, 0.08))]
    return Robot("atlas", J, fixed)


def random_tree(n: int, seed: int = 0, branching: float = 0.35, prismatic: float = 0.2,
                axis_aligned: float = 0.5) -> Robot:
    """Random fixed-base tree for property tests: random parents, joint kinds, axes, inertias."""
    rng = np.random.default_rng(seed)
    joints = []
    for i in range(n):
        if i == 0:
            parent = -1
        elif rng.random() < branching:
            parent = int(rng.integers(-1, i))
        else:
            parent = i - 1
        kind = "prismatic" if rng.random() < prismatic else "revolute"
        if rng.random() < axis_aligned:
            axis = np.zeros(3)
            axis[int(rng.integers(0, 3))] = 1.0 if rng.random() < 0.5 else -1.0
        else:
            axis = rng.normal(size=3)
        A = rng.normal(size=(3, 3)) * 0.1
        Ic = A @ A.T + 0.01 * np.eye(3)
        joints.append(JointSpec(
            "j%d" % i, parent, kind, axis,
            xyz=rng.uniform(-0.3, 0.3, size=3), rpy=rng.uniform(-_PI, _PI, size=3),
            mass=float(rng.uniform(0.2, 5.0)), com=rng.uniform(-0.1, 0.1, size=3),
            inertia_diag=(Ic[0, 0], Ic[1, 1], Ic[2, 2]),
            inertia_offdiag=(Ic[0, 1], Ic[0, 2], Ic[1, 2]),
            damping=float(rng.uniform(0.0, 1.0))))
    return Robot("random_tree_n%d_s%d" % (n, seed), joints)


def quat_rotation(quat_xyzw: Sequence[float]) -> np.ndarray:
    """Rotation matrix R (body -> world) of a unit quaternion given as (x, y, z, w)."""
    x, y, z, w = (float(t) for t in quat_xyzw)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


class FloatingBaseRobot:
    """A fixed-base `Robot` put on a free-floating base body, with the getter conventions the
    reference's floating-base branches rely on (RBDReference.py:585, :591, :652-691, :761-779,
    :1141-1147, :1212-1218, :1267-1282, :1309-1315):

    * body 0 is the base, joined to the world by one 6-DoF joint with S = eye(6); body i >= 1 is
      body i-1 of the wrapped robot, bodies that were roots now hang off the base;
    * q has NB + 6 entries: q[0:7] = base position (3) and unit quaternion (x, y, z, w), q[i + 6]
      the angle of joint i; qd / qdd / c have NB + 5 entries: [0:6] = base twist in base
      coordinates [angular; linear], [i + 5] = joint i;
    * `get_joint_index_q/v/f(0)` return index lists, `get_Xmat_Func_by_id(0)` takes q[0:7] and
      returns X = xrot(R(quat)^T) @ xlt(position).

    URDFParser is not available here, so these conventions are this stand-in's; the
    engine re-discovers the layout by probing `get_Xmat_Func_by_id(0)` (model.compile_fb_model).
    """

    floating_base = True

    def __init__(self, base: Robot, base_mass=20.0, base_com=(0.01, -0.02, 0.03),
                 base_inertia=(0.9, 1.1, 1.3), base_inertia_offdiag=(0.02, -0.03, 0.01), name=None):
        self.base = base
        self.name = name or (base.name + "_fb")
        self._nb = base.get_num_bodies() + 1
        ixx, iyy, izz = base_inertia
        ixy, ixz, iyz = base_inertia_offdiag
        self._I0 = mcI(base_mass, base_com, np.array([[ixx, ixy, ixz], [ixy, iyy, iyz], [ixz, iyz, izz]]))

    @staticmethod
    def base_transform(q7) -> np.ndarray:
        q7 = np.asarray(q7, dtype=float).reshape(7)
        return xrot(quat_rotation(q7[3:7]).T) @ xlt(q7[0:3])

    def get_num_bodies(self) -> int:
        return self._nb

    def get_num_joints(self) -> int:
        return self._nb

    def get_num_vel(self) -> int:
        return self._nb + 5

    def get_num_pos(self) -> int:
        return self._nb + 6

    def get_parent_id(self, i: int) -> int:
        return -1 if i == 0 else self.base.get_parent_id(i - 1) + 1

    def get_S_by_id(self, i: int) -> np.ndarray:
        return np.eye(6) if i == 0 else self.base.get_S_by_id(i - 1)

    def get_joint_index_q(self, i: int):
        return list(range(7)) if i == 0 else i + 6

    def get_joint_index_v(self, i: int):
        return list(range(6)) if i == 0 else i + 5

    def get_joint_index_f(self, i: int):
        return list(range(6)) if i == 0 else i + 5

    def get_Xmat_Func_by_id(self, i: int):
        return self.base_transform if i == 0 else self.base.get_Xmat_Func_by_id(i - 1)

    def get_Imat_by_id(self, i: int) -> np.ndarray:
        return self._I0.copy() if i == 0 else self.base.get_Imat_by_id(i - 1)

    def get_Imats_dict_by_id(self) -> Dict[int, np.ndarray]:
        return {i: self.get_Imat_by_id(i) for i in range(self._nb)}

    def get_subtree_by_id(self, i: int) -> List[int]:
        if i == 0:
            return list(range(self._nb))
        return [j + 1 for j in self.base.get_subtree_by_id(i - 1)]

    def get_ancestors_by_id(self, i: int) -> List[int]:
        out, p = [], self.get_parent_id(i)
        while p != -1:
            out.append(p)
            p = self.get_parent_id(p)
        return out

    def get_damping_by_id(self, i: int) -> float:
        return 0.3 if i == 0 else self.base.get_damping_by_id(i - 1)

    def random_state(self, rng, B=None):
        """Seeded (q, qd, qdd) with a unit base quaternion; leading axis B when given."""
        shape = () if B is None else (B,)
        nq, nv = self.get_num_pos(), self.get_num_vel()
        q = rng.uniform(-np.pi, np.pi, shape + (nq,))
        q[..., 0:3] = rng.uniform(-1.0, 1.0, shape + (3,))
        quat = rng.normal(size=shape + (4,))
        q[..., 3:7] = quat / np.linalg.norm(quat, axis=-1, keepdims=True)
        return q, rng.uniform(-1, 1, shape + (nv,)), rng.uniform(-1, 1, shape + (nv,))


def by_name(name: str):
    if name.endswith("_fb"):
        return FloatingBaseRobot(by_name(name[:-3]))
    table = {"iiwa14": iiwa14, "iiwa": iiwa14, "hyq": hyq, "atlas": atlas}
    if name not in table:
        raise KeyError("unknown robot %r (have %s)" % (name, sorted(table)))
    return table[name]()
