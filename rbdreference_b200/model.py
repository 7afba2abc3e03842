"""Model compiler: URDFParser-style robot object -> flat, device-ready tables.

The reference re-queries the robot object per body per call
(/root/reference/RBDReference.py:570-574, :595, :662, :666).  Here the robot is
compiled ONCE into plain arrays that the CUDA library turns into a kernel-parameter
(constant-bank) model:

* topology: parent[], ancestor bit masks, subtree lists, level schedule;
* per joint: motion subspace S (6), spatial inertia I (6x6), damping;
* per joint transform coefficients: every 1-DoF joint satisfies
      X(q) = A + B*cos(q) + C*sin(q)   (revolute)      X(q) = A + B*q   (prismatic)
  so A, B, C are recovered by PROBING the robot's own `get_Xmat_Func_by_id(i)` callable at
  q = 0, pi/2, pi (or 0, 1) and verified on random angles.  This reproduces whatever
  convention the parser used without re-deriving URDF semantics.

Only the 18 structurally non-zero entries of a Pluecker transform are kept:
X = [[E, 0], [L, E]]  ->  E (3x3 row-major) followed by L (3x3 row-major).
"""
from __future__ import annotations

import dataclasses
from typing import List

import numpy as np

MAX_DOF = 32  # kernel-parameter model limit (include/rbd_b200.h: RBD_MAX_DOF)

MAX_EE = 32   # include/rbd_b200.h: RBD_MAX_EE

__all__ = ["RobotModel", "EeModel", "compile_model", "compile_ee_model", "compile_fb_model", "FbModel", "select_end_effector_joints", "MAX_DOF", "MAX_EE"]


@dataclasses.dataclass
class RobotModel:
    name: str
    n: int
    parent: np.ndarray        # (n,) int32, -1 = fixed base
    kind: np.ndarray          # (n,) int32, 0 = revolute (cos/sin), 1 = prismatic (affine)
    S: np.ndarray             # (n,6) float64
    XA: np.ndarray            # (n,18) float64  [E | L]
    XB: np.ndarray            # (n,18)
    XC: np.ndarray            # (n,18)
    I: np.ndarray             # (n,36) float64 row-major 6x6
    damping: np.ndarray       # (n,) float64
    anc_mask: np.ndarray      # (n,) uint32: bit c set iff c is an ancestor of i or c == i
    subtree: List[List[int]]  # subtree(i) including i
    depth: np.ndarray         # (n,) int32, roots have depth 0
    levels: List[List[int]]   # bodies per depth

    # ---- derived counts used by bench.py / DESIGN.md (SURVEY.md 8d formulae) -------------
    def flops(self, op: str) -> int:
        """Algorithmic flop count per evaluation (ancestor-sparse, block-dense)."""
        n = self.n
        P = self.depth.astype(np.int64)                  # number of ancestors
        A = P + 1
        ST = np.array([len(s) for s in self.subtree], dtype=np.int64)
        Bi = A + ST - 1
        nr = self.parent >= 0
        root_of = np.arange(n)
        for i in range(n):
            root_of[i] = i if self.parent[i] < 0 else root_of[self.parent[i]]
        C = np.array([np.sum((root_of == root_of[i]) & (np.arange(n) >= i)) for i in range(n)], dtype=np.int64)
        rnea = int(np.sum(54 + np.where(nr, 132, 6) + 10 + 168) + 72 * np.sum(nr))
        if op == "rnea":
            return rnea
        if op == "rnea_grad":
            return int(rnea + 2 * np.sum(132 * P + 212 * A + 4) + 2 * 72 * np.sum(Bi[nr]) + 72 * np.sum(nr))
        if op == "minv":
            return int(np.sum(55 + 2 * ST) + np.sum(906 + 84 * ST[nr]) + np.sum(66 + 81 * C[nr]))
        if op == "fd":
            # forward_dynamics (:1369-1372): rnea + minv + (u - c) and one n x n mat-vec
            return int(self.flops("rnea") + self.flops("minv") + n + 2 * n * n)
        if op == "fd_grad":
            # forward_dynamics_grad (:1374-1384): fd + rnea_grad + the n x n by n x 2n product
            # (the reference's second minv evaluation at :1381 is not counted)
            return int(self.flops("fd") + self.flops("rnea_grad") + 2 * n * n * 2 * n)
        if op == "crba":
            # X build 54 per body; X^T IC X + add = 2 * 396 + 36 per non-root body (:1100-1103);
            # IC S and S.fh = 66 + 11 per body, X^T fh and S.fh = 66 + 11 per (body, ancestor) pair (:1108-1122)
            return int(54 * n + 828 * np.sum(nr) + np.sum(77 + 77 * P))
        if op == "ee_grad":
            # end_effector_pose_gradient over every leaf (:295-386): per chain joint a 4x4 build (24) and,
            # per derivative column, a full chain of 3x4 affine products (63 each, :312-324) plus the
            # column extraction (:327-351, ~40)
            leaves = [i for i in range(n) if not np.any(self.parent == i)]
            L = A[leaves]
            return int(np.sum(24 * L + L * (63 * L + 40)))
        raise KeyError(op)

    def io_bytes(self, op: str, itemsize: int = 8, full_rnea: bool = False) -> int:
        """Compulsory HBM bytes per evaluation: dense inputs + outputs."""
        n = self.n
        if op == "rnea":
            return (3 * n + n + (18 * n if full_rnea else 0)) * itemsize
        if op == "rnea_grad":
            return (3 * n + 2 * n * n) * itemsize
        if op in ("minv", "crba"):
            return (n + n * n) * itemsize
        if op == "fd":
            return (3 * n + n) * itemsize
        if op == "fd_grad":
            return (3 * n + 2 * n * n) * itemsize
        if op == "ee_grad":
            n_ee = sum(1 for i in range(n) if not np.any(self.parent == i))
            return (n + n_ee * 6 * n) * itemsize
        raise KeyError(op)


def _as_matrix(x) -> np.ndarray:
    return np.asarray(x, dtype=np.float64).reshape(6, 6)


def _pack18(X: np.ndarray, tol: float, what: str) -> np.ndarray:
    """Keep [E | L] of a Pluecker-structured 6x6 and check the structure."""
    if np.max(np.abs(X[:3, 3:])) > tol:
        raise ValueError("%s: upper-right 3x3 block is not zero - not a Pluecker motion transform" % what)
    if np.max(np.abs(X[:3, :3] - X[3:, 3:])) > tol:
        raise ValueError("%s: rotation blocks differ - not a Pluecker motion transform" % what)
    return np.concatenate((X[:3, :3].reshape(-1), X[3:, :3].reshape(-1)))


def compile_model(robot, name: str | None = None, check_points: int = 6) -> RobotModel:
    """Compile a URDFParser-style fixed-base robot into flat tables (see module docstring)."""
    if getattr(robot, "floating_base", False):
        raise NotImplementedError(
            "floating-base robots are outside the accelerated hot path (SURVEY.md 8f rank 3)")
    n = int(robot.get_num_bodies())
    if int(robot.get_num_vel()) != n:
        raise ValueError("hot path expects one 1-DoF joint per body (num_vel == num_bodies)")
    if not (1 <= n <= MAX_DOF):
        raise ValueError("robot has %d DoF; supported range is 1..%d" % (n, MAX_DOF))

    parent = np.array([int(robot.get_parent_id(i)) for i in range(n)], dtype=np.int32)
    for i in range(n):
        if not (-1 <= parent[i] < i):
            raise ValueError("body ids must be topologically ordered (parent id < child id)")
        for getter in ("get_joint_index_q", "get_joint_index_v", "get_joint_index_f"):
            fn = getattr(robot, getter, None)
            if fn is not None and int(fn(i)) != i:
                raise ValueError("%s(%d) != %d: only identity joint indexing is supported" % (getter, i, i))

    S = np.zeros((n, 6))
    kind = np.zeros(n, dtype=np.int32)
    XA = np.zeros((n, 18)); XB = np.zeros((n, 18)); XC = np.zeros((n, 18))
    I = np.zeros((n, 36))
    damping = np.zeros(n)
    rng = np.random.default_rng(20261018)
    for i in range(n):
        Si = np.asarray(robot.get_S_by_id(i), dtype=np.float64).reshape(-1)   # (6,), (6,1) or np.matrix
        if Si.shape != (6,):
            raise ValueError("S of joint %d must have 6 entries" % i)
        S[i] = Si
        ang, lin = np.any(Si[:3] != 0), np.any(Si[3:] != 0)
        if ang and lin:
            raise ValueError("joint %d: helical/compound motion subspaces are not supported" % i)
        kind[i] = 0 if ang else 1
        fn = robot.get_Xmat_Func_by_id(i)
        X0 = _as_matrix(fn(0.0))
        scale = max(1.0, float(np.max(np.abs(X0))))
        tol = 1e-11 * scale
        if kind[i] == 0:
            Xh, Xp = _as_matrix(fn(np.pi / 2)), _as_matrix(fn(np.pi))
            A, Bm = 0.5 * (X0 + Xp), 0.5 * (X0 - Xp)
            C = Xh - A
        else:
            A, Bm, C = X0, _as_matrix(fn(1.0)) - X0, np.zeros((6, 6))
        for t in rng.uniform(-3.0, 3.0, size=check_points):
            f1, f2 = (np.cos(t), np.sin(t)) if kind[i] == 0 else (t, 0.0)
            if np.max(np.abs(A + f1 * Bm + f2 * C - _as_matrix(fn(t)))) > tol:
                raise ValueError("joint %d: Xmat(q) is not A + B*cos(q) + C*sin(q) / A + B*q" % i)
        XA[i] = _pack18(A, tol, "joint %d (A)" % i)
        XB[i] = _pack18(Bm, tol, "joint %d (B)" % i)
        XC[i] = _pack18(C, tol, "joint %d (C)" % i)
        # snap numerical dust (|x| < 1e-15) so structural zeros are exact zeros
        for arr in (XA, XB, XC):
            arr[i][np.abs(arr[i]) < 1e-15 * scale] = 0.0
        I[i] = np.asarray(robot.get_Imat_by_id(i), dtype=np.float64).reshape(36)
        dget = getattr(robot, "get_damping_by_id", None)
        damping[i] = float(dget(i)) if dget is not None else 0.0

    # subtree(i) includes i (RBDReference.py:720-723); cross-check with the robot's own answer
    subtree: List[List[int]] = [[i] for i in range(n)]
    for i in range(n - 1, -1, -1):
        if parent[i] >= 0:
            subtree[parent[i]].extend(subtree[i])
    subtree = [sorted(s) for s in subtree]
    sget = getattr(robot, "get_subtree_by_id", None)
    if sget is not None:
        for i in range(n):
            if sorted(int(x) for x in sget(i)) != subtree[i]:
                raise ValueError("get_subtree_by_id(%d) disagrees with the parent array" % i)

    depth = np.zeros(n, dtype=np.int32)
    anc_mask = np.zeros(n, dtype=np.uint32)
    for i in range(n):
        if parent[i] >= 0:
            depth[i] = depth[parent[i]] + 1
            anc_mask[i] = anc_mask[parent[i]]
        anc_mask[i] |= np.uint32(1) << np.uint32(i)
    levels = [[int(i) for i in range(n) if depth[i] == d] for d in range(int(depth.max()) + 1)]

    return RobotModel(name=name or getattr(robot, "name", "robot"), n=n, parent=parent, kind=kind, S=S,
                      XA=XA, XB=XB, XC=XC, I=I, damping=damping, anc_mask=anc_mask, subtree=subtree,
                      depth=depth, levels=levels)


# ----------------------------------------------------------------------------------------------
# end-effector kinematics (RBDReference.py:190-386)
# ----------------------------------------------------------------------------------------------
@dataclasses.dataclass
class EeModel:
    """Flat tables behind rbd_ee_model_create (include/rbd_b200.h: RbdEeDesc)."""
    n: int
    parent: np.ndarray      # (n,) int32
    kind: np.ndarray        # (n,) int32
    TA: np.ndarray          # (n,12) rows 0..2 of get_Xmat_hom_Func_by_id(i)(q) = TA + TB f1 + TC f2
    TB: np.ndarray
    TC: np.ndarray
    DA: np.ndarray          # (n,12) the same for get_dXmat_hom_Func_by_id
    DB: np.ndarray
    DC: np.ndarray
    ee_joint: np.ndarray    # (n_ee,) int32: moving joint each chain starts from
    ee_final: np.ndarray    # (n_ee,12) fixed-joint transform closing the chain (identity rows otherwise)
    offset: np.ndarray      # (4,) ee_offsets[0]

    @property
    def n_ee(self) -> int:
        return int(self.ee_joint.shape[0])


def select_end_effector_joints(robot, ee_joint_names):
    """RBDReference.py:190-211: all leaf joints by default; otherwise the named moving joints followed
    by the named fixed joints.  Raises the reference's ValueError for an unknown name (:208)."""
    if ee_joint_names is None:
        return [int(j) for j in robot.get_leaf_nodes()], []
    ee_jids, fixed_jids = [], []
    for name in ee_joint_names:
        joint = robot.get_joint_by_name(name)
        if joint is not None:
            ee_jids.append(int(joint.get_id()))
        else:
            fjoint = robot.get_fixed_joint_by_name(name)
            if fjoint is None:
                raise ValueError("Could not find joint or fixed joint named: " + name)
            fixed_jids.append(int(fjoint.get_id()))
    return ee_jids, fixed_jids


def _fit_hom(fn, affine: bool, rng, what: str, last_row) -> tuple:
    """Recover A, B, C of a 4x4 joint function (A + B cos q + C sin q, or A + B q) by probing."""
    M0 = np.asarray(fn(0.0), dtype=np.float64).reshape(4, 4)
    if affine:
        A, Bm, C = M0, np.asarray(fn(1.0), dtype=np.float64).reshape(4, 4) - M0, np.zeros((4, 4))
    else:
        Mh = np.asarray(fn(np.pi / 2), dtype=np.float64).reshape(4, 4)
        Mp = np.asarray(fn(np.pi), dtype=np.float64).reshape(4, 4)
        A, Bm = 0.5 * (M0 + Mp), 0.5 * (M0 - Mp)
        C = Mh - A
    tol = 1e-11 * max(1.0, float(np.max(np.abs(M0))))
    for t in rng.uniform(-3.0, 3.0, size=6):
        f1, f2 = (t, 0.0) if affine else (np.cos(t), np.sin(t))
        Mt = np.asarray(fn(t), dtype=np.float64).reshape(4, 4)
        if np.max(np.abs(A + f1 * Bm + f2 * C - Mt)) > tol:
            raise ValueError("%s is not A + B*cos(q) + C*sin(q) / A + B*q" % what)
        if np.max(np.abs(Mt[3] - np.asarray(last_row, dtype=np.float64))) > tol:
            raise ValueError("%s: fourth row must be %s" % (what, list(last_row)))
    out = []
    for M in (A, Bm, C):
        M = M[:3].reshape(12).copy()
        M[np.abs(M) < 1e-15] = 0.0
        out.append(M)
    return tuple(out)


def compile_ee_model(robot, ee_joint_names=None, ee_offsets=None) -> EeModel:
    """Compile an end-effector selection (the per-call arguments of end_effector_pose /
    end_effector_pose_gradient, RBDReference.py:220, :295) into flat tables."""
    if getattr(robot, "floating_base", False):
        raise NotImplementedError("end-effector kinematics: floating base is a TODO upstream (RBDReference.py:217, :287)")
    n = int(robot.get_num_joints())
    if not (1 <= n <= MAX_DOF):
        raise ValueError("robot has %d joints; supported range is 1..%d" % (n, MAX_DOF))
    parent = np.array([int(robot.get_parent_id(i)) for i in range(n)], dtype=np.int32)
    for i in range(n):
        if not (-1 <= parent[i] < i):
            raise ValueError("joint ids must be topologically ordered (parent id < child id)")
    rng = np.random.default_rng(20261018)
    kind = np.zeros(n, dtype=np.int32)
    tabs = {k: np.zeros((n, 12)) for k in ("TA", "TB", "TC", "DA", "DB", "DC")}
    for i in range(n):
        S = np.asarray(robot.get_S_by_id(i), dtype=np.float64).reshape(-1)
        kind[i] = 0 if np.any(S[:3] != 0) else 1
        tabs["TA"][i], tabs["TB"][i], tabs["TC"][i] = _fit_hom(
            robot.get_Xmat_hom_Func_by_id(i), kind[i] == 1, rng, "joint %d: Xmat_hom(q)" % i, (0, 0, 0, 1))
        tabs["DA"][i], tabs["DB"][i], tabs["DC"][i] = _fit_hom(
            robot.get_dXmat_hom_Func_by_id(i), kind[i] == 1, rng, "joint %d: dXmat_hom(q)" % i, (0, 0, 0, 0))
    ee_jids, fixed_jids = select_end_effector_joints(robot, ee_joint_names)
    ee_joint, ee_final = [], []
    for jid in ee_jids:
        ee_joint.append(jid)
        ee_final.append(np.eye(4)[:3].reshape(12))
    for fjid in fixed_jids:                       # RBDReference.py:276-281
        fj = robot.get_fixed_joint_by_id(fjid)
        pj = robot.get_joint_by_name(fj.parent_name)
        F = np.asarray(fj.get_transformation_matrix_hom(), dtype=np.float64).reshape(4, 4)
        if np.max(np.abs(F[3] - np.array([0.0, 0.0, 0.0, 1.0]))) > 1e-12:
            raise ValueError("fixed joint %r: fourth row of its transform must be [0, 0, 0, 1]" % fj.name)
        ee_joint.append(int(pj.get_id()))
        ee_final.append(F[:3].reshape(12))
    if not (1 <= len(ee_joint) <= MAX_EE):
        raise ValueError("%d end effectors requested; supported range is 1..%d" % (len(ee_joint), MAX_EE))
    if ee_offsets is None:
        offset = np.array([0.0, 0.0, 0.0, 1.0])
    else:
        offset = np.asarray(ee_offsets[0], dtype=np.float64).reshape(-1)      # only offsets[0] is used upstream
        if offset.shape != (4,):
            raise ValueError("ee_offsets[0] must have 4 entries (x, y, z, 1)")
    return EeModel(n=n, parent=parent, kind=kind, ee_joint=np.array(ee_joint, dtype=np.int32),
                   ee_final=np.stack(ee_final), offset=offset, **tabs)


# ----------------------------------------------------------------------------------------------
# floating base (RBDReference.py: the `self.robot.floating_base` branches)
# ----------------------------------------------------------------------------------------------
@dataclasses.dataclass
class FbModel:
    """Flat tables behind rbd_fb_model_create (include/rbd_b200.h: RbdFbModelDesc).  Entry 0 of the
    per-body tables is the base (only its inertia and damping are used)."""
    name: str
    NB: int
    parent: np.ndarray
    kind: np.ndarray
    S: np.ndarray
    XA: np.ndarray
    XB: np.ndarray
    XC: np.ndarray
    I: np.ndarray
    damping: np.ndarray
    pos_off: int
    quat_off: int
    w_first: int
    transpose: int

    @property
    def nv(self) -> int:
        return self.NB + 5

    @property
    def nq(self) -> int:
        return self.NB + 6

    def flops(self, op: str) -> int:
        """Algorithmic flops per evaluation: the SURVEY.md 8d formulae with the base counted as one
        body that owns six derivative / matrix columns (S = eye(6)) instead of one."""
        NB = self.NB
        depth = np.zeros(NB, dtype=np.int64)
        size = np.ones(NB, dtype=np.int64)
        for i in range(1, NB):
            depth[i] = depth[self.parent[i]] + 1
        for i in range(NB - 1, 0, -1):
            size[self.parent[i]] += size[i]
        own = np.where(np.arange(NB) == 0, 6, 1)
        P = np.where(np.arange(NB) == 0, 0, 5 + depth)           # columns owned by ancestors
        A = P + own
        STc = size + np.where(np.arange(NB) == 0, 5, 0)          # columns owned by the subtree
        Bi = A + STc - own
        nr = np.arange(NB) > 0
        rnea = int(NB * (54 + 132 + 10 + 168) + 72 * np.sum(nr))
        if op == "rnea":
            return rnea
        if op == "rnea_grad":
            return int(rnea + 2 * np.sum(132 * P + 212 * A + 4) + 2 * 72 * np.sum(Bi[nr]) + 72 * np.sum(nr))
        if op == "minv":
            nv = NB + 5
            return int(np.sum(55 + 2 * STc) + np.sum(906 + 84 * STc[nr]) + np.sum(66 + 81 * nv) * np.sum(nr) + 6 * 6 * 6 * 2)
        raise KeyError(op)

    def io_bytes(self, op: str, itemsize: int = 8, full_rnea: bool = False) -> int:
        nv, nq = self.nv, self.nq
        if op == "rnea":
            return (nq + 2 * nv + nv + (18 * self.NB if full_rnea else 0)) * itemsize
        if op == "rnea_grad":
            return (nq + 2 * nv + 2 * nv * nv) * itemsize
        if op == "minv":
            return (nq + nv * nv) * itemsize
        raise KeyError(op)


def _fb_candidate(q7, pos_off, quat_off, w_first, transpose):
    p = q7[pos_off:pos_off + 3]
    qq = q7[quat_off:quat_off + 4]
    w, x, y, z = (qq[0], qq[1], qq[2], qq[3]) if w_first else (qq[3], qq[0], qq[1], qq[2])
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    E = R if transpose else R.T
    px = np.array([[0.0, -p[2], p[1]], [p[2], 0.0, -p[0]], [-p[1], p[0], 0.0]])
    X = np.zeros((6, 6))
    X[:3, :3] = E
    X[3:, 3:] = E
    X[3:, :3] = -E @ px
    return X


def compile_fb_model(robot, name: str | None = None) -> FbModel:
    """Compile a floating-base robot: 1-DoF joints as in `compile_model`, base layout by probing
    `get_Xmat_Func_by_id(0)` against X0 = xrot(E(quat)) xlt(position) in its eight layouts."""
    NB = int(robot.get_num_bodies())
    if not (2 <= NB <= MAX_DOF):
        raise ValueError("floating-base robot has %d bodies; supported range is 2..%d" % (NB, MAX_DOF))
    if int(robot.get_num_vel()) != NB + 5:
        raise ValueError("floating base: expected num_vel == num_bodies + 5 (RBDReference.py:653)")
    parent = np.array([int(robot.get_parent_id(i)) for i in range(NB)], dtype=np.int32)
    if parent[0] != -1 or np.any(parent[1:] < 0) or np.any(parent[1:] >= np.arange(1, NB)):
        raise ValueError("floating base: body 0 must be the only root and ids topologically ordered")
    S0 = np.asarray(robot.get_S_by_id(0), dtype=np.float64)
    if S0.shape != (6, 6) or not np.array_equal(S0, np.eye(6)):
        raise ValueError("floating base: S of body 0 must be eye(6)")
    if list(robot.get_joint_index_q(0)) != list(range(7)) or list(robot.get_joint_index_v(0)) != list(range(6)):
        raise ValueError("floating base: the base must own q[0:7] and qd[0:6]")
    for i in range(1, NB):
        if int(robot.get_joint_index_q(i)) != i + 6 or int(robot.get_joint_index_v(i)) != i + 5:
            raise ValueError("floating base: joint i must read q[i+6] and qd[i+5] (RBDReference.py:1143-1144)")

    class _Joints:                                   # bodies 1.. through the fixed-base compiler
        floating_base = False

        def get_num_bodies(self): return NB - 1
        def get_num_vel(self): return NB - 1
        def get_parent_id(self, i): return int(robot.get_parent_id(i + 1)) - 1
        def get_S_by_id(self, i): return robot.get_S_by_id(i + 1)
        def get_Xmat_Func_by_id(self, i): return robot.get_Xmat_Func_by_id(i + 1)
        def get_Imat_by_id(self, i): return robot.get_Imat_by_id(i + 1)
        def get_damping_by_id(self, i): return robot.get_damping_by_id(i + 1)

    jm = compile_model(_Joints(), name="joints")

    def pad(a, first):
        return np.concatenate((np.asarray(first, dtype=a.dtype).reshape((1,) + a.shape[1:]), a))

    rng = np.random.default_rng(20261018)
    f0 = robot.get_Xmat_Func_by_id(0)
    samples = []
    for _ in range(6):
        q7 = rng.uniform(-1.0, 1.0, 7)
        samples.append(q7)
    layout = None
    for pos_off, quat_off in ((0, 3), (4, 0)):
        for w_first in (0, 1):
            for transpose in (0, 1):
                ok = True
                for q7 in samples:
                    q7 = q7.copy()
                    q7[quat_off:quat_off + 4] /= np.linalg.norm(q7[quat_off:quat_off + 4])
                    X = np.asarray(f0(q7), dtype=np.float64).reshape(6, 6)
                    if np.max(np.abs(X - _fb_candidate(q7, pos_off, quat_off, w_first, transpose))) > 1e-11 * max(1.0, np.max(np.abs(X))):
                        ok = False
                        break
                if ok and layout is None:
                    layout = (pos_off, quat_off, w_first, transpose)
    if layout is None:
        raise ValueError("floating base: get_Xmat_Func_by_id(0) is not xrot(E(quaternion)) @ xlt(position) in any "
                         "supported layout (position/quaternion order, w first/last, E = R or R^T)")
    dget = getattr(robot, "get_damping_by_id", None)
    return FbModel(name=name or getattr(robot, "name", "robot"), NB=NB, parent=parent,
                   kind=pad(jm.kind, [0]), S=pad(jm.S, np.zeros(6)), XA=pad(jm.XA, np.zeros(18)),
                   XB=pad(jm.XB, np.zeros(18)), XC=pad(jm.XC, np.zeros(18)),
                   I=pad(jm.I, np.asarray(robot.get_Imat_by_id(0), dtype=np.float64).reshape(36)),
                   damping=pad(jm.damping, [float(dget(0)) if dget is not None else 0.0]),
                   pos_off=layout[0], quat_off=layout[1], w_first=layout[2], transpose=layout[3])
