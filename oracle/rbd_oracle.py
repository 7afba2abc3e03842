"""CPU oracle for the rnea / rnea_grad / minv hot path.  TEST INFRASTRUCTURE ONLY.

This module is a from-scratch numpy restatement of the algorithms in
/root/reference/RBDReference.py.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it; the product
package `rbdreference_b200` never does (its ops raise if the CUDA library is missing).

Pinning: the reference ships no tests, fixtures or golden vectors,
so the oracle is pinned against outputs of the unmodified reference run in the build
container: `oracle/make_golden.py` imports /root/reference/RBDReference.py, evaluates
every hot-path function on seeded states of iiwa14 / HyQ / Atlas / random trees and
commits the results to `tests/golden/*.npz`; `tests/test_oracle.py` checks this module
against those vectors (and, when the reference is importable, against it live).

Two layers:

* `ScalarOracle`  - one state per call, same method names, argument order, return
  shapes and in-place-mutation behaviour as the reference class, written as plain
  loops over bodies.  Each method cites the reference lines it follows.
* `BatchOracle`   - the same recursions vectorised over a leading batch axis B so that
  parity at 10^4..10^6 knot points finishes in seconds.  Outputs put the batch first:
  c (B,n); v,a,f (B,6,NB); dc_du (B,n,2n); Minv (B,n,n).

Spatial vectors are [angular; linear]; X maps parent motion coordinates to the child
(RBDReference.py:566,:580,:618).
"""
from __future__ import annotations

import numpy as np

__all__ = ["ScalarOracle", "BatchOracle", "crm", "crf"]


# ----------------------------------------------------------------------------------------
# spatial cross-product operators
# ----------------------------------------------------------------------------------------
def _skew(w):
    return np.array([[0.0, -w[2], w[1]], [w[2], 0.0, -w[0]], [-w[1], w[0], 0.0]])


def crm(v):
    """Motion cross-product matrix [w x, 0; v x, w x]   (RBDReference.py:9-21)."""
    v = np.asarray(v, dtype=float).reshape(6)
    out = np.zeros((6, 6))
    out[:3, :3] = _skew(v[:3])
    out[3:, :3] = _skew(v[3:])
    out[3:, 3:] = _skew(v[:3])
    return out


def crf(v):
    """Force cross-product matrix = -crm(v)^T   (RBDReference.py:23-25)."""
    return -crm(v).T


def _flat6(S):
    return np.asarray(S, dtype=float).reshape(-1)


# ----------------------------------------------------------------------------------------
# scalar oracle: one knot point per call, reference shapes
# ----------------------------------------------------------------------------------------
class ScalarOracle:
    """Single-state restatement with the reference's public surface (README.md:9-17)."""

    def __init__(self, robot):
        self.robot = robot
        if getattr(robot, "floating_base", False):
            raise NotImplementedError("oracle covers the fixed-base hot path only (SURVEY.md 8f)")
        self.NB = robot.get_num_bodies()
        self.n = robot.get_num_vel()
        self.parent = [robot.get_parent_id(i) for i in range(self.NB)]
        self.S = [_flat6(robot.get_S_by_id(i)) for i in range(self.NB)]
        self.I = [np.array(robot.get_Imat_by_id(i), dtype=float) for i in range(self.NB)]
        self.subtree = [list(robot.get_subtree_by_id(i)) for i in range(self.NB)]

    def _X(self, i, q):
        return np.asarray(self.robot.get_Xmat_Func_by_id(i)(q[self.robot.get_joint_index_q(i)]), dtype=float)

    @staticmethod
    def _gravity(GRAVITY):
        g = np.zeros(6)
        g[5] = -GRAVITY  # RBDReference.py:566
        return g

    # -- RNEA -----------------------------------------------------------------------------
    def rnea_fpass(self, q, qd, qdd=None, GRAVITY=-9.81):
        """Forward sweep, RBDReference.py:559-598."""
        NB = self.NB
        v = np.zeros((6, NB))
        a = np.zeros((6, NB))
        f = np.zeros((6, NB))
        a_base = self._gravity(GRAVITY)
        for i in range(NB):
            X = self._X(i, q)
            p = self.parent[i]
            S = self.S[i]
            if p == -1:
                vi = np.zeros(6)                       # :577
                ai = X @ a_base                        # :578
            else:
                vi = X @ v[:, p]                       # :580
                ai = X @ a[:, p]                       # :581
            vJ = S * qd[i]                             # :586
            vi = vi + vJ                               # :587
            ai = ai + crm(vi) @ vJ                     # :588 (uses v_i after adding vJ)
            if qdd is not None:
                ai = ai + S * qdd[i]                   # :589-593
            Ii = self.I[i]
            f[:, i] = Ii @ ai + crf(vi) @ (Ii @ vi)    # :596 (vxIv :170-182)
            v[:, i] = vi
            a[:, i] = ai
        return v, a, f

    def rnea_bpass(self, q, f):
        """Backward sweep, RBDReference.py:600-621.  `f` is accumulated in place."""
        c = np.zeros(self.n)
        for i in range(self.NB - 1, -1, -1):
            c[i] = self.S[i] @ f[:, i]                 # :612
            p = self.parent[i]
            if p != -1:
                f[:, p] = f[:, p] + self._X(i, q).T @ f[:, i]   # :617-619
        return c, f

    def rnea(self, q, qd, qdd=None, GRAVITY=-9.81, f_ext=None):
        """RBDReference.py:623-628 (f_ext accepted and ignored, f returned accumulated)."""
        v, a, f = self.rnea_fpass(q, qd, qdd, GRAVITY)
        c, f = self.rnea_bpass(q, f)
        return c, v, a, f

    # -- Minv -----------------------------------------------------------------------------
    def minv_bpass(self, q):
        """Backward pass of the analytical inverse, RBDReference.py:630-735."""
        n, NB = self.n, self.NB
        Minv = np.zeros((n, n))
        F = np.zeros((n, 6, n))
        U = np.zeros((n, 6))
        D = np.zeros(n)                                # the reference calls this Dinv (:698)
        IA = {i: self.I[i].copy() for i in range(NB)}  # :662
        for i in range(NB - 1, -1, -1):
            S = self.S[i]
            sub = self.subtree[i]
            U[i] = IA[i] @ S                           # :697
            D[i] = S @ U[i]                            # :698
            Minv[i, i] = 1.0 / D[i]                    # :700
            Minv[i, sub] -= (1.0 / D[i]) * (S @ F[i][:, sub])   # :702-708
            p = self.parent[i]
            if p != -1:
                X = self._X(i, q)
                for j in sub:                          # :720-726
                    F[i][:, j] += U[i] * Minv[i, j]
                    F[p][:, j] += X.T @ F[i][:, j]
                Ia = IA[i] - np.outer(U[i], U[i]) / D[i]   # :728-731
                IA[p] = IA[p] + X.T @ Ia @ X               # :732-733
        return Minv, F, U, D

    def minv_fpass(self, q, Minv, F, U, Dinv):
        """Forward pass, RBDReference.py:737-783.  Mutates and returns `Minv`; rewrites F."""
        for i in range(self.NB):
            p = self.parent[i]
            S = self.S[i]
            if p != -1:
                X = self._X(i, q)
                Minv[i, :] -= (1.0 / Dinv[i]) * ((U[i] @ X) @ F[p])   # :771-773
                F[i] = X @ F[p] + np.outer(S, Minv[i, :])              # :774-776
            else:
                F[i] = np.outer(S, Minv[i, :])                         # :781
        return Minv

    def minv(self, q, output_dense=True):
        """RBDReference.py:785-806."""
        Minv, F, U, D = self.minv_bpass(q)
        Minv = self.minv_fpass(q, Minv, F, U, D)
        if output_dense:
            iu = np.triu_indices(self.NB, 1)
            Minv[iu[1], iu[0]] = Minv[iu]              # :799-804 (lower <- upper)
        return Minv

    # -- CRBA (fixed base) - used by identity tests only ----------------------------------
    def crba(self, q):
        """Joint-space inertia matrix, RBDReference.py:1090-1124."""
        n = self.n
        IC = {i: self.I[i].copy() for i in range(n)}
        Xs = [self._X(i, q) for i in range(n)]
        for i in range(n - 1, -1, -1):
            p = self.parent[i]
            if p != -1:
                IC[p] = IC[p] + Xs[i].T @ IC[i] @ Xs[i]
        H = np.zeros((n, n))
        for i in range(n):
            fh = IC[i] @ self.S[i]
            H[i, i] = self.S[i] @ fh
            j = i
            while self.parent[j] > -1:
                fh = Xs[j].T @ fh
                j = self.parent[j]
                H[i, j] = self.S[j] @ fh
                H[j, i] = H[i, j]
        return H

    # -- RNEA gradient --------------------------------------------------------------------
    def rnea_grad_fpass_dq(self, q, qd, v, a, GRAVITY=-9.81):
        """d(v,a,f)/dq per body, RBDReference.py:1127-1187."""
        n, NB = self.n, self.NB
        dv = np.zeros((6, n, NB))
        da = np.zeros((6, n, NB))
        df = np.zeros((6, n, NB))
        a_base = self._gravity(GRAVITY)
        for i in range(NB):
            p = self.parent[i]
            X = self._X(i, q)
            S = self.S[i]
            if p != -1:
                dv[:, :, i] = X @ dv[:, :, p]                      # :1158
                dv[:, i, i] += crm(X @ v[:, p]) @ S                # :1159
                da[:, :, i] = X @ da[:, :, p]                      # :1163
            for c in range(n):
                da[:, c, i] += qd[i] * (crm(dv[:, c, i]) @ S)      # :1170
            a_par = a[:, p] if p != -1 else a_base
            da[:, i, i] += crm(X @ a_par) @ S                      # :1173,:1175
            Ii = self.I[i]
            Iv = Ii @ v[:, i]                                      # :1180
            df[:, :, i] = Ii @ da[:, :, i]                         # :1179
            for c in range(n):                                     # :1182-1185
                df[:, c, i] += crf(dv[:, c, i]) @ Iv + crf(v[:, i]) @ (Ii @ dv[:, c, i])
        return dv, da, df

    def rnea_grad_fpass_dqd(self, q, qd, v):
        """d(v,a,f)/dqd per body, RBDReference.py:1189-1255."""
        n, NB = self.n, self.NB
        dv = np.zeros((6, n, NB))
        da = np.zeros((6, n, NB))
        df = np.zeros((6, n, NB))
        for i in range(NB):
            p = self.parent[i]
            X = self._X(i, q)
            S = self.S[i]
            if p != -1:
                dv[:, :, i] = X @ dv[:, :, p]                      # :1230
                da[:, :, i] = X @ da[:, :, p]                      # :1234
            dv[:, i, i] += S                                       # :1231
            for c in range(n):
                da[:, c, i] += qd[i] * (crm(dv[:, c, i]) @ S)      # :1240
            da[:, i, i] += crm(v[:, i]) @ S                        # :1243
            Ii = self.I[i]
            Iv = Ii @ v[:, i]
            df[:, :, i] = Ii @ da[:, :, i]                         # :1247
            for c in range(n):                                     # :1249-1252
                df[:, c, i] += crf(dv[:, c, i]) @ Iv + crf(v[:, i]) @ (Ii @ dv[:, c, i])
        return dv, da, df

    def rnea_grad_bpass_dq(self, q, f, df_dq):
        """dc/dq, RBDReference.py:1257-1297.  `df_dq` is accumulated in place."""
        n = self.n
        dc = np.zeros((n, n))
        for i in range(self.NB - 1, -1, -1):
            S = self.S[i]
            dc[i, :] = S @ df_dq[:, :, i]                          # :1284
            p = self.parent[i]
            if p != -1:
                X = self._X(i, q)
                df_dq[:, :, p] += X.T @ df_dq[:, :, i]             # :1291
                df_dq[:, i, p] += X.T @ (-(crm(f[:, i]) @ S))      # :1292-1294 (fxS :166-168)
        return dc

    def rnea_grad_bpass_dqd(self, q, df_dqd, USE_VELOCITY_DAMPING=False):
        """dc/dqd, RBDReference.py:1299-1343.  `df_dqd` is accumulated in place."""
        n = self.n
        dc = np.zeros((n, n))
        for i in range(self.NB - 1, -1, -1):
            dc[i, :] = self.S[i] @ df_dqd[:, :, i]                 # :1325
            p = self.parent[i]
            if p != -1:
                df_dqd[:, :, p] += self._X(i, q).T @ df_dqd[:, :, i]   # :1331
        if USE_VELOCITY_DAMPING:
            for i in range(self.NB):
                dc[i, i] += self.robot.get_damping_by_id(i)        # :1341
        return dc

    def rnea_grad(self, q, qd, qdd=None, GRAVITY=-9.81, USE_VELOCITY_DAMPING=False):
        """RBDReference.py:1345-1368 -> dc_du = [dc_dq | dc_dqd], shape (n, 2n)."""
        c, v, a, f = self.rnea(q, qd, qdd, GRAVITY)
        _, _, df_dq = self.rnea_grad_fpass_dq(q, qd, v, a, GRAVITY)
        _, _, df_dqd = self.rnea_grad_fpass_dqd(q, qd, v)
        dc_dq = self.rnea_grad_bpass_dq(q, f, df_dq)
        dc_dqd = self.rnea_grad_bpass_dqd(q, df_dqd, USE_VELOCITY_DAMPING)
        return np.hstack((dc_dq, dc_dqd))

    # -- compositions ------------------------------------------------
    def forward_dynamics(self, q, qd, u):
        """RBDReference.py:1371-1374 (rnea is called with qdd=None)."""
        c = self.rnea(q, qd)[0]
        return self.minv(q) @ (u - c)

    def forward_dynamics_grad(self, q, qd, u):
        """RBDReference.py:1376-1384."""
        qdd = self.forward_dynamics(q, qd, u)
        dc_du = self.rnea_grad(q, qd, qdd)
        Minv = self.minv(q)
        return -Minv @ dc_du[:, : self.n], -Minv @ dc_du[:, self.n:]

    # -- ABA (fixed base) - SURVEY.md 8f rank 4 --------------------------------------------
    def aba(self, q, qd, tau, f_ext=None, GRAVITY=-9.81):
        """Articulated-body algorithm, RBDReference.py:817 (fixed-base branch :940-1024).

        Follows the reference to the letter, including :984 where the bias force of every
        body is set to ELEMENT 0 of crf(v) I v broadcast over all six entries (`...[0]` on a
        1-D product).  f_ext is ignored by the fixed-base branch."""
        n = self.n
        v = np.zeros((6, n)); c = np.zeros((6, n)); a = np.zeros((6, n))
        d = np.zeros(n); U = np.zeros((6, n)); u = np.zeros(n)
        IA = [None] * n
        pA = np.zeros((6, n))
        qdd = np.zeros(n)
        Xs = [self._X(i, q) for i in range(n)]
        for i in range(n):
            p = self.parent[i]
            S = self.S[i]
            if p == -1:
                v[:, i] = S * qd[i]                                     # :957
            else:
                v[:, i] = Xs[i] @ v[:, p] + S * qd[i]                   # :960-961
                c[:, i] = qd[i] * (crm(v[:, i]) @ S)                    # :962 (_mxS :56-59)
            IA[i] = self.I[i].copy()                                    # :966
            pA[:, i] = (crf(v[:, i]) @ self.I[i] @ v[:, i])[0]          # :978-984 (element 0, broadcast)
        for i in range(n - 1, -1, -1):
            S = self.S[i]
            p = self.parent[i]
            U[:, i] = IA[i] @ S                                         # :990
            d[i] = S @ U[:, i]                                          # :991
            u[i] = tau[i] - S @ pA[:, i]                                # :992
            if p != -1:
                Ia = IA[i] - np.outer(U[:, i], U[:, i]) / d[i]          # :996-997
                pa = pA[:, i] + Ia @ c[:, i] + U[:, i] * u[i] / d[i]    # :999
                IA[p] = IA[p] + Xs[i].T @ Ia @ Xs[i]                    # :1001-1004
                pA[:, p] = pA[:, p] + Xs[i].T @ pa                      # :1006-1007
        g = self._gravity(GRAVITY)
        for i in range(n):
            p = self.parent[i]
            if p == -1:
                a[:, i] = Xs[i] @ g + c[:, i]                           # :1015
            else:
                a[:, i] = Xs[i] @ a[:, p] + c[:, i]                     # :1017
            qdd[i] = (u[i] - U[:, i] @ a[:, i]) / d[i]                  # :1020-1021
            a[:, i] = a[:, i] + qdd[i] * self.S[i]                      # :1022
        return qdd

    # -- end-effector kinematics ---------------------------------------
    def select_end_effector_joints(self, ee_joint_names):
        """RBDReference.py:190-211: leaf joints by default, else moving joints first, then fixed joints."""
        if ee_joint_names is None:
            return list(self.robot.get_leaf_nodes()), []
        ee_jids, fixed_jids = [], []
        for name in ee_joint_names:
            joint = self.robot.get_joint_by_name(name)
            if joint is not None:
                ee_jids.append(joint.get_id())
            else:
                fjoint = self.robot.get_fixed_joint_by_name(name)
                if fjoint is None:
                    raise ValueError("Could not find joint or fixed joint named: " + name)   # :208
                fixed_jids.append(fjoint.get_id())
        return ee_jids, fixed_jids

    def _ee_targets(self, ee_joint_names):
        """(leaf joint id, final 4x4) per requested end effector, in the reference's output order."""
        ee_jids, fixed_jids = self.select_end_effector_joints(ee_joint_names)
        out = [(jid, np.eye(4)) for jid in ee_jids]
        for fjid in fixed_jids:
            fj = self.robot.get_fixed_joint_by_id(fjid)                                        # :277
            parent = self.robot.get_joint_by_name(fj.parent_name)                              # :278
            out.append((parent.get_id(), np.asarray(fj.get_transformation_matrix_hom(), dtype=float)))
        return out

    @staticmethod
    def _ee_offset(ee_offsets):
        return np.asarray(ee_offsets[0], dtype=float).reshape(4)    # only offsets[0] is used (:248, :335)

    def _ee_chain(self, jid, q, final, dind=None):
        """backwardChain / dbackward_chain, RBDReference.py:235-243, :312-324."""
        X = np.array(final, dtype=float)
        dX = np.array(final, dtype=float)
        cur = jid
        while cur != -1:
            T = np.asarray(self.robot.get_Xmat_hom_Func_by_id(cur)(q[cur]), dtype=float)
            dT = np.asarray(self.robot.get_dXmat_hom_Func_by_id(cur)(q[cur]), dtype=float) if cur == dind else T
            dX = dT @ dX
            X = T @ X
            cur = self.parent[cur]
        return X, dX

    @staticmethod
    def _ee_pose_from(X, off):
        """xyz of the offset point and roll-pitch-yaw of the rotation, RBDReference.py:247-260."""
        xyz = (X @ off)[:3]
        roll = np.arctan2(X[2, 1], X[2, 2])
        pitch = np.arctan2(-X[2, 0], np.sqrt(X[2, 2] * X[2, 2] + X[2, 1] * X[2, 1]))
        yaw = np.arctan2(X[1, 0], X[0, 0])
        return np.concatenate((xyz, [roll, pitch, yaw]))

    @staticmethod
    def _ee_dpose_from(X, dX, off):
        """One column of the pose gradient, RBDReference.py:327-351."""
        def darctan2(y, x, yp, xp):
            return (-xp * y + x * yp) / (x * x + y * y)
        dxyz = (dX @ off)[:3]
        droll = darctan2(X[2, 1], X[2, 2], dX[2, 1], dX[2, 2])
        sq = np.sqrt(X[2, 2] * X[2, 2] + X[2, 1] * X[2, 1])
        dsq = (X[2, 2] * dX[2, 2] + X[2, 1] * dX[2, 1]) / sq
        dpitch = darctan2(-X[2, 0], sq, -dX[2, 0], dsq)
        dyaw = darctan2(X[1, 0], X[0, 0], dX[1, 0], dX[0, 0])
        return np.concatenate((dxyz, [droll, dpitch, dyaw]))

    def end_effector_pose(self, q, ee_joint_names=None, ee_offsets=((0, 0, 0, 1),)):
        """RBDReference.py:220-283 -> list over end effectors of (6, 1) [x y z roll pitch yaw]."""
        off = self._ee_offset(ee_offsets)
        return [self._ee_pose_from(self._ee_chain(jid, q, fin)[0], off).reshape(6, 1)
                for jid, fin in self._ee_targets(ee_joint_names)]

    def end_effector_pose_gradient(self, q, ee_joint_names=None, ee_offsets=((0, 0, 0, 1),)):
        """RBDReference.py:295-386 -> list over end effectors of (6, n); zero columns off the chain."""
        off = self._ee_offset(ee_offsets)
        n = self.robot.get_num_joints()
        out = []
        for jid, fin in self._ee_targets(ee_joint_names):
            chain = set(self.robot.get_ancestors_by_id(jid)) | {jid}
            G = np.zeros((6, n))
            for dind in range(n):
                if dind in chain:
                    X, dX = self._ee_chain(jid, q, fin, dind)
                    G[:, dind] = self._ee_dpose_from(X, dX, off)
            out.append(G)
        return out


# ----------------------------------------------------------------------------------------
# batched oracle: the same recursions vectorised over a leading batch axis
# ----------------------------------------------------------------------------------------
def _crm_b(v):
    """Batched crm: v (..., 6) -> (..., 6, 6)."""
    out = np.zeros(v.shape[:-1] + (6, 6), dtype=v.dtype)
    w, u = v[..., :3], v[..., 3:]
    for blk, x in (((0, 0), w), ((3, 0), u), ((3, 3), w)):
        r, c = blk
        out[..., r + 0, c + 1] = -x[..., 2]
        out[..., r + 0, c + 2] = x[..., 1]
        out[..., r + 1, c + 0] = x[..., 2]
        out[..., r + 1, c + 2] = -x[..., 0]
        out[..., r + 2, c + 0] = -x[..., 1]
        out[..., r + 2, c + 1] = x[..., 0]
    return out


def _mv(M, x):
    """(..., 6, 6) @ (..., 6)."""
    return np.einsum("...ij,...j->...i", M, x)


def _mtv(M, x):
    """(..., 6, 6)^T @ (..., 6)."""
    return np.einsum("...ji,...j->...i", M, x)


class BatchOracle:
    """Vectorised numpy restatement; dtype float64 (or float32 to study rounding)."""

    def __init__(self, robot, dtype=np.float64):
        self.robot = robot
        self.dtype = np.dtype(dtype)
        self.NB = robot.get_num_bodies()
        self.n = robot.get_num_vel()
        self.parent = [robot.get_parent_id(i) for i in range(self.NB)]
        self.S = [_flat6(robot.get_S_by_id(i)).astype(self.dtype) for i in range(self.NB)]
        self.I = [np.array(robot.get_Imat_by_id(i), dtype=self.dtype) for i in range(self.NB)]
        self.subtree = [list(robot.get_subtree_by_id(i)) for i in range(self.NB)]
        self.damping = np.array([robot.get_damping_by_id(i) for i in range(self.NB)], dtype=self.dtype)
        self._fit_transforms()

    # X(q) of a 1-DoF joint is A + B*cos(q) + C*sin(q) (revolute) or A + B*q (prismatic);
    # the coefficients are recovered by probing the robot's own callable and verified.
    def _fit_transforms(self):
        self.kind, self.XA, self.XB, self.XC = [], [], [], []
        rng = np.random.default_rng(12345)
        for i in range(self.NB):
            fn = self.robot.get_Xmat_Func_by_id(i)
            S = _flat6(self.robot.get_S_by_id(i))
            revolute = bool(np.any(S[:3] != 0))
            X0 = np.asarray(fn(0.0), dtype=float)
            if revolute:
                Xh = np.asarray(fn(np.pi / 2), dtype=float)
                Xp = np.asarray(fn(np.pi), dtype=float)
                A = 0.5 * (X0 + Xp)
                B = 0.5 * (X0 - Xp)
                C = Xh - A
            else:
                X1 = np.asarray(fn(1.0), dtype=float)
                A, B, C = X0, X1 - X0, np.zeros((6, 6))
            for t in rng.uniform(-3.0, 3.0, size=4):
                f1, f2 = (np.cos(t), np.sin(t)) if revolute else (t, 0.0)
                err = np.max(np.abs(A + B * f1 + C * f2 - np.asarray(fn(t), dtype=float)))
                if err > 1e-12 * max(1.0, np.max(np.abs(X0))):
                    raise ValueError("joint %d: Xmat(q) is not of 1-DoF revolute/prismatic form" % i)
            self.kind.append(revolute)
            self.XA.append(A.astype(self.dtype))
            self.XB.append(B.astype(self.dtype))
            self.XC.append(C.astype(self.dtype))

    def _Xs(self, q):
        """All joint transforms for a batch: list over bodies of (B,6,6)."""
        out = []
        for i in range(self.NB):
            qi = q[:, i]
            if self.kind[i]:
                f1, f2 = np.cos(qi), np.sin(qi)
                out.append(self.XA[i] + f1[:, None, None] * self.XB[i] + f2[:, None, None] * self.XC[i])
            else:
                out.append(self.XA[i] + qi[:, None, None] * self.XB[i])
        return out

    def _prep(self, *arrs):
        return [None if x is None else np.ascontiguousarray(x, dtype=self.dtype) for x in arrs]

    # -- RNEA -----------------------------------------------------------------------------
    def rnea(self, q, qd, qdd=None, GRAVITY=-9.81, return_X=False):
        q, qd, qdd = self._prep(q, qd, qdd)
        B, NB = q.shape[0], self.NB
        Xs = self._Xs(q)
        v = np.zeros((B, 6, NB), dtype=self.dtype)
        a = np.zeros_like(v)
        f = np.zeros_like(v)
        a_base = np.zeros(6, dtype=self.dtype)
        a_base[5] = -GRAVITY
        for i in range(NB):
            p, S, X = self.parent[i], self.S[i], Xs[i]
            if p == -1:
                vi = np.zeros((B, 6), dtype=self.dtype)
                ai = _mv(X, np.broadcast_to(a_base, (B, 6)))
            else:
                vi = _mv(X, v[:, :, p])
                ai = _mv(X, a[:, :, p])
            vJ = qd[:, i, None] * S
            vi = vi + vJ
            ai = ai + _mv(_crm_b(vi), vJ)
            if qdd is not None:
                ai = ai + qdd[:, i, None] * S
            Iv = vi @ self.I[i].T
            f[:, :, i] = ai @ self.I[i].T + _mtv(-_crm_b(vi), Iv)
            v[:, :, i] = vi
            a[:, :, i] = ai
        c = np.zeros((B, self.n), dtype=self.dtype)
        for i in range(NB - 1, -1, -1):
            c[:, i] = f[:, :, i] @ self.S[i]
            p = self.parent[i]
            if p != -1:
                f[:, :, p] += _mtv(Xs[i], f[:, :, i])
        if return_X:
            return c, v, a, f, Xs
        return c, v, a, f

    # -- RNEA gradient --------------------------------------------------------------------
    def rnea_grad(self, q, qd, qdd=None, GRAVITY=-9.81, USE_VELOCITY_DAMPING=False, return_parts=False):
        q, qd, qdd = self._prep(q, qd, qdd)
        B, NB, n = q.shape[0], self.NB, self.n
        c, v, a, f, Xs = self.rnea(q, qd, qdd, GRAVITY, return_X=True)
        a_base = np.zeros(6, dtype=self.dtype)
        a_base[5] = -GRAVITY
        shape = (B, 6, n, NB)
        dv_q = np.zeros(shape, dtype=self.dtype)
        da_q = np.zeros(shape, dtype=self.dtype)
        df_q = np.zeros(shape, dtype=self.dtype)
        dv_d = np.zeros(shape, dtype=self.dtype)
        da_d = np.zeros(shape, dtype=self.dtype)
        df_d = np.zeros(shape, dtype=self.dtype)
        for i in range(NB):
            p, S, X, Ii = self.parent[i], self.S[i], Xs[i], self.I[i]
            qdi = qd[:, i]
            vi = v[:, :, i]
            crmS = lambda x: _mv(_crm_b(x), np.broadcast_to(S, x.shape))  # crm(x) @ S
            if p != -1:
                dv_q[:, :, :, i] = np.einsum("bij,bjc->bic", X, dv_q[:, :, :, p])
                dv_q[:, :, i, i] += crmS(_mv(X, v[:, :, p]))
                da_q[:, :, :, i] = np.einsum("bij,bjc->bic", X, da_q[:, :, :, p])
                dv_d[:, :, :, i] = np.einsum("bij,bjc->bic", X, dv_d[:, :, :, p])
                da_d[:, :, :, i] = np.einsum("bij,bjc->bic", X, da_d[:, :, :, p])
                Xa = _mv(X, a[:, :, p])
            else:
                Xa = _mv(X, np.broadcast_to(a_base, (B, 6)))
            dv_d[:, :, i, i] += S
            for (dv, da) in ((dv_q, da_q), (dv_d, da_d)):
                # da[:, c, i] += qd_i * crm(dv[:, c, i]) @ S  for every column c
                cols = np.swapaxes(dv[:, :, :, i], 1, 2)            # (B, n, 6)
                da[:, :, :, i] += qdi[:, None, None] * np.swapaxes(crmS(cols), 1, 2)
            da_q[:, :, i, i] += crmS(Xa)
            da_d[:, :, i, i] += crmS(vi)
            Iv = vi @ Ii.T
            crf_v = -np.swapaxes(_crm_b(vi), 1, 2)                  # (B,6,6)
            for (dv, da, df) in ((dv_q, da_q, df_q), (dv_d, da_d, df_d)):
                dvi = dv[:, :, :, i]                                # (B,6,n)
                df[:, :, :, i] = np.einsum("ij,bjc->bic", Ii, da[:, :, :, i])
                cols = np.swapaxes(dvi, 1, 2)                       # (B,n,6)
                crf_dv = -np.swapaxes(_crm_b(cols), 2, 3)           # (B,n,6,6)
                t1 = np.einsum("bcij,bj->bic", crf_dv, Iv)
                t2 = np.einsum("bij,jk,bkc->bic", crf_v, Ii, dvi)
                df[:, :, :, i] += t1 + t2
        parts = None
        if return_parts:
            parts = dict(dv_dq=dv_q.copy(), da_dq=da_q.copy(), df_dq=df_q.copy(),
                         dv_dqd=dv_d.copy(), da_dqd=da_d.copy(), df_dqd=df_d.copy(),
                         c=c, v=v, a=a, f=f)
        dc_dq = np.zeros((B, n, n), dtype=self.dtype)
        dc_dqd = np.zeros((B, n, n), dtype=self.dtype)
        for i in range(NB - 1, -1, -1):
            p, S, X = self.parent[i], self.S[i], Xs[i]
            dc_dq[:, i, :] = np.einsum("j,bjc->bc", S, df_q[:, :, :, i])
            dc_dqd[:, i, :] = np.einsum("j,bjc->bc", S, df_d[:, :, :, i])
            if p != -1:
                df_q[:, :, :, p] += np.einsum("bji,bjc->bic", X, df_q[:, :, :, i])
                df_d[:, :, :, p] += np.einsum("bji,bjc->bic", X, df_d[:, :, :, i])
                fxS = -_mv(_crm_b(f[:, :, i]), np.broadcast_to(S, (B, 6)))
                df_q[:, :, i, p] += _mtv(X, fxS)
        if USE_VELOCITY_DAMPING:
            idx = np.arange(n)
            dc_dqd[:, idx, idx] += self.damping
        dc_du = np.concatenate((dc_dq, dc_dqd), axis=2)
        if return_parts:
            parts.update(df_dq_acc=df_q, df_dqd_acc=df_d, dc_dq=dc_dq, dc_dqd=dc_dqd)
            return dc_du, parts
        return dc_du

    # -- Minv -----------------------------------------------------------------------------
    def minv(self, q, output_dense=True, return_parts=False):
        (q,) = self._prep(q)
        B, NB, n = q.shape[0], self.NB, self.n
        Xs = self._Xs(q)
        Minv = np.zeros((B, n, n), dtype=self.dtype)
        F = np.zeros((B, n, 6, n), dtype=self.dtype)
        U = np.zeros((B, n, 6), dtype=self.dtype)
        D = np.zeros((B, n), dtype=self.dtype)
        IA = [np.broadcast_to(self.I[i], (B, 6, 6)).copy() for i in range(NB)]
        for i in range(NB - 1, -1, -1):
            S, sub, p = self.S[i], self.subtree[i], self.parent[i]
            U[:, i] = IA[i] @ S
            D[:, i] = U[:, i] @ S
            Minv[:, i, i] = 1.0 / D[:, i]
            Minv[:, i, sub] -= (1.0 / D[:, i])[:, None] * np.einsum("j,bjs->bs", S, F[:, i][:, :, sub])
            if p != -1:
                X = Xs[i]
                F[:, i][:, :, sub] += U[:, i][:, :, None] * Minv[:, i, sub][:, None, :]
                F[:, p][:, :, sub] += np.einsum("bji,bjs->bis", X, F[:, i][:, :, sub])
                Ia = IA[i] - np.einsum("bi,bj->bij", U[:, i], U[:, i]) / D[:, i][:, None, None]
                IA[p] = IA[p] + np.einsum("bji,bjk,bkl->bil", X, Ia, X)
        parts = None
        if return_parts:
            parts = dict(Minv_b=Minv.copy(), F_b=F.copy(), U=U.copy(), D=D.copy())
        for i in range(NB):
            p, S = self.parent[i], self.S[i]
            if p != -1:
                X = Xs[i]
                UX = np.einsum("bj,bjk->bk", U[:, i], X)
                Minv[:, i, :] -= (1.0 / D[:, i])[:, None] * np.einsum("bk,bkc->bc", UX, F[:, p])
                F[:, i] = np.einsum("bij,bjc->bic", X, F[:, p]) + S[None, :, None] * Minv[:, i, None, :]
            else:
                F[:, i] = S[None, :, None] * Minv[:, i, None, :]
        if output_dense:
            iu = np.triu_indices(NB, 1)
            Minv[:, iu[1], iu[0]] = Minv[:, iu[0], iu[1]]
        if return_parts:
            parts.update(F_f=F)
            return Minv, parts
        return Minv

    # -- CRBA (identity tests) ------------------------------------------------------------
    def crba(self, q):
        (q,) = self._prep(q)
        B, n = q.shape[0], self.n
        Xs = self._Xs(q)
        IC = [np.broadcast_to(self.I[i], (B, 6, 6)).copy() for i in range(n)]
        for i in range(n - 1, -1, -1):
            p = self.parent[i]
            if p != -1:
                IC[p] = IC[p] + np.einsum("bji,bjk,bkl->bil", Xs[i], IC[i], Xs[i])
        H = np.zeros((B, n, n), dtype=self.dtype)
        for i in range(n):
            fh = IC[i] @ self.S[i]
            H[:, i, i] = fh @ self.S[i]
            j = i
            while self.parent[j] > -1:
                fh = _mtv(Xs[j], fh)
                j = self.parent[j]
                H[:, i, j] = fh @ self.S[j]
                H[:, j, i] = H[:, i, j]
        return H

    # -- ABA (fixed base), same quirk at :984 as ScalarOracle.aba -------------------------
    def aba(self, q, qd, tau, GRAVITY=-9.81):
        q, qd, tau = self._prep(q, qd, tau)
        B, n = q.shape[0], self.n
        Xs = self._Xs(q)
        crm_b = lambda x: np.stack([crm(x[k]) for k in range(x.shape[0])]).astype(self.dtype)
        v = [None] * n; c = [np.zeros((B, 6), dtype=self.dtype) for _ in range(n)]
        IA = [np.broadcast_to(self.I[i], (B, 6, 6)).copy() for i in range(n)]
        pA = [None] * n
        for i in range(n):
            p, S = self.parent[i], self.S[i]
            if p == -1:
                v[i] = qd[:, i:i + 1] * S
            else:
                v[i] = np.einsum("bij,bj->bi", Xs[i], v[p]) + qd[:, i:i + 1] * S
                c[i] = qd[:, i:i + 1] * (crm_b(v[i]) @ S)
            Iv = v[i] @ self.I[i].T
            full = np.einsum("bji,bj->bi", -crm_b(v[i]), Iv)               # crf(v) I v = -crm(v)^T I v
            pA[i] = np.repeat(full[:, :1], 6, axis=1)                      # :984
        U = [None] * n; d = [None] * n; u = [None] * n
        for i in range(n - 1, -1, -1):
            p, S = self.parent[i], self.S[i]
            U[i] = IA[i] @ S
            d[i] = U[i] @ S
            u[i] = tau[:, i] - pA[i] @ S
            if p != -1:
                Ia = IA[i] - U[i][:, :, None] * U[i][:, None, :] / d[i][:, None, None]
                pa = pA[i] + np.einsum("bij,bj->bi", Ia, c[i]) + U[i] * (u[i] / d[i])[:, None]
                IA[p] = IA[p] + np.einsum("bji,bjk,bkl->bil", Xs[i], Ia, Xs[i])
                pA[p] = pA[p] + np.einsum("bji,bj->bi", Xs[i], pa)
        g = np.zeros(6, dtype=self.dtype); g[5] = -GRAVITY
        a = [None] * n
        qdd = np.zeros((B, n), dtype=self.dtype)
        for i in range(n):
            p = self.parent[i]
            ap = np.broadcast_to(g, (B, 6)) if p == -1 else a[p]
            ai = np.einsum("bij,bj->bi", Xs[i], ap) + c[i]
            qdd[:, i] = (u[i] - np.einsum("bi,bi->b", U[i], ai)) / d[i]
            a[i] = ai + qdd[:, i:i + 1] * self.S[i]
        return qdd

    # -- end-effector kinematics, vectorised (same chain products as ScalarOracle) ---------------
    def _fit_hom(self):
        """T(q) and dT(q) of every joint as A + B cos q + C sin q / A + B q, by probing."""
        if hasattr(self, "_hom"):
            return self._hom
        rng = np.random.default_rng(999)
        out = []
        for i in range(self.NB):
            coefs = []
            for getter in (self.robot.get_Xmat_hom_Func_by_id, self.robot.get_dXmat_hom_Func_by_id):
                fn = getter(i)
                T0 = np.asarray(fn(0.0), dtype=float)
                if self.kind[i]:
                    Th, Tp = np.asarray(fn(np.pi / 2), dtype=float), np.asarray(fn(np.pi), dtype=float)
                    A, Bm = 0.5 * (T0 + Tp), 0.5 * (T0 - Tp)
                    C = Th - A
                else:
                    A, Bm, C = T0, np.asarray(fn(1.0), dtype=float) - T0, np.zeros((4, 4))
                for t in rng.uniform(-3.0, 3.0, size=4):
                    f1, f2 = (np.cos(t), np.sin(t)) if self.kind[i] else (t, 0.0)
                    if np.max(np.abs(A + Bm * f1 + C * f2 - np.asarray(fn(t), dtype=float))) > 1e-12 * max(1.0, np.max(np.abs(T0))):
                        raise ValueError("joint %d: homogeneous transform is not of 1-DoF form" % i)
                coefs.append((A.astype(self.dtype), Bm.astype(self.dtype), C.astype(self.dtype)))
            out.append(coefs)
        self._hom = out
        return out

    def _hom_eval(self, i, qi, which):
        A, Bm, C = self._fit_hom()[i][which]
        if self.kind[i]:
            return A + np.cos(qi)[:, None, None] * Bm + np.sin(qi)[:, None, None] * C
        return A + qi[:, None, None] * Bm

    def end_effector_pose(self, q, ee_joint_names=None, ee_offsets=((0, 0, 0, 1),), gradient=False):
        """-> pose (B, n_ee, 6) [and gradient (B, n_ee, 6, n)]."""
        (q,) = self._prep(q)
        B, n = q.shape[0], self.n
        so = ScalarOracle(self.robot)
        targets = so._ee_targets(ee_joint_names)
        off = so._ee_offset(ee_offsets).astype(self.dtype)
        pose = np.zeros((B, len(targets), 6), dtype=self.dtype)
        grad = np.zeros((B, len(targets), 6, n), dtype=self.dtype)

        def darctan2(y, x, yp, xp):
            return (-xp * y + x * yp) / (x * x + y * y)

        for e, (jid, fin) in enumerate(targets):
            chain = [jid] + list(self.robot.get_ancestors_by_id(jid))
            chain = sorted(chain, reverse=True)            # leaf first (ids are topologically ordered)
            T = {k: self._hom_eval(k, q[:, k], 0) for k in chain}
            X = np.broadcast_to(np.asarray(fin, dtype=self.dtype), (B, 4, 4))
            for k in chain:
                X = T[k] @ X
            pose[:, e, :3] = (X @ off)[:, :3]
            pose[:, e, 3] = np.arctan2(X[:, 2, 1], X[:, 2, 2])
            sq = np.sqrt(X[:, 2, 2] * X[:, 2, 2] + X[:, 2, 1] * X[:, 2, 1])
            pose[:, e, 4] = np.arctan2(-X[:, 2, 0], sq)
            pose[:, e, 5] = np.arctan2(X[:, 1, 0], X[:, 0, 0])
            if not gradient:
                continue
            for dind in chain:
                dX = np.broadcast_to(np.asarray(fin, dtype=self.dtype), (B, 4, 4))
                for k in chain:
                    dX = (self._hom_eval(k, q[:, k], 1) if k == dind else T[k]) @ dX
                grad[:, e, :3, dind] = (dX @ off)[:, :3]
                grad[:, e, 3, dind] = darctan2(X[:, 2, 1], X[:, 2, 2], dX[:, 2, 1], dX[:, 2, 2])
                dsq = (X[:, 2, 2] * dX[:, 2, 2] + X[:, 2, 1] * dX[:, 2, 1]) / sq
                grad[:, e, 4, dind] = darctan2(-X[:, 2, 0], sq, -dX[:, 2, 0], dsq)
                grad[:, e, 5, dind] = darctan2(X[:, 1, 0], X[:, 0, 0], dX[:, 1, 0], dX[:, 0, 0])
        return (pose, grad) if gradient else pose

    def end_effector_pose_gradient(self, q, ee_joint_names=None, ee_offsets=((0, 0, 0, 1),)):
        return self.end_effector_pose(q, ee_joint_names, ee_offsets, gradient=True)[1]
