"""CPU oracle for the FLOATING-BASE branches of rnea / rnea_grad / minv.  TEST INFRASTRUCTURE ONLY.

From-scratch numpy restatement of the `self.robot.floating_base` branches of
/root/reference/RBDReference.py: `rnea_fpass` :585/:591, `minv_bpass`
:652-691, `minv_fpass` :761-779, `rnea_grad_fpass_dq` :1141-1168, `rnea_grad_fpass_dqd`
:1212-1238, `rnea_grad_bpass_dq` :1267-1282, `rnea_grad_bpass_dqd` :1309-1341.  Only tests may
import it (same rule as oracle/rbd_oracle.py).

Pinning: tests/golden/fb_*.npz hold outputs of the unmodified reference for floating-base robots
(oracle/make_golden.py --fb); tests/test_oracle.py checks this module against them.

Conventions the reference's branches assume (taken from its index arithmetic): body 0 is the base
with S = eye(6), `get_joint_index_q(0)` = 7 indices, `get_joint_index_v(0)` = 6 indices; body
i >= 1 owns row / column i + 5 of every joint-space quantity.  n = NB + 5.

Reference behaviour kept on purpose (each is what the cited line does, not what one would write):
  * :1166-1168  the base's q-columns add `_mxS(S[ii], dv_dq[:, c, ii], qd[ii])` into BODY ii for
    ii = 0..5 while every dv_dq is still zero: a no-op that raises IndexError when NB < 6;
  * :799-804    `output_dense` mirrors only the leading NB x NB block (matrix indices, not bodies);
  * :1338-1341  velocity damping is added at [ind, ind] (body index, not ind + 5) and, for the
    base, to the whole block [0:5, 0:5].
"""
from __future__ import annotations

import numpy as np

from .rbd_oracle import crf, crm, _flat6

__all__ = ["FloatingScalarOracle"]


class FloatingScalarOracle:
    """Single-state restatement of the reference's floating-base paths (reference shapes)."""

    def __init__(self, robot):
        if not getattr(robot, "floating_base", False):
            raise ValueError("FloatingScalarOracle needs a floating-base robot")
        self.robot = robot
        self.NB = robot.get_num_bodies()
        self.n = self.NB + 5                                     # :653
        self.parent = [robot.get_parent_id(i) for i in range(self.NB)]
        if self.parent[0] != -1 or any(p < 0 for p in self.parent[1:]):
            raise ValueError("body 0 must be the only root (the floating base)")
        self.S = [None] + [_flat6(robot.get_S_by_id(i)) for i in range(1, self.NB)]
        self.I = [np.array(robot.get_Imat_by_id(i), dtype=float) for i in range(self.NB)]
        self.subtree = [list(robot.get_subtree_by_id(i)) for i in range(self.NB)]

    def _X(self, i, q):
        return np.asarray(self.robot.get_Xmat_Func_by_id(i)(q[self.robot.get_joint_index_q(i)]), dtype=float)

    @staticmethod
    def _gravity(GRAVITY):
        g = np.zeros(6)
        g[5] = -GRAVITY                                          # :566
        return g

    # -- RNEA -----------------------------------------------------------------------------
    def rnea_fpass(self, q, qd, qdd=None, GRAVITY=-9.81):
        """RBDReference.py:559-598 with the `floating_base and curr_id == 0` branches."""
        NB = self.NB
        v, a, f = np.zeros((6, NB)), np.zeros((6, NB)), np.zeros((6, NB))
        for i in range(NB):
            X = self._X(i, q)
            p = self.parent[i]
            if p == -1:
                vi, ai = np.zeros(6), X @ self._gravity(GRAVITY)                 # :577-578
            else:
                vi, ai = X @ v[:, p], X @ a[:, p]                                # :580-581
            vJ = qd[0:6].copy() if i == 0 else self.S[i] * qd[i + 5]             # :585-586 (S = eye(6))
            vi = vi + vJ
            ai = ai + crm(vi) @ vJ                                               # :588
            if qdd is not None:
                ai = ai + (qdd[0:6] if i == 0 else self.S[i] * qdd[i + 5])       # :591-593
            Ii = self.I[i]
            f[:, i] = Ii @ ai + crf(vi) @ (Ii @ vi)                              # :596
            v[:, i], a[:, i] = vi, ai
        return v, a, f

    def rnea_bpass(self, q, f):
        """RBDReference.py:600-621; c[0:6] = S^T f_0 = f_0 for the base."""
        c = np.zeros(self.n)
        for i in range(self.NB - 1, -1, -1):
            if i == 0:
                c[0:6] = f[:, 0]                                                 # :612 with S = eye(6)
            else:
                c[i + 5] = self.S[i] @ f[:, i]
                p = self.parent[i]
                f[:, p] = f[:, p] + self._X(i, q).T @ f[:, i]                    # :617-619
        return c, f

    def rnea(self, q, qd, qdd=None, GRAVITY=-9.81, f_ext=None):
        v, a, f = self.rnea_fpass(q, qd, qdd, GRAVITY)
        c, f = self.rnea_bpass(q, f)
        return c, v, a, f

    # -- Minv -----------------------------------------------------------------------------
    def minv_bpass(self, q):
        """RBDReference.py:630-735."""
        n, NB = self.n, self.NB
        Minv, F, U, D = np.zeros((n, n)), np.zeros((n, 6, n)), np.zeros((n, 6)), np.zeros(n)
        IA = {i: self.I[i].copy() for i in range(NB)}
        for i in range(NB - 1, -1, -1):
            adj = [j + 5 for j in self.subtree[i]]                               # :668-671
            if i == 0:
                U[0:6, :] = IA[0]                                                # :681 (S = eye(6))
                fb_Dinv = np.linalg.inv(U[0:6, :])                               # :682-684
                Minv[0:6, 0:6] = Minv[0, 0] + fb_Dinv                            # :686
                Minv[0:6, adj] -= fb_Dinv @ F[5][:, adj]                         # :687-691 (the [-1] picks F[5])
                continue
            mi, S = i + 5, self.S[i]
            U[mi] = IA[i] @ S                                                    # :697
            D[mi] = S @ U[mi]                                                    # :698
            Minv[mi, mi] = 1.0 / D[mi]                                           # :700
            Minv[mi, adj] -= (1.0 / D[mi]) * (S @ F[mi][:, adj])                 # :702-708
            mp = self.parent[i] + 5                                              # :713
            X = self._X(i, q)
            for j in adj:                                                        # :720-726
                F[mi][:, j] += U[mi] * Minv[mi, j]
                F[mp][:, j] += X.T @ F[mi][:, j]
            Ia = IA[i] - np.outer(U[mi], U[mi]) / D[mi]                          # :728-731
            IA[self.parent[i]] = IA[self.parent[i]] + X.T @ Ia @ X               # :732-733
        return Minv, F, U, D

    def minv_fpass(self, q, Minv, F, U, Dinv):
        """RBDReference.py:737-783; F is re-used with BODY indices (F[ind], F[parent_ind])."""
        for i in range(self.NB):
            if i == 0:
                F[0] = Minv[0:6, 0:]                                             # :779 (S = eye(6))
                continue
            mi, p = i + 5, self.parent[i]
            X = self._X(i, q)
            Minv[mi, :] -= (1.0 / Dinv[mi]) * ((U[mi] @ X) @ F[p])               # :771-773
            F[i] = X @ F[p] + np.outer(self.S[i], Minv[mi, :])                   # :774-776
        return Minv

    def minv(self, q, output_dense=True):
        Minv, F, U, D = self.minv_bpass(q)
        Minv = self.minv_fpass(q, Minv, F, U, D)
        if output_dense:
            iu = np.triu_indices(self.NB, 1)                                     # :799-804: range(NB), not n
            Minv[iu[1], iu[0]] = Minv[iu]
        return Minv

    # -- RNEA gradient ----------------------------------------------------------------------
    def _df(self, i, v, dv, da):
        Ii = self.I[i]
        Iv = Ii @ v[:, i]
        df = Ii @ da
        for c in range(self.n):                                                  # :1182-1185 / :1249-1252
            df[:, c] += crf(dv[:, c]) @ Iv + crf(v[:, i]) @ (Ii @ dv[:, c])
        return df

    def rnea_grad_fpass_dq(self, q, qd, v, a, GRAVITY=-9.81):
        """RBDReference.py:1127-1187."""
        n, NB = self.n, self.NB
        if NB < 6:
            raise IndexError("the reference indexes bodies 0..5 at :1168 - it needs NB >= 6")
        dv, da, df = np.zeros((6, n, NB)), np.zeros((6, n, NB)), np.zeros((6, n, NB))
        for i in range(NB):
            X = self._X(i, q)
            if i == 0:
                # :1164-1168 adds zeros (every dv_dq is still zero); :1175 with S = eye(6)
                da[:, 0:6, 0] += crm(X @ self._gravity(GRAVITY))
            else:
                p, S, idx = self.parent[i], self.S[i], i + 5
                dv[:, :, i] = X @ dv[:, :, p]                                    # :1158
                dv[:, idx, i] += crm(X @ v[:, p]) @ S                            # :1159
                da[:, :, i] = X @ da[:, :, p]                                    # :1163
                for c in range(n):
                    da[:, c, i] += qd[idx] * (crm(dv[:, c, i]) @ S)              # :1170
                da[:, idx, i] += crm(X @ a[:, p]) @ S                            # :1173
            df[:, :, i] = self._df(i, v, dv[:, :, i], da[:, :, i])
        return dv, da, df

    def rnea_grad_fpass_dqd(self, q, qd, v):
        """RBDReference.py:1189-1255."""
        n, NB = self.n, self.NB
        dv, da, df = np.zeros((6, n, NB)), np.zeros((6, n, NB)), np.zeros((6, n, NB))
        for i in range(NB):
            X = self._X(i, q)
            if i == 0:
                dv[:, 0:6, 0] += np.eye(6)                                       # :1231
                for c in range(n):
                    for ii in range(6):                                          # :1236-1238
                        da[:, c, 0] += qd[ii] * crm(dv[:, c, 0])[:, ii]
                da[:, 0:6, 0] += crm(v[:, 0])                                    # :1243 with S = eye(6)
            else:
                p, S, idx = self.parent[i], self.S[i], i + 5
                dv[:, :, i] = X @ dv[:, :, p]                                    # :1230
                dv[:, idx, i] += S                                               # :1231
                da[:, :, i] = X @ da[:, :, p]                                    # :1234
                for c in range(n):
                    da[:, c, i] += qd[idx] * (crm(dv[:, c, i]) @ S)              # :1240
                da[:, idx, i] += crm(v[:, i]) @ S                                # :1243
            df[:, :, i] = self._df(i, v, dv[:, :, i], da[:, :, i])
        return dv, da, df

    def rnea_grad_bpass_dq(self, q, f, df_dq):
        """RBDReference.py:1257-1297."""
        n = self.n
        dc = np.zeros((n, n))
        for i in range(self.NB - 1, -1, -1):
            if i == 0:
                dc[:6] = df_dq[:, :, 0]                                          # :1282
                continue
            S, idx, p = self.S[i], i + 5, self.parent[i]
            dc[idx, :] = S @ df_dq[:, :, i]                                      # :1284
            X = self._X(i, q)
            df_dq[:, :, p] += X.T @ df_dq[:, :, i]                               # :1291
            df_dq[:, idx, p] += X.T @ (-(crm(f[:, i]) @ S))                      # :1292-1294
        return dc

    def rnea_grad_bpass_dqd(self, q, df_dqd, USE_VELOCITY_DAMPING=False):
        """RBDReference.py:1299-1343."""
        n = self.n
        dc = np.zeros((n, n))
        for i in range(self.NB - 1, -1, -1):
            if i == 0:
                dc[0:6, :] = df_dqd[:, :, 0]                                     # :1325 with S = eye(6)
                continue
            dc[i + 5, :] = self.S[i] @ df_dqd[:, :, i]
            df_dqd[:, :, self.parent[i]] += self._X(i, q).T @ df_dqd[:, :, i]    # :1331
        if USE_VELOCITY_DAMPING:
            for i in range(self.NB):                                             # :1336-1341
                if i == 0:
                    dc[0:5, 0:5] += self.robot.get_damping_by_id(0)
                else:
                    dc[i, i] += self.robot.get_damping_by_id(i)
        return dc

    def rnea_grad(self, q, qd, qdd=None, GRAVITY=-9.81, USE_VELOCITY_DAMPING=False):
        c, v, a, f = self.rnea(q, qd, qdd, GRAVITY)
        _, _, df_dq = self.rnea_grad_fpass_dq(q, qd, v, a, GRAVITY)
        _, _, df_dqd = self.rnea_grad_fpass_dqd(q, qd, v)
        dc_dq = self.rnea_grad_bpass_dq(q, f, df_dq)
        dc_dqd = self.rnea_grad_bpass_dqd(q, df_dqd, USE_VELOCITY_DAMPING)
        return np.hstack((dc_dq, dc_dqd))

    # -- compositions (robot-agnostic upstream, RBDReference.py:1369-1384) ---------------------
    def forward_dynamics(self, q, qd, u):
        c = self.rnea(q, qd)[0]                                                  # :1370 (qdd omitted)
        return self.minv(q) @ (u - c)                                            # :1371-1372

    def forward_dynamics_grad(self, q, qd, u):
        qdd = self.forward_dynamics(q, qd, u)                                    # :1377
        dc_du = self.rnea_grad(q, qd, qdd)                                       # :1378
        Minv = self.minv(q)                                                      # :1381
        return -Minv @ dc_du[:, : self.n], -Minv @ dc_du[:, self.n:]             # :1379 splits at len(qd)
