// rbd_fused_kernels.cuh - fused drivers: rnea, rnea_grad, minv in one launch each.
//
// Generic-topology versions (any fixed-base tree with n <= RBD_MAX_DOF): one thread per knot
// point, joint transforms built once per joint from one sincos, all intermediates on chip
// (registers + per-thread local arrays); only q/qd/qdd are read and only the requested outputs
// are written.  Derivative columns are ancestor-sparse: column c is only propagated through
// subtree(c) on the way down and subtree(c) + ancestors(c) on the way up
// (SURVEY.md Appendix A.2).
#pragma once
#include "rbd_common.cuh"

namespace rbd {

constexpr int kFusedThreads = 128;

// =============================================================================================
// rnea (RBDReference.py:623-628 = :559-598 + :600-621)
// =============================================================================================
template <typename T>
__global__ void __launch_bounds__(kFusedThreads)
rnea_fused_kernel(const __grid_constant__ DevModel<T> m, int64_t B, const T* __restrict__ q,
                  const T* __restrict__ qd, const T* __restrict__ qdd, T gravity,
                  T* __restrict__ c, T* __restrict__ v, T* __restrict__ a, T* __restrict__ f) {
  __shared__ T tiles[kFusedThreads / 32][32 * (kFlushChunk + 1)];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t b0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) - lane;
  if (b0 >= B) return;                                     // whole warp past the end
  const int nlive = (int)(B - b0 < 32 ? B - b0 : 32);
  const bool live = lane < nlive;
  const int64_t b = b0 + (live ? lane : 0);                // idle lanes shadow a valid knot point
  const int n = m.n;
  const T* qb = q + b * n;
  const T* qdb = qd + b * n;
  const T* qddb = qdd ? qdd + b * n : nullptr;
  T lv[6 * RBD_MAX_DOF], la[6 * RBD_MAX_DOF], lf[6 * RBD_MAX_DOF], lb[RBD_MAX_DOF][2];   // (6, NB): (r, i) at r*n + i
  for (int i = 0; i < n; ++i) {
    T X[18], f1, f2;
    joint_basis(m, i, qb[i], f1, f2);
    lb[i][0] = f1; lb[i][1] = f2;
    build_X(m, i, f1, f2, X);
    const int p = m.parent[i];
    T vi[6], ai[6], par[6];
    if (p < 0) {
#pragma unroll
      for (int r = 0; r < 6; ++r) { vi[r] = T(0); par[r] = T(0); }
      par[5] = -gravity;
      X_apply(X, par, ai);
    } else {
#pragma unroll
      for (int r = 0; r < 6; ++r) par[r] = lv[r * n + p];
      X_apply(X, par, vi);
#pragma unroll
      for (int r = 0; r < 6; ++r) par[r] = la[r * n + p];
      X_apply(X, par, ai);
    }
    T vJ[6], t[6];
    const T qdi = qdb[i];
#pragma unroll
    for (int r = 0; r < 6; ++r) { vJ[r] = m.S[i][r] * qdi; vi[r] += vJ[r]; }
    crm_mul(vi, vJ, t);
#pragma unroll
    for (int r = 0; r < 6; ++r) ai[r] += t[r];
    if (qddb) {
      const T qddi = qddb[i];
#pragma unroll
      for (int r = 0; r < 6; ++r) ai[r] = fma_t(m.S[i][r], qddi, ai[r]);
    }
    T Ia[6], Iv[6], vxIv[6];
    mat6_apply(m.I[i], ai, Ia);
    mat6_apply(m.I[i], vi, Iv);
    crf_mul(vi, Iv, vxIv);
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      lv[r * n + i] = vi[r];
      la[r * n + i] = ai[r];
      lf[r * n + i] = Ia[r] + vxIv[r];
    }
  }
  T* cb = c + b * n;
  for (int i = n - 1; i >= 0; --i) {
    T fi[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) fi[r] = lf[r * n + i];
    if (live) cb[i] = dot6(m.S[i], fi);
    const int p = m.parent[i];
    if (p >= 0) {
      T X[18], t[6];
      build_X(m, i, lb[i][0], lb[i][1], X);
      XT_apply(X, fi, t);
#pragma unroll
      for (int r = 0; r < 6; ++r) lf[r * n + p] += t[r];
    }
  }
  // v / a / f: coalesced 256-byte rows through a shared-memory transpose (warp_flush_blocks)
  if (v) warp_flush_blocks(lv, 6 * n, tiles[warp], v + b0 * 6 * n, nlive, lane);
  if (a) warp_flush_blocks(la, 6 * n, tiles[warp], a + b0 * 6 * n, nlive, lane);
  if (f) warp_flush_blocks(lf, 6 * n, tiles[warp], f + b0 * 6 * n, nlive, lane);
}

// =============================================================================================
// rnea_grad (RBDReference.py:1345-1368): rnea + the four gradient passes, one launch.
// Column c (both the d/dq and the d/dqd column) is swept down subtree(c) and back up to the
// root; body-frame recursion exactly as :1127-1343.
// =============================================================================================
template <typename T>
__global__ void __launch_bounds__(kFusedThreads)
rnea_grad_fused_kernel(const __grid_constant__ DevModel<T> m, int64_t B, const T* __restrict__ q,
                       const T* __restrict__ qd, const T* __restrict__ qdd, T gravity,
                       int use_damping, T* __restrict__ dc_du, T* __restrict__ c_out) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int n = m.n;
  const T* qb = q + b * n;
  const T* qdb = qd + b * n;
  const T* qddb = qdd ? qdd + b * n : nullptr;
  // ---- RNEA state kept for the gradient sweeps ----
  T lv[RBD_MAX_DOF][6], la[RBD_MAX_DOF][6], lf[RBD_MAX_DOF][6], lb[RBD_MAX_DOF][3];
  for (int i = 0; i < n; ++i) {
    T X[18], f1, f2;
    joint_basis(m, i, qb[i], f1, f2);
    const T qdi = qdb[i];
    lb[i][0] = f1; lb[i][1] = f2; lb[i][2] = qdi;
    build_X(m, i, f1, f2, X);
    const int p = m.parent[i];
    T vi[6], ai[6], par[6];
    if (p < 0) {
#pragma unroll
      for (int r = 0; r < 6; ++r) { vi[r] = T(0); par[r] = T(0); }
      par[5] = -gravity;
      X_apply(X, par, ai);
    } else {
#pragma unroll
      for (int r = 0; r < 6; ++r) par[r] = lv[p][r];
      X_apply(X, par, vi);
#pragma unroll
      for (int r = 0; r < 6; ++r) par[r] = la[p][r];
      X_apply(X, par, ai);
    }
    T vJ[6], t[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) { vJ[r] = m.S[i][r] * qdi; vi[r] += vJ[r]; }
    crm_mul(vi, vJ, t);
#pragma unroll
    for (int r = 0; r < 6; ++r) ai[r] += t[r];
    if (qddb) {
      const T qddi = qddb[i];
#pragma unroll
      for (int r = 0; r < 6; ++r) ai[r] = fma_t(m.S[i][r], qddi, ai[r]);
    }
    T Ia[6], Iv[6], vxIv[6];
    mat6_apply(m.I[i], ai, Ia);
    mat6_apply(m.I[i], vi, Iv);
    crf_mul(vi, Iv, vxIv);
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      lv[i][r] = vi[r];
      la[i][r] = ai[r];
      lf[i][r] = Ia[r] + vxIv[r];
    }
  }
  for (int i = n - 1; i >= 0; --i) {
    T fi[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) fi[r] = lf[i][r];
    if (c_out) c_out[b * n + i] = dot6(m.S[i], fi);
    const int p = m.parent[i];
    if (p >= 0) {
      T X[18], t[6];
      build_X(m, i, lb[i][0], lb[i][1], X);
      XT_apply(X, fi, t);
#pragma unroll
      for (int r = 0; r < 6; ++r) lf[p][r] += t[r];
    }
  }
  // ---- gradient columns ----
  T* out = dc_du + b * (int64_t)2 * n * n;      // (n, 2n): [i][c] = dc_dq, [i][n+c] = dc_dqd
  T sdv[RBD_MAX_DOF][12];                        // per body: dv_dq(6) | dv_dqd(6) for column c
  T sda[RBD_MAX_DOF][12];
  T sdf[RBD_MAX_DOF][12];
  for (int c = 0; c < n; ++c) {
    const unsigned sub = m.sub_mask[c];
    // forward over subtree(c), ascending ids (parents first)
    for (int i = c; i < n; ++i) {
      if (!((sub >> i) & 1u)) continue;
      T X[18], S[6], vi[6];
      build_X(m, i, lb[i][0], lb[i][1], X);
      const T qdi = lb[i][2];
#pragma unroll
      for (int r = 0; r < 6; ++r) { S[r] = m.S[i][r]; vi[r] = lv[i][r]; }
      T dvq[6], daq[6], dvd[6], dad[6], t[6];
      const int p = m.parent[i];
      if (i == c) {
        // d/dq seeds: dv = crm(X v_p) S (:1159), da = crm(X a_p) S (:1173/:1175)
        T par[6], xp[6];
        if (p >= 0) {
#pragma unroll
          for (int r = 0; r < 6; ++r) par[r] = lv[p][r];
          X_apply(X, par, xp);
          crm_mul(xp, S, dvq);
#pragma unroll
          for (int r = 0; r < 6; ++r) par[r] = la[p][r];
        } else {
#pragma unroll
          for (int r = 0; r < 6; ++r) { dvq[r] = T(0); par[r] = T(0); }
          par[5] = -gravity;
        }
        X_apply(X, par, xp);
        crm_mul(xp, S, daq);
        // d/dqd seeds: dv = S (:1231), da = crm(v_i) S (:1243)
#pragma unroll
        for (int r = 0; r < 6; ++r) dvd[r] = S[r];
        crm_mul(vi, S, dad);
      } else {
        T pv[6];
#pragma unroll
        for (int r = 0; r < 6; ++r) pv[r] = sdv[p][r];
        X_apply(X, pv, dvq);
#pragma unroll
        for (int r = 0; r < 6; ++r) pv[r] = sda[p][r];
        X_apply(X, pv, daq);
#pragma unroll
        for (int r = 0; r < 6; ++r) pv[r] = sdv[p][6 + r];
        X_apply(X, pv, dvd);
#pragma unroll
        for (int r = 0; r < 6; ++r) pv[r] = sda[p][6 + r];
        X_apply(X, pv, dad);
        for (int r = 0; r < 6; ++r) t[r] = T(0);
      }
      // da += qd_i * crm(dv) S   (:1170 / :1240)
      crm_mul(dvq, S, t);
#pragma unroll
      for (int r = 0; r < 6; ++r) daq[r] = fma_t(qdi, t[r], daq[r]);
      crm_mul(dvd, S, t);
#pragma unroll
      for (int r = 0; r < 6; ++r) dad[r] = fma_t(qdi, t[r], dad[r]);
      // df = I da + crf(dv) Iv + crf(v) I dv   (:1179-1185 / :1247-1252)
      T Iv[6], Ida[6], Idv[6], t1[6], t2[6];
      mat6_apply(m.I[i], vi, Iv);
      mat6_apply(m.I[i], daq, Ida);
      mat6_apply(m.I[i], dvq, Idv);
      crf_mul(dvq, Iv, t1);
      crf_mul(vi, Idv, t2);
#pragma unroll
      for (int r = 0; r < 6; ++r) {
        sdv[i][r] = dvq[r];
        sda[i][r] = daq[r];
        sdf[i][r] = Ida[r] + t1[r] + t2[r];
      }
      mat6_apply(m.I[i], dad, Ida);
      mat6_apply(m.I[i], dvd, Idv);
      crf_mul(dvd, Iv, t1);
      crf_mul(vi, Idv, t2);
#pragma unroll
      for (int r = 0; r < 6; ++r) {
        sdv[i][6 + r] = dvd[r];
        sda[i][6 + r] = dad[r];
        sdf[i][6 + r] = Ida[r] + t1[r] + t2[r];
      }
    }
    // backward over subtree(c), descending ids, then up the ancestors of c
    T Fq[6], Fd[6];
    for (int i = n - 1; i >= c; --i) {
      if (!((sub >> i) & 1u)) continue;
#pragma unroll
      for (int r = 0; r < 6; ++r) { Fq[r] = sdf[i][r]; Fd[r] = sdf[i][6 + r]; }
      out[i * 2 * n + c] = dot6(m.S[i], Fq);                                   // :1284
      T dd = dot6(m.S[i], Fd);                                                 // :1325
      if (use_damping && i == c) dd += m.damping[i];                           // :1341
      out[i * 2 * n + n + c] = dd;
      const int p = m.parent[i];
      if (p < 0) break;   // only possible for i == c (root column)
      T X[18], tq[6], td[6];
      build_X(m, i, lb[i][0], lb[i][1], X);
      if (i == c) {
        // extra term of column idx == i pushed to the parent: X^T(-crm(f_i) S)   (:1292-1294)
        T fi[6], S[6], fxs[6];
#pragma unroll
        for (int r = 0; r < 6; ++r) { fi[r] = lf[i][r]; S[r] = m.S[i][r]; }
        crm_mul(fi, S, fxs);
#pragma unroll
        for (int r = 0; r < 6; ++r) Fq[r] -= fxs[r];
      }
      XT_apply(X, Fq, tq);                                                     // :1291
      XT_apply(X, Fd, td);                                                     // :1331
      if (i == c) {
#pragma unroll
        for (int r = 0; r < 6; ++r) { Fq[r] = tq[r]; Fd[r] = td[r]; }
      } else {
#pragma unroll
        for (int r = 0; r < 6; ++r) { sdf[p][r] += tq[r]; sdf[p][6 + r] += td[r]; }
      }
    }
    // ancestors of c: the column only carries what came up from subtree(c)
    for (int j = m.parent[c]; j >= 0;) {
      out[j * 2 * n + c] = dot6(m.S[j], Fq);
      out[j * 2 * n + n + c] = dot6(m.S[j], Fd);
      const int pj = m.parent[j];
      if (pj >= 0) {
        T X[18], tq[6], td[6];
        build_X(m, j, lb[j][0], lb[j][1], X);
        XT_apply(X, Fq, tq);
        XT_apply(X, Fd, td);
#pragma unroll
        for (int r = 0; r < 6; ++r) { Fq[r] = tq[r]; Fd[r] = td[r]; }
      }
      j = pj;
    }
    // rows on other branches are structurally zero
    const unsigned touched = sub | m.anc_mask[c];
    for (int i = 0; i < n; ++i) {
      if ((touched >> i) & 1u) continue;
      out[i * 2 * n + c] = T(0);
      out[i * 2 * n + n + c] = T(0);
    }
  }
}

// =============================================================================================
// minv (RBDReference.py:785-806 = :630-735 + :737-783 + mirror)
// Phase A: articulated-inertia recursion leaf->root gives U_i, D_i.
// Phase B: every column j is an independent chain walk up (backward pass) and a sweep down
// (forward pass); F is a 6-vector per visited body instead of the reference's (n,6,n) tensor.
// =============================================================================================
template <typename T>
__global__ void __launch_bounds__(kFusedThreads)
minv_fused_kernel(const __grid_constant__ DevModel<T> m, int64_t B, const T* __restrict__ q,
                  int output_dense, T* __restrict__ Minv) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int n = m.n;
  const T* qb = q + b * n;
  T* Mb = Minv + b * (int64_t)n * n;
  T IA[RBD_MAX_DOF][36];
  T lb[RBD_MAX_DOF][2];
  T lU[RBD_MAX_DOF][6], lUX[RBD_MAX_DOF][6], linvD[RBD_MAX_DOF];
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int k = 0; k < 36; ++k) IA[i][k] = m.I[i][k];
    T f1, f2;
    joint_basis(m, i, qb[i], f1, f2);
    lb[i][0] = f1; lb[i][1] = f2;
  }
  // ---- phase A ----
  for (int i = n - 1; i >= 0; --i) {
    T S[6], Ui[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) S[r] = m.S[i][r];
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      T acc = T(0);
#pragma unroll
      for (int k = 0; k < 6; ++k) acc = fma_t(IA[i][6 * r + k], S[k], acc);
      Ui[r] = acc;
      lU[i][r] = acc;
    }
    const T invD = T(1) / dot6(S, Ui);
    linvD[i] = invD;
    const int p = m.parent[i];
    if (p >= 0) {
      T X[18], UX[6];
      build_X(m, i, lb[i][0], lb[i][1], X);
      XT_apply(X, Ui, UX);
#pragma unroll
      for (int r = 0; r < 6; ++r) lUX[i][r] = UX[r];
      T Xf[6][6];
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
          Xf[r][cc] = X[3 * r + cc];
          Xf[r][3 + cc] = T(0);
          Xf[3 + r][cc] = X[9 + 3 * r + cc];
          Xf[3 + r][3 + cc] = X[3 * r + cc];
        }
      T Ia[6][6], tmp[6][6];
#pragma unroll
      for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int cc = 0; cc < 6; ++cc) Ia[r][cc] = IA[i][6 * r + cc] - Ui[r] * (invD * Ui[cc]);
#pragma unroll
      for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int cc = 0; cc < 6; ++cc) {
          T acc = T(0);
#pragma unroll
          for (int k = 0; k < 6; ++k) acc = fma_t(Ia[r][k], Xf[k][cc], acc);
          tmp[r][cc] = acc;
        }
#pragma unroll
      for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int cc = 0; cc < 6; ++cc) {
          T acc = T(0);
#pragma unroll
          for (int k = 0; k < 6; ++k) acc = fma_t(Xf[k][r], tmp[k][cc], acc);
          IA[p][6 * r + cc] += acc;
        }
    }
  }
  // ---- phase B: one column at a time ----
  T colM[RBD_MAX_DOF];      // Minv[i][j] for the current column j
  T colF[RBD_MAX_DOF][6];   // forward-pass F[i][:, j]
  for (int j = 0; j < n; ++j) {
    for (int i = 0; i < n; ++i) colM[i] = T(0);
    // backward pass restricted to column j: walk j -> root   (:697-726)
    {
      T F[6];
#pragma unroll
      for (int r = 0; r < 6; ++r) F[r] = T(0);
      for (int i = j; i >= 0;) {
        const T invD = linvD[i];
        const T mij = (i == j ? invD : T(0)) - invD * dot6(m.S[i], F);
        colM[i] = mij;
        const int p = m.parent[i];
        if (p >= 0) {
          T X[18], t[6];
#pragma unroll
          for (int r = 0; r < 6; ++r) F[r] = fma_t(lU[i][r], mij, F[r]);
          build_X(m, i, lb[i][0], lb[i][1], X);
          XT_apply(X, F, t);
#pragma unroll
          for (int r = 0; r < 6; ++r) F[r] = t[r];
        }
        i = p;
      }
    }
    // forward pass restricted to column j   (:760-781).  Dense output only needs rows i <= j
    // (the mirror overwrites the rest); output_dense == 0 keeps every row as computed.
    const int last = output_dense ? j : n - 1;
    for (int i = 0; i <= last; ++i) {
      const int p = m.parent[i];
      T mij = colM[i];
      if (p >= 0) {
        T X[18], Fp[6], Fi[6];
#pragma unroll
        for (int r = 0; r < 6; ++r) Fp[r] = colF[p][r];
        mij -= linvD[i] * dot6(lUX[i], Fp);
        build_X(m, i, lb[i][0], lb[i][1], X);
        X_apply(X, Fp, Fi);
#pragma unroll
        for (int r = 0; r < 6; ++r) colF[i][r] = fma_t(m.S[i][r], mij, Fi[r]);
      } else {
#pragma unroll
        for (int r = 0; r < 6; ++r) colF[i][r] = m.S[i][r] * mij;
      }
      Mb[i * n + j] = mij;
      if (output_dense && i < j) Mb[j * n + i] = mij;                          // :799-804
    }
  }
}

// =============================================================================================
// crba, fixed-base branch (RBDReference.py:1090-1124), generic inertias: composite inertias as
// dense 6x6 in per-thread local memory, IC_p += X^T IC_i X leaf -> root (:1096-1103), then
// fh = IC_i S_i carried up the root path, H[i,j] = H[j,i] = S_j . fh (:1108-1122).
// =============================================================================================
template <typename T>
__global__ void __launch_bounds__(kFusedThreads)
crba_fused_kernel(const __grid_constant__ DevModel<T> m, int64_t B, const T* __restrict__ q, T* __restrict__ H) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int n = m.n;
  const T* qb = q + b * n;
  T IC[RBD_MAX_DOF][36], lb[RBD_MAX_DOF][2];
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int k = 0; k < 36; ++k) IC[i][k] = m.I[i][k];
    joint_basis(m, i, qb[i], lb[i][0], lb[i][1]);
  }
  for (int i = n - 1; i >= 0; --i) {
    const int p = m.parent[i];
    if (p < 0) continue;
    T X[18];
    build_X(m, i, lb[i][0], lb[i][1], X);
    // column k of IC_i X, then X^T of it, accumulated into column k of IC_p
    for (int k = 0; k < 6; ++k) {
      T xk[6], col[6], t[6];
#pragma unroll
      for (int r = 0; r < 6; ++r) xk[r] = T(0);
      // column k of X = [[E,0],[L,E]]
      if (k < 3) {
#pragma unroll
        for (int r = 0; r < 3; ++r) { xk[r] = X[3 * r + k]; xk[3 + r] = X[9 + 3 * r + k]; }
      } else {
#pragma unroll
        for (int r = 0; r < 3; ++r) xk[3 + r] = X[3 * r + (k - 3)];
      }
#pragma unroll
      for (int r = 0; r < 6; ++r) {
        T acc = T(0);
#pragma unroll
        for (int c = 0; c < 6; ++c) acc = fma_t(IC[i][6 * r + c], xk[c], acc);
        col[r] = acc;
      }
      XT_apply(X, col, t);
#pragma unroll
      for (int r = 0; r < 6; ++r) IC[p][6 * r + k] += t[r];
    }
  }
  T* Hb = H + b * (int64_t)n * n;
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) Hb[i * n + j] = T(0);
  for (int i = 0; i < n; ++i) {
    T fh[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      T acc = T(0);
#pragma unroll
      for (int c = 0; c < 6; ++c) acc = fma_t(IC[i][6 * r + c], m.S[i][c], acc);
      fh[r] = acc;
    }
    Hb[i * n + i] = dot6(m.S[i], fh);
    int j = i;
    while (m.parent[j] >= 0) {
      T X[18], t[6];
      build_X(m, j, lb[j][0], lb[j][1], X);
      XT_apply(X, fh, t);
#pragma unroll
      for (int r = 0; r < 6; ++r) fh[r] = t[r];
      j = m.parent[j];
      const T h = dot6(m.S[j], fh);
      Hb[i * n + j] = h;
      Hb[j * n + i] = h;
    }
  }
}

// =============================================================================================
// aba, fixed-base branch (RBDReference.py:817, :940-1024), one knot point per thread, the
// reference's recursion to the letter: v / c forward (:951-984), articulated inertia IA (dense 6x6,
// per-thread local memory) and bias force pA leaf -> root (:986-1007), accelerations root -> leaf
// (:1009-1022).  Reference quirk kept: :984 assigns ELEMENT 0 of crf(v) I v to all six entries of
// pA (`np.matmul(temp, v)[0]` on a 1-D product); f_ext is ignored by this branch.
// =============================================================================================
template <typename T>
__device__ __forceinline__ void congruence_add(const T* __restrict__ A, const T (&X)[18], T* __restrict__ P) {
  // P += X^T A X for dense row-major 6x6 A, P and X = [[E,0],[L,E]]
  for (int k = 0; k < 6; ++k) {
    T xk[6], col[6], t[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) xk[r] = T(0);
    if (k < 3) {
#pragma unroll
      for (int r = 0; r < 3; ++r) { xk[r] = X[3 * r + k]; xk[3 + r] = X[9 + 3 * r + k]; }
    } else {
#pragma unroll
      for (int r = 0; r < 3; ++r) xk[3 + r] = X[3 * r + (k - 3)];
    }
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      T acc = T(0);
#pragma unroll
      for (int c = 0; c < 6; ++c) acc = fma_t(A[6 * r + c], xk[c], acc);
      col[r] = acc;
    }
    XT_apply(X, col, t);
#pragma unroll
    for (int r = 0; r < 6; ++r) P[6 * r + k] += t[r];
  }
}

template <typename T>
__global__ void __launch_bounds__(kFusedThreads)
aba_fused_kernel(const __grid_constant__ DevModel<T> m, int64_t B, const T* __restrict__ q,
                 const T* __restrict__ qd, const T* __restrict__ tau, T gravity, T* __restrict__ qdd) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int n = m.n;
  const T* qb = q + b * n;
  const T* qdb = qd + b * n;
  const T* taub = tau + b * n;
  T IA[RBD_MAX_DOF][36], lv[RBD_MAX_DOF][6], lc[RBD_MAX_DOF][6], lp[RBD_MAX_DOF][6], lU[RBD_MAX_DOF][6];
  T lb[RBD_MAX_DOF][2], ld[RBD_MAX_DOF], lu[RBD_MAX_DOF];
  for (int i = 0; i < n; ++i) {
    joint_basis(m, i, qb[i], lb[i][0], lb[i][1]);
    const int p = m.parent[i];
    const T qdi = qdb[i];
    T vi[6], ci[6];
    if (p < 0) {
#pragma unroll
      for (int r = 0; r < 6; ++r) { vi[r] = m.S[i][r] * qdi; ci[r] = T(0); }                  // :957
    } else {
      T X[18], vp[6], t[6];
      build_X(m, i, lb[i][0], lb[i][1], X);
#pragma unroll
      for (int r = 0; r < 6; ++r) vp[r] = lv[p][r];
      X_apply(X, vp, vi);
#pragma unroll
      for (int r = 0; r < 6; ++r) vi[r] = fma_t(m.S[i][r], qdi, vi[r]);                        // :960-961
      crm_mul(vi, m.S[i], t);
#pragma unroll
      for (int r = 0; r < 6; ++r) ci[r] = qdi * t[r];                                          // :962
    }
    T Iv[6], x[6];
    mat6_apply(m.I[i], vi, Iv);
    crf_mul(vi, Iv, x);
#pragma unroll
    for (int r = 0; r < 6; ++r) { lv[i][r] = vi[r]; lc[i][r] = ci[r]; lp[i][r] = x[0]; }       // :984 (element 0)
#pragma unroll
    for (int k = 0; k < 36; ++k) IA[i][k] = m.I[i][k];                                         // :966
  }
  for (int i = n - 1; i >= 0; --i) {
    const int p = m.parent[i];
    T U[6], pa[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      T acc = T(0);
#pragma unroll
      for (int c = 0; c < 6; ++c) acc = fma_t(IA[i][6 * r + c], m.S[i][c], acc);
      U[r] = acc;                                                                              // :990
      pa[r] = lp[i][r];
    }
    const T d = dot6(m.S[i], U);                                                               // :991
    const T u = taub[i] - dot6(m.S[i], pa);                                                    // :992
#pragma unroll
    for (int r = 0; r < 6; ++r) lU[i][r] = U[r];
    ld[i] = d; lu[i] = u;
    if (p >= 0) {
      T Ia[36];
#pragma unroll
      for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int c = 0; c < 6; ++c) Ia[6 * r + c] = IA[i][6 * r + c] - U[r] * U[c] / d;        // :996-997
      const T ud = u / d;
#pragma unroll
      for (int r = 0; r < 6; ++r) {
        T acc = pa[r];
#pragma unroll
        for (int c = 0; c < 6; ++c) acc = fma_t(Ia[6 * r + c], lc[i][c], acc);
        pa[r] = fma_t(U[r], ud, acc);                                                          // :999
      }
      T X[18], t[6];
      build_X(m, i, lb[i][0], lb[i][1], X);
      congruence_add(Ia, X, IA[p]);                                                            // :1001-1004
      XT_apply(X, pa, t);
#pragma unroll
      for (int r = 0; r < 6; ++r) lp[p][r] += t[r];                                            // :1006-1007
    }
  }
  T* out = qdd + b * n;
  for (int i = 0; i < n; ++i) {
    const int p = m.parent[i];
    T X[18], ap[6], ai[6];
    build_X(m, i, lb[i][0], lb[i][1], X);
    if (p < 0) {
#pragma unroll
      for (int r = 0; r < 6; ++r) ap[r] = T(0);
      ap[5] = -gravity;                                                                        // :941-942
    } else {
#pragma unroll
      for (int r = 0; r < 6; ++r) ap[r] = lv[p][r];                                            // lv re-used for a
    }
    X_apply(X, ap, ai);
    T U[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) { ai[r] += lc[i][r]; U[r] = lU[i][r]; }                        // :1015 / :1017
    const T qddi = (lu[i] - dot6(U, ai)) / ld[i];                                              // :1020-1021
    out[i] = qddi;
#pragma unroll
    for (int r = 0; r < 6; ++r) lv[i][r] = fma_t(m.S[i][r], qddi, ai[r]);                      // :1022
  }
}

// =============================================================================================
// FMA peak micro-benchmark: 8 independent dependent-chains per thread.
// =============================================================================================
template <typename T>
__global__ void __launch_bounds__(256) fma_peak_kernel(T* out, int iters, T x0) {
  T a0 = x0, a1 = x0 + T(1), a2 = x0 + T(2), a3 = x0 + T(3), a4 = x0 + T(4), a5 = x0 + T(5),
    a6 = x0 + T(6), a7 = x0 + T(7);
  const T m1 = T(0.999999), c1 = T(1e-6) * T(threadIdx.x);
  for (int k = 0; k < iters; ++k) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      a0 = fma_t(a0, m1, c1); a1 = fma_t(a1, m1, c1); a2 = fma_t(a2, m1, c1); a3 = fma_t(a3, m1, c1);
      a4 = fma_t(a4, m1, c1); a5 = fma_t(a5, m1, c1); a6 = fma_t(a6, m1, c1); a7 = fma_t(a7, m1, c1);
    }
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

}  // namespace rbd
