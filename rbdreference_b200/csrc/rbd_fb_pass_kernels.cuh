// rbd_fb_pass_kernels.cuh - the eight per-pass helpers for FLOATING-BASE robots.
//
// The `self.robot.floating_base` branches of rnea_fpass RBDReference.py:559-598 (:585, :591), rnea_bpass
// :600-621, minv_bpass :630-735 (:652-691), minv_fpass :737-783 (:761-779), rnea_grad_fpass_dq :1127-1187
// (:1141-1168), rnea_grad_fpass_dqd :1189-1255 (:1212-1238), rnea_grad_bpass_dq :1257-1297 (:1267-1282) and
// rnea_grad_bpass_dqd :1299-1343 (:1309-1341), with every intermediate array in the reference's own shape and
// index convention: body 0 is the base (S = eye(6), q[0:7], qd[0:6]), body i >= 1 owns row / column i + 5,
// n = NB + 5.  The reference's behaviour is kept to the letter where it is unusual:
//   * minv_bpass :687-691 subtracts fb_Dinv F[5][:, adj] for the base (the trailing [-1] picks F[5]);
//   * minv_fpass :771-781 re-uses F with BODY indices (F[ind], F[parent_ind]), F[0] = Minv[0:6, :];
//   * rnea_grad_fpass_dq :1166-1168 adds zeros for the base (dv_dq is still zero there);
//   * rnea_grad_bpass_dqd :1336-1341 adds the base's damping to the block [0:5, 0:5] and body i's to [i, i].
//
// These entry points exist so that downstream accelerators can be checked pass by pass (README.md:19): one
// knot point per thread, the passes' arrays in HBM are the working storage exactly as in the reference
// (the in-place contracts - f of rnea_bpass, Minv and F of minv_fpass, df of the gradient bpasses - come for
// free).  They are HBM-bound by construction (each array is written once and read back by the same thread).
#pragma once
#include "rbd_common.cuh"
#include "rbd_fb_kernels.cuh"
#include "rbd_coop_pass_kernels.cuh"
#include "rbd_pass_kernels.cuh"
#include "rbd_coop_minv_kernels.cuh"

namespace rbd {

constexpr int kFbPassThreads = 64;

// X of body i from q: the base from position + quaternion, joint i >= 1 from q[i + 6]
template <typename T>
__device__ __forceinline__ void fbp_X(const FbModel<T>& m, int i, const T* __restrict__ qb, T (&X)[18]) {
  if (i == 0) fb_base_X(m, qb, X);
  else build_X_from_q(m.d, i, qb[i + 6], X);
}

// dense 6x6 (row-major) of the 18-value [E | L] layout:  X = [[E, 0], [L, E]]
template <typename T>
__device__ __forceinline__ void fbp_dense(const T (&X)[18], T (&D)[36]) {
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      D[6 * r + c] = X[3 * r + c];
      D[6 * r + 3 + c] = T(0);
      D[6 * (3 + r) + c] = X[9 + 3 * r + c];
      D[6 * (3 + r) + 3 + c] = X[3 * r + c];
    }
}

// ---- rnea_fpass (:559-598) -------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kFbPassThreads)
fbp_rnea_fpass_kernel(const __grid_constant__ FbModel<T> m, int64_t B, const T* __restrict__ q, const T* __restrict__ qd,
                      const T* __restrict__ qdd, T gravity, T* __restrict__ v, T* __restrict__ a, T* __restrict__ f) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int NB = m.d.n, nv = NB + 5;
  const T* qb = q + b * (NB + 6);
  const T* qdb = qd + b * nv;
  const T* qddb = qdd ? qdd + b * nv : nullptr;
  T* vb = v + b * 6 * NB;
  T* ab = a + b * 6 * NB;
  T* fb = f + b * 6 * NB;
  for (int i = 0; i < NB; ++i) {
    T X[18], vi[6], ai[6], par[6], vJ[6], t[6];
    fbp_X(m, i, qb, X);
    const int p = m.d.parent[i];
    if (p < 0) {
#pragma unroll
      for (int r = 0; r < 6; ++r) { vi[r] = T(0); par[r] = T(0); }
      par[5] = -gravity;                                                       // :566
      X_apply(X, par, ai);                                                     // :578
    } else {
#pragma unroll
      for (int r = 0; r < 6; ++r) par[r] = vb[r * NB + p];
      X_apply(X, par, vi);                                                     // :580
#pragma unroll
      for (int r = 0; r < 6; ++r) par[r] = ab[r * NB + p];
      X_apply(X, par, ai);                                                     // :581
    }
#pragma unroll
    for (int r = 0; r < 6; ++r) vJ[r] = i == 0 ? qdb[r] : m.d.S[i][r] * qdb[i + 5];   // :585-586
#pragma unroll
    for (int r = 0; r < 6; ++r) vi[r] += vJ[r];
    crm_mul(vi, vJ, t);                                                        // :588
#pragma unroll
    for (int r = 0; r < 6; ++r) ai[r] += t[r];
    if (qddb) {
#pragma unroll
      for (int r = 0; r < 6; ++r) ai[r] += i == 0 ? qddb[r] : m.d.S[i][r] * qddb[i + 5];   // :591-593
    }
    T Ia[6], Iv[6], vxIv[6];
    mat6_apply(m.d.I[i], ai, Ia);
    mat6_apply(m.d.I[i], vi, Iv);
    crf_mul(vi, Iv, vxIv);                                                     // :596
#pragma unroll
    for (int r = 0; r < 6; ++r) { vb[r * NB + i] = vi[r]; ab[r * NB + i] = ai[r]; fb[r * NB + i] = Ia[r] + vxIv[r]; }
  }
}

// ---- rnea_bpass (:600-621): f accumulated in place ---------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kFbPassThreads)
fbp_rnea_bpass_kernel(const __grid_constant__ FbModel<T> m, int64_t B, const T* __restrict__ q, T* __restrict__ f,
                      T* __restrict__ c) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int NB = m.d.n, nv = NB + 5;
  const T* qb = q + b * (NB + 6);
  T* fb = f + b * 6 * NB;
  T* cb = c + b * nv;
  for (int i = NB - 1; i >= 1; --i) {
    T X[18], fi[6], t[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) fi[r] = fb[r * NB + i];
    cb[i + 5] = dot6(m.d.S[i], fi);                                            // :612
    fbp_X(m, i, qb, X);
    XT_apply(X, fi, t);
    const int p = m.d.parent[i];
#pragma unroll
    for (int r = 0; r < 6; ++r) fb[r * NB + p] += t[r];                        // :617-619
  }
#pragma unroll
  for (int r = 0; r < 6; ++r) cb[r] = fb[r * NB];                              // :612 with S = eye(6)
}

// ---- minv_bpass (:630-735) -----------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kFbPassThreads)
fbp_minv_bpass_kernel(const __grid_constant__ FbModel<T> m, int64_t B, const T* __restrict__ q, T* __restrict__ Minv,
                      T* __restrict__ F, T* __restrict__ U, T* __restrict__ Dinv) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int NB = m.d.n, n = NB + 5;
  const T* qb = q + b * (NB + 6);
  T* Mb = Minv + b * (int64_t)n * n;
  T* Fb = F + b * (int64_t)n * 6 * n;                       // F[mi][r][col] = Fb[(mi * 6 + r) * n + col]
  T* Ub = U + b * (int64_t)n * 6;
  T* Db = Dinv + b * (int64_t)n;
  for (int k = 0; k < n * n; ++k) Mb[k] = T(0);
  for (int k = 0; k < n * 6 * n; ++k) Fb[k] = T(0);
  for (int k = 0; k < n * 6; ++k) Ub[k] = T(0);
  for (int k = 0; k < n; ++k) Db[k] = T(0);
  T IA[RBD_MAX_DOF][36];                                     // :662
  for (int i = 0; i < NB; ++i)
#pragma unroll
    for (int k = 0; k < 36; ++k) IA[i][k] = m.d.I[i][k];
  for (int i = NB - 1; i >= 1; --i) {
    const int mi = i + 5, p = m.d.parent[i], mp = p + 5;
    const unsigned sub = m.d.sub_mask[i];
    T Ui[6];
    mat6_apply(IA[i], m.d.S[i], Ui);                                           // :697
    const T D = dot6(m.d.S[i], Ui);                                            // :698
    const T invD = T(1) / D;
#pragma unroll
    for (int r = 0; r < 6; ++r) Ub[mi * 6 + r] = Ui[r];
    Db[mi] = D;
    Mb[mi * n + mi] = invD;                                                    // :700
    for (int j = i; j < NB; ++j) {                                             // :702-708
      if (!((sub >> j) & 1u)) continue;
      T sF = T(0);
#pragma unroll
      for (int r = 0; r < 6; ++r) sF = fma_t(m.d.S[i][r], Fb[(mi * 6 + r) * n + j + 5], sF);
      Mb[mi * n + j + 5] -= invD * sF;
    }
    T X[18];
    fbp_X(m, i, qb, X);
    for (int j = i; j < NB; ++j) {                                             // :720-726
      if (!((sub >> j) & 1u)) continue;
      const T mij = Mb[mi * n + j + 5];
      T Fi[6], t[6];
#pragma unroll
      for (int r = 0; r < 6; ++r) {
        Fi[r] = fma_t(Ui[r], mij, Fb[(mi * 6 + r) * n + j + 5]);
        Fb[(mi * 6 + r) * n + j + 5] = Fi[r];
      }
      XT_apply(X, Fi, t);
#pragma unroll
      for (int r = 0; r < 6; ++r) Fb[(mp * 6 + r) * n + j + 5] += t[r];
    }
    // IA_parent += X^T (IA - U U^T / D) X   (:728-733)
    T Xd[36], tmp[36];
    fbp_dense(X, Xd);
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
      for (int cc = 0; cc < 6; ++cc) {
        T acc = T(0);
#pragma unroll
        for (int k = 0; k < 6; ++k) acc = fma_t(IA[i][6 * r + k] - Ui[r] * (invD * Ui[k]), Xd[6 * k + cc], acc);
        tmp[6 * r + cc] = acc;
      }
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
      for (int cc = 0; cc < 6; ++cc) {
        T acc = T(0);
#pragma unroll
        for (int k = 0; k < 6; ++k) acc = fma_t(Xd[6 * k + r], tmp[6 * k + cc], acc);
        IA[p][6 * r + cc] += acc;
      }
  }
  // base (:677-691): U[0:6] = IA_0, fb_Dinv = inv(IA_0), Minv[0:6, 0:6] = fb_Dinv, Minv[0:6, adj] -= fb_Dinv F[5][:, adj]
#pragma unroll
  for (int k = 0; k < 36; ++k) Ub[k] = IA[0][k];
  T Di[36];
  {
    T A[36];
#pragma unroll
    for (int k = 0; k < 36; ++k) { A[k] = IA[0][k]; Di[k] = (k % 7 == 0) ? T(1) : T(0); }
#pragma unroll
    for (int pv = 0; pv < 6; ++pv) {
      const T inv = T(1) / A[7 * pv];
#pragma unroll
      for (int cc = 0; cc < 6; ++cc) { A[6 * pv + cc] *= inv; Di[6 * pv + cc] *= inv; }
#pragma unroll
      for (int r = 0; r < 6; ++r) {
        if (r == pv) continue;
        const T fct = A[6 * r + pv];
#pragma unroll
        for (int cc = 0; cc < 6; ++cc) {
          A[6 * r + cc] = fma_t(-fct, A[6 * pv + cc], A[6 * r + cc]);
          Di[6 * r + cc] = fma_t(-fct, Di[6 * pv + cc], Di[6 * r + cc]);
        }
      }
    }
  }
  const T m00 = Mb[0];
#pragma unroll
  for (int r = 0; r < 6; ++r)
#pragma unroll
    for (int cc = 0; cc < 6; ++cc) Mb[r * n + cc] = m00 + Di[6 * r + cc];      // :686
  for (int col = 5; col < n; ++col) {                                          // adj = subtree(0) + 5 = 5 .. n-1
    T Fc[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) Fc[r] = Fb[(5 * 6 + r) * n + col];
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      T acc = T(0);
#pragma unroll
      for (int k = 0; k < 6; ++k) acc = fma_t(Di[6 * r + k], Fc[k], acc);
      Mb[r * n + col] -= acc;
    }
  }
}

// ---- minv_fpass (:737-783): Minv and F updated in place ---------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kFbPassThreads)
fbp_minv_fpass_kernel(const __grid_constant__ FbModel<T> m, int64_t B, const T* __restrict__ q, T* __restrict__ Minv,
                      T* __restrict__ F, const T* __restrict__ U, const T* __restrict__ Dinv) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int NB = m.d.n, n = NB + 5;
  const T* qb = q + b * (NB + 6);
  T* Mb = Minv + b * (int64_t)n * n;
  T* Fb = F + b * (int64_t)n * 6 * n;
  const T* Ub = U + b * (int64_t)n * 6;
  const T* Db = Dinv + b * (int64_t)n;
  for (int r = 0; r < 6; ++r)
    for (int col = 0; col < n; ++col) Fb[(0 * 6 + r) * n + col] = Mb[r * n + col];   // :779  F[0] = S Minv[0:6, 0:]
  for (int i = 1; i < NB; ++i) {
    const int mi = i + 5, p = m.d.parent[i];
    T X[18], Ui[6], UX[6];
    fbp_X(m, i, qb, X);
#pragma unroll
    for (int r = 0; r < 6; ++r) Ui[r] = Ub[mi * 6 + r];
    XT_apply(X, Ui, UX);                                                       // (U^T X)^T = X^T U
    const T invD = T(1) / Db[mi];
    for (int col = 0; col < n; ++col) {
      T Fp[6], XF[6];
#pragma unroll
      for (int r = 0; r < 6; ++r) Fp[r] = Fb[(p * 6 + r) * n + col];
      const T mij = Mb[mi * n + col] - invD * dot6(UX, Fp);                    // :771-773
      Mb[mi * n + col] = mij;
      X_apply(X, Fp, XF);
#pragma unroll
      for (int r = 0; r < 6; ++r) Fb[(i * 6 + r) * n + col] = fma_t(m.d.S[i][r], mij, XF[r]);   // :774-776
    }
  }
}

// df[:, c] = I da_c + crf(dv_c) (I v) + crf(v) (I dv_c)   (:1179-1185 / :1247-1252)
template <typename T>
__device__ __forceinline__ void fbp_df(const T* __restrict__ I, const T (&vi)[6], const T (&Iv)[6], const T (&dvc)[6],
                                       const T (&dac)[6], T (&dfc)[6]) {
  T t1[6], t2[6], Idv[6];
  mat6_apply(I, dac, dfc);
  crf_mul(dvc, Iv, t1);
  mat6_apply(I, dvc, Idv);
  crf_mul(vi, Idv, t2);
#pragma unroll
  for (int r = 0; r < 6; ++r) dfc[r] += t1[r] + t2[r];
}

// ---- rnea_grad_fpass_dq (:1127-1187) and rnea_grad_fpass_dqd (:1189-1255) --------------------------------------
template <typename T, bool DQ>
__global__ void __launch_bounds__(kFbPassThreads)
fbp_grad_fpass_kernel(const __grid_constant__ FbModel<T> m, int64_t B, const T* __restrict__ q, const T* __restrict__ qd,
                      const T* __restrict__ v, const T* __restrict__ a, T gravity, T* __restrict__ dv, T* __restrict__ da,
                      T* __restrict__ df) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int NB = m.d.n, n = NB + 5;
  const T* qb = q + b * (NB + 6);
  const T* qdb = qd + b * n;
  const T* vb = v + b * 6 * NB;
  const T* ab = DQ ? a + b * 6 * NB : nullptr;
  const int64_t sz = (int64_t)6 * n * NB;                   // [r][c][body] = (r * n + c) * NB + body
  T* dvb = dv + b * sz;
  T* dab = da + b * sz;
  T* dfb = df + b * sz;
  for (int i = 0; i < NB; ++i) {
    T X[18], vi[6], Iv[6];
    fbp_X(m, i, qb, X);
#pragma unroll
    for (int r = 0; r < 6; ++r) vi[r] = vb[r * NB + i];
    mat6_apply(m.d.I[i], vi, Iv);
    const int p = m.d.parent[i], idx = i + 5;
    T seed_v[6], seed_a[6];                                   // what column idx gets on top (i >= 1)
    T xg[6];                                                  // base, dq: X0 g
    if (i == 0) {
      if (DQ) {
        T g6[6] = {T(0), T(0), T(0), T(0), T(0), -gravity};
        X_apply(X, g6, xg);
      }
    } else {
      T par[6], t[6];
      if (DQ) {
#pragma unroll
        for (int r = 0; r < 6; ++r) par[r] = vb[r * NB + p];
        X_apply(X, par, t);
        crm_mul(t, m.d.S[i], seed_v);                                          // :1159
#pragma unroll
        for (int r = 0; r < 6; ++r) par[r] = ab[r * NB + p];
        X_apply(X, par, t);
        crm_mul(t, m.d.S[i], seed_a);                                          // :1173
      } else {
#pragma unroll
        for (int r = 0; r < 6; ++r) seed_v[r] = m.d.S[i][r];                   // :1231
        crm_mul(vi, m.d.S[i], seed_a);                                         // :1243
      }
    }
    for (int c = 0; c < n; ++c) {
      T dvc[6], dac[6], dfc[6];
      if (i == 0) {
#pragma unroll
        for (int r = 0; r < 6; ++r) { dvc[r] = T(0); dac[r] = T(0); }
        if (c < 6) {
          T e[6] = {T(0), T(0), T(0), T(0), T(0), T(0)};
          e[c] = T(1);
          if (DQ) {
            crm_mul(xg, e, dac);                                               // :1175 with S = eye(6); :1166-1168 adds zeros
          } else {
            dvc[c] = T(1);                                                     // :1231
            T qb6[6], t1[6], t2[6];
#pragma unroll
            for (int r = 0; r < 6; ++r) qb6[r] = qdb[r];
            crm_mul(e, qb6, t1);                                               // :1236-1238  sum_ii qd[ii] crm(dv_c)[:, ii]
            crm_mul(vi, e, t2);                                                // :1243
#pragma unroll
            for (int r = 0; r < 6; ++r) dac[r] = t1[r] + t2[r];
          }
        }
      } else {
        T par[6], t[6];
#pragma unroll
        for (int r = 0; r < 6; ++r) par[r] = dvb[(r * n + c) * NB + p];
        X_apply(X, par, dvc);                                                  // :1158 / :1230
#pragma unroll
        for (int r = 0; r < 6; ++r) par[r] = dab[(r * n + c) * NB + p];
        X_apply(X, par, dac);                                                  // :1163 / :1234
        if (c == idx) {
#pragma unroll
          for (int r = 0; r < 6; ++r) dvc[r] += seed_v[r];
        }
        crm_mul(dvc, m.d.S[i], t);                                             // :1170 / :1240
        const T qdi = qdb[idx];
#pragma unroll
        for (int r = 0; r < 6; ++r) dac[r] = fma_t(qdi, t[r], dac[r]);
        if (c == idx) {
#pragma unroll
          for (int r = 0; r < 6; ++r) dac[r] += seed_a[r];
        }
      }
      fbp_df(m.d.I[i], vi, Iv, dvc, dac, dfc);
#pragma unroll
      for (int r = 0; r < 6; ++r) {
        dvb[(r * n + c) * NB + i] = dvc[r];
        dab[(r * n + c) * NB + i] = dac[r];
        dfb[(r * n + c) * NB + i] = dfc[r];
      }
    }
  }
}

// ---- rnea_grad_fpass_dq / _dqd, one BODY per lane (the scheme of grad_fpass_level_kernel, rbd_coop_pass_kernels.cuh) ----
// Lane i of a group of G lanes is body i (body 0 = the base).  A body's non-zero columns are those of its joint ancestors
// (column a + 5 for ancestor body a >= 1, itself included) and the base's six.  Round d handles the ancestor at distance d:
//   * d < depth(i): one pair, the parent's pair of round d - 1 arrives by warp shuffle (as for a fixed base);
//   * the base's six columns (every body has them) are a second phase, level by level from the base: the 6 x (bodies of
//     the level) pairs of a level are spread over the lanes of the group - a lane works for ANOTHER body there, whose
//     X, I v and qd wait in shared memory next to the results, where the parent's pair of the previous level is read
//     too.  The base's own pairs are the seeds (:1175 / :1231-1243 with S = eye(6)).  (First version: every lane ran the
//     six pairs of its own body in the round in which its ancestor is the base - a six-fold serial loop in almost every
//     round: Atlas + base 6.2 ms per 2^16.)
// The three (6, n, NB) slabs then leave in one coalesced pass through a per-CTA map (structural zeros where it says -1).
constexpr int kFbLvlBody = 37;          // per body and knot point: X (18) | I v (6) | qd | X_0 a_grav or qd[0:6] (base row only) ... odd stride
__host__ __device__ inline int fbp_level_knot_vals(int NB, int npairs) { return (12 * NB + 18 * npairs + kFbLvlBody * NB + 3) & ~3; }
__host__ __device__ inline size_t fbp_level_head_bytes(int NB, int G, size_t tsize) {
  const size_t slab = (size_t)6 * (NB + 5) * NB;
  return (((size_t)NB * 6 * sizeof(int) + 64 * sizeof(int) + (size_t)(32 / G) * slab * sizeof(short) + 15) & ~(size_t)15) +
         (((size_t)NB * kCpLvlMdl + 3) & ~(size_t)3) * tsize;
}

template <typename T, int G, bool DQ>
__global__ void __launch_bounds__(kCpLvlMaxWarps * 32)
fbp_grad_fpass_level_kernel(const __grid_constant__ FbModel<T> m, int npairs, int64_t B, const T* __restrict__ q,
                            const T* __restrict__ qd, const T* __restrict__ v, const T* __restrict__ a, T gravity,
                            T* __restrict__ dv, T* __restrict__ da, T* __restrict__ df) {
  constexpr int IPW = 32 / G;
  typedef typename Vec2<T>::type V2;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int NB = m.d.n, n = NB + 5, nq = NB + 6;
  const int cn = n * NB;                                   // entries of one tensor row: [column][body]
  const int slab = 6 * cn;                                 // values of one tensor of one knot point (even)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int g = lane / G, i = lane - g * G;
  const int gbase = g * G;
  const bool valid = i < NB;
  int* topo = reinterpret_cast<int*>(smem_raw);                            // [NB][4]: parent kind depth first-pair
  int* lvl_body = topo + 4 * NB;                                           // [NB]: bodies sorted by depth
  int* lvl_begin = lvl_body + 2 * NB;                                      // [maxdepth + 2] (<= 34 entries of the 64)
  short* pmap = reinterpret_cast<short*>(topo + 6 * NB + 64);               // [IPW][6][n][NB]: where the value waits, or -1
  T* mdl = reinterpret_cast<T*>(smem_raw + fbp_level_head_bytes(NB, G, 0)); // [NB][97]: XA XB XC S I
  const int knot_vals = fbp_level_knot_vals(NB, npairs);
  T* ws = mdl + (((size_t)NB * kCpLvlMdl + 3) & ~(size_t)3) + (size_t)warp * IPW * knot_vals;
  T* sv = ws + g * knot_vals;                              // [6][NB] of this lane's knot point
  T* sa = sv + 6 * NB;
  T* res = sa + 6 * NB;                                    // [3][6][npairs]
  T* bd = res + 18 * npairs;                               // [NB][37]: X, I v, qd of every body of this knot point
  for (int k = threadIdx.x; k < NB * kCpLvlMdl; k += blockDim.x) {
    const int b = k / kCpLvlMdl, w = k - b * kCpLvlMdl;
    mdl[k] = w < 18 ? m.d.XA[b][w] : w < 36 ? m.d.XB[b][w - 18] : w < 54 ? m.d.XC[b][w - 36] : w < 60 ? m.d.S[b][w - 54] : w < 96 ? m.d.I[b][w - 60] : T(0);
  }
  for (int k = threadIdx.x; k < IPW * slab; k += blockDim.x) pmap[k] = (short)-1;
  int maxdepth = 0;
  {
    int first = 0;
    for (int b = 0; b < NB; ++b) {                         // (every thread walks the same entries of the constant bank)
      int d = 0;
      for (int p = m.d.parent[b]; p >= 0; p = m.d.parent[p]) ++d;
      if (threadIdx.x == 0) { topo[4 * b] = m.d.parent[b]; topo[4 * b + 1] = m.d.kind[b]; topo[4 * b + 2] = d; topo[4 * b + 3] = first; }
      first += d + 6;                                      // d joint ancestors (itself included) + the base's six columns
      maxdepth = d > maxdepth ? d : maxdepth;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {                                   // bodies by depth
    int pos = 0;
    for (int L = 0; L <= maxdepth; ++L) {
      lvl_begin[L] = pos;
      for (int b = 0; b < NB; ++b)
        if (topo[4 * b + 2] == L) lvl_body[pos++] = b;
    }
    lvl_begin[maxdepth + 1] = pos;
  }
  if (threadIdx.x < NB) {
    const int b = threadIdx.x;
    int idx = topo[4 * b + 3];
    for (int c = b; c >= 1; c = topo[4 * c]) {             // joint ancestors: column c + 5
      for (int kk = 0; kk < IPW; ++kk)
        for (int r = 0; r < 6; ++r) pmap[(kk * 6 + r) * cn + (c + 5) * NB + b] = (short)(kk * knot_vals + r * npairs + idx);
      ++idx;
    }
    for (int k = 0; k < 6; ++k) {                          // the base's columns
      for (int kk = 0; kk < IPW; ++kk)
        for (int r = 0; r < 6; ++r) pmap[(kk * 6 + r) * cn + k * NB + b] = (short)(kk * knot_vals + r * npairs + idx);
      ++idx;
    }
  }
  __syncthreads();
  const int ib = valid ? i : 0;
  const int par = topo[4 * ib], kind = topo[4 * ib + 1];
  const int depth = valid ? topo[4 * ib + 2] : -1;
  const int pair0 = topo[4 * ib + 3];
  const T* mc = mdl + ib * kCpLvlMdl;
  T S[6], Im[36];                                          // lane = body for the whole kernel
#pragma unroll
  for (int r = 0; r < 6; ++r) S[r] = mc[54 + r];
#pragma unroll
  for (int k = 0; k < 36; ++k) Im[k] = mc[60 + k];
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(dv) | reinterpret_cast<uintptr_t>(da) | reinterpret_cast<uintptr_t>(df)) & (2 * sizeof(T) - 1)) == 0;
  const int64_t ngroups = (B + IPW - 1) / IPW;
  for (int64_t grp = (int64_t)blockIdx.x * nwarps + warp; grp < ngroups; grp += (int64_t)gridDim.x * nwarps) {
    const int64_t first = grp * IPW;
    const int nk = (int)((B - first) < IPW ? (B - first) : IPW);
    int64_t b = first + g;
    if (b >= B) b = B - 1;                                  // duplicate work, never stored
    int64_t bn = first + (int64_t)gridDim.x * nwarps * IPW + g;
    const bool more = bn < B;
    for (int e = i; e < 6 * NB; e += G) {
      sv[e] = v[b * 6 * NB + e];
      if (DQ) sa[e] = a[b * 6 * NB + e];
      if (more) {
        prefetch_l2(v + bn * 6 * NB + e);
        if (DQ) prefetch_l2(a + bn * 6 * NB + e);
      }
    }
    T X[18], vi[6], Iv[6], qdi = T(0);
    T xg[6], qd0[6];                                        // base lane: X_0 a_grav (dq) / qd[0:6] (dqd)
#pragma unroll
    for (int r = 0; r < 6; ++r) { xg[r] = T(0); qd0[r] = T(0); }
    if (valid && i == 0) {
      fb_base_X(m, q + b * nq, X);
      if (DQ) {
        T g6[6] = {T(0), T(0), T(0), T(0), T(0), -gravity};
        X_apply(X, g6, xg);
      } else {
#pragma unroll
        for (int r = 0; r < 6; ++r) qd0[r] = qd[b * n + r];
      }
    } else {
      T f1 = T(0), f2 = T(0);
      if (valid) {
        const T qi = q[b * nq + i + 6];
        qdi = qd[b * n + i + 5];
        if (kind == 0) sincos_t(qi, &f2, &f1);
        else f1 = qi;
      }
#pragma unroll
      for (int k = 0; k < 18; ++k) X[k] = fma_t(mc[36 + k], f2, fma_t(mc[18 + k], f1, mc[k]));
    }
    if (more && valid) { prefetch_l2(q + bn * nq + i + 6); prefetch_l2(qd + bn * n + i + 5); }
    __syncwarp();                                           // staged rows are in place; the previous slabs have been read
#pragma unroll
    for (int r = 0; r < 6; ++r) vi[r] = sv[r * NB + ib];
    mat6_apply(Im, vi, Iv);                                                      // :1180 / :1248
    if (valid) {                                            // what the second phase needs of this body
      T* mybd = bd + i * kFbLvlBody;
#pragma unroll
      for (int k = 0; k < 18; ++k) mybd[k] = X[k];
#pragma unroll
      for (int r = 0; r < 6; ++r) { mybd[18 + r] = Iv[r]; mybd[25 + r] = DQ ? xg[r] : qd0[r]; }
      mybd[24] = qdi;
    }
    T cdv[6], cda[6];                                       // dv / da of this lane's joint pair of the previous round
#pragma unroll
    for (int r = 0; r < 6; ++r) { cdv[r] = T(0); cda[r] = T(0); }
#pragma unroll 1
    for (int d = 0; d <= maxdepth; ++d) {
      T pv[6], pa[6];
      const int src = gbase + (par >= 0 ? par : 0);
#pragma unroll
      for (int r = 0; r < 6; ++r) { pv[r] = __shfl_sync(0xffffffffu, cdv[r], src); pa[r] = __shfl_sync(0xffffffffu, cda[r], src); }
      if (depth > d) {
        // ---- a joint ancestor at distance d (d = 0: the body's own column i + 5)
        T dvc[6], dac[6], t[6];
        if (d == 0) {
          T seed_a[6];
          if (DQ) {
            T pr[6], xp[6];
#pragma unroll
            for (int r = 0; r < 6; ++r) pr[r] = sv[r * NB + par];
            X_apply(X, pr, xp);
            crm_mul(xp, S, dvc);                                                 // :1159
#pragma unroll
            for (int r = 0; r < 6; ++r) pr[r] = sa[r * NB + par];
            X_apply(X, pr, xp);
            crm_mul(xp, S, seed_a);                                              // :1173
          } else {
#pragma unroll
            for (int r = 0; r < 6; ++r) dvc[r] = S[r];                           // :1231
            crm_mul(vi, S, seed_a);                                              // :1243
          }
          crm_mul(dvc, S, t);                                                    // :1170 / :1240
#pragma unroll
          for (int r = 0; r < 6; ++r) dac[r] = fma_t(qdi, t[r], seed_a[r]);
        } else {
          X_apply(X, pv, dvc);                                                   // :1158 / :1230
          X_apply(X, pa, dac);                                                   // :1163 / :1234
          crm_mul(dvc, S, t);
#pragma unroll
          for (int r = 0; r < 6; ++r) dac[r] = fma_t(qdi, t[r], dac[r]);
        }
        T dfc[6];
        fbp_df(Im, vi, Iv, dvc, dac, dfc);
        T* rp = res + pair0 + d;
#pragma unroll
        for (int r = 0; r < 6; ++r) {
          rp[r * npairs] = dvc[r];
          rp[(6 + r) * npairs] = dac[r];
          rp[(12 + r) * npairs] = dfc[r];
          cdv[r] = dvc[r];
          cda[r] = dac[r];
        }
      }
    }
    __syncwarp();
    // ---- the base's six columns of every body, level by level: task t of a level = (body lvl[t / 6], column t % 6)
#pragma unroll 1
    for (int L = 0; L <= maxdepth; ++L) {
      const int l0 = lvl_begin[L], ntask = 6 * (lvl_begin[L + 1] - l0);
#pragma unroll 1
      for (int t = i; t < ntask; t += G) {
        const int bb = lvl_body[l0 + t / 6], k = t - 6 * (t / 6);
        const T* b2 = bd + bb * kFbLvlBody;
        const T* mc2 = mdl + bb * kCpLvlMdl;
        T dvc[6], dac[6], v2[6], Iv2[6];
#pragma unroll
        for (int r = 0; r < 6; ++r) { v2[r] = sv[r * NB + bb]; Iv2[r] = b2[18 + r]; }
        if (bb == 0) {
          T e[6], x0[6];
#pragma unroll
          for (int r = 0; r < 6; ++r) { e[r] = r == k ? T(1) : T(0); dvc[r] = T(0); x0[r] = b2[25 + r]; }
          if (DQ) {
            crm_mul(x0, e, dac);                                                 // :1175 with S = eye(6); :1166-1168 adds zeros
          } else {
            T t1[6], t2[6];
#pragma unroll
            for (int r = 0; r < 6; ++r) dvc[r] = e[r];                           // :1231
            crm_mul(e, x0, t1);                                                  // :1236-1238
            crm_mul(v2, e, t2);                                                  // :1243
#pragma unroll
            for (int r = 0; r < 6; ++r) dac[r] = t1[r] + t2[r];
          }
        } else {
          const int pb = topo[4 * bb];
          const T* pr = res + topo[4 * pb + 3] + topo[4 * pb + 2] + k;           // the parent's pair of column k (previous level)
          T X2[18], S2[6], ppv[6], ppa[6], tt[6];
#pragma unroll
          for (int q2 = 0; q2 < 18; ++q2) X2[q2] = b2[q2];
#pragma unroll
          for (int r = 0; r < 6; ++r) { S2[r] = mc2[54 + r]; ppv[r] = pr[r * npairs]; ppa[r] = pr[(6 + r) * npairs]; }
          X_apply(X2, ppv, dvc);
          X_apply(X2, ppa, dac);
          crm_mul(dvc, S2, tt);
          const T qd2 = b2[24];
#pragma unroll
          for (int r = 0; r < 6; ++r) dac[r] = fma_t(qd2, tt[r], dac[r]);
        }
        T dfc[6];
        fbp_df(mc2 + 60, v2, Iv2, dvc, dac, dfc);
        T* rp = res + topo[4 * bb + 3] + topo[4 * bb + 2] + k;
#pragma unroll
        for (int r = 0; r < 6; ++r) {
          rp[r * npairs] = dvc[r];
          rp[(6 + r) * npairs] = dac[r];
          rp[(12 + r) * npairs] = dfc[r];
        }
      }
      __syncwarp();                                         // the level's results are visible to the next level
    }
    // ---- the warp's slabs of the three tensors (contiguous: consecutive knot points), every sector once
    const int total = nk * slab;
#pragma unroll 1
    for (int w = 0; w < 3; ++w) {
      T* out = (w == 0 ? dv : (w == 1 ? da : df)) + first * slab;
      const T* rk = ws + 12 * NB + (size_t)w * 6 * npairs;  // tensor w of the warp's first knot point
      if (vec_ok) {
        for (int f2 = lane; f2 < (total >> 1); f2 += 32) {
          const short2 pp = reinterpret_cast<const short2*>(pmap)[f2];
          V2 x;
          x.x = pp.x >= 0 ? rk[pp.x] : T(0);
          x.y = pp.y >= 0 ? rk[pp.y] : T(0);
          __stcs(reinterpret_cast<V2*>(out) + f2, x);
        }
      } else {
        for (int f = lane; f < total; f += 32) {
          const int pp = pmap[f];
          __stcs(out + f, pp >= 0 ? rk[pp] : T(0));
        }
      }
    }
    __syncwarp();                                           // sv / sa / res are rewritten for the next knot points
  }
}

// ---- rnea_grad_bpass_dq (:1257-1297) and rnea_grad_bpass_dqd (:1299-1343): df accumulated in place --------------
template <typename T, bool DQ>
__global__ void __launch_bounds__(kFbPassThreads)
fbp_grad_bpass_kernel(const __grid_constant__ FbModel<T> m, int64_t B, const T* __restrict__ q, const T* __restrict__ f,
                      T* __restrict__ df, int use_damping, T* __restrict__ dc) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int NB = m.d.n, n = NB + 5;
  const T* qb = q + b * (NB + 6);
  const T* fb = DQ ? f + b * 6 * NB : nullptr;
  T* dfb = df + b * (int64_t)6 * n * NB;
  T* dcb = dc + b * (int64_t)n * n;
  for (int i = NB - 1; i >= 1; --i) {
    const int idx = i + 5, p = m.d.parent[i];
    T X[18], extra[6];
    fbp_X(m, i, qb, X);
    if (DQ) {
      T fi[6], t[6];
#pragma unroll
      for (int r = 0; r < 6; ++r) fi[r] = fb[r * NB + i];
      crm_mul(fi, m.d.S[i], t);                                                // fxS = -crm(f) S (:166-168)
#pragma unroll
      for (int r = 0; r < 6; ++r) t[r] = -t[r];
      XT_apply(X, t, extra);                                                   // :1292
    }
    for (int c = 0; c < n; ++c) {
      T col[6], t[6];
#pragma unroll
      for (int r = 0; r < 6; ++r) col[r] = dfb[(r * n + c) * NB + i];
      dcb[idx * n + c] = dot6(m.d.S[i], col);                                  // :1284 / :1325
      XT_apply(X, col, t);
      if (DQ && c == idx) {
#pragma unroll
        for (int r = 0; r < 6; ++r) t[r] += extra[r];                          // :1293-1294
      }
#pragma unroll
      for (int r = 0; r < 6; ++r) dfb[(r * n + c) * NB + p] += t[r];           // :1291 / :1331
    }
  }
  for (int r = 0; r < 6; ++r)
    for (int c = 0; c < n; ++c) dcb[r * n + c] = dfb[(r * n + c) * NB];        // :1282 / :1325 with S = eye(6)
  if (!DQ && use_damping) {                                                    // :1336-1341
    for (int r = 0; r < 5; ++r)
      for (int c = 0; c < 5; ++c) dcb[r * n + c] += m.d.damping[0];
    for (int i = 1; i < NB; ++i) dcb[i * n + i] += m.d.damping[i];
  }
}

// ---- rnea_grad_bpass_dq / _dqd, one COLUMN per lane (the scheme of grad_bpass_coop_kernel) ------------------------------
// The pass never mixes columns (:1284-1294 / :1325-1331 act column by column).  A warp owns one knot point: its df slab
// (6, n, NB) - contiguous in HBM - is pulled into shared memory with one cp.async.bulk (mbarrier completion), lane c
// carries column c (and c + 32 when n > 32) through the bodies leaf -> base inside the tile, dc (n, n) is assembled next
// to it, and both leave with bulk stores (element loops where the 16-byte rule of the instruction is not met).  The
// thread-per-knot-point kernel above used df in HBM as read-modify-write working storage, 8 bytes per 32-byte sector.
__host__ __device__ inline int fbp_bpass_warp_vals(int NB) {
  const int n = NB + 5;
  return (((6 * n * NB + 3) & ~3) + ((n * n + 3) & ~3) + 8 * NB + 2 + 3) & ~3;  // df tile | dc tile | f | (f1, f2) | mbarrier
}

template <typename T, bool DQ>
__global__ void __launch_bounds__(kCpMaxWarps * 32)
fbp_grad_bpass_coop_kernel(const __grid_constant__ FbModel<T> m, int64_t B, const T* __restrict__ q, const T* __restrict__ f,
                           T* __restrict__ df, int use_damping, T* __restrict__ dc) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int NB = m.d.n, n = NB + 5, nq = NB + 6, nn = n * n;
  const int slab = 6 * n * NB;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  T* ws = reinterpret_cast<T*>(smem_raw) + (size_t)warp * fbp_bpass_warp_vals(NB);
  T* td = ws;                                              // [6][n][NB]: the slab itself
  T* tc = td + ((slab + 3) & ~3);                          // [n][n]
  T* sf = tc + ((nn + 3) & ~3);                            // [6][NB]
  T* sj = sf + 6 * NB;                                     // [NB][2]
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(ws + fbp_bpass_warp_vals(NB) - 2);
  unsigned phase = 0;
  warp_bulk_bar_init(bar, lane);
  for (int64_t b = (int64_t)blockIdx.x * nwarps + warp; b < B; b += (int64_t)gridDim.x * nwarps) {
    T* dfb = df + b * (int64_t)slab;
    T* dcb = dc + b * (int64_t)nn;
    warp_bulk_store_wait(lane);                            // the previous knot point's tiles have left
    const bool bulk_in = warp_bulk_load(td, dfb, slab, bar, lane);
    if (lane >= 1 && lane < NB) {
      T f1, f2;
      joint_basis(m.d, lane, q[b * nq + lane + 6], f1, f2);
      sj[2 * lane] = f1; sj[2 * lane + 1] = f2;
    }
    if (DQ) {
      for (int e = lane; e < 6 * NB; e += 32) sf[e] = f[b * 6 * NB + e];
    }
    if (bulk_in) {
      warp_bulk_load_wait(bar, phase);
      phase ^= 1u;
    } else {
      for (int e = lane; e < slab; e += 32) td[e] = dfb[e];
    }
    __syncwarp();
    for (int c = lane; c < n; c += 32) {
      T* md = td + c * NB;                                 // + r n NB + body
      const int rstride = n * NB;
      for (int i = NB - 1; i >= 1; --i) {
        const int p = m.d.parent[i];
        T col[6], X[18], t[6];
#pragma unroll
        for (int r = 0; r < 6; ++r) col[r] = md[r * rstride + i];
        tc[(i + 5) * n + c] = dot6(m.d.S[i], col);                             // :1284 / :1325
        build_X(m.d, i, sj[2 * i], sj[2 * i + 1], X);
        XT_apply(X, col, t);                                                   // :1291 / :1331
        if (DQ && c == i + 5) {
          T fi[6], S[6], fxs[6], t2[6];
#pragma unroll
          for (int r = 0; r < 6; ++r) { fi[r] = sf[r * NB + i]; S[r] = m.d.S[i][r]; }
          crm_mul(fi, S, fxs);
#pragma unroll
          for (int r = 0; r < 6; ++r) fxs[r] = -fxs[r];                        // fxS :166-168
          XT_apply(X, fxs, t2);                                                // :1292
#pragma unroll
          for (int r = 0; r < 6; ++r) t[r] += t2[r];                           // :1293-1294
        }
#pragma unroll
        for (int r = 0; r < 6; ++r) md[r * rstride + p] += t[r];
      }
#pragma unroll
      for (int r = 0; r < 6; ++r) tc[r * n + c] = md[r * rstride];             // :1282 / :1325 with S = eye(6)
    }
    __syncwarp();
    if (!DQ && use_damping) {                                                  // :1336-1341, in the reference's order
      if (lane < 25) tc[(lane / 5) * n + lane % 5] += m.d.damping[0];
      __syncwarp();
      if (lane >= 1 && lane < NB) tc[lane * n + lane] += m.d.damping[lane];
      __syncwarp();
    }
    if (!warp_bulk_store(dfb, td, slab, lane))
      for (int e = lane; e < slab; e += 32) dfb[e] = td[e];
    if (!warp_bulk_store(dcb, tc, nn, lane))
      for (int e = lane; e < nn; e += 32) __stcs(dcb + e, tc[e]);
    __syncwarp();
  }
  warp_bulk_store_wait(lane);                              // shared memory must outlive the copies
}

// ---- minv_fpass, one COLUMN per lane (the scheme of minv_fpass_col_kernel, rbd_pass_kernels.cuh) -----------------------
// The forward pass never mixes columns (:771-776 act on whole rows, entry by entry).  A warp owns one knot point and lane
// c carries column c (and c + 32 when n > 32) through the bodies: every access to Minv[i + 5, :] and F[i][r, :] is one
// contiguous row per knot point, the parent's F stays in registers along chains and is re-read (the lane's own store)
// at branch points.  F keeps the reference's indexing: F[0] <- Minv[0:6, :] (:779), body i reads F[parent BODY index]
// and writes F[i] (:771-781).
template <typename T>
__global__ void __launch_bounds__(kPassThreads)
fbp_minv_fpass_col_kernel(const __grid_constant__ FbModel<T> m, int64_t B, const T* __restrict__ q, T* __restrict__ Minv,
                          T* __restrict__ F, const T* __restrict__ U, const T* __restrict__ Dinv) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int NB = m.d.n, n = NB + 5, nq = NB + 6;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  T* sj = reinterpret_cast<T*>(smem_raw) + (size_t)warp * 2 * RBD_MAX_DOF;    // [NB][2]: (f1, f2) of every joint
  for (int64_t b = (int64_t)blockIdx.x * nwarps + warp; b < B; b += (int64_t)gridDim.x * nwarps) {
    T* Mb = Minv + b * (int64_t)n * n;
    T* Fb = F + b * (int64_t)n * 6 * n;
    const T* Ub = U + b * (int64_t)n * 6;
    const T* Db = Dinv + b * (int64_t)n;
    __syncwarp();
    if (lane >= 1 && lane < NB) {
      T f1, f2;
      joint_basis(m.d, lane, q[b * nq + lane + 6], f1, f2);
      sj[2 * lane] = f1; sj[2 * lane + 1] = f2;
    }
    __syncwarp();
    for (int col = lane; col < n; col += 32) {
      T Fprev[6];                                          // F[i - 1][:, col]
#pragma unroll
      for (int r = 0; r < 6; ++r) {
        Fprev[r] = Mb[r * n + col];
        Fb[r * n + col] = Fprev[r];                                            // :779  F[0] = S Minv[0:6, :]
      }
      for (int i = 1; i < NB; ++i) {
        const int mi = i + 5, p = m.d.parent[i];
        T X[18], Ui[6], UX[6], Fp[6], XF[6];
        build_X(m.d, i, sj[2 * i], sj[2 * i + 1], X);
#pragma unroll
        for (int r = 0; r < 6; ++r) Ui[r] = Ub[mi * 6 + r];
        XT_apply(X, Ui, UX);                                                   // (U^T X)^T = X^T U
        const T invD = T(1) / Db[mi];
        if (p == i - 1) {
#pragma unroll
          for (int r = 0; r < 6; ++r) Fp[r] = Fprev[r];
        } else {
#pragma unroll
          for (int r = 0; r < 6; ++r) Fp[r] = Fb[(p * 6 + r) * n + col];       // written by this lane earlier
        }
        const T mij = Mb[mi * n + col] - invD * dot6(UX, Fp);                  // :771-773
        Mb[mi * n + col] = mij;
        X_apply(X, Fp, XF);
#pragma unroll
        for (int r = 0; r < 6; ++r) {
          Fprev[r] = fma_t(m.d.S[i][r], mij, XF[r]);                           // :774-776
          Fb[(i * 6 + r) * n + col] = Fprev[r];
        }
      }
    }
  }
}

// ---- minv_bpass, one COLUMN per lane (the scheme of minv_bpass_col_kernel, rbd_pass_kernels.cuh) -----------------------
// A warp owns one knot point.  Phase 1: the articulated inertias leaf -> base (:694-733), shared by the lanes through
// shared memory (lane c < 6 owns column c of IA_parent += X^T (IA - U U^T / D) X); every body's U and 1 / D stay in shared
// memory, U and D also go to the caller's arrays, the base's U[0:6] = IA_0 and fb_Dinv = inv(IA_0) (:677-684).  Phase 2:
// columns never mix in :700-726, so lane c carries matrix column c (and c + 32 when n > 32) from its body to the base with
// a running F in six registers and writes every row of Minv and F once, contiguously across the lanes (zeros outside the
// subtree; the reference's arrays start as zeros).  The reference's indexing is kept: rows / columns i + 5 for body i,
// children of the base accumulate into F[5] (:724-726 with parent index 0 + 5), which is what the base step reads
// (:687-691).
__host__ __device__ inline int fbp_minv_bpass_warp_vals(int NB) { return (45 * NB + 36 + 3) & ~3; }   // IA | U | 1/D | (f1, f2) | fb_Dinv

template <typename T>
__global__ void __launch_bounds__(kPassThreads)
fbp_minv_bpass_col_kernel(const __grid_constant__ FbModel<T> m, int64_t B, const T* __restrict__ q, T* __restrict__ Minv,
                          T* __restrict__ F, T* __restrict__ U, T* __restrict__ Dinv) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int NB = m.d.n, n = NB + 5, nq = NB + 6;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  T* ws = reinterpret_cast<T*>(smem_raw) + (size_t)warp * fbp_minv_bpass_warp_vals(NB);
  T* IA = ws;                    // [NB][36]
  T* Us = IA + 36 * NB;          // [NB][6]
  T* iD = Us + 6 * NB;           // [NB]
  T* sj = iD + NB;               // [NB][2]
  T* Di = sj + 2 * NB;           // [36]
  for (int64_t b = (int64_t)blockIdx.x * nwarps + warp; b < B; b += (int64_t)gridDim.x * nwarps) {
    T* Mb = Minv + b * (int64_t)n * n;
    T* Fb = F + b * (int64_t)n * 6 * n;                     // F[mi][r][col] = Fb[(mi * 6 + r) * n + col]
    T* Ub = U + b * (int64_t)n * 6;
    T* Db = Dinv + b * (int64_t)n;
    __syncwarp();
    for (int e = lane; e < NB * 36; e += 32) IA[e] = m.d.I[e / 36][e % 36];    // :662
    if (lane >= 1 && lane < NB) {
      T f1, f2;
      joint_basis(m.d, lane, q[b * nq + lane + 6], f1, f2);
      sj[2 * lane] = f1; sj[2 * lane + 1] = f2;
    }
    if (lane < 6) Db[lane] = T(0);
    __syncwarp();
    // ---------------------------------------------------------------- phase 1: articulated inertias
    for (int i = NB - 1; i >= 1; --i) {
      const int p = m.d.parent[i], mi = i + 5;
      T X[18];
      build_X(m.d, i, sj[2 * i], sj[2 * i + 1], X);
      if (lane < 6) {                                       // U = IA_i S (:697): lane r computes U[r]
        T acc = T(0);
#pragma unroll
        for (int k = 0; k < 6; ++k) acc = fma_t(IA[i * 36 + 6 * lane + k], m.d.S[i][k], acc);
        Us[i * 6 + lane] = acc;
      }
      __syncwarp();
      T Ui[6];
#pragma unroll
      for (int r = 0; r < 6; ++r) Ui[r] = Us[i * 6 + r];
      const T D = dot6(m.d.S[i], Ui);                                          // :698
      const T invD = T(1) / D;
      if (lane < 6) Ub[mi * 6 + lane] = Ui[lane];
      if (lane == 6) { Db[mi] = D; iD[i] = invD; }
      if (lane < 6) {                                       // IA_p += X^T (IA_i - U U^T / D) X (:728-733), column `lane`
        const int cc = lane < 3 ? lane : lane - 3;
        T xk[6];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const T e = cc == 0 ? X[3 * r] : (cc == 1 ? X[3 * r + 1] : X[3 * r + 2]);
          const T lo = cc == 0 ? X[9 + 3 * r] : (cc == 1 ? X[9 + 3 * r + 1] : X[9 + 3 * r + 2]);
          xk[r] = lane < 3 ? e : T(0);
          xk[3 + r] = lane < 3 ? lo : e;
        }
        const T ux = invD * dot6(Ui, xk);
        T colv[6], t[6];
#pragma unroll
        for (int r = 0; r < 6; ++r) {
          T acc = T(0);
#pragma unroll
          for (int k = 0; k < 6; ++k) acc = fma_t(IA[i * 36 + 6 * r + k], xk[k], acc);
          colv[r] = acc - Ui[r] * ux;
        }
        XT_apply(X, colv, t);
#pragma unroll
        for (int r = 0; r < 6; ++r) IA[p * 36 + 6 * r + lane] += t[r];
      }
      __syncwarp();
    }
    // the base (:677-684): U[0:6] = IA_0, fb_Dinv = inv(IA_0)
    for (int e = lane; e < 36; e += 32) Ub[e] = IA[e];
    if (lane == 0) {
      T s[21];
#pragma unroll
      for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int c = r; c < 6; ++c) s[sym6_idx(r, c)] = IA[6 * r + c];
      sym6_invert(s);
#pragma unroll
      for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int c = 0; c < 6; ++c) Di[6 * r + c] = s[sym6_idx(r, c)];
    }
    __syncwarp();
    // ---------------------------------------------------------------- phase 2: matrix column `col`, body col - 5
    for (int col = lane; col < n; col += 32) {
      const int jb = col - 5;
      T Frun[6] = {T(0), T(0), T(0), T(0), T(0), T(0)};
      for (int i = NB - 1; i >= 1; --i) {
        const int mi = i + 5;
        const bool mine = jb >= 1 && ((m.d.sub_mask[i] >> jb) & 1u);
        T mij = T(0), Fout[6] = {T(0), T(0), T(0), T(0), T(0), T(0)};
        if (mine) {
          const T invD = iD[i];
          mij = (jb == i ? invD : T(0)) - invD * dot6(m.d.S[i], Frun);         // :700-708
          T X[18];
#pragma unroll
          for (int r = 0; r < 6; ++r) Fout[r] = fma_t(Us[i * 6 + r], mij, Frun[r]);   // :721-723
          build_X(m.d, i, sj[2 * i], sj[2 * i + 1], X);
          XT_apply(X, Fout, Frun);                                             // now F[parent + 5][:, col] (:724-726)
        }
        Mb[mi * n + col] = mij;
#pragma unroll
        for (int r = 0; r < 6; ++r) Fb[(mi * 6 + r) * n + col] = Fout[r];
      }
      // rows 0..5: F[0..4] = 0, F[5] = what the base's children left (:724-726), Minv[0:6, col] (:686-691)
#pragma unroll
      for (int r = 0; r < 6; ++r) {
#pragma unroll
        for (int mi = 0; mi < 5; ++mi) Fb[(mi * 6 + r) * n + col] = T(0);
        Fb[(5 * 6 + r) * n + col] = Frun[r];
        T acc = T(0);
#pragma unroll
        for (int k = 0; k < 6; ++k) acc = fma_t(Di[6 * r + k], Frun[k], acc);
        Mb[r * n + col] = (col < 6 ? Di[6 * r + col] : T(0)) - acc;
      }
    }
  }
}

}  // namespace rbd
