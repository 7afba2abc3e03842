// rbd_fb_kernels.cuh - floating-base rnea / rnea_grad / minv.
//
// The `self.robot.floating_base` branches of the reference: rnea_fpass RBDReference.py:585/:591,
// minv_bpass :652-691, minv_fpass :761-779, rnea_grad_fpass_dq :1141-1168, rnea_grad_fpass_dqd
// :1212-1238, rnea_grad_bpass_dq :1267-1282, rnea_grad_bpass_dqd :1309-1341.
// Body 0 is the base: one 6-DoF joint with S = eye(6) whose transform is built from q[0:7]
// (position + unit quaternion, layout found by the model compiler); body i >= 1 is a 1-DoF joint
// that reads q[i + 6], qd[i + 5] and owns row / column i + 5 of every joint-space quantity.
// nv = NB + 5, nq = NB + 6.
//
// First correct path for this row of the scope table: one knot point per thread, the reference's
// body-frame recursion with per-thread local arrays (the structure of rbd_fused_kernels.cuh),
// gradient and inverse worked through one column at a time, results collected four columns at a
// time so that they leave as 32-byte pieces of each row.  What bounds it (ncu, HyQ + base rnea_grad):
// the per-thread state (36 values per visited body and column) does not stay in L2 - 6 GB read and
// 12 GB written for 1.4 GB of results; capping the resident CTAs did not change the time.
#pragma once
#include "rbd_common.cuh"

namespace rbd {

constexpr int kFbThreads = 64;
// resident CTAs the register allocation aims at (measured on B200, 2^18 knot points): minv gains from six
// (168 registers; Atlas + base 23.9 -> 17.9 ms, eight spills and is slower), rnea_grad does not (4.9 -> 5.1 ms)
constexpr int kFbMinvMinCtas = 6;
constexpr int kFbMaxNv = RBD_MAX_DOF + 5;

template <typename T>
struct FbModel {
  DevModel<T> d;        // n = NB bodies; entry 0 = base (only I and damping are used)
  int pos_off;          // q[pos_off .. +3]  base position
  int quat_off;         // q[quat_off .. +4] unit quaternion
  int w_first;          // 1: (w, x, y, z), 0: (x, y, z, w)
  int transpose;        // 0: E = R(quat)^T (coordinate transform world -> base), 1: E = R(quat)
  unsigned store_mask;  // bit i: body i has a child other than i + 1, so its per-column state must be kept in
                        // local memory; along chains (parent == i - 1) the state travels in registers
};

// X0 = xrot(E) xlt(p) in the 18-value layout [E | L], L = -E p^x
template <typename T>
__device__ __forceinline__ void fb_base_X(const FbModel<T>& m, const T* __restrict__ qb, T (&X)[18]) {
  const T px = qb[m.pos_off], py = qb[m.pos_off + 1], pz = qb[m.pos_off + 2];
  const T* qq = qb + m.quat_off;
  const T w = m.w_first ? qq[0] : qq[3];
  const T x = m.w_first ? qq[1] : qq[0], y = m.w_first ? qq[2] : qq[1], z = m.w_first ? qq[3] : qq[2];
  T R[9];
  R[0] = T(1) - T(2) * (y * y + z * z); R[1] = T(2) * (x * y - z * w); R[2] = T(2) * (x * z + y * w);
  R[3] = T(2) * (x * y + z * w); R[4] = T(1) - T(2) * (x * x + z * z); R[5] = T(2) * (y * z - x * w);
  R[6] = T(2) * (x * z - y * w); R[7] = T(2) * (y * z + x * w); R[8] = T(1) - T(2) * (x * x + y * y);
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) X[3 * r + c] = m.transpose ? R[3 * r + c] : R[3 * c + r];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const T e0 = X[3 * r], e1 = X[3 * r + 1], e2 = X[3 * r + 2];
    X[9 + 3 * r] = -(e1 * pz - e2 * py);
    X[9 + 3 * r + 1] = -(e2 * px - e0 * pz);
    X[9 + 3 * r + 2] = -(e0 * py - e1 * px);
  }
}

// Per-body constants in shared memory: [body][96] = XA(18) XB(18) XC(18) S(6) I(36).  The body index is a run-time
// value in every sweep, and indexed constant-bank loads (LDC with a register index) stalled the FMA chains
// (ncu: short scoreboard 12.6 stalls per issue in minv); a broadcast shared-memory load does not.
constexpr int kFbSmStride = 96;
template <typename T>
__device__ __forceinline__ void fb_stage_model(const FbModel<T>& m, T* sm) {
  const int NB = m.d.n;
  for (int k = threadIdx.x; k < NB * kFbSmStride; k += blockDim.x) {
    const int i = k / kFbSmStride, w = k - i * kFbSmStride;
    sm[k] = w < 18 ? m.d.XA[i][w] : w < 36 ? m.d.XB[i][w - 18] : w < 54 ? m.d.XC[i][w - 36] : w < 60 ? m.d.S[i][w - 54] : m.d.I[i][w - 60];
  }
  __syncthreads();
}
template <typename T>
__device__ __forceinline__ void fb_build_X(const T* sm, int i, T f1, T f2, T (&X)[18]) {
  const T* c = sm + i * kFbSmStride;
#pragma unroll
  for (int k = 0; k < 18; ++k) X[k] = fma_t(c[36 + k], f2, fma_t(c[18 + k], f1, c[k]));
}
template <typename T> __device__ __forceinline__ const T* fb_S(const T* sm, int i) { return sm + i * kFbSmStride + 54; }
template <typename T> __device__ __forceinline__ const T* fb_I(const T* sm, int i) { return sm + i * kFbSmStride + 60; }

// forward + backward sweep of rnea (:559-621) shared by the rnea and rnea_grad kernels
// SM: per-body constants from shared memory (rnea_grad) or from the constant bank (rnea: one sweep per knot
// point does not repay the staging, measured 10-20 % slower)
template <typename T, bool SM>
__device__ __forceinline__ void fb_rnea_state(const FbModel<T>& m, const T* sm, const T* qb, const T* qdb, const T* qddb, T gravity,
                                              const T (&X0)[18], T (*lv)[6], T (*la)[6], T (*lf)[6], T (*lb)[2]) {
  const int NB = m.d.n;
  for (int i = 0; i < NB; ++i) {
    T X[18], vi[6], ai[6], par[6], vJ[6], t[6];
    if (i == 0) {
      lb[0][0] = T(0); lb[0][1] = T(0);
#pragma unroll
      for (int r = 0; r < 6; ++r) { vi[r] = T(0); par[r] = T(0); }
      par[5] = -gravity;
      X_apply(X0, par, ai);                                                    // :578
#pragma unroll
      for (int r = 0; r < 6; ++r) vJ[r] = qdb[r];                              // :585, S = eye(6)
    } else {
      joint_basis(m.d, i, qb[i + 6], lb[i][0], lb[i][1]);
      if (SM) fb_build_X(sm, i, lb[i][0], lb[i][1], X); else build_X(m.d, i, lb[i][0], lb[i][1], X);
      const int p = m.d.parent[i];
#pragma unroll
      for (int r = 0; r < 6; ++r) par[r] = lv[p][r];
      X_apply(X, par, vi);
#pragma unroll
      for (int r = 0; r < 6; ++r) par[r] = la[p][r];
      X_apply(X, par, ai);
      const T qdi = qdb[i + 5];
#pragma unroll
      for (int r = 0; r < 6; ++r) vJ[r] = (SM ? fb_S(sm, i) : m.d.S[i])[r] * qdi;
    }
#pragma unroll
    for (int r = 0; r < 6; ++r) vi[r] += vJ[r];
    crm_mul(vi, vJ, t);                                                        // :588
#pragma unroll
    for (int r = 0; r < 6; ++r) ai[r] += t[r];
    if (qddb) {
      if (i == 0) {
#pragma unroll
        for (int r = 0; r < 6; ++r) ai[r] += qddb[r];                          // :591
      } else {
        const T qddi = qddb[i + 5];
#pragma unroll
        for (int r = 0; r < 6; ++r) ai[r] = fma_t((SM ? fb_S(sm, i) : m.d.S[i])[r], qddi, ai[r]);
      }
    }
    T Ia[6], Iv[6], vxIv[6];
    mat6_apply(SM ? fb_I(sm, i) : m.d.I[i], ai, Ia);
    mat6_apply(SM ? fb_I(sm, i) : m.d.I[i], vi, Iv);
    crf_mul(vi, Iv, vxIv);
#pragma unroll
    for (int r = 0; r < 6; ++r) { lv[i][r] = vi[r]; la[i][r] = ai[r]; lf[i][r] = Ia[r] + vxIv[r]; }
  }
  for (int i = NB - 1; i >= 1; --i) {                                          // :607-619
    T X[18], fi[6], t[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) fi[r] = lf[i][r];
    if (SM) fb_build_X(sm, i, lb[i][0], lb[i][1], X); else build_X(m.d, i, lb[i][0], lb[i][1], X);
    XT_apply(X, fi, t);
    const int p = m.d.parent[i];
#pragma unroll
    for (int r = 0; r < 6; ++r) lf[p][r] += t[r];
  }
}

// =============================================================================================
// rnea: q (B, NB+6), qd / qdd / c (B, NB+5), v / a / f (B, 6, NB)
// =============================================================================================
template <typename T>
__global__ void __launch_bounds__(kFbThreads)
fb_rnea_kernel(const __grid_constant__ FbModel<T> m, int64_t B, const T* __restrict__ q, const T* __restrict__ qd,
               const T* __restrict__ qdd, T gravity, T* __restrict__ c, T* __restrict__ v, T* __restrict__ a,
               T* __restrict__ f) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int NB = m.d.n, nv = NB + 5, nq = NB + 6;
  const T* qb = q + b * nq;
  T X0[18];
  fb_base_X(m, qb, X0);
  T lv[RBD_MAX_DOF][6], la[RBD_MAX_DOF][6], lf[RBD_MAX_DOF][6], lb[RBD_MAX_DOF][2];
  fb_rnea_state<T, false>(m, nullptr, qb, qd + b * nv, qdd ? qdd + b * nv : nullptr, gravity, X0, lv, la, lf, lb);
  T* cb = c + b * nv;
#pragma unroll
  for (int r = 0; r < 6; ++r) cb[r] = lf[0][r];                                // :612 with S = eye(6)
  for (int i = 1; i < NB; ++i) cb[i + 5] = dot6(m.d.S[i], lf[i]);
  T* outs[3] = {v, a, f};
  for (int w = 0; w < 3; ++w) {
    if (!outs[w]) continue;
    T* ob = outs[w] + b * 6 * NB;
    for (int i = 0; i < NB; ++i)
#pragma unroll
      for (int r = 0; r < 6; ++r) ob[r * NB + i] = (w == 0 ? lv : (w == 1 ? la : lf))[i][r];
  }
}

// =============================================================================================
// rnea_grad: dc_du (B, nv, 2 nv) = [dc_dq | dc_dqd]; columns 0..5 belong to the base (unit twists
// in base coordinates), column c >= 6 to body c - 5.
// =============================================================================================
template <typename T>
__global__ void __launch_bounds__(kFbThreads)
fb_rnea_grad_kernel(const __grid_constant__ FbModel<T> m, int64_t B, const T* __restrict__ q,
                    const T* __restrict__ qd, const T* __restrict__ qdd, T gravity, int use_damping,
                    T* __restrict__ dc_du, T* __restrict__ c_out) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  extern __shared__ __align__(16) unsigned char fb_smem_raw[];
  T* sm = reinterpret_cast<T*>(fb_smem_raw);
  fb_stage_model(m, sm);                      // every thread of the CTA takes part (barrier inside)
  if (b >= B) return;
  const int NB = m.d.n, nv = NB + 5, nq = NB + 6;
  const T* qb = q + b * nq;
  const T* qdb = qd + b * nv;
  T X0[18];
  fb_base_X(m, qb, X0);
  T lv[RBD_MAX_DOF][6], la[RBD_MAX_DOF][6], lf[RBD_MAX_DOF][6], lb[RBD_MAX_DOF][2];
  fb_rnea_state<T, true>(m, sm, qb, qdb, qdd ? qdd + b * nv : nullptr, gravity, X0, lv, la, lf, lb);
  if (c_out) {
    T* cb = c_out + b * nv;
#pragma unroll
    for (int r = 0; r < 6; ++r) cb[r] = lf[0][r];
    for (int i = 1; i < NB; ++i) cb[i + 5] = dot6(fb_S(sm, i), lf[i]);
  }
  T* gout = dc_du + b * (int64_t)2 * nv * nv;
  const int ld = 2 * nv;
  // Results are produced one column at a time but stored row-major: four columns are collected in
  // obuf[row][dq | dqd][column & 3] and flushed as 32-byte pieces of each row, so that the 8-byte
  // stores of one thread fill whole sectors (ncu, before: 22 GB of DRAM traffic for 1.4 GB of results).
  T obuf[kFbMaxNv][8];
  T sdv[RBD_MAX_DOF][12], sda[RBD_MAX_DOF][12], sdf[RBD_MAX_DOF][12];   // per body: d/dq (6) | d/dqd (6) of column c
  T Xg[6];                                                                  // X0 a_base (:1175)
  {
    T g6[6] = {T(0), T(0), T(0), T(0), T(0), -gravity};
    X_apply(X0, g6, Xg);
  }
  for (int c = 0; c < nv; ++c) {
    const int bc = c < 6 ? 0 : c - 5;
    const int cg = c & 3;
    const unsigned sub = m.d.sub_mask[bc];
    T pvq[6], paq[6], pvd[6], pad[6];          // the previously visited body's dv / da (d/dq | d/dqd)
    int prev = -1;
    for (int i = bc; i < NB; ++i) {
      if (!((sub >> i) & 1u)) continue;
      T dvq[6], daq[6], dvd[6], dad[6], t[6], vi[6];
#pragma unroll
      for (int r = 0; r < 6; ++r) vi[r] = lv[i][r];
      if (i == 0) {
        // base column k = c: S = eye(6), so the seeds are columns of crm(.) (:1175, :1231-1243)
        T ek[6] = {T(0), T(0), T(0), T(0), T(0), T(0)};
        ek[c] = T(1);
        T qd0[6];
#pragma unroll
        for (int r = 0; r < 6; ++r) { dvq[r] = T(0); dvd[r] = ek[r]; qd0[r] = qdb[r]; }
        crm_mul(Xg, ek, daq);
        crm_mul(dvd, qd0, dad);                    // sum_ii qd[ii] crm(dv)[:, ii]   (:1236-1238)
        crm_mul(vi, ek, t);                        // crm(v_0)[:, k]                 (:1243)
#pragma unroll
        for (int r = 0; r < 6; ++r) dad[r] += t[r];
      } else {
        T X[18], S[6];
        fb_build_X(sm, i, lb[i][0], lb[i][1], X);
        const T qdi = qdb[i + 5];
        const int p = m.d.parent[i];
#pragma unroll
        for (int r = 0; r < 6; ++r) S[r] = fb_S(sm, i)[r];
        if (i == bc) {
          T par[6], xp[6];
#pragma unroll
          for (int r = 0; r < 6; ++r) par[r] = lv[p][r];
          X_apply(X, par, xp);
          crm_mul(xp, S, dvq);                                                 // :1159
#pragma unroll
          for (int r = 0; r < 6; ++r) par[r] = la[p][r];
          X_apply(X, par, xp);
          crm_mul(xp, S, daq);                                                 // :1173
#pragma unroll
          for (int r = 0; r < 6; ++r) dvd[r] = S[r];                           // :1231
          crm_mul(vi, S, dad);                                                 // :1243
        } else {
          if (p != prev) {                     // parent is a branch point: its state was kept in local memory
#pragma unroll
            for (int r = 0; r < 6; ++r) { pvq[r] = sdv[p][r]; paq[r] = sda[p][r]; pvd[r] = sdv[p][6 + r]; pad[r] = sda[p][6 + r]; }
          }
          X_apply(X, pvq, dvq);
          X_apply(X, paq, daq);
          X_apply(X, pvd, dvd);
          X_apply(X, pad, dad);
        }
        crm_mul(dvq, S, t);                                                    // :1170
#pragma unroll
        for (int r = 0; r < 6; ++r) daq[r] = fma_t(qdi, t[r], daq[r]);
        crm_mul(dvd, S, t);                                                    // :1240
#pragma unroll
        for (int r = 0; r < 6; ++r) dad[r] = fma_t(qdi, t[r], dad[r]);
      }
      T Iv[6], Ida[6], Idv[6], t1[6], t2[6];
      mat6_apply(fb_I(sm, i), vi, Iv);
      mat6_apply(fb_I(sm, i), daq, Ida);
      mat6_apply(fb_I(sm, i), dvq, Idv);
      crf_mul(dvq, Iv, t1);
      crf_mul(vi, Idv, t2);
#pragma unroll
      for (int r = 0; r < 6; ++r) sdf[i][r] = Ida[r] + t1[r] + t2[r];
      mat6_apply(fb_I(sm, i), dad, Ida);
      mat6_apply(fb_I(sm, i), dvd, Idv);
      crf_mul(dvd, Iv, t1);
      crf_mul(vi, Idv, t2);
#pragma unroll
      for (int r = 0; r < 6; ++r) sdf[i][6 + r] = Ida[r] + t1[r] + t2[r];
      if ((m.store_mask >> i) & 1u) {
#pragma unroll
        for (int r = 0; r < 6; ++r) { sdv[i][r] = dvq[r]; sda[i][r] = daq[r]; sdv[i][6 + r] = dvd[r]; sda[i][6 + r] = dad[r]; }
      }
#pragma unroll
      for (int r = 0; r < 6; ++r) { pvq[r] = dvq[r]; paq[r] = daq[r]; pvd[r] = dvd[r]; pad[r] = dad[r]; }
      prev = i;
    }
    // backward over subtree(bc), then up the ancestors of bc to the base
    T Fq[6], Fd[6];
    T cq[6], cd[6];                             // X^T F of a chain child (parent == child - 1), carried in registers
    int carry_to = -1;
    for (int i = NB - 1; i >= bc; --i) {
      if (!((sub >> i) & 1u)) continue;
#pragma unroll
      for (int r = 0; r < 6; ++r) { Fq[r] = sdf[i][r]; Fd[r] = sdf[i][6 + r]; }
      if (carry_to == i) {
#pragma unroll
        for (int r = 0; r < 6; ++r) { Fq[r] += cq[r]; Fd[r] += cd[r]; }
      }
      if (i == 0) break;                        // base column: rows 0..5 are written below
      obuf[i + 5][cg] = dot6(fb_S(sm, i), Fq);                                    // :1284
      obuf[i + 5][4 + cg] = dot6(fb_S(sm, i), Fd);                                // :1325
      T X[18], tq[6], td[6];
      fb_build_X(sm, i, lb[i][0], lb[i][1], X);
      if (i == bc) {                                                           // :1292-1294
        T fi[6], S[6], fxs[6];
#pragma unroll
        for (int r = 0; r < 6; ++r) { fi[r] = lf[i][r]; S[r] = fb_S(sm, i)[r]; }
        crm_mul(fi, S, fxs);
#pragma unroll
        for (int r = 0; r < 6; ++r) Fq[r] -= fxs[r];
      }
      XT_apply(X, Fq, tq);                                                     // :1291
      XT_apply(X, Fd, td);                                                     // :1331
      if (i == bc) {
#pragma unroll
        for (int r = 0; r < 6; ++r) { Fq[r] = tq[r]; Fd[r] = td[r]; }
      } else {
        const int p = m.d.parent[i];
        if (p == i - 1) {
#pragma unroll
          for (int r = 0; r < 6; ++r) { cq[r] = tq[r]; cd[r] = td[r]; }
          carry_to = p;
        } else {
#pragma unroll
          for (int r = 0; r < 6; ++r) { sdf[p][r] += tq[r]; sdf[p][6 + r] += td[r]; }
        }
      }
    }
    unsigned touched = sub;
    if (bc != 0) {
      for (int j = m.d.parent[bc]; j > 0; j = m.d.parent[j]) {
        touched |= 1u << j;
        obuf[j + 5][cg] = dot6(fb_S(sm, j), Fq);
        obuf[j + 5][4 + cg] = dot6(fb_S(sm, j), Fd);
        T X[18], tq[6], td[6];
        fb_build_X(sm, j, lb[j][0], lb[j][1], X);
        XT_apply(X, Fq, tq);
        XT_apply(X, Fd, td);
#pragma unroll
        for (int r = 0; r < 6; ++r) { Fq[r] = tq[r]; Fd[r] = td[r]; }
      }
    }
#pragma unroll
    for (int r = 0; r < 6; ++r) {                                              // :1282, :1325 with S = eye(6)
      obuf[r][cg] = Fq[r];
      obuf[r][4 + cg] = Fd[r];
    }
    for (int i = 1; i < NB; ++i) {
      if ((touched >> i) & 1u) continue;
      obuf[i + 5][cg] = T(0);
      obuf[i + 5][4 + cg] = T(0);
    }
    if (cg == 3 || c == nv - 1) {
      const int c0 = c - cg;
      if (cg == 3) {
        // full group: all eight loads of a row are issued before its stores (the generic loop below waited
        // on one local-memory load at a time: 22 % of the stall samples)
#pragma unroll 2
        for (int row = 0; row < nv; ++row) {
          T* orow = gout + row * ld + c0;
          T w[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) w[k] = obuf[row][k];
#pragma unroll
          for (int k = 0; k < 4; ++k) { orow[k] = w[k]; orow[nv + k] = w[4 + k]; }
        }
      } else {
        for (int row = 0; row < nv; ++row) {
          T* orow = gout + row * ld + c0;
          for (int k = 0; k <= cg; ++k) orow[k] = obuf[row][k];
          for (int k = 0; k <= cg; ++k) orow[nv + k] = obuf[row][4 + k];
        }
      }
    }
  }
  T* out = gout;
  if (use_damping) {                                                           // :1336-1341, to the letter
    for (int r = 0; r < 5; ++r)
      for (int cc = 0; cc < 5; ++cc) out[r * ld + nv + cc] += m.d.damping[0];
    for (int i = 1; i < NB; ++i) out[i * ld + nv + i] += m.d.damping[i];
  }
}

// =============================================================================================
// minv: Minv (B, nv, nv).  Phase A: articulated inertias leaf -> base, base block inverted as a
// dense 6x6 (:681-684).  Then one column at a time: walk from the column's body to the base
// (:697-726, :686-691) and sweep every body root -> leaves (:760-781, whole rows as upstream).
// =============================================================================================
template <typename T>
__global__ void __launch_bounds__(kFbThreads, kFbMinvMinCtas)
fb_minv_kernel(const __grid_constant__ FbModel<T> m, int64_t B, const T* __restrict__ q, int output_dense,
               T* __restrict__ Minv) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  extern __shared__ __align__(16) unsigned char fb_smem_raw[];
  T* sm = reinterpret_cast<T*>(fb_smem_raw);
  fb_stage_model(m, sm);                      // every thread of the CTA takes part (barrier inside)
  if (b >= B) return;
  const int NB = m.d.n, nv = NB + 5, nq = NB + 6;
  const T* qb = q + b * nq;
  T* Mb = Minv + b * (int64_t)nv * nv;
  T IA[RBD_MAX_DOF][36];
  T lb[RBD_MAX_DOF][2], lU[RBD_MAX_DOF][6], lUX[RBD_MAX_DOF][6], linvD[RBD_MAX_DOF];
  for (int i = 0; i < NB; ++i) {
#pragma unroll
    for (int k = 0; k < 36; ++k) IA[i][k] = fb_I(sm, i)[k];
    if (i > 0) joint_basis(m.d, i, qb[i + 6], lb[i][0], lb[i][1]);
  }
  for (int i = NB - 1; i >= 1; --i) {
    T S[6], Ui[6], X[18], UX[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) S[r] = fb_S(sm, i)[r];
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      T acc = T(0);
#pragma unroll
      for (int k = 0; k < 6; ++k) acc = fma_t(IA[i][6 * r + k], S[k], acc);
      Ui[r] = acc;
      lU[i][r] = acc;                                                          // :697
    }
    const T invD = T(1) / dot6(S, Ui);                                         // :698
    linvD[i] = invD;
    fb_build_X(sm, i, lb[i][0], lb[i][1], X);
    XT_apply(X, Ui, UX);
#pragma unroll
    for (int r = 0; r < 6; ++r) lUX[i][r] = UX[r];
    // IA_p += X^T (IA - U U^T / D) X   (:728-733), column by column
    const int p = m.d.parent[i];
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
      const T ux = invD * dot6(Ui, xk);
#pragma unroll
      for (int r = 0; r < 6; ++r) {
        T acc = T(0);
#pragma unroll
        for (int cc = 0; cc < 6; ++cc) acc = fma_t(IA[i][6 * r + cc], xk[cc], acc);
        col[r] = acc - Ui[r] * ux;
      }
      XT_apply(X, col, t);
#pragma unroll
      for (int r = 0; r < 6; ++r) IA[p][6 * r + k] += t[r];
    }
  }
  // base: fb_Dinv = inv(IA_0) (:681-684), Gauss-Jordan on the symmetric positive definite 6x6
  T Di[36];
  {
    T A[36];
#pragma unroll
    for (int k = 0; k < 36; ++k) { A[k] = IA[0][k]; Di[k] = (k % 7 == 0) ? T(1) : T(0); }
#pragma unroll
    for (int p = 0; p < 6; ++p) {
      const T inv = T(1) / A[7 * p];
#pragma unroll
      for (int cc = 0; cc < 6; ++cc) { A[6 * p + cc] *= inv; Di[6 * p + cc] *= inv; }
#pragma unroll
      for (int r = 0; r < 6; ++r) {
        if (r == p) continue;
        const T fct = A[6 * r + p];
#pragma unroll
        for (int cc = 0; cc < 6; ++cc) {
          A[6 * r + cc] = fma_t(-fct, A[6 * p + cc], A[6 * r + cc]);
          Di[6 * r + cc] = fma_t(-fct, Di[6 * p + cc], Di[6 * r + cc]);
        }
      }
    }
  }
  T colM[kFbMaxNv];
  T colF[RBD_MAX_DOF][6];
  T obuf[kFbMaxNv][4];      // four finished columns, flushed as 32-byte pieces of each row
  for (int j = 0; j < nv; ++j) {
    const int jg = j & 3;
    for (int i = 0; i < nv; ++i) colM[i] = T(0);
    if (j < 6) {
#pragma unroll
      for (int r = 0; r < 6; ++r) colM[r] = Di[6 * r + j];                     // :686
    } else {
      T F[6] = {T(0), T(0), T(0), T(0), T(0), T(0)};
      for (int i = j - 5; i > 0; i = m.d.parent[i]) {                          // :697-726 restricted to column j
        const T invD = linvD[i];
        const T mij = (i + 5 == j ? invD : T(0)) - invD * dot6(fb_S(sm, i), F);
        colM[i + 5] = mij;
        T X[18], t[6];
#pragma unroll
        for (int r = 0; r < 6; ++r) F[r] = fma_t(lU[i][r], mij, F[r]);
        fb_build_X(sm, i, lb[i][0], lb[i][1], X);
        XT_apply(X, F, t);
#pragma unroll
        for (int r = 0; r < 6; ++r) F[r] = t[r];
      }
#pragma unroll
      for (int r = 0; r < 6; ++r) {                                            // :687-691
        T acc = T(0);
#pragma unroll
        for (int k = 0; k < 6; ++k) acc = fma_t(Di[6 * r + k], F[k], acc);
        colM[r] = -acc;
      }
    }
    // forward pass restricted to column j (:760-781)
#pragma unroll
    for (int r = 0; r < 6; ++r) { colF[0][r] = colM[r]; obuf[r][jg] = colM[r]; }     // :779 (the base always keeps its F)
    T Fprev[6];                                  // F of body i - 1, in registers along chains
#pragma unroll
    for (int r = 0; r < 6; ++r) Fprev[r] = colM[r];
    for (int i = 1; i < NB; ++i) {
      const int p = m.d.parent[i];
      T X[18], Fp[6], Fi[6];
#pragma unroll
      for (int r = 0; r < 6; ++r) Fp[r] = (p == i - 1) ? Fprev[r] : colF[p][r];
      const T mij = colM[i + 5] - linvD[i] * dot6(lUX[i], Fp);                 // :771-773
      fb_build_X(sm, i, lb[i][0], lb[i][1], X);
      X_apply(X, Fp, Fi);
#pragma unroll
      for (int r = 0; r < 6; ++r) Fprev[r] = fma_t(fb_S(sm, i)[r], mij, Fi[r]);   // :774-776
      if ((m.store_mask >> i) & 1u) {
#pragma unroll
        for (int r = 0; r < 6; ++r) colF[i][r] = Fprev[r];
      }
      obuf[i + 5][jg] = mij;
    }
    // :799-804 mirrors the leading NB x NB block (range(NB), not nv): Minv[row, col] = Minv[col, row]
    // for col < row < NB.  Column j is final above its diagonal, so it also IS the left part of row j,
    // which is contiguous in memory; entries of the block below the diagonal are not stored from the
    // column pass at all.
    if (output_dense && j < NB)
      for (int i = 0; i < j; ++i) Mb[j * nv + i] = obuf[i][jg];
    if (jg == 3 || j == nv - 1) {
      const int j0 = j - jg;
#pragma unroll 2
      for (int row = 0; row < nv; ++row) {
        T* orow = Mb + row * nv + j0;
        T w[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) w[k] = obuf[row][k];                       // loads first (entries past jg are unused)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const bool mirrored = output_dense && row < NB && row > j0 + k;      // j0 + k < row < NB: filled by row `row`'s pass
          if (k <= jg && !mirrored) orow[k] = w[k];
        }
      }
    }
  }
}

}  // namespace rbd
