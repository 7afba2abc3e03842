// rbd_ee_kernels.cuh - batched end-effector pose and pose gradient.
//
// Reference: end_effector_pose (RBDReference.py:220-283) and end_effector_pose_gradient (:295-386).
// For every requested end effector the reference multiplies the joints' 4x4 homogeneous transforms
// from the leaf up to the base (backwardChain :235-243), reads x, y, z of an offset point and
// roll / pitch / yaw of the rotation (:247-260), and for the gradient repeats the chain once per
// joint of the chain with that joint's transform replaced by its derivative (:312-324, :327-351).
//
// Here: one knot point per lane.  T_k(q) = A + B cos q + C sin q (or A + B q) comes from the model
// compiler, which probes the robot's own get_Xmat_hom_Func_by_id / get_dXmat_hom_Func_by_id.
// One leaf->base sweep leaves the suffix products S_t = T_t ... T_leaf F in local memory, one
// base->leaf sweep carries the prefix P_t = T_0 ... T_{t-1}; column t of the gradient then comes from
// P_t (dT_t S_{t+1}) - three 3x4 products per joint instead of the reference's O(depth) per column.
// Results leave through a shared-memory tile laid out like the warp's slab in HBM (a knot point's
// (6, n) block is contiguous), so every store instruction writes full lines.
#pragma once
#include "rbd_common.cuh"

#define RBD_MAX_EE 32

namespace rbd {

template <typename T>
struct EeModel {
  int n, n_ee;
  int kind[RBD_MAX_DOF];                      // 0: cos/sin basis, 1: affine in q
  int chain_len[RBD_MAX_EE];
  unsigned char chain[RBD_MAX_EE][RBD_MAX_DOF];  // joint ids base -> leaf
  signed char chain_pos[RBD_MAX_EE][RBD_MAX_DOF]; // position of joint j in chain e, -1 when off the chain
  T off[4];                                   // ee_offsets[0] = (x, y, z, w)   (:248, :335)
  T TA[RBD_MAX_DOF][12], TB[RBD_MAX_DOF][12], TC[RBD_MAX_DOF][12];   // rows 0..2 of T_k(q)
  T DA[RBD_MAX_DOF][12], DB[RBD_MAX_DOF][12], DC[RBD_MAX_DOF][12];   // rows 0..2 of dT_k/dq
  T fin[RBD_MAX_EE][12];                      // fixed-joint transform closing the chain (identity for a moving joint)
};

constexpr int kEeMaxWarps = 4;

__device__ __forceinline__ double atan2_t(double y, double x) { return atan2(y, x); }
__device__ __forceinline__ float atan2_t(float y, float x) { return atan2f(y, x); }
__device__ __forceinline__ double sqrt_t(double x) { return sqrt(x); }
__device__ __forceinline__ float sqrt_t(float x) { return sqrtf(x); }

// C = A B for 3x4 blocks of 4x4 matrices; the hidden fourth row of B is (0, 0, 0, wB)
template <typename T>
__device__ __forceinline__ void mul34(const T (&A)[12], const T (&Bm)[12], T wB, T (&C)[12]) {
#pragma unroll
  for (int r = 0; r < 3; ++r) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      T acc = A[4 * r] * Bm[c];
      acc = fma_t(A[4 * r + 1], Bm[4 + c], acc);
      acc = fma_t(A[4 * r + 2], Bm[8 + c], acc);
      if (c == 3) acc = fma_t(A[4 * r + 3], wB, acc);
      C[4 * r + c] = acc;
    }
  }
}

// d/dz atan2(y(z), x(z))  (RBDReference.py:337-338)
template <typename T>
__device__ __forceinline__ T darctan2(T y, T x, T yp, T xp) { return (-xp * y + x * yp) / (x * x + y * y); }

// pose_out   (B, n_ee, 6)      [x y z roll pitch yaw], may be null when GRAD
// grad_out   (B, n_ee, 6, n)   only when GRAD
// COEF_SMEM: joint coefficients staged in shared memory (small robots); otherwise they are read from
// the constant bank (large robots, where 72 n values per CTA would cost more occupancy than they save).
template <typename T, bool GRAD, bool COEF_SMEM>
__global__ void __launch_bounds__(kEeMaxWarps * 32)
ee_pose_kernel(const __grid_constant__ EeModel<T> m, int64_t B, const T* __restrict__ q, T* __restrict__ pose_out,
               T* __restrict__ grad_out, int grad_pitch, int compact) {
  extern __shared__ __align__(16) unsigned char ee_smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int n = m.n, n_ee = m.n_ee;
  // joint coefficients: [n][72] = TA | TB | TC | DA | DB | DC.  The joint index comes from the chain
  // table at run time; indexed constant-bank loads stalled the FMA chain (ncu: short scoreboard), a
  // broadcast shared-memory load does not.
  // compact tile ([lane][6][len], several end effectors): the copy-out looks the chain position of
  // every column up with a lane-dependent index - shared memory, not the constant bank
  __shared__ signed char s_chain_pos[RBD_MAX_EE][RBD_MAX_DOF];
  if (GRAD && compact)
    for (int k = threadIdx.x; k < RBD_MAX_EE * RBD_MAX_DOF; k += blockDim.x)
      s_chain_pos[k / RBD_MAX_DOF][k % RBD_MAX_DOF] = m.chain_pos[k / RBD_MAX_DOF][k % RBD_MAX_DOF];
  T* coef = reinterpret_cast<T*>(ee_smem_raw);
  if (COEF_SMEM) {
    for (int k = threadIdx.x; k < n * 72; k += blockDim.x) {
      const int j = k / 72, w = k - j * 72, which = w / 12, idx = w - which * 12;
      const T* src = which == 0 ? m.TA[j] : which == 1 ? m.TB[j] : which == 2 ? m.TC[j] : which == 3 ? m.DA[j] : which == 4 ? m.DB[j] : m.DC[j];
      coef[k] = src[idx];
    }
  }
  __syncthreads();
  const int pose_vals = 32 * n_ee * 6;
  const int per_warp = pose_vals + (GRAD ? 32 * grad_pitch : 0);
  T* pose_tile = coef + (COEF_SMEM ? ((n * 72 + 1) & ~1) : 0) + (size_t)warp * per_warp;
  T* grad_tile = pose_tile + pose_vals;
  const int64_t ntask = (B + 31) / 32;
  for (int64_t task = (int64_t)blockIdx.x * nwarps + warp; task < ntask; task += (int64_t)gridDim.x * nwarps) {
    const int64_t b0 = task * 32;
    const int64_t b = b0 + lane;
    const bool live = b < B;
    const int nlive = (int)((B - b0) < 32 ? (B - b0) : 32);
    const T* qb = q + (live ? b : b0) * n;
    for (int e = 0; e < n_ee; ++e) {
      const int len = m.chain_len[e];
      T suf[RBD_MAX_DOF][12];
      T f1s[RBD_MAX_DOF], f2s[RBD_MAX_DOF];
      T M[12];
#pragma unroll
      for (int k = 0; k < 12; ++k) M[k] = m.fin[e][k];
      // leaf -> base: suffix products (backwardChain, :235-243)
      for (int t = len - 1; t >= 0; --t) {
        const int j = m.chain[e][t];
        T f1, f2;
        if (m.kind[j] == 0) sincos_t(qb[j], &f2, &f1); else { f1 = qb[j]; f2 = T(0); }
        f1s[t] = f1; f2s[t] = f2;
        T Tk[12], Nw[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) { suf[t][k] = M[k]; Tk[k] = COEF_SMEM ? fma_t(coef[j * 72 + 24 + k], f2, fma_t(coef[j * 72 + 12 + k], f1, coef[j * 72 + k]))
                                                                  : fma_t(m.TC[j][k], f2, fma_t(m.TB[j][k], f1, m.TA[j][k])); }
        mul34(Tk, M, T(1), Nw);
#pragma unroll
        for (int k = 0; k < 12; ++k) M[k] = Nw[k];
      }
      // pose (:247-260): M rows are [R | p]; X[2,1] = M[9], X[2,2] = M[10], X[2,0] = M[8], X[1,0] = M[4], X[0,0] = M[0]
      const T sq = sqrt_t(M[10] * M[10] + M[9] * M[9]);
      {
        T* pt = pose_tile + (lane * n_ee + e) * 6;
#pragma unroll
        for (int r = 0; r < 3; ++r)
          pt[r] = fma_t(M[4 * r + 3], m.off[3], fma_t(M[4 * r + 2], m.off[2], fma_t(M[4 * r + 1], m.off[1], M[4 * r] * m.off[0])));
        pt[3] = atan2_t(M[9], M[10]);
        pt[4] = atan2_t(-M[8], sq);
        pt[5] = atan2_t(M[4], M[0]);
      }
      if (GRAD) {
        T* gt = grad_tile + lane * grad_pitch;                 // dense [6][n] or compact [6][len]
        const int gld = compact ? len : n;
        if (!compact)
          for (int k = 0; k < 6 * n; ++k) gt[k] = T(0);        // columns off the chain stay zero (:359-361)
        T P[12] = {T(1), T(0), T(0), T(0), T(0), T(1), T(0), T(0), T(0), T(0), T(1), T(0)};
        for (int t = 0; t < len; ++t) {
          const int j = m.chain[e][t];
          const T f1 = f1s[t], f2 = f2s[t];
          T dT[12], Tk[12], W[12], dX[12], S[12];
#pragma unroll
          for (int k = 0; k < 12; ++k) {
            S[k] = suf[t][k];
            dT[k] = COEF_SMEM ? fma_t(coef[j * 72 + 60 + k], f2, fma_t(coef[j * 72 + 48 + k], f1, coef[j * 72 + 36 + k]))
                              : fma_t(m.DC[j][k], f2, fma_t(m.DB[j][k], f1, m.DA[j][k]));
            Tk[k] = COEF_SMEM ? fma_t(coef[j * 72 + 24 + k], f2, fma_t(coef[j * 72 + 12 + k], f1, coef[j * 72 + k]))
                              : fma_t(m.TC[j][k], f2, fma_t(m.TB[j][k], f1, m.TA[j][k]));
          }
          mul34(dT, S, T(1), W);       // hidden row of dT is zero, of S it is (0,0,0,1)
          mul34(P, W, T(0), dX);
          // one gradient column (:327-351)
#pragma unroll
          for (int r = 0; r < 3; ++r)
            gt[r * gld + (compact ? t : j)] = fma_t(dX[4 * r + 3], m.off[3], fma_t(dX[4 * r + 2], m.off[2], fma_t(dX[4 * r + 1], m.off[1], dX[4 * r] * m.off[0])));
          const int gc = compact ? t : j;
          gt[3 * gld + gc] = darctan2(M[9], M[10], dX[9], dX[10]);
          const T dsq = (M[10] * dX[10] + M[9] * dX[9]) / sq;
          gt[4 * gld + gc] = darctan2(-M[8], sq, -dX[8], dsq);
          gt[5 * gld + gc] = darctan2(M[4], M[0], dX[4], dX[0]);
          mul34(P, Tk, T(1), W);
#pragma unroll
          for (int k = 0; k < 12; ++k) P[k] = W[k];
        }
        __syncwarp();
        // a knot point's (6, n) block of this end effector is contiguous in HBM
        const int row = 6 * n;
        if (n_ee == 1) {
          // the whole warp's slab is contiguous: full 256-byte stores whatever the row length
          T* dst = grad_out + b0 * (int64_t)row;
          const int total = nlive * row;
          int r = 0, c = lane;
          while (c >= row) { c -= row; ++r; }
          for (int k = lane; k < total; k += 32) {
            __stcs(dst + k, grad_tile[r * grad_pitch + c]);
            c += 32;
            while (c >= row) { c -= row; ++r; }
          }
        } else if (compact) {
          // where element k = lane + 32 it of a (6, n) block comes from in the compact tile (-1: a zero
          // column, :359-361); the same for every knot point, so the divisions are done once per task
          int src[6];
#pragma unroll
          for (int it = 0; it < 6; ++it) {
            const int k = lane + 32 * it;
            src[it] = -2;
            if (k < row) {
              const int comp = k / n, j = k - comp * n, pos = s_chain_pos[e][j];
              src[it] = pos >= 0 ? comp * len + pos : -1;
            }
          }
          for (int r = 0; r < nlive; ++r) {
            T* dst = grad_out + ((b0 + r) * n_ee + e) * (int64_t)row;
            const T* tl = grad_tile + r * grad_pitch;
#pragma unroll
            for (int it = 0; it < 6; ++it)
              if (src[it] > -2) __stcs(dst + lane + 32 * it, src[it] >= 0 ? tl[src[it]] : T(0));
          }
        } else {
          for (int r = 0; r < nlive; ++r) {
            T* dst = grad_out + ((b0 + r) * n_ee + e) * (int64_t)row;
            const T* src = grad_tile + r * grad_pitch;
            for (int k = lane; k < row; k += 32) __stcs(dst + k, src[k]);
          }
        }
        __syncwarp();
      }
    }
    if (pose_out) {
      __syncwarp();
      T* dst = pose_out + b0 * n_ee * 6;
      const int cnt = nlive * n_ee * 6;
      for (int k = lane; k < cnt; k += 32) __stcs(dst + k, pose_tile[k]);
    }
    __syncwarp();
  }
}

}  // namespace rbd
