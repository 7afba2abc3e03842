// rbd_coop_kernels.cuh - warp-cooperative fused rnea_grad: one BODY per lane.
//
// Same world-frame composite formulation as rbd_grad_kernels.cuh, mapped the other way round:
// a group of G = 8 / 16 / 32 lanes owns one knot point and lane i owns body i (bodies in
// depth-first preorder), so a warp evaluates 32 / G knot points at a time.
//
//   * independent branches of the tree run concurrently by construction (every body is a lane);
//   * the root-path recursions (pose, v, a) are pointer-jumping scans over the ancestor chain:
//     ceil(log2(depth + 1)) rounds of warp shuffles instead of `depth` sequential steps;
//   * subtree composites (inertia, momentum, Coriolis, force: 28 numbers) are a suffix scan over
//     the contiguous preorder range of the subtree: comp_i = PS(i) - PS(subtree_end(i));
//   * each lane walks its own ancestors for the four dot products per (body, ancestor) pair,
//     reading the ancestor's S / Psi_dot / Psi_ddot from a per-warp shared-memory table;
//   * dc_du of the warp's knot points is assembled in shared memory and written to HBM as one
//     contiguous, fully coalesced slab.
// Per-thread state is ~O(1) six-vectors, so residency is no longer limited by the O(n) working
// set of a knot point (the limit of the thread-per-knot-point kernels on Atlas-sized trees).
#pragma once
#include "rbd_common.cuh"
#include "rbd_grad_kernels.cuh"
#include "rbd_minv_kernels.cuh"

namespace rbd {

constexpr int kCoopWarps = 4;          // warps per CTA
constexpr int kCoopMdlStride = 51;     // per-body constants in shared memory (odd stride)
constexpr int kCoopIntStride = 10;     // parent kind sub_end orig jump[0..4] pad
constexpr int kCoopVecStride = 19;     // S(6) Psi_dot(6) Psi_ddot(6) + pad (odd stride)

struct CoopPlan {
  int nsteps;                          // pointer-jumping rounds
  int maxdepth;                        // longest ancestor chain
  int ncomp;                           // root components; component c = bodies [comp_begin[c], comp_begin[c+1])
  int comp_begin[RBD_MAX_DOF + 1];
  int jump[5][RBD_MAX_DOF];            // jump[s][i] = ancestor of i at distance 2^s (DFS ids) or -1
};

// Floating base (FB = true; RBDReference.py :585/:591, :1141-1168, :1212-1238, :1267-1282,
// :1309-1341): body 0 of the model is the base, a 6-DoF joint with S = eye(6); body i >= 1 reads q[i + 6], qd[i + 5] and
// owns row / column i + 5.  The kernel then works in BASE coordinates instead of world coordinates: the base's pose only
// enters through the direction of gravity (X_0 a_grav), its motion subspace is the unit vectors, its own velocity and
// acceleration (qd[0:6], X_0 a_grav + qdd[0:6]) seed the root-path sums, and it is one more lane of the group whose
// subtree composite is the whole robot.  Columns 0..5 are derivatives along unit twists of the base in base
// coordinates (Psi_dot = 0, Psi_ddot = (X_0 a_grav) x e_k, S_dot = v_0 x e_k).
struct FbBaseLayout {
  int quat_off;                        // q[quat_off .. +4] unit quaternion
  int w_first;                         // 1: (w, x, y, z), 0: (x, y, z, w)
  int transpose;                       // 0: E = R(quat)^T, 1: E = R(quat)
};

template <typename T>
__device__ __forceinline__ T shfl_t(T x, int src) { return __shfl_sync(0xffffffffu, x, src); }

constexpr int kCoopScanStride = 29;    // 28 composite values per body, odd stride
__host__ __device__ inline int coop_grad_tile_stride(int n, int ipw, bool split) {
  const int t = ipw * n * (split ? n : 2 * n), s = 32 * kCoopScanStride;
  return ((t > s ? t : s) + 1) & ~1;
}

// CONLY = true: rnea only (c into c_out; dc_du is not touched): the forward scans, the composite
// force f^C by a 6-value segmented scan and c_i = S_i . f^C_i.  Used for small (MPC-sized) batches,
// where the knot-point-per-lane rnea kernel is bound by the latency of one 32-knot-point task.
template <typename T, int G, bool SPLIT, bool CONLY = false, bool FB = false>
__global__ void __launch_bounds__(kCoopWarps * 32)
rnea_grad_coop_kernel(const __grid_constant__ FastModel<T> m, const __grid_constant__ DfsPlan plan,
                      const __grid_constant__ CoopPlan cp, int64_t B, const T* __restrict__ q,
                      const T* __restrict__ qd, const T* __restrict__ qdd, T gravity, int use_damping,
                      T* __restrict__ dc_du, T* __restrict__ c_out, const FbBaseLayout fbl) {
  static_assert(!(FB && CONLY), "the floating-base rnea has its own kernel");
  constexpr int IPW = 32 / G;                          // knot points per warp
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = m.n;                                   // bodies (FB: the base is body 0)
  const int nv = FB ? n + 5 : n;                       // joint-space size: rows of dc_du, entries of qd / qdd / c
  const int nq = FB ? n + 6 : n;                       // entries of q
  const int n2 = 2 * nv;
  T* mdl = reinterpret_cast<T*>(smem_raw);                                   // [n][51]
  T* vec_all = mdl + ((n * kCoopMdlStride + 1) & ~1);                        // [warps][32][19]
  T* tile_all = vec_all + kCoopWarps * 32 * kCoopVecStride;                  // [warps][tile_vals (+1)]
  const int tile_stride = coop_grad_tile_stride(nv, IPW, SPLIT);  // also holds the 32 x 29 composite-scan buffer
  int* imdl = reinterpret_cast<int*>(tile_all + kCoopWarps * tile_stride);   // [n][10]

  // ---- robot constants -> shared memory (once per CTA)
  for (int idx = threadIdx.x; idx < n * kCoopMdlStride; idx += blockDim.x) {
    const int i = idx / kCoopMdlStride, k = idx - i * kCoopMdlStride;
    T val = T(0);
    if (k < 9) val = m.EA[i][k];
    else if (k < 18) val = m.EB[i][k - 9];
    else if (k < 27) val = m.EC[i][k - 18];
    else if (k < 30) val = m.rA[i][k - 27];
    else if (k < 33) val = m.rB[i][k - 30];
    else if (k < 36) val = m.rC[i][k - 33];
    else if (k < 39) val = m.axis[i][k - 36];
    else if (k == 39) val = m.mass[i];
    else if (k < 43) val = m.h[i][k - 40];
    else if (k < 49) val = m.Ib[i][k - 43];
    else if (k == 49) val = m.damping[i];
    mdl[idx] = val;
  }
  for (int idx = threadIdx.x; idx < n * kCoopIntStride; idx += blockDim.x) {
    const int i = idx / kCoopIntStride, k = idx - i * kCoopIntStride;
    int val = 0;
    if (k == 0) val = m.parent[i];
    else if (k == 1) val = m.kind[i];
    else if (k == 2) val = plan.sub_end[i];
    else if (k == 3) val = plan.orig[i];
    else if (k < 9) val = cp.jump[k - 4][i];
    imdl[idx] = val;
  }
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane / G, i = lane - g * G;
  const bool valid = i < n;
  const int ib = valid ? i : 0;
  const int gbase = g * G;
  const T* mb = mdl + ib * kCoopMdlStride;
  const int* ip = imdl + ib * kCoopIntStride;
  const int par = valid ? ip[0] : -1;
  const int kind = ip[1];
  const int sub_end = ip[2];
  const bool cut = valid && sub_end < plan.comp_end[ib];   // bodies of the same component follow the subtree
  const int oi = ip[3];
  const bool base = FB && valid && i == 0;             // the lane of the floating base
  const int qoff = FB ? oi + 6 : oi;                   // this body's entry of q
  const int voff = FB ? oi + 5 : oi;                   // ... of qd / qdd / c, and its row / column of dc_du
  T* vec = vec_all + warp * 32 * kCoopVecStride;
  T* tile = tile_all + warp * tile_stride;
  T* myvec = vec + lane * kCoopVecStride;
  const int nsteps = cp.nsteps;

  const int64_t ngroups = (B + IPW - 1) / IPW;
  const int64_t gstride = (int64_t)gridDim.x * kCoopWarps;
  // the inputs of the next knot points are requested one iteration ahead (HBM latency is hidden
  // behind a whole evaluation)
  T q_nx = T(0), qd_nx = T(0), qdd_nx = T(0);
  // FB: the base's 16 inputs (quaternion, qd[0:6], qdd[0:6]) are fetched by the lanes of the group, NBX each, also one
  // iteration ahead: input t = i + u G is q[quat_off + t] (t < 4), qd[t - 4] (t < 10) or qdd[t - 10]
  constexpr int NBX = FB ? (G == 8 ? 2 : 1) : 1;
  T bx_nx[NBX];
  auto base_input = [&](int64_t bb, int t) -> T {
    if (t < 4) return q[bb * nq + fbl.quat_off + t];
    if (t < 10) return qd[bb * nv + (t - 4)];
    return (qdd && t < 16) ? qdd[bb * nv + (t - 10)] : T(0);
  };
#pragma unroll
  for (int u = 0; u < NBX; ++u) bx_nx[u] = T(0);
  {
    const int64_t grp0 = (int64_t)blockIdx.x * kCoopWarps + warp;
    if (grp0 < ngroups) {
      int64_t b = grp0 * IPW + g;
      if (b >= B) b = B - 1;
      q_nx = q[b * nq + qoff];
      qd_nx = qd[b * nv + voff];
      qdd_nx = qdd ? qdd[b * nv + voff] : T(0);
      if (FB) {
#pragma unroll
        for (int u = 0; u < NBX; ++u) bx_nx[u] = base_input(b, i + u * G);
      }
    }
  }
  for (int64_t grp = (int64_t)blockIdx.x * kCoopWarps + warp; grp < ngroups; grp += gstride) {
    int64_t b = grp * IPW + g;
    if (b >= B) b = B - 1;                                    // duplicate work, never stored

    // ------------------------------------------------------------------ forward
    T E[9], p[3], S[6], Pd[6], Pdd[6], v[6], a[6];
    T qdi, qddi;
    {
      const T qi = q_nx;
      qdi = qd_nx;
      qddi = qdd_nx;
      if (grp + gstride < ngroups) {
        int64_t bn = (grp + gstride) * IPW + g;
        if (bn >= B) bn = B - 1;
        q_nx = q[bn * nq + qoff];
        qd_nx = qd[bn * nv + voff];
        qdd_nx = qdd ? qdd[bn * nv + voff] : T(0);
      }
      if (FB) {
        // raw inputs of the base -> its row of the table: qd[0:6] at 0..5 (= v_0, :585 with S = eye(6)), the
        // quaternion at 6..9, qdd[0:6] at 12..17
        T* vb = vec + gbase * kCoopVecStride;
        T bx[NBX];
#pragma unroll
        for (int u = 0; u < NBX; ++u) bx[u] = bx_nx[u];
        if (grp + gstride < ngroups) {
          int64_t bn = (grp + gstride) * IPW + g;
          if (bn >= B) bn = B - 1;
#pragma unroll
          for (int u = 0; u < NBX; ++u) bx_nx[u] = base_input(bn, i + u * G);
        }
#pragma unroll
        for (int u = 0; u < NBX; ++u) {
          const int t = i + u * G;
          if (t < 16) vb[t < 4 ? 6 + t : (t < 10 ? t - 4 : t + 2)] = bx[u];
        }
        __syncwarp();
      }
      if (FB && base) {
        // X_0 a_grav (:578; only the rotation matters: a_grav is a pure linear acceleration) and
        // a_0 = X_0 a_grav + qdd[0:6] (:591; crm(v_0) v_0 = 0) replace the raw values
        const T* qq = myvec + 6;
        const T qw = fbl.w_first ? qq[0] : qq[3];
        const T qx = fbl.w_first ? qq[1] : qq[0], qy = fbl.w_first ? qq[2] : qq[1], qz = fbl.w_first ? qq[3] : qq[2];
        T e2[3];                                            // third column of E: the world's z axis in base coordinates
        e2[0] = fbl.transpose ? T(2) * (qx * qz + qy * qw) : T(2) * (qx * qz - qy * qw);
        e2[1] = fbl.transpose ? T(2) * (qy * qz - qx * qw) : T(2) * (qy * qz + qx * qw);
        e2[2] = T(1) - T(2) * (qx * qx + qy * qy);
#pragma unroll
        for (int k = 0; k < 6; ++k) {
          const T xg = k < 3 ? T(0) : -gravity * e2[k - 3];
          myvec[6 + k] = xg;
          myvec[12 + k] += xg;
        }
      }
      T f1, f2;
      if (kind == 0) sincos_t(qi, &f2, &f1);
      else { f1 = qi; f2 = T(0); }
#pragma unroll
      for (int k = 0; k < 9; ++k) E[k] = fma_t(mb[18 + k], f2, fma_t(mb[9 + k], f1, mb[k]));
#pragma unroll
      for (int k = 0; k < 3; ++k) p[k] = fma_t(mb[33 + k], f2, fma_t(mb[30 + k], f1, mb[27 + k]));
    }
    // pose scan over the ancestor chain: (E, p) <- (E, p) o (E_anc, p_anc)
    for (int s = 0; s < nsteps; ++s) {
      const int src = valid ? ip[4 + s] : -1;
      const int sl = gbase + (src >= 0 ? src : 0);
      T E2[9], p2[3];
#pragma unroll
      for (int k = 0; k < 9; ++k) E2[k] = shfl_t(E[k], sl);
#pragma unroll
      for (int k = 0; k < 3; ++k) p2[k] = shfl_t(p[k], sl);
      if (src >= 0) {
        T En[9], pn[3];
#pragma unroll
        for (int rr = 0; rr < 3; ++rr)
#pragma unroll
          for (int cc = 0; cc < 3; ++cc)
            En[3 * rr + cc] = E[3 * rr] * E2[cc] + E[3 * rr + 1] * E2[3 + cc] + E[3 * rr + 2] * E2[6 + cc];
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) pn[cc] = p2[cc] + E2[cc] * p[0] + E2[3 + cc] * p[1] + E2[6 + cc] * p[2];
#pragma unroll
        for (int k = 0; k < 9; ++k) E[k] = En[k];
#pragma unroll
        for (int k = 0; k < 3; ++k) p[k] = pn[k];
      }
    }
    // world joint axis
    {
      T w[3];
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) w[cc] = E[cc] * mb[36] + E[3 + cc] * mb[37] + E[6 + cc] * mb[38];
      if (kind == 0) {
        S[0] = w[0]; S[1] = w[1]; S[2] = w[2];
        cross3(p, w, S + 3);
      } else {
        S[0] = S[1] = S[2] = T(0);
        S[3] = w[0]; S[4] = w[1]; S[5] = w[2];
      }
      if (!valid || base) {
#pragma unroll
        for (int k = 0; k < 6; ++k) S[k] = T(0);
      }
    }
    // v_i = sum over the root path of S_j qd_j
#pragma unroll
    for (int k = 0; k < 6; ++k) v[k] = S[k] * qdi;
    if (FB && base) {
#pragma unroll
      for (int k = 0; k < 6; ++k) v[k] = myvec[k];
    }
    for (int s = 0; s < nsteps; ++s) {
      const int src = valid ? ip[4 + s] : -1;
      const int sl = gbase + (src >= 0 ? src : 0);
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        const T t = shfl_t(v[k], sl);
        if (src >= 0) v[k] += t;
      }
    }
    T vl[6], al[6];   // parent's velocity / acceleration
#pragma unroll
    for (int k = 0; k < 6; ++k) vl[k] = fma_t(-S[k], qdi, v[k]);
    crm_mul(vl, S, Pd);
#pragma unroll
    for (int k = 0; k < 6; ++k) a[k] = fma_t(Pd[k], qdi, S[k] * qddi);
    if (FB && base) {
#pragma unroll
      for (int k = 0; k < 6; ++k) a[k] = myvec[12 + k];
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) al[k] = a[k];           // own increment, subtracted back below
    for (int s = 0; s < nsteps; ++s) {
      const int src = valid ? ip[4 + s] : -1;
      const int sl = gbase + (src >= 0 ? src : 0);
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        const T t = shfl_t(a[k], sl);
        if (src >= 0) a[k] += t;
      }
    }
    if (!FB) a[5] -= gravity;                             // a_base = [0,0,0,0,0,-GRAVITY] (RBDReference.py:566); FB: inside a_0
#pragma unroll
    for (int k = 0; k < 6; ++k) al[k] = a[k] - al[k];
    {
      T t6[6];
      crm_mul(al, S, Pdd);
      crm_mul(vl, Pd, t6);
#pragma unroll
      for (int k = 0; k < 6; ++k) Pdd[k] += t6[k];
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) { if (!CONLY && !(FB && base)) { myvec[k] = S[k]; myvec[6 + k] = Pd[k]; myvec[12 + k] = Pdd[k]; } }

    // ------------------------------------------------------------------ own terms -> subtree composites
    // 0 m | 1..3 h | 4..9 Ibar | 10..15 Sym | 16..18 n | 19..21 l | 22..27 f
    T acc[28];
    {
      const T mi = valid ? mb[39] : T(0);
      T hr[3], hw[3];
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) {
        hr[cc] = E[cc] * mb[40] + E[3 + cc] * mb[41] + E[6 + cc] * mb[42];
        hw[cc] = fma_t(mi, p[cc], hr[cc]);
      }
      T IbE[9];
      {
        const T xx = mb[43], xy = mb[44], xz = mb[45], yy = mb[46], yz = mb[47], zz = mb[48];
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
          IbE[cc] = xx * E[cc] + xy * E[3 + cc] + xz * E[6 + cc];
          IbE[3 + cc] = xy * E[cc] + yy * E[3 + cc] + yz * E[6 + cc];
          IbE[6 + cc] = xz * E[cc] + yz * E[3 + cc] + zz * E[6 + cc];
        }
      }
      T Iw[6];
      {
        const T tr = (hr[0] + hw[0]) * p[0] + (hr[1] + hw[1]) * p[1] + (hr[2] + hw[2]) * p[2];
        int idx = 0;
#pragma unroll
        for (int rr = 0; rr < 3; ++rr)
#pragma unroll
          for (int cc = rr; cc < 3; ++cc) {
            T val = E[rr] * IbE[cc] + E[3 + rr] * IbE[3 + cc] + E[6 + rr] * IbE[6 + cc];
            val -= hr[rr] * p[cc] + p[rr] * hw[cc];
            if (rr == cc) val += tr;
            Iw[idx++] = val;
          }
      }
      T mom[6], fo[6], t6[6];
      rigid_mul(mi, hw, Iw, v, mom);
      rigid_mul(mi, hw, Iw, a, fo);
      crf_mul(v, mom, t6);
      const T* wv = v;
      const T* uv = v + 3;
      T M[9];
      {
        const T Im[9] = {Iw[0], Iw[1], Iw[2], Iw[1], Iw[3], Iw[4], Iw[2], Iw[4], Iw[5]};
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
          M[cc] = wv[1] * Im[6 + cc] - wv[2] * Im[3 + cc];
          M[3 + cc] = wv[2] * Im[cc] - wv[0] * Im[6 + cc];
          M[6 + cc] = wv[0] * Im[3 + cc] - wv[1] * Im[cc];
        }
      }
      const T uh2 = T(2) * (uv[0] * hw[0] + uv[1] * hw[1] + uv[2] * hw[2]);
      acc[0] = mi;
      acc[1] = hw[0]; acc[2] = hw[1]; acc[3] = hw[2];
#pragma unroll
      for (int k = 0; k < 6; ++k) acc[4 + k] = Iw[k];
      acc[10] = T(2) * M[0] - T(2) * hw[0] * uv[0] + uh2;
      acc[11] = M[1] + M[3] - (hw[0] * uv[1] + uv[0] * hw[1]);
      acc[12] = M[2] + M[6] - (hw[0] * uv[2] + uv[0] * hw[2]);
      acc[13] = T(2) * M[4] - T(2) * hw[1] * uv[1] + uh2;
      acc[14] = M[5] + M[7] - (hw[1] * uv[2] + uv[1] * hw[2]);
      acc[15] = T(2) * M[8] - T(2) * hw[2] * uv[2] + uh2;
#pragma unroll
      for (int k = 0; k < 6; ++k) { acc[16 + k] = mom[k]; acc[22 + k] = fo[k] + t6[k]; }
      if (!valid) {
#pragma unroll
        for (int k = 0; k < 28; ++k) acc[k] = T(0);
      }
    }
    // subtree composites = suffix sums over the contiguous preorder range of the subtree, restarted
    // at every root component.  Transposed through shared memory: every lane parks its 28 own
    // terms, lane c < 28 then runs the sequential suffix sum of component c over the bodies of each
    // knot point of the warp, and every body reads back PS(i) - PS(subtree_end(i)).
    {
      constexpr int K0 = CONLY ? 22 : 0;                      // rnea only needs the force composite (22..27)
      if (!CONLY) warp_bulk_store_wait(lane);                 // the previous slab has left the tile
      T* cs = tile;                                           // [32][29], the tile is not live yet
      T* mine = cs + lane * kCoopScanStride;
#pragma unroll
      for (int k = K0; k < 28; ++k) mine[k] = acc[k];
      __syncwarp();
      if (lane >= K0 && lane < 28) {
#pragma unroll
        for (int gg = 0; gg < IPW; ++gg) {
          T* col = cs + (gg * G) * kCoopScanStride + lane;
          for (int cidx = 0; cidx < cp.ncomp; ++cidx) {
            T run = T(0);
            const int first = cp.comp_begin[cidx];
            for (int bdy = cp.comp_begin[cidx + 1] - 1; bdy >= first; --bdy) {
              run += col[bdy * kCoopScanStride];
              col[bdy * kCoopScanStride] = run;
            }
          }
        }
      }
      __syncwarp();
#pragma unroll
      for (int k = K0; k < 28; ++k) acc[k] = mine[k];
      if (cut) {
        const T* beyond = cs + (gbase + sub_end) * kCoopScanStride;
#pragma unroll
        for (int k = K0; k < 28; ++k) acc[k] -= beyond[k];
      }
      __syncwarp();                                           // the buffer becomes the output tile again
    }
    if (CONLY) {
      if (valid && (grp * IPW + g) < B) c_out[b * n + oi] = dot6s(S, acc + 22);     // :613
      continue;
    }

    // ------------------------------------------------------------------ F vectors, diagonal, zero fill
    const T mC = acc[0];
    const T* hC = acc + 1;
    const T* IC = acc + 4;
    const T* SyC = acc + 10;
    const T* nC = acc + 16;
    const T* lC = acc + 19;
    const T* fC = acc + 22;
    T F1[6], F2[6], F3[3], F4[6];
    rigid_mul(mC, hC, IC, S, F4);
    {
      T t3[3];
      sym3_mul(SyC, S, F3);
      cross3_add(nC, S, F3);
      cross3(lC, S + 3, t3);
#pragma unroll
      for (int k = 0; k < 3; ++k) F3[k] = fma_t(T(0.5), F3[k], t3[k]);
    }
    {
      T t6[6], tb[3], tl[3];
      rigid_mul(mC, hC, IC, Pdd, F1);
      crf_mul(S, fC, t6);
      sym3_mul(SyC, Pd, tb);
      cross3(nC, Pd, tl);
#pragma unroll
      for (int k = 0; k < 3; ++k) F1[k] += t6[k] + tb[k] - tl[k];
      cross3(lC, Pd, tl);
#pragma unroll
      for (int k = 0; k < 3; ++k) F1[3 + k] += t6[3 + k] - T(2) * tl[k];
      rigid_mul(mC, hC, IC, Pd, F2);
      sym3_mul(SyC, S, tb);
      cross3(nC, S, tl);
#pragma unroll
      for (int k = 0; k < 3; ++k) F2[k] = T(2) * F2[k] + tb[k] - tl[k];
      cross3(lC, S, tl);
#pragma unroll
      for (int k = 0; k < 3; ++k) F2[3 + k] = T(2) * (F2[3 + k] - tl[k]);
    }
    const bool store = valid && (grp * IPW + g) < B;
    if (c_out && store) {
      if (FB && base) {
#pragma unroll
        for (int k = 0; k < 6; ++k) c_out[b * nv + k] = fC[k];                      // :612 with S = eye(6)
      } else {
        c_out[b * nv + voff] = dot6s(S, fC);
      }
    }
    T F1q[6];                                                 // F1 with the prismatic quirk of :1292 folded in
#pragma unroll
    for (int k = 0; k < 6; ++k) F1q[k] = F1[k];
    if (kind == 1) {
      // reference quirk for prismatic joints: X^T(-crm(f)S) instead of X^T(S x* f) (:1292)
      T nrot[3], dl[3], da[3], t3[3];
      cross3(p, fC + 3, t3);
#pragma unroll
      for (int k = 0; k < 3; ++k) nrot[k] = fC[k] - t3[k];
      cross3(S + 3, nrot, dl);
      cross3(S + 3, fC + 3, da);
      cross3(p, dl, t3);
#pragma unroll
      for (int k = 0; k < 3; ++k) { F1q[k] += t3[k] - da[k]; F1q[3 + k] += dl[k]; }
    }
    // The tile holds dc_du of the warp's knot points (SPLIT = false) or one half of it at a time
    // (SPLIT = true: dc_dq, then dc_dqd - half the shared memory, for large robots).
    constexpr int NPASS = SPLIT ? 2 : 1;
    const int tw = SPLIT ? nv : n2;                           // tile row width
    const int tvals = IPW * nv * tw;
    T* mytile2 = tile + g * nv * tw;
#pragma unroll
    for (int half = 0; half < NPASS; ++half) {
      const bool do_q = !SPLIT || half == 0, do_qd = !SPLIT || half == 1;
      const int qd_off = SPLIT ? 0 : nv;                      // column offset of the dc_dqd block inside the tile
      {
        typedef typename Vec2<T>::type V2;
        V2 z; z.x = T(0); z.y = T(0);
        for (int k = lane; k < ((tvals + 1) >> 1); k += 32) reinterpret_cast<V2*>(tile)[k] = z;   // structural zeros
      }
      __syncwarp();
      if (valid && !base) {
        if (do_q) mytile2[voff * tw + voff] = dot6s(S, F1);
        if (do_qd) {
          T ddd = dot6s(S, F2);
          if (!FB && use_damping) ddd += mb[49];              // RBDReference.py:1341
          mytile2[voff * tw + qd_off + voff] = ddd;
        }
      }
      // ---------------------------------------------------------------- ancestors of i
      {
        int j = par;
        for (int t = 0; t < cp.maxdepth; ++t) {
          if (j >= (FB ? 1 : 0)) {                            // FB: the base (body 0) is handled below
            const T* vj = vec + (gbase + j) * kCoopVecStride;
            T Sj[6], Pdj[6];
#pragma unroll
            for (int k = 0; k < 6; ++k) { Sj[k] = vj[k]; Pdj[k] = vj[6 + k]; }
            const int oj = imdl[j * kCoopIntStride + 3] + (FB ? 5 : 0);
            if (do_q) {
              T Pddj[6];
#pragma unroll
              for (int k = 0; k < 6; ++k) Pddj[k] = vj[12 + k];
              mytile2[oj * tw + voff] = dot6s(Sj, F1q);
              mytile2[voff * tw + oj] = fma_t(T(2), dot3s(F3, Pdj), dot6s(F4, Pddj));
            }
            if (do_qd) {
              mytile2[oj * tw + qd_off + voff] = dot6s(Sj, F2);
              mytile2[voff * tw + qd_off + oj] = T(2) * (dot6s(F4, Pdj) + dot3s(F3, Sj));
            }
            j = imdl[j * kCoopIntStride];
          }
        }
      }
      if (FB) {
        // ---------------------------------------------------------------- the base's six rows and columns
        const T* vb = vec + gbase * kCoopVecStride;           // the base's row of the table: v_0 | X_0 a_grav
        T v0[6], xg[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) { v0[k] = vb[k]; xg[k] = vb[6 + k]; }
        if (valid && !base) {
          // rows 0..5: S_0 = eye(6), so the entries are the components of F1 / F2 (:1282, :1325); columns 0..5:
          // F4 . (a x e_k) = -(a x* F4)[k] with a = X_0 a_grav (dq: Psi_ddot_0k) or v_0 (dqd: Psi_dot_0k + S_dot_0k)
          T t6[6];
          if (do_q) {
            crf_mul(xg, F4, t6);
#pragma unroll
            for (int k = 0; k < 6; ++k) { mytile2[k * tw + voff] = F1q[k]; mytile2[voff * tw + k] = -t6[k]; }
          }
          if (do_qd) {
            crf_mul(v0, F4, t6);
#pragma unroll
            for (int k = 0; k < 6; ++k) {
              mytile2[k * tw + qd_off + voff] = F2[k];
              mytile2[voff * tw + qd_off + k] = k < 3 ? fma_t(T(2), F3[k], -t6[k]) : -t6[k];
            }
          }
        }
        // base x base block: lane k < 6 of the group owns column k and works on the base's composite
        T cb[22];
#pragma unroll
        for (int k = 0; k < 22; ++k) cb[k] = shfl_t(acc[k], gbase);
        if (i < 6) {
          T ek[6], x6[6], y6[6];
#pragma unroll
          for (int k = 0; k < 6; ++k) ek[k] = k == i ? T(1) : T(0);
          if (do_q) {
            crm_mul(xg, ek, x6);
            rigid_mul(cb[0], cb + 1, cb + 4, x6, y6);          // I^C_0 ((X_0 a_grav) x e_k)
#pragma unroll
            for (int r = 0; r < 6; ++r) mytile2[r * tw + i] = y6[r];
          }
          if (do_qd) {
            T tb[3], tl[3], tl2[3];
            crm_mul(v0, ek, x6);
            rigid_mul(cb[0], cb + 1, cb + 4, x6, y6);          // I^C_0 (v_0 x e_k) + 2 B^C_0 e_k
            sym3_mul(cb + 10, ek, tb);
            cross3(cb + 16, ek, tl);
            cross3(cb + 19, ek, tl2);
#pragma unroll
            for (int k = 0; k < 3; ++k) { y6[k] += tb[k] - tl[k]; y6[3 + k] -= T(2) * tl2[k]; }
#pragma unroll
            for (int r = 0; r < 6; ++r) mytile2[r * tw + qd_off + i] = y6[r];
          }
        }
        if (use_damping && do_qd) {
          // :1336-1341 to the letter: the base's damping on the whole block [0:5, 0:5], body i's on [i, i] (body
          // index, not i + 5)
          __syncwarp();
          for (int e = i; e < 25; e += G) mytile2[(e / 5) * tw + qd_off + (e % 5)] += mdl[49];
          __syncwarp();
          if (valid && !base) mytile2[oi * tw + qd_off + oi] += mb[49];
        }
      }
      __syncwarp();
      // ---------------------------------------------------------------- coalesced write
      {
        const int64_t first = grp * IPW;
        const int nk = (int)((B - first) < IPW ? (B - first) : IPW);
        T* dst = dc_du + first * nv * n2;
        if (!SPLIT && warp_bulk_store(dst, tile, nk * nv * n2, lane)) {
          // one cp.async.bulk for the warp's slab (awaited before the tile is written again)
        } else if (!SPLIT) {
          typedef typename Vec2<T>::type V2;
          const int count = nk * nv * n2;
          if (((IPW * nv * n2) & 1) == 0 && nk == IPW && (reinterpret_cast<uintptr_t>(dc_du) & (sizeof(V2) - 1)) == 0) {
            for (int k = lane; k < (count >> 1); k += 32) __stcs(reinterpret_cast<V2*>(dst) + k, reinterpret_cast<const V2*>(tile)[k]);
          } else {
            for (int k = lane; k < count; k += 32) __stcs(dst + k, tile[k]);
          }
        } else {
          // rows of n values go to columns [half*n, half*n + n) of the (n, 2n) result
          const int count = nk * nv * nv;
          for (int k = lane; k < count; k += 32) {
            const int r = k / nv, c = k - r * nv;             // r runs over (knot, row)
            __stcs(dst + (int64_t)r * n2 + half * nv + c, tile[k]);
          }
        }
      }
      __syncwarp();
    }
  }
  if (!CONLY) warp_bulk_store_wait(lane);                     // shared memory must outlive the copies
}

}  // namespace rbd
