// rbd_coop_minv_kernels.cuh - warp-cooperative fused minv (RBDReference.py:785-806).
//
// Same recursion as minv_bpass (:630-735) + minv_fpass (:737-783) + mirror (:799-804), mapped so
// that one knot point's whole working set lives in one warp's registers and shared memory:
// a group of G = 8 / 16 / 32 lanes owns a knot point; lane i is BODY i for the articulated-inertia
// recursion and COLUMN i of Minv for the two column sweeps (bodies in depth-first preorder).
//
// Frames: every body's quantities are expressed in WORLD-ALIGNED axes about the body's OWN origin
// ("local world aligned").  A parent<-child transfer is then a pure translation by
// r_i = p_i - p_parent (no rotation, no 6x6 congruence with a dense X), the joint axis is
// S_i = [w_i; 0] (revolute) or [0; w_i] (prismatic), and - unlike coordinates about the world
// origin - no m|p|^2 terms appear, so single precision keeps its digits on light distal links.
//
//   phase 0  rotation scan: E_i = prod of joint rotations along the root path (pointer jumping,
//            ceil(log2(depth+1)) shuffle rounds); r_i and w_i follow from E_i alone
//   phase A  articulated inertia, level by level from the deepest bodies to the roots
//            (all bodies of one depth in parallel):  U = IA S, D = S.U,
//            IA_parent += T(r)^T (IA - U U^T / D) T(r)                                  (:694-733)
//   phase B  lane j walks from body j to its root:  Minv[i,j] = (delta_ij - S_i.F_j) / D_i,
//            F_j += U_i Minv[i,j], shifted to the parent's origin                       (:700-726)
//   phase C  lane j visits the bodies i <= j of its root component in preorder:
//            Minv[i,j] -= (U_i . G_parent) / D_i,  G_i = G_parent + S_i Minv[i,j]       (:771-781)
//            and writes Minv[i,j] = Minv[j,i] into the warp's output tile (mirror :799-804)
//   the tile (all knot points of the warp, contiguous in HBM) is stored with coalesced writes.
#pragma once
#include "rbd_common.cuh"
#include "rbd_grad_kernels.cuh"
#include "rbd_minv_kernels.cuh"
#include "rbd_coop_kernels.cuh"

namespace rbd {

constexpr int kCmTabStride = 14;     // w(3) invD | U(6) | r(3) pad  (seven aligned pairs)
constexpr int kCmIaStride = 22;      // articulated inertia handed to the parent (21, eleven pairs)
constexpr int kCmMaxWarps = 8;       // warps per CTA is chosen at launch (blockDim.x / 32)

struct CoopMinvPlan {
  int maxcomp;                       // largest root component
  int depth[RBD_MAX_DOF];
  int comp_root[RBD_MAX_DOF];
};

// per-warp shared memory, in values of T:  tab | mb | big, where `big` holds the children's
// inertias in phase A and is re-used for the G stashes and the output tile in phase C
__host__ __device__ inline int coop_minv_warp_vals(int n, int G, int maxdepth, int nslot) {
  const int ipw = 32 / G;
  const int a = 32 * kCmIaStride;
  const int c = 6 * nslot * 32 + ((ipw * n * n + 1) & ~1);
  return 32 * kCmTabStride + (maxdepth + 1) * 32 + (a > c ? a : c);
}
__host__ __device__ inline size_t coop_minv_smem_bytes(int n, int G, int maxdepth, int nslot, int warps, size_t tsize) {
  return (size_t)(((n * kCoopMdlStride + 1) & ~1) + warps * coop_minv_warp_vals(n, G, maxdepth, nslot)) * tsize +
         (size_t)n * 8 * sizeof(int);
}

template <typename T, int G>
__global__ void __launch_bounds__(kCmMaxWarps * 32)
minv_coop_kernel(const __grid_constant__ FastModel<T> m, const __grid_constant__ DfsPlan plan,
                 const __grid_constant__ CoopPlan cp, const __grid_constant__ CoopMinvPlan mp, int64_t B,
                 const T* __restrict__ q, T* __restrict__ Minv) {
  constexpr int IPW = 32 / G;
  typedef typename Vec2<T>::type V2;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = m.n;
  const int nn = n * n;
  const int maxdepth = cp.maxdepth;
  const int nwarps = blockDim.x >> 5;
  const int warp_vals = coop_minv_warp_vals(n, G, maxdepth, m.n_slot_a);
  int4* imdl = reinterpret_cast<int4*>(smem_raw);                            // [n][2]
  T* mdl = reinterpret_cast<T*>(smem_raw + (size_t)n * 2 * sizeof(int4));    // [n][51]
  T* warp_all = mdl + ((n * kCoopMdlStride + 1) & ~1);                       // [nwarps][warp_vals]

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
    mdl[idx] = val;
  }
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int p = m.parent[i];
    imdl[2 * i] = make_int4(p, m.kind[i], plan.sub_end[i], plan.orig[i]);
    // slot of this body's G (kept for a later, non-first child) and of the parent's
    imdl[2 * i + 1] = make_int4(mp.depth[i], mp.comp_root[i], m.slot_a[i], p >= 0 ? m.slot_a[p] : -1);
  }
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane / G, i = lane - g * G;
  const bool valid = i < n;
  const int ib = valid ? i : 0;
  const int gbase = g * G;
  const T* mb = mdl + ib * kCoopMdlStride;
  const int4 myA = imdl[2 * ib], myB = imdl[2 * ib + 1];
  const int par = valid ? myA.x : -1;
  const int kind = myA.y;
  const int sub_end = myA.z;
  const int oi = myA.w;
  const int depth = valid ? myB.x : -1;
  const int comp_root = myB.y;
  int jump[5];
#pragma unroll
  for (int s = 0; s < 5; ++s) jump[s] = valid ? cp.jump[s][ib] : -1;
  T* tab = warp_all + warp * warp_vals;                   // [32][14]
  T* mbw = tab + 32 * kCmTabStride;                       // [maxdepth+1][32]
  T* big = mbw + (maxdepth + 1) * 32;                     // phase A: [32][22]; phase C: G stashes [slot][6][32] | tile
  T* tile = big + 6 * m.n_slot_a * 32;                    // [IPW][n][n]
  T* mytile = tile + g * nn;
  T* mytab = tab + lane * kCmTabStride;
  const int nsteps = cp.nsteps;
  const int tile_vals = IPW * nn;
  const bool pair_ok = (tile_vals & 1) == 0;              // slab is a whole number of aligned pairs

  const int64_t ngroups = (B + IPW - 1) / IPW;
  for (int64_t grp = (int64_t)blockIdx.x * nwarps + warp; grp < ngroups; grp += (int64_t)gridDim.x * nwarps) {
    int64_t b = grp * IPW + g;
    if (b >= B) b = B - 1;                                // duplicate work, never stored

    // ------------------------------------------------------------------ phase 0: rotations
    T E[9], rw[3], w[3];
    {
      const T qi = q[b * n + oi];
      T f1, f2;
      if (kind == 0) sincos_t(qi, &f2, &f1);
      else { f1 = qi; f2 = T(0); }
      T r[3];
#pragma unroll
      for (int k = 0; k < 9; ++k) E[k] = fma_t(mb[18 + k], f2, fma_t(mb[9 + k], f1, mb[k]));
#pragma unroll
      for (int k = 0; k < 3; ++k) r[k] = fma_t(mb[33 + k], f2, fma_t(mb[30 + k], f1, mb[27 + k]));
      // t = E_J r : the joint offset seen from body i; rotated to world axes after the scan
#pragma unroll
      for (int k = 0; k < 3; ++k) rw[k] = E[3 * k] * r[0] + E[3 * k + 1] * r[1] + E[3 * k + 2] * r[2];
    }
#pragma unroll
    for (int s = 0; s < 5; ++s) {
      if (s < nsteps) {
        const int src = jump[s];
        const int sl = gbase + (src >= 0 ? src : 0);
        T E2[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) E2[k] = shfl_t(E[k], sl);
        if (src >= 0) {
          T En[9];
#pragma unroll
          for (int rr = 0; rr < 3; ++rr)
#pragma unroll
            for (int cc = 0; cc < 3; ++cc)
              En[3 * rr + cc] = E[3 * rr] * E2[cc] + E[3 * rr + 1] * E2[3 + cc] + E[3 * rr + 2] * E2[6 + cc];
#pragma unroll
          for (int k = 0; k < 9; ++k) E[k] = En[k];
        }
      }
    }
    {
      T t[3] = {rw[0], rw[1], rw[2]};
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) {
        rw[cc] = E[cc] * t[0] + E[3 + cc] * t[1] + E[6 + cc] * t[2];           // r_i = p_i - p_parent, world axes
        w[cc] = E[cc] * mb[36] + E[3 + cc] * mb[37] + E[6 + cc] * mb[38];       // joint axis, world axes
      }
    }

    // ------------------------------------------------------------------ own rigid inertia about p_i
    // IA = [[A, Bm], [Bm^T, C]] : A sym (0..5: xx xy xz yy yz zz), Bm 3x3 row-major (6..14), C sym (15..20)
    T IA[22];
    {
      const T mi = mb[39];
      T hr[3];
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) hr[cc] = E[cc] * mb[40] + E[3 + cc] * mb[41] + E[6 + cc] * mb[42];
      T IbE[9];
      const T xx = mb[43], xy = mb[44], xz = mb[45], yy = mb[46], yz = mb[47], zz = mb[48];
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) {
        IbE[cc] = xx * E[cc] + xy * E[3 + cc] + xz * E[6 + cc];
        IbE[3 + cc] = xy * E[cc] + yy * E[3 + cc] + yz * E[6 + cc];
        IbE[6 + cc] = xz * E[cc] + yz * E[3 + cc] + zz * E[6 + cc];
      }
      int idx = 0;
#pragma unroll
      for (int rr = 0; rr < 3; ++rr)
#pragma unroll
        for (int cc = rr; cc < 3; ++cc)
          IA[idx++] = E[rr] * IbE[cc] + E[3 + rr] * IbE[3 + cc] + E[6 + rr] * IbE[6 + cc];
      IA[6] = T(0); IA[7] = -hr[2]; IA[8] = hr[1];
      IA[9] = hr[2]; IA[10] = T(0); IA[11] = -hr[0];
      IA[12] = -hr[1]; IA[13] = hr[0]; IA[14] = T(0);
      IA[15] = mi; IA[16] = T(0); IA[17] = T(0); IA[18] = mi; IA[19] = T(0); IA[20] = mi;
      IA[21] = T(0);
    }

    // ------------------------------------------------------------------ phase A: articulated inertias
    for (int d = maxdepth; d >= 0; --d) {
      if (depth == d) {
        for (int c = i + 1; c < sub_end; c = imdl[2 * c].z) {                   // children of i
          const V2* src = reinterpret_cast<const V2*>(big + (gbase + c) * kCmIaStride);
#pragma unroll
          for (int k = 0; k < 11; ++k) { const V2 t = src[k]; IA[2 * k] += t.x; IA[2 * k + 1] += t.y; }
        }
        T U[6];
        if (kind == 0) {
          sym3_mul(IA, w, U);                                                  // A w
#pragma unroll
          for (int cc = 0; cc < 3; ++cc) U[3 + cc] = IA[6 + cc] * w[0] + IA[9 + cc] * w[1] + IA[12 + cc] * w[2];   // Bm^T w
        } else {
#pragma unroll
          for (int rr = 0; rr < 3; ++rr) U[rr] = IA[6 + 3 * rr] * w[0] + IA[7 + 3 * rr] * w[1] + IA[8 + 3 * rr] * w[2];   // Bm w
          sym3_mul(IA + 15, w, U + 3);                                         // C w
        }
        const T D = kind == 0 ? dot3s(w, U) : dot3s(w, U + 3);
        const T invD = T(1) / D;                                               // RBDReference.py:698-700
        {
          V2* dst = reinterpret_cast<V2*>(mytab);
          V2 t;
          t.x = w[0]; t.y = w[1]; dst[0] = t;
          t.x = w[2]; t.y = invD; dst[1] = t;
          t.x = U[0]; t.y = U[1]; dst[2] = t;
          t.x = U[2]; t.y = U[3]; dst[3] = t;
          t.x = U[4]; t.y = U[5]; dst[4] = t;
          t.x = rw[0]; t.y = rw[1]; dst[5] = t;
          t.x = rw[2]; t.y = T(0); dst[6] = t;
        }
        if (par >= 0) {
          // IA -= U U^T / D                                                   (:728-731)
          T Us[6];
#pragma unroll
          for (int k = 0; k < 6; ++k) Us[k] = U[k] * invD;
          IA[0] -= U[0] * Us[0]; IA[1] -= U[0] * Us[1]; IA[2] -= U[0] * Us[2];
          IA[3] -= U[1] * Us[1]; IA[4] -= U[1] * Us[2]; IA[5] -= U[2] * Us[2];
#pragma unroll
          for (int rr = 0; rr < 3; ++rr)
#pragma unroll
            for (int cc = 0; cc < 3; ++cc) IA[6 + 3 * rr + cc] -= U[rr] * Us[3 + cc];
          IA[15] -= U[3] * Us[3]; IA[16] -= U[3] * Us[4]; IA[17] -= U[3] * Us[5];
          IA[18] -= U[4] * Us[4]; IA[19] -= U[4] * Us[5]; IA[20] -= U[5] * Us[5];
          // translate to the parent's origin (:732-733 with X = [[1,0],[-r x,1]]):
          //   Bm' = Bm + R C,  A' = A + R W^T + W R^T,  W = Bm + R C / 2,  R = r x
          const T Cm[9] = {IA[15], IA[16], IA[17], IA[16], IA[18], IA[19], IA[17], IA[19], IA[20]};
          T RC[9], W[9];
#pragma unroll
          for (int cc = 0; cc < 3; ++cc) {
            RC[cc] = rw[1] * Cm[6 + cc] - rw[2] * Cm[3 + cc];
            RC[3 + cc] = rw[2] * Cm[cc] - rw[0] * Cm[6 + cc];
            RC[6 + cc] = rw[0] * Cm[3 + cc] - rw[1] * Cm[cc];
          }
#pragma unroll
          for (int k = 0; k < 9; ++k) { W[k] = fma_t(T(0.5), RC[k], IA[6 + k]); IA[6 + k] += RC[k]; }
          // (R W^T)[a][b] = (r x W[b,:])[a]
          T RW[9];
#pragma unroll
          for (int bb = 0; bb < 3; ++bb) {
            RW[bb] = rw[1] * W[3 * bb + 2] - rw[2] * W[3 * bb + 1];            // a = 0
            RW[3 + bb] = rw[2] * W[3 * bb] - rw[0] * W[3 * bb + 2];            // a = 1
            RW[6 + bb] = rw[0] * W[3 * bb + 1] - rw[1] * W[3 * bb];            // a = 2
          }
          IA[0] += T(2) * RW[0];
          IA[1] += RW[1] + RW[3];
          IA[2] += RW[2] + RW[6];
          IA[3] += T(2) * RW[4];
          IA[4] += RW[5] + RW[7];
          IA[5] += T(2) * RW[8];
          V2* dst = reinterpret_cast<V2*>(big + lane * kCmIaStride);
#pragma unroll
          for (int k = 0; k < 11; ++k) { V2 t; t.x = IA[2 * k]; t.y = IA[2 * k + 1]; dst[k] = t; }
        }
      }
      __syncwarp();
    }

    // ------------------------------------------------------------------ phase B: walk to the root
    {
      T F[6] = {T(0), T(0), T(0), T(0), T(0), T(0)};
      int a = valid ? i : -1;
      for (int t = 0; t <= maxdepth; ++t) {
        if (a >= 0) {
          const V2* ta = reinterpret_cast<const V2*>(tab + (gbase + a) * kCmTabStride);
          const int4 ia = imdl[2 * a];
          const int da = imdl[2 * a + 1].x;
          const V2 t0 = ta[0], t1 = ta[1];
          const T wa[3] = {t0.x, t0.y, t1.x};
          const T invD = t1.y;
          const T sF = ia.y == 0 ? dot3s(wa, F) : dot3s(wa, F + 3);
          const T mij = (a == i ? invD : T(0)) - invD * sF;                    // :700-708
          mbw[da * 32 + lane] = mij;
          if (ia.x >= 0) {
            const V2 t2 = ta[2], t3 = ta[3], t4 = ta[4], t5 = ta[5], t6 = ta[6];
            F[0] = fma_t(t2.x, mij, F[0]); F[1] = fma_t(t2.y, mij, F[1]); F[2] = fma_t(t3.x, mij, F[2]);   // :721-726
            F[3] = fma_t(t3.y, mij, F[3]); F[4] = fma_t(t4.x, mij, F[4]); F[5] = fma_t(t4.y, mij, F[5]);
            const T ra[3] = {t5.x, t5.y, t6.x};
            cross3_add(ra, F + 3, F);                                          // moment about the parent's origin
          }
          a = ia.x;
        }
      }
    }
    __syncwarp();
    // zero the output tile (entries between different root components stay zero)
    if (pair_ok) {
      V2 z; z.x = T(0); z.y = T(0);
      for (int k = lane; k < (tile_vals >> 1); k += 32) reinterpret_cast<V2*>(tile)[k] = z;
    } else {
      for (int k = lane; k < tile_vals; k += 32) tile[k] = T(0);
    }
    __syncwarp();

    // ------------------------------------------------------------------ phase C: preorder sweep
    {
      T Gv[6] = {T(0), T(0), T(0), T(0), T(0), T(0)};
      const int oin = oi * n;
      for (int t = 0; t < mp.maxcomp; ++t) {
        const int a = comp_root + t;
        if (valid && a <= i) {
          const V2* ta = reinterpret_cast<const V2*>(tab + (gbase + a) * kCmTabStride);
          const int4 ia = imdl[2 * a], ja = imdl[2 * a + 1];
          const V2 t0 = ta[0], t1 = ta[1];
          T mij = (i < ia.z) ? mbw[ja.x * 32 + lane] : T(0);
          if (ia.x < 0) {
#pragma unroll
            for (int k = 0; k < 6; ++k) Gv[k] = T(0);
          } else {
            if (ia.x != a - 1) {
              const T* gs = big + (ja.w * 6) * 32 + lane;
#pragma unroll
              for (int k = 0; k < 6; ++k) Gv[k] = gs[k * 32];
            }
            const V2 t2 = ta[2], t3 = ta[3], t4 = ta[4], t5 = ta[5], t6 = ta[6];
            const T ra[3] = {t5.x, t5.y, t6.x};
            cross3_add(Gv, ra, Gv + 3);                                        // velocity at p_a: v += w x r
            const T lo = fma_t(t3.x, Gv[2], fma_t(t2.y, Gv[1], t2.x * Gv[0]));
            const T hi = fma_t(t4.y, Gv[5], fma_t(t4.x, Gv[4], t3.y * Gv[3]));
            mij = fma_t(-t1.y, lo + hi, mij);                                  // :771-773
          }
          if (ia.y == 0) {
            Gv[0] = fma_t(t0.x, mij, Gv[0]); Gv[1] = fma_t(t0.y, mij, Gv[1]); Gv[2] = fma_t(t1.x, mij, Gv[2]);   // :774-781
          } else {
            Gv[3] = fma_t(t0.x, mij, Gv[3]); Gv[4] = fma_t(t0.y, mij, Gv[4]); Gv[5] = fma_t(t1.x, mij, Gv[5]);
          }
          if (ja.z >= 0) {
            T* gs = big + (ja.z * 6) * 32 + lane;
#pragma unroll
            for (int k = 0; k < 6; ++k) gs[k * 32] = Gv[k];
          }
          mytile[ia.w * n + oi] = mij;
          mytile[oin + ia.w] = mij;                                            // :799-804
        }
      }
    }
    __syncwarp();
    // ------------------------------------------------------------------ coalesced slab write
    {
      const int64_t first = grp * IPW;
      const int count = (int)((B - first) < IPW ? (B - first) : IPW) * nn;
      T* dst = Minv + first * nn;
      if (pair_ok && count == tile_vals) {
        for (int k = lane; k < (tile_vals >> 1); k += 32) reinterpret_cast<V2*>(dst)[k] = reinterpret_cast<const V2*>(tile)[k];
      } else {
        for (int k = lane; k < count; k += 32) dst[k] = tile[k];
      }
    }
    __syncwarp();
  }
}

}  // namespace rbd
