// rbd_lane_minv_kernels.cuh - fused minv (RBDReference.py:785-806) for robots whose whole
// per-warp working set fits in shared memory (n <= ~12): ONE KNOT POINT PER LANE in every phase.
//
// Same recursion and the same local world-aligned frames as rbd_coop_minv_kernels.cuh (every
// body's quantities in world-aligned axes about the body's own origin, so a parent<-child
// transfer is a pure translation by r_i = p_i - p_parent):
//
//   stage 1  rotation sweep root -> leaf, articulated-inertia sweep leaf -> root (:694-733);
//            each body's (w, 1/D, U, r) goes to a [body][field][lane] table in shared memory
//            (lane-contiguous: conflict free), registers hold the running IA and rotation.
//   stage 2  columns of Minv in groups of GC consecutive (depth-first) indices; the group's
//            F_j (phase B, :700-726) / G_j (phase C, :771-781) six-vectors stay in registers while
//            the bodies are swept, so one table row read serves up to GC (body, column) pairs.
//            Rows of the result are assembled (with the mirror, :799-804) in a per-warp tile that
//            has the exact layout of the warp's contiguous slab of Minv in HBM and is written
//            with coalesced streaming stores.
//
// Compared with the column-per-lane mapping no lane idles in the triangular sweeps and no table
// row is fetched more than ceil(n / GC) times per phase; the price is (13 n + n^2) values of shared
// memory per knot point, which bounds the robots this kernel serves.
#pragma once
#include "rbd_common.cuh"
#include "rbd_grad_kernels.cuh"
#include "rbd_minv_kernels.cuh"
#include "rbd_coop_minv_kernels.cuh"

namespace rbd {

constexpr int kLmTab = 13;           // w(3) invD U(6) r(3)
constexpr int kLmMaxWarps = 8;

__host__ __device__ inline int lane_minv_tile_stride(int n) { return (n * n) | 1; }   // odd: conflict-free lanes
template <int GC>
__host__ __device__ inline int lane_minv_warp_vals(int n, int nslot_a, int nslot_b) {
  const int s2 = 32 * lane_minv_tile_stride(n) + nslot_a * GC * 6 * 32;       // tile | G stashes
  const int s1 = (2 * n + 22 * nslot_a + 9 * nslot_b) * 32;                   // f1 f2 | stage-1 stashes
  return n * kLmTab * 32 + (s1 > s2 ? s1 : s2);
}
template <typename T, int GC>
__host__ __device__ inline size_t lane_minv_smem_bytes(int n, int nslot_a, int nslot_b, int warps) {
  return (size_t)warps * lane_minv_warp_vals<GC>(n, nslot_a, nslot_b) * sizeof(T);
}

// NMAX >= n > 0: the stage-1 loops over bodies are fully unrolled (guarded by the runtime n), so
// model constants become immediates; NMAX = 0: rolled loops.  NMAX2: the same for the stage-2 loops
// over column groups and bodies (measured on B200, iiwa14: unrolling stage 2 as well overflows the
// instruction cache - 1.18e9 instead of 1.34e9 evals/s FP64).
template <typename T, int GC, bool PRISM, int NMAX, int NMAX2>
__global__ void __launch_bounds__(kLmMaxWarps * 32)
minv_lane_kernel(const __grid_constant__ FastModel<T> m, const __grid_constant__ DfsPlan plan,
                 const __grid_constant__ CoopMinvPlan mp, int64_t B, const T* __restrict__ q, T* __restrict__ Minv) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = m.n;
  const int nn = n * n;
  const int nwarps = blockDim.x >> 5;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int warp_vals = lane_minv_warp_vals<GC>(n, m.n_slot_a, m.n_slot_b);
  const int tstride = lane_minv_tile_stride(n);
  T* ws = reinterpret_cast<T*>(smem_raw) + (size_t)warp * warp_vals;
  T* tab = ws;                                            // [n][13][32]
  T* big = tab + n * kLmTab * 32;
  // stage-2 views
  T* tile = big;                                          // [32][tstride]
  T* gst = tile + 32 * tstride;                           // [slot_a][GC][6][32]
  // stage-1 views (same memory, earlier in time)
  T* ffq = big;                                           // [n][2][32]: q, then (f1, f2)
  T* sta = ffq + 2 * n * 32;                              // [slot_a][22][32]
  T* stb = sta + m.n_slot_a * 22 * 32;                    // [slot_b][9][32]
#define LTAB(i, k) tab[((i) * kLmTab + (k)) * 32 + lane]
#define LSTA(s, k) sta[((s) * 22 + (k)) * 32 + lane]
#define LSTB(s, k) stb[((s) * 9 + (k)) * 32 + lane]
#define LGST(s, c, k) gst[(((s) * GC + (c)) * 6 + (k)) * 32 + lane]
  T* mytile = tile + lane * tstride;

  const int64_t ntasks = (B + 31) / 32;
  for (int64_t task = (int64_t)blockIdx.x * nwarps + warp; task < ntasks; task += (int64_t)gridDim.x * nwarps) {
    const int64_t first = task * 32;
    const int nk = (int)((B - first) < 32 ? (B - first) : 32);
    // ---------------------------------------------------------------- stage q (coalesced)
    {
      const T* src = q + first * n;
      const int count = nk * n;
      int kn = lane / n, jn = lane - kn * n;              // element e = kn * n + jn of the slab
      const int dk = 32 / n, dj = 32 - dk * n;
      for (int e = lane; e < 32 * n; e += 32) {
        ffq[(plan.pos[jn] * 2) * 32 + kn] = e < count ? __ldg(src + e) : T(0);
        kn += dk; jn += dj;
        if (jn >= n) { jn -= n; kn += 1; }
      }
    }
    {
      // pull the next task's slab of q towards L2 while this one is processed
      const int64_t nxt = task + (int64_t)gridDim.x * nwarps;
      if (nxt < ntasks) {
        const char* p = reinterpret_cast<const char*>(q + nxt * 32 * n);
        const int bytes = 32 * n * (int)sizeof(T);
        for (int off = lane * 128; off < bytes; off += 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + off));
      }
    }
    __syncwarp();
    // ================================================================ stage 1
    {
      T E[9];
      // ---- rotations, root -> leaf
#pragma unroll(NMAX > 0 ? NMAX : 1)
      for (int i = 0; i < (NMAX > 0 ? NMAX : n); ++i) {
        if (NMAX > 0 && i >= n) break;
        T f1, f2;
        {
          const T qi = ffq[(i * 2) * 32 + lane];
          if (!PRISM || m.kind[i] == 0) sincos_t(qi, &f2, &f1);
          else { f1 = qi; f2 = T(0); }
        }
        ffq[(i * 2) * 32 + lane] = f1;
        ffq[(i * 2 + 1) * 32 + lane] = f2;
        const int par = m.parent[i];
        T Ep[9];
        if (par < 0) {
#pragma unroll
          for (int k = 0; k < 9; ++k) Ep[k] = (k % 4 == 0) ? T(1) : T(0);
        } else if (par != i - 1) {
          const int s = m.slot_a[par];
#pragma unroll
          for (int k = 0; k < 9; ++k) Ep[k] = LSTA(s, k);
        } else {
#pragma unroll
          for (int k = 0; k < 9; ++k) Ep[k] = E[k];
        }
        T Ej[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) Ej[k] = fma_t(m.EC[i][k], f2, fma_t(m.EB[i][k], f1, m.EA[i][k]));
#pragma unroll
        for (int rr = 0; rr < 3; ++rr)
#pragma unroll
          for (int cc = 0; cc < 3; ++cc)
            E[3 * rr + cc] = Ej[3 * rr] * Ep[cc] + Ej[3 * rr + 1] * Ep[3 + cc] + Ej[3 * rr + 2] * Ep[6 + cc];
        const int sa = m.slot_a[i], sb = m.slot_b[i];
        if (sa >= 0) {
#pragma unroll
          for (int k = 0; k < 9; ++k) LSTA(sa, k) = E[k];
        }
        if (sb >= 0) {
#pragma unroll
          for (int k = 0; k < 9; ++k) LSTB(sb, k) = E[k];
        }
      }
      // ---- articulated inertias, leaf -> root
      for (int s = 0; s < m.n_slot_a; ++s)
#pragma unroll
        for (int k = 0; k < 22; ++k) LSTA(s, k) = T(0);
      // IA = [[A, Bm], [Bm^T, C]] : A sym (0..5), Bm 3x3 row-major (6..14), C sym (15..20)
      T IA[21];
#pragma unroll(NMAX > 0 ? NMAX : 1)
      for (int i = (NMAX > 0 ? NMAX : n) - 1; i >= 0; --i) {
        if (NMAX > 0 && i >= n) continue;
        const bool chained = (i != n - 1) && (m.parent[i + 1 < RBD_MAX_DOF ? i + 1 : i] == i);
        if (!chained && i != n - 1) {
          const int s = m.slot_b[i];
#pragma unroll
          for (int k = 0; k < 9; ++k) E[k] = LSTB(s, k);
        }
        const T f1 = ffq[(i * 2) * 32 + lane], f2 = ffq[(i * 2 + 1) * 32 + lane];
        const int kind = PRISM ? m.kind[i] : 0;
        const int par = m.parent[i];
        // own rigid inertia about p_i, world-aligned axes
        {
          const T mi = m.mass[i];
          T hr[3];
#pragma unroll
          for (int cc = 0; cc < 3; ++cc) hr[cc] = E[cc] * m.h[i][0] + E[3 + cc] * m.h[i][1] + E[6 + cc] * m.h[i][2];
          T IbE[9];
          const T xx = m.Ib[i][0], xy = m.Ib[i][1], xz = m.Ib[i][2], yy = m.Ib[i][3], yz = m.Ib[i][4], zz = m.Ib[i][5];
#pragma unroll
          for (int cc = 0; cc < 3; ++cc) {
            IbE[cc] = xx * E[cc] + xy * E[3 + cc] + xz * E[6 + cc];
            IbE[3 + cc] = xy * E[cc] + yy * E[3 + cc] + yz * E[6 + cc];
            IbE[6 + cc] = xz * E[cc] + yz * E[3 + cc] + zz * E[6 + cc];
          }
          T own[6];
          int idx = 0;
#pragma unroll
          for (int rr = 0; rr < 3; ++rr)
#pragma unroll
            for (int cc = rr; cc < 3; ++cc)
              own[idx++] = E[rr] * IbE[cc] + E[3 + rr] * IbE[3 + cc] + E[6 + rr] * IbE[6 + cc];
          if (chained) {
#pragma unroll
            for (int k = 0; k < 6; ++k) IA[k] += own[k];
            IA[7] -= hr[2]; IA[8] += hr[1]; IA[9] += hr[2]; IA[11] -= hr[0]; IA[12] -= hr[1]; IA[13] += hr[0];
            IA[15] += mi; IA[18] += mi; IA[20] += mi;
          } else {
#pragma unroll
            for (int k = 0; k < 6; ++k) IA[k] = own[k];
            IA[6] = T(0); IA[7] = -hr[2]; IA[8] = hr[1];
            IA[9] = hr[2]; IA[10] = T(0); IA[11] = -hr[0];
            IA[12] = -hr[1]; IA[13] = hr[0]; IA[14] = T(0);
            IA[15] = mi; IA[16] = T(0); IA[17] = T(0); IA[18] = mi; IA[19] = T(0); IA[20] = mi;
          }
        }
        const int sa = m.slot_a[i];
        if (sa >= 0) {
#pragma unroll
          for (int k = 0; k < 21; ++k) IA[k] += LSTA(sa, k);
        }
        T w[3], rw[3], Ej[9];
        {
          T r[3], t[3];
#pragma unroll
          for (int k = 0; k < 9; ++k) Ej[k] = fma_t(m.EC[i][k], f2, fma_t(m.EB[i][k], f1, m.EA[i][k]));
#pragma unroll
          for (int k = 0; k < 3; ++k) r[k] = fma_t(m.rC[i][k], f2, fma_t(m.rB[i][k], f1, m.rA[i][k]));
#pragma unroll
          for (int k = 0; k < 3; ++k) t[k] = Ej[3 * k] * r[0] + Ej[3 * k + 1] * r[1] + Ej[3 * k + 2] * r[2];
#pragma unroll
          for (int cc = 0; cc < 3; ++cc) {
            rw[cc] = E[cc] * t[0] + E[3 + cc] * t[1] + E[6 + cc] * t[2];       // r_i = p_i - p_parent, world axes
            w[cc] = E[cc] * m.axis[i][0] + E[3 + cc] * m.axis[i][1] + E[6 + cc] * m.axis[i][2];
          }
        }
        T U[6];
        if (kind == 0) {
          sym3_mul(IA, w, U);
#pragma unroll
          for (int cc = 0; cc < 3; ++cc) U[3 + cc] = IA[6 + cc] * w[0] + IA[9 + cc] * w[1] + IA[12 + cc] * w[2];
        } else {
#pragma unroll
          for (int rr = 0; rr < 3; ++rr) U[rr] = IA[6 + 3 * rr] * w[0] + IA[7 + 3 * rr] * w[1] + IA[8 + 3 * rr] * w[2];
          sym3_mul(IA + 15, w, U + 3);
        }
        const T D = kind == 0 ? dot3s(w, U) : dot3s(w, U + 3);
        const T invD = T(1) / D;                                               // RBDReference.py:698-700
        LTAB(i, 0) = w[0]; LTAB(i, 1) = w[1]; LTAB(i, 2) = w[2]; LTAB(i, 3) = invD;
#pragma unroll
        for (int k = 0; k < 6; ++k) LTAB(i, 4 + k) = U[k];
        LTAB(i, 10) = rw[0]; LTAB(i, 11) = rw[1]; LTAB(i, 12) = rw[2];
        if (par >= 0) {
          // IA -= U U^T / D (:728-731), then translate to the parent's origin (:732-733)
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
          T RW[9];
#pragma unroll
          for (int bb = 0; bb < 3; ++bb) {
            RW[bb] = rw[1] * W[3 * bb + 2] - rw[2] * W[3 * bb + 1];
            RW[3 + bb] = rw[2] * W[3 * bb] - rw[0] * W[3 * bb + 2];
            RW[6 + bb] = rw[0] * W[3 * bb + 1] - rw[1] * W[3 * bb];
          }
          IA[0] += T(2) * RW[0];
          IA[1] += RW[1] + RW[3];
          IA[2] += RW[2] + RW[6];
          IA[3] += T(2) * RW[4];
          IA[4] += RW[5] + RW[7];
          IA[5] += T(2) * RW[8];
          if (par != i - 1) {
            const int s = m.slot_a[par];
#pragma unroll
            for (int k = 0; k < 21; ++k) LSTA(s, k) += IA[k];
          } else {
            T Ep[9];                                       // E_parent = E_J^T E
#pragma unroll
            for (int rr = 0; rr < 3; ++rr)
#pragma unroll
              for (int cc = 0; cc < 3; ++cc)
                Ep[3 * rr + cc] = Ej[rr] * E[cc] + Ej[3 + rr] * E[3 + cc] + Ej[6 + rr] * E[6 + cc];
#pragma unroll
            for (int k = 0; k < 9; ++k) E[k] = Ep[k];
          }
        }
      }
    }
    __syncwarp();
    // ================================================================ stage 2 (the tile aliases f1 f2)
    for (int k = 0; k < nn; ++k) mytile[k] = T(0);        // entries between root components stay zero
#pragma unroll(NMAX2 > 0 ? (NMAX2 + GC - 1) / GC : 1)
    for (int j0 = 0; j0 < (NMAX2 > 0 ? NMAX2 : n); j0 += GC) {
      if (NMAX2 > 0 && j0 >= n) break;
      const int jtop = (j0 + GC < n ? j0 + GC : n) - 1;   // last column of the group
      int oj[GC];
#pragma unroll
      for (int c = 0; c < GC; ++c) oj[c] = plan.orig[j0 + c < n ? j0 + c : n - 1];
      T V[GC][6];
      // ---------------------------------------------------------------- phase B: leaf -> root
#pragma unroll
      for (int c = 0; c < GC; ++c)
#pragma unroll
        for (int k = 0; k < 6; ++k) V[c][k] = T(0);
#pragma unroll(NMAX2 > 0 ? NMAX2 : 1)
      for (int a = (NMAX2 > 0 ? (j0 + GC < NMAX2 ? j0 + GC : NMAX2) - 1 : jtop); a >= 0; --a) {
        if (NMAX2 > 0 && a > jtop) continue;
        const int send = plan.sub_end[a];
        if (send <= j0) continue;                         // no column of the group below body a
        T w[3], U[6], r[3];
        const T invD = LTAB(a, 3);
#pragma unroll
        for (int k = 0; k < 3; ++k) { w[k] = LTAB(a, k); r[k] = LTAB(a, 10 + k); }
#pragma unroll
        for (int k = 0; k < 6; ++k) U[k] = LTAB(a, 4 + k);
        const bool pris = PRISM && m.kind[a] != 0;
        const int oa = plan.orig[a];
        if (jtop < send) {
          // every started column of the group hangs below body a: straight-line code, GC independent
          // chains.  Columns j < a have not started (F = 0 gives Minv = 0 and leaves F = 0).
#pragma unroll
          for (int c = 0; c < GC; ++c) {
            const int j = j0 + c;
            const T sF = pris ? dot3s(w, V[c] + 3) : dot3s(w, V[c]);
            const T mij = (j == a ? invD : T(0)) - invD * sF;                  // :700-708
            if (j >= a && j <= jtop) mytile[oa * n + oj[c]] = mij;
#pragma unroll
            for (int k = 0; k < 6; ++k) V[c][k] = fma_t(U[k], mij, V[c][k]);   // :721-726
            cross3_add(r, V[c] + 3, V[c]);                                     // moment about the parent's origin
          }
        } else {
#pragma unroll
          for (int c = 0; c < GC; ++c) {
            const int j = j0 + c;
            if (j >= a && j < send) {                     // a is j or an ancestor of j (warp-uniform)
              const T sF = pris ? dot3s(w, V[c] + 3) : dot3s(w, V[c]);
              const T mij = (j == a ? invD : T(0)) - invD * sF;
              mytile[oa * n + oj[c]] = mij;
#pragma unroll
              for (int k = 0; k < 6; ++k) V[c][k] = fma_t(U[k], mij, V[c][k]);
              cross3_add(r, V[c] + 3, V[c]);
            }
          }
        }
      }
      // ---------------------------------------------------------------- phase C: root -> leaf
      const int cr0 = mp.comp_root[j0];                   // first body any column of the group couples with
#pragma unroll(NMAX2 > 0 ? NMAX2 : 1)
      for (int a = (NMAX2 > 0 ? 0 : cr0); a <= (NMAX2 > 0 ? (j0 + GC < NMAX2 ? j0 + GC : NMAX2) - 1 : jtop); ++a) {
        if (NMAX2 > 0 && (a < cr0 || a > jtop)) continue;
        const int cend = plan.comp_end[a];
        if (cend <= j0) continue;                         // a's root component ends before the group
        const int send = plan.sub_end[a];
        const int par = m.parent[a];
        T w[3], U[6], r[3];
        const T invD = LTAB(a, 3);
#pragma unroll
        for (int k = 0; k < 3; ++k) { w[k] = LTAB(a, k); r[k] = LTAB(a, 10 + k); }
#pragma unroll
        for (int k = 0; k < 6; ++k) U[k] = LTAB(a, 4 + k);
        const bool pris = PRISM && m.kind[a] != 0;
        const int oa = plan.orig[a];
        const int sl = m.slot_a[a];
        const int psl = (par >= 0 && par != a - 1) ? m.slot_a[par] : -1;
        // Straight-line over the GC columns (independent chains).  A column that is not coupled with
        // body a (j < a: finished; j >= cend: a later root component, restarted at its root) only
        // computes garbage that is never stored.
        T mij[GC];
#pragma unroll
        for (int c = 0; c < GC; ++c) mij[c] = (j0 + c < send) ? mytile[oa * n + oj[c]] : T(0);
        if (par < 0) {
          // the root of a component: no parent term (:778-781)
#pragma unroll
          for (int c = 0; c < GC; ++c)
#pragma unroll
            for (int k = 0; k < 6; ++k) V[c][k] = T(0);
        } else {
          if (psl >= 0) {                                 // parent is a branch point: its G was stashed
#pragma unroll
            for (int c = 0; c < GC; ++c)
#pragma unroll
              for (int k = 0; k < 6; ++k) V[c][k] = LGST(psl, c, k);
          }
#pragma unroll
          for (int c = 0; c < GC; ++c) {
            cross3_add(V[c], r, V[c] + 3);                                     // velocity at p_a: v += w x r
            mij[c] = fma_t(-invD, dot6s(U, V[c]), mij[c]);                     // :771-773
          }
        }
#pragma unroll
        for (int c = 0; c < GC; ++c) {
          const int j = j0 + c;
          if (pris) {
#pragma unroll
            for (int k = 0; k < 3; ++k) V[c][3 + k] = fma_t(w[k], mij[c], V[c][3 + k]);
          } else {
#pragma unroll
            for (int k = 0; k < 3; ++k) V[c][k] = fma_t(w[k], mij[c], V[c][k]);   // :774-781
          }
          if (j >= a && j < cend) {                       // same root component, upper triangle (warp-uniform)
            if (sl >= 0) {
#pragma unroll
              for (int k = 0; k < 6; ++k) LGST(sl, c, k) = V[c][k];
            }
            mytile[oa * n + oj[c]] = mij[c];
            mytile[oj[c] * n + oa] = mij[c];                                   // :799-804
          }
        }
      }
    }
    __syncwarp();
    // ---------------------------------------------------------------- coalesced slab write
    {
      T* dst = Minv + first * (int64_t)nn;
      const int count = nk * nn;
      int kn = lane / nn, idx = lane - kn * nn;
      const int dk = 32 / nn, di = 32 - dk * nn;
      for (int e = lane; e < count; e += 32) {
        __stcs(dst + e, tile[kn * tstride + idx]);
        kn += dk; idx += di;
        if (idx >= nn) { idx -= nn; kn += 1; }
      }
    }
    __syncwarp();
  }
#undef LTAB
#undef LSTA
#undef LSTB
#undef LGST
}

}  // namespace rbd
