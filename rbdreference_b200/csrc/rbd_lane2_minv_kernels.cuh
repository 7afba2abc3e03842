// rbd_lane2_minv_kernels.cuh - fused minv (RBDReference.py:785-806) for LARGE robots (n > 16) with
// ONE KNOT POINT PER LANE in every phase.
//
// Same recursion, frames and column grouping as rbd_lane_minv_kernels.cuh, but the per-body table
// (w, 1/D, U, r: 13 values per body and knot point) of a warp's 32 knot points does not fit in
// shared memory (100 KB in FP64 for Atlas), so it lives in a warp-private global scratch block in
// [body][16-byte vector][lane] order: every table access of a warp is a fully coalesced 512-byte
// segment that is written and re-read within the same task, i.e. served by L2 as long as the
// resident warps' blocks fit there (the launcher bounds the resident warps).
// Stage 2 follows a host-made flat schedule of (group, phase, body) steps; the table row of step
// s + 3 is copied global -> shared with 16-byte cp.async.cg while step s is computed (a 4-slot ring
// per warp, each lane moves and reads only its own values).  It keeps the group's F / G six-vectors
// in registers and writes
// Minv[a, j] and its mirror straight to the caller's tensor; entries between different root
// components are zero-filled cooperatively.  No lane idles in the triangular sweeps and no table
// row is broadcast-gathered, which is what bounds the column-per-lane kernels.
//
// MEASURED (B200, Atlas, 2^18 knot points): 0.97e8 evals/s FP64 against 1.25e8 for the hybrid kernel.
// A knot point's 7.2 KB result cannot be staged on chip for 32 knot points at once, so every result
// is an 8-byte store 7.2 KB away from its neighbour lane's: partial-sector writes that L2 turns
// into DRAM read-modify-write (ncu: 2.5 GB read + 3.4 GB written for 1.95 GB of compulsory
// traffic) and that the memory pipeline serialises.  Kept as variant 6 (parity-tested); not selected.
#pragma once
#include "rbd_common.cuh"
#include "rbd_grad_kernels.cuh"
#include "rbd_minv_kernels.cuh"
#include "rbd_coop_minv_kernels.cuh"

namespace rbd {

constexpr int kL2GC = 4;                                  // columns per group
constexpr int kL2MaxGroups = RBD_MAX_DOF / kL2GC;
constexpr int kL2MaxDepth = 16;                           // deepest root path served (else: hybrid kernel)
constexpr int kL2MaxSlots = 4;                            // branch points whose G is stashed
constexpr int kL2Tab = 13;                                // w(3) invD U(6) r(3)
// 16-byte vectors: VW values each; NVT vectors hold a table row, one more holds (f1, f2) of stage 1
template <typename T> struct L2Vec;
template <> struct L2Vec<double> { typedef double2 V; static constexpr int VW = 2, NVT = 7, NVS = 8; };
template <> struct L2Vec<float> { typedef float4 V; static constexpr int VW = 4, NVT = 4, NVS = 5; };
constexpr int kL2MaxWarps = 8;
constexpr int kL2Ahead = 3;                               // table rows in flight ahead of the one in use
constexpr int kL2Ring = kL2Ahead + 1;                     // shared-memory ring slots per warp
constexpr int kL2MaxSteps = 448;                          // (group, phase, body) steps of stage 2

struct Lane2Plan {
  int ok;                                                 // 0: robot outside the limits above
  int ngroups;
  int nsteps;
  // stage-2 schedule: body (bits 0-4) | phase C (bit 5) | first step of its phase (bit 6) | group (bits 7-9).
  // Per group: phase B visits, leaf -> root, every body with a column of the group in its subtree;
  // phase C visits, root -> leaf, every body <= the group's last column in the columns' root components.
  unsigned short seq[kL2MaxSteps];
};

// shared memory per warp, in values of T: table-row ring | max(stage-1 stashes, Mb rows + G stashes)
template <typename T>
__host__ __device__ inline int lane2_warp_vals(int maxdepth, int nslot_a, int nslot_b) {
  const int s1 = (22 * nslot_a + 9 * nslot_b) * 32;
  const int s2 = (kL2GC * (maxdepth + 1) + nslot_a * kL2GC * 6) * 32;
  return kL2Ring * L2Vec<T>::NVT * L2Vec<T>::VW * 32 + (((s1 > s2 ? s1 : s2) + 3) & ~3);
}
template <typename T>
__host__ __device__ inline size_t lane2_scratch_vals_per_warp(int n) { return (size_t)n * L2Vec<T>::NVS * L2Vec<T>::VW * 32; }

template <typename T> struct L2Row { T w[3], invD, U[6], r[3]; };

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}

template <typename T, bool PRISM>
__global__ void __launch_bounds__(kL2MaxWarps * 32)
minv_lane2_kernel(const __grid_constant__ FastModel<T> m, const __grid_constant__ DfsPlan plan,
                  const __grid_constant__ CoopMinvPlan mp, const __grid_constant__ Lane2Plan lp, int maxdepth,
                  int64_t B, const T* __restrict__ q, T* __restrict__ Minv, T* __restrict__ scratch) {
  constexpr int GC = kL2GC;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = m.n;
  const int nn = n * n;
  const int nwarps = blockDim.x >> 5;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int mbd = maxdepth + 1;                           // Mb rows per column
  typedef typename L2Vec<T>::V VT;
  constexpr int VW = L2Vec<T>::VW, NVT = L2Vec<T>::NVT, NVS = L2Vec<T>::NVS;
  T* ring = reinterpret_cast<T*>(smem_raw) + (size_t)warp * lane2_warp_vals<T>(maxdepth, m.n_slot_a, m.n_slot_b);
  T* big = ring + kL2Ring * NVT * VW * 32;                // ring: [slot][NVT][32] vectors
  T* sta = big;                                           // stage 1: [slot_a][22][32]
  T* stb = sta + m.n_slot_a * 22 * 32;                    //          [slot_b][9][32]
  T* mbs = big;                                           // stage 2: [GC][maxdepth + 1][32]
  T* gst = mbs + GC * mbd * 32;                           //          [slot_a][GC][6][32]
#define HSTA(s, k) sta[((s) * 22 + (k)) * 32 + lane]
#define HSTB(s, k) stb[((s) * 9 + (k)) * 32 + lane]
#define LMB(c, d) mbs[((c) * mbd + (d)) * 32 + lane]
#define LGST(s, c, k) gst[(((s) * GC + (c)) * 6 + (k)) * 32 + lane]
  // scratch: [body][NVS][32 lanes] vectors; vector v of body i for this lane:
  VT* scr = reinterpret_cast<VT*>(scratch + (size_t)(blockIdx.x * nwarps + warp) * lane2_scratch_vals_per_warp<T>(n)) + lane;
#define SVEC(i, v) scr[((i) * NVS + (v)) * 32]

  const int64_t ntasks = (B + 31) / 32;
  for (int64_t task = (int64_t)blockIdx.x * nwarps + warp; task < ntasks; task += (int64_t)gridDim.x * nwarps) {
    const int64_t first = task * 32;
    const int nk = (int)((B - first) < 32 ? (B - first) : 32);
    // ================================================================ stage 1: lane = knot point
    {
      int64_t b = first + lane;
      if (b >= B) b = B - 1;                              // duplicate work, never stored
      const T* qb = q + b * n;
      T E[9];
      // ---- rotations, root -> leaf; q is fetched four bodies ahead
      T qpre[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) qpre[u] = qb[plan.orig[u < n ? u : 0]];
      if (lane < 2) {
        const int64_t nxt = task + (int64_t)gridDim.x * nwarps;
        if (nxt < ntasks) {
          const char* pq = reinterpret_cast<const char*>(q + nxt * 32 * n);
          const int bytes = 32 * n * (int)sizeof(T);
          for (int off = lane * 128; off < bytes; off += 2 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(pq + off));
        }
      }
#pragma unroll 1
      for (int i0 = 0; i0 < n; i0 += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = i0 + u;
          if (i >= n) break;
          T f1, f2;
          {
            const T qi = qpre[u];
            if (i + 4 < n) qpre[u] = qb[plan.orig[i + 4]];
            if (!PRISM || m.kind[i] == 0) sincos_t(qi, &f2, &f1);
            else { f1 = qi; f2 = T(0); }
          }
          {
            T pk[VW] = {};
            pk[0] = f1; pk[1] = f2;
            __stcg(&SVEC(i, NVT), *reinterpret_cast<const VT*>(pk));
          }
          const int par = m.parent[i];
          T Ep[9];
          if (par < 0) {
#pragma unroll
            for (int k = 0; k < 9; ++k) Ep[k] = (k % 4 == 0) ? T(1) : T(0);
          } else if (par != i - 1) {
            const int s = m.slot_a[par];
#pragma unroll
            for (int k = 0; k < 9; ++k) Ep[k] = HSTA(s, k);
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
            for (int k = 0; k < 9; ++k) HSTA(sa, k) = E[k];
          }
          if (sb >= 0) {
#pragma unroll
            for (int k = 0; k < 9; ++k) HSTB(sb, k) = E[k];
          }
        }
      }
      // ---- articulated inertias, leaf -> root (:694-733)
      for (int s = 0; s < m.n_slot_a; ++s)
#pragma unroll
        for (int k = 0; k < 22; ++k) HSTA(s, k) = T(0);
      // IA = [[A, Bm], [Bm^T, C]] : A sym (0..5), Bm 3x3 row-major (6..14), C sym (15..20)
      T IA[21];
      VT ffn = __ldcg(&SVEC(n - 1, NVT));                                     // (f1, f2) one body ahead
#pragma unroll 1
      for (int i = n - 1; i >= 0; --i) {
        const bool chained = (i != n - 1) && (m.parent[i + 1] == i);
        if (!chained && i != n - 1) {
          const int s = m.slot_b[i];
#pragma unroll
          for (int k = 0; k < 9; ++k) E[k] = HSTB(s, k);
        }
        const T f1 = ffn.x, f2 = ffn.y;
        if (i > 0) ffn = __ldcg(&SVEC(i - 1, NVT));
        const int kind = PRISM ? m.kind[i] : 0;
        const int par = m.parent[i];
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
          for (int k = 0; k < 21; ++k) IA[k] += HSTA(sa, k);
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
        {
          T pk[NVT * VW] = {};
          pk[0] = w[0]; pk[1] = w[1]; pk[2] = w[2]; pk[3] = invD;
#pragma unroll
          for (int k = 0; k < 6; ++k) pk[4 + k] = U[k];
          pk[10] = rw[0]; pk[11] = rw[1]; pk[12] = rw[2];
#pragma unroll
          for (int v = 0; v < NVT; ++v) __stcg(&SVEC(i, v), reinterpret_cast<const VT*>(pk)[v]);
        }
        if (par >= 0) {
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
            for (int k = 0; k < 21; ++k) HSTA(s, k) += IA[k];
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
    // ================================================================ zeros between root components
    {
      T* dst = Minv + first * (int64_t)nn;
      const int ol = plan.orig[lane < n ? lane : 0];
      for (int i = 0; i < n; ++i) {
        // row orig[i]: columns of bodies outside [comp_root(i), comp_end(i)) are structurally zero
        const bool z = lane < n && (lane < mp.comp_root[i] || lane >= plan.comp_end[i]);
        if (!__any_sync(0xffffffffu, z)) continue;
        T* row = dst + plan.orig[i] * n + ol;
        if (z) {
          for (int k = 0; k < nk; ++k) __stcs(row + (int64_t)k * nn, T(0));
        }
      }
    }
    // ================================================================ stage 2: lane = knot point
    // One flat, host-made sequence of (group, phase, body) steps.  Row s + kL2Ahead of the table is
    // copied global -> shared with cp.async (each lane its own 13 values: no cross-lane traffic,
    // no registers) while step s is computed, across group and phase boundaries.
    T* out = Minv + (first + (lane < nk ? lane : 0)) * (int64_t)nn;
    const bool store = lane < nk;
    const int nsteps = lp.nsteps;
    auto prefetch = [&](int s) {
      if (s < nsteps) {
        const VT* src = &SVEC(lp.seq[s] & 31, 0);
        VT* dst = reinterpret_cast<VT*>(ring) + (s % kL2Ring) * NVT * 32 + lane;
#pragma unroll
        for (int v = 0; v < NVT; ++v) cp_async16(dst + v * 32, src + v * 32);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
#pragma unroll
    for (int s = 0; s < kL2Ahead; ++s) prefetch(s);
    int j0 = 0, jtop = 0;
    int oj[GC];
    T V[GC][6];
#pragma unroll 1
    for (int s = 0; s < nsteps; ++s) {
      const int code = lp.seq[s];
      const int a = code & 31;
      const bool phaseC = (code >> 5) & 1;
      if ((code >> 6) & 1) {                              // first step of a phase
        if (!phaseC) {                                    // new column group
          j0 = (code >> 7) * GC;
          jtop = (j0 + GC < n ? j0 + GC : n) - 1;
#pragma unroll
          for (int c = 0; c < GC; ++c) oj[c] = plan.orig[j0 + c < n ? j0 + c : n - 1];
#pragma unroll
          for (int c = 0; c < GC; ++c)
#pragma unroll
            for (int k = 0; k < 6; ++k) V[c][k] = T(0);
        }
      }
      asm volatile("cp.async.wait_group %0;" ::"n"(kL2Ahead - 1) : "memory");   // row s has landed
      prefetch(s + kL2Ahead);
      L2Row<T> e;
      {
        const VT* rp = reinterpret_cast<const VT*>(ring) + (s % kL2Ring) * NVT * 32 + lane;
        T pk[NVT * VW];
#pragma unroll
        for (int v = 0; v < NVT; ++v) reinterpret_cast<VT*>(pk)[v] = rp[v * 32];
#pragma unroll
        for (int k = 0; k < 3; ++k) { e.w[k] = pk[k]; e.r[k] = pk[10 + k]; }
        e.invD = pk[3];
#pragma unroll
        for (int k = 0; k < 6; ++k) e.U[k] = pk[4 + k];
      }
      const int send = plan.sub_end[a];
      const int da = mp.depth[a];
      const bool pris = PRISM && m.kind[a] != 0;
      if (!phaseC) {
        // -------------------------------------------------------------- phase B: leaf -> root (:700-726)
        if (jtop < send) {
          // every started column of the group hangs below body a (columns j < a: F = 0 stays 0)
#pragma unroll
          for (int c = 0; c < GC; ++c) {
            const int j = j0 + c;
            const T sF = pris ? dot3s(e.w, V[c] + 3) : dot3s(e.w, V[c]);
            const T mij = (j == a ? e.invD : T(0)) - e.invD * sF;
            if (j >= a && j <= jtop) LMB(c, da) = mij;
#pragma unroll
            for (int k = 0; k < 6; ++k) V[c][k] = fma_t(e.U[k], mij, V[c][k]);
            cross3_add(e.r, V[c] + 3, V[c]);
          }
        } else {
#pragma unroll
          for (int c = 0; c < GC; ++c) {
            const int j = j0 + c;
            if (j >= a && j < send) {
              const T sF = pris ? dot3s(e.w, V[c] + 3) : dot3s(e.w, V[c]);
              const T mij = (j == a ? e.invD : T(0)) - e.invD * sF;
              LMB(c, da) = mij;
#pragma unroll
              for (int k = 0; k < 6; ++k) V[c][k] = fma_t(e.U[k], mij, V[c][k]);
              cross3_add(e.r, V[c] + 3, V[c]);
            }
          }
        }
      } else {
        // -------------------------------------------------------------- phase C: root -> leaf (:771-781)
        const int cend = plan.comp_end[a];
        const int par = m.parent[a];
        const int oa = plan.orig[a];
        const int sl = m.slot_a[a];
        const int psl = (par >= 0 && par != a - 1) ? m.slot_a[par] : -1;
        // a column that is not coupled with body a (j < a: finished; j >= cend: a later root
        // component, restarted at its root) only computes garbage that is never stored
        T mij[GC];
#pragma unroll
        for (int c = 0; c < GC; ++c) mij[c] = (j0 + c >= a && j0 + c < send) ? LMB(c, da) : T(0);
        if (par < 0) {
#pragma unroll
          for (int c = 0; c < GC; ++c)
#pragma unroll
            for (int k = 0; k < 6; ++k) V[c][k] = T(0);
        } else {
          if (psl >= 0) {
#pragma unroll
            for (int c = 0; c < GC; ++c)
#pragma unroll
              for (int k = 0; k < 6; ++k) V[c][k] = LGST(psl, c, k);
          }
#pragma unroll
          for (int c = 0; c < GC; ++c) {
            cross3_add(V[c], e.r, V[c] + 3);
            mij[c] = fma_t(-e.invD, dot6s(e.U, V[c]), mij[c]);
          }
        }
        T* orow = out + oa * n;
        T* ocol = out + oa;
#pragma unroll
        for (int c = 0; c < GC; ++c) {
          const int j = j0 + c;
          if (pris) {
#pragma unroll
            for (int k = 0; k < 3; ++k) V[c][3 + k] = fma_t(e.w[k], mij[c], V[c][3 + k]);
          } else {
#pragma unroll
            for (int k = 0; k < 3; ++k) V[c][k] = fma_t(e.w[k], mij[c], V[c][k]);
          }
          if (j >= a && j < cend) {
            if (sl >= 0) {
#pragma unroll
              for (int k = 0; k < 6; ++k) LGST(sl, c, k) = V[c][k];
            }
            if (store) {
              __stcs(orow + oj[c], mij[c]);
              if (j != a) __stcs(ocol + oj[c] * n, mij[c]);                      // :799-804
            }
          }
        }
      }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
  }
#undef HSTA
#undef HSTB
#undef LMB
#undef LGST
#undef SVEC
}

}  // namespace rbd
