// rbd_tile_minv_kernels.cuh - fused minv (RBDReference.py:785-806) for large trees (Atlas): one CTA
// owns a TILE of 32 knot points, one knot point per lane in every phase, and the per-body table
// (w, 1/D, U, r) of the tile never leaves shared memory.
//
// Same recursion and the same local world-aligned frames as rbd_lane_minv_kernels.cuh; what changes
// is who does what:
//
//   stage 1  (articulated inertias, :694-733) is parallel over the BRANCHES of the tree: the tree is
//            cut into chains (maximal runs i, i+1 with parent[i+1] = i of the depth-first numbering),
//            every chain is a warp's job (lane = knot point), chains that do not depend on each other
//            run concurrently on different warps and __syncthreads() separates the dependency levels.
//            The forward sweep leaves w_i, r_i and two rows of the rotation E_i in body i's table row;
//            the backward sweep reads them back (third row = cross product), so nothing is re-derived
//            and no per-warp stash exists.  A chain hands its articulated inertia to the parent chain
//            through its own 21-value slot (no two warps ever add into the same memory).
//   stage 2  (rows of Minv, :700-726 and :771-781) is parallel over COLUMN GROUPS: up to GC = 4
//            consecutive columns of one root component per warp (host-balanced), four independent
//            dependency chains per lane, one table row read serves four (body, column) pairs.
//            Every group sweeps ALL bodies of its component, which yields whole columns = (by symmetry,
//            :799-804) whole row segments: lane k writes Minv[k][a][j0 .. j0+3] as 32 contiguous
//            bytes straight from registers.  No output tile, no mirror pass, no zero fill pass.
//
// Against the hybrid kernel it replaces for n > 16: no scratch hand-off through L2 / HBM (2.6x the
// compulsory DRAM traffic there), no idle lanes in the triangular column sweeps.
#pragma once
#include "rbd_common.cuh"
#include "rbd_grad_kernels.cuh"
#include "rbd_minv_kernels.cuh"
#include "rbd_coop_minv_kernels.cuh"

namespace rbd {

constexpr int kTmTab = 13;           // w(3) invD U(6) r(3); stage 1 keeps q / E rows 0, 1 in the invD / U slots
constexpr int kTmMaxGC = 4;          // columns per group: 4 (8 warps per CTA) or 2 (16 warps per CTA)
constexpr int kTmMaxWarps = 16;
constexpr int kTmMaxGroups = RBD_MAX_DOF;

constexpr int kTmMaxSteps = 30 * RBD_MAX_DOF;   // step words of all column groups

// One step of a column group, packed by the host (rbd_capi.cu build_tile_plan):
//   bits 0..4 body a | 5..8 depth(a) | 9..12 columns of the group that hang below a (a is the column or an
//   ancestor) | 13..15 column of the group that IS body a (7: none) | 16..19 slot_a(a) + 1 |
//   20..23 slot_a(parent) + 1 if the parent is a branch point whose G has to be reloaded | 24 a is a root |
//   25 a is prismatic | 26..30 row of body a in the caller's numbering
__host__ __device__ inline int tm_pack_step(int a, int depth, int mask, int self, int sl, int psl, int root, int pris, int orow) {
  return a | (depth << 5) | (mask << 9) | (self << 13) | ((sl + 1) << 16) | ((psl + 1) << 20) | (root << 24) | (pris << 25) | (orow << 26);
}

struct TilePlan {
  int ok;                            // 0: the robot does not fit this kernel (too many slots / deep trees)
  int nwarps;                        // warps per CTA the work lists were built for
  int gc;                            // columns per group the step tables were built for
  int nchain;
  int chain_begin[RBD_MAX_DOF];      // bodies [begin, end) in depth-first numbering
  int chain_end[RBD_MAX_DOF];
  int nflevel, nblevel;              // dependency levels of the forward / backward sweep
  // work lists: chains of (level, warp) are f_item[f_begin[level * nwarps + warp] .. f_begin[.. + 1])
  int f_begin[RBD_MAX_DOF * kTmMaxWarps + 1];
  int f_item[RBD_MAX_DOF];
  int b_begin[RBD_MAX_DOF * kTmMaxWarps + 1];
  int b_item[RBD_MAX_DOF];
  int nslot;                         // hand-off slots (one per chain whose head has a parent)
  int out_slot[RBD_MAX_DOF];         // per chain: slot its head writes, -1 for a root chain
  int in_begin[RBD_MAX_DOF + 1];     // per body: slots to add, in_slot[in_begin[i] .. in_begin[i+1])
  int in_slot[RBD_MAX_DOF];
  int maxdepth;
  int ngroup;
  int g_begin[kTmMaxWarps + 1];      // groups of warp w: g_item[g_begin[w] .. g_begin[w+1])
  int g_item[kTmMaxGroups];
  int g_first[kTmMaxGroups];         // first column of the group (depth-first numbering)
  int g_ncols[kTmMaxGroups];
  int g_ocol[kTmMaxGroups];          // first column in the caller's numbering, -1: columns not consecutive there
  int g_sb[kTmMaxGroups];            // steps[g_sb .. g_sc): phase B (leaf -> root), steps[g_sc .. g_sz): phase C (preorder)
  int g_sc[kTmMaxGroups];
  int g_sz[kTmMaxGroups];            // steps[g_sz .. g_se): rows of the other root components (zeros)
  int g_se[kTmMaxGroups];
  int steps[kTmMaxSteps];
  int nslot_g;                       // G stashes (= FastModel.n_slot_a)
};

// shared memory of one CTA, in values of T:  table | max(stage-1 slots, stage-2 per-warp scratch)
__host__ __device__ inline int tile_minv_warp_vals(int maxdepth, int nslot_g, int gc) {
  return ((maxdepth + 1) * gc + nslot_g * gc * 6) * 32;                 // mb | G stashes
}
constexpr int kTmMdl = 52;           // per-body constants in shared memory: EA EB EC (27) rA rB rC (9) axis (3) m h (4) Ib (6), padded
__host__ __device__ inline size_t tile_minv_smem_vals(int n, int nslot, int maxdepth, int nslot_g, int nwarps, int gc) {
  const size_t s1 = (size_t)nslot * 21 * 32;
  const size_t s2 = (size_t)nwarps * tile_minv_warp_vals(maxdepth, nslot_g, gc);
  return (size_t)n * kTmMdl + (size_t)n * kTmTab * 32 + (s1 > s2 ? s1 : s2);
}

// Minv[row][j0 .. j0 + nc) <- v[0 .. nc)   (p points at column j0; par = 0: p is 2-value aligned, 1: p + 1 is,
// 2: nothing is known - the result's n*n is odd or its base is not 2-value aligned)
template <typename T, int GC>
__device__ __forceinline__ void tm_store_row(T* p, const T* v, int nc, int par) {
  typedef typename Vec2<T>::type V2;
  if (GC == 4 && nc == 4 && par == 0) {
    V2 a, b;
    a.x = v[0]; a.y = v[1]; b.x = v[GC - 2]; b.y = v[GC - 1];
    __stcs(reinterpret_cast<V2*>(p), a);
    __stcs(reinterpret_cast<V2*>(p) + 1, b);
  } else if (GC == 4 && nc == 4 && par == 1) {
    V2 a;
    a.x = v[1]; a.y = v[GC - 2];
    __stcs(p, v[0]);
    __stcs(reinterpret_cast<V2*>(p + 1), a);
    __stcs(p + 3, v[GC - 1]);
  } else if (GC == 2 && nc == 2 && par == 0) {
    V2 a;
    a.x = v[0]; a.y = v[1];
    __stcs(reinterpret_cast<V2*>(p), a);
  } else {
#pragma unroll
    for (int c = 0; c < GC; ++c)
      if (c < nc) __stcs(p + c, v[c]);
  }
}

template <typename T, bool PRISM, int GC>
__global__ void __launch_bounds__(GC == 4 ? 256 : 512)
minv_tile_kernel(const __grid_constant__ FastModel<T> m, const __grid_constant__ DfsPlan plan,
                 const __grid_constant__ TilePlan tp, int64_t B,
                 const T* __restrict__ q, T* __restrict__ Minv) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = m.n;
  const int nn = n * n;
  const int nwarps = blockDim.x >> 5;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // The robot's constants: indexed constant-bank reads (a run-time body index per warp, several warps on different
  // bodies) miss the small first-level constant cache and serialise the dependent FMAs; a broadcast shared-memory
  // load does not.
  T* mdl = reinterpret_cast<T*>(smem_raw);                // [n][52]
  T* tab = mdl + (size_t)n * kTmMdl;                      // [n][13][32]
  T* big = tab + (size_t)n * kTmTab * 32;
  for (int idx = threadIdx.x; idx < n * kTmMdl; idx += blockDim.x) {
    const int i = idx / kTmMdl, k = idx - i * kTmMdl;
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
  __syncthreads();
  T* slots = big;                                         // stage 1: [nslot][21][32]
  T* mbw = big + (size_t)warp * tile_minv_warp_vals(tp.maxdepth, tp.nslot_g, GC);   // stage 2: [depth][GC][32]
  T* gst = mbw + (tp.maxdepth + 1) * GC * 32;             //          [slot][GC][6][32]
#define TTAB(i, k) tab[((i) * kTmTab + (k)) * 32 + lane]
#define TSLOT(s, k) slots[((s) * 21 + (k)) * 32 + lane]
#define TMB(d, c) mbw[((d) * GC + (c)) * 32 + lane]
#define TGST(s, c, k) gst[(((s) * GC + (c)) * 6 + (k)) * 32 + lane]
  const bool vec_ok = (nn & 1) == 0 && (reinterpret_cast<uintptr_t>(Minv) & (2 * sizeof(T) - 1)) == 0;

  const int64_t ntiles = (B + 31) / 32;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t first = tile * 32;
    const int nk = (int)((B - first) < 32 ? (B - first) : 32);
    // ---------------------------------------------------------------- q of the tile (coalesced) -> (cos, sin) in slots 3, 0
    {
      const T* src = q + first * n;
      const int count = nk * n;
      for (int e = threadIdx.x; e < 32 * n; e += blockDim.x) {
        const int kn = e / n, jn = e - kn * n;
        const int i = plan.pos[jn];
        T qi[1], sv[1], cv[1];
        qi[0] = e < count ? __ldg(src + e) : T(0);
        T f1 = qi[0], f2 = T(0);
        if (!PRISM || m.kind[i] == 0) { sincos_batch<1>(qi, sv, cv); f1 = cv[0]; f2 = sv[0]; }
        tab[(i * kTmTab + 3) * 32 + kn] = f1;
        tab[(i * kTmTab + 0) * 32 + kn] = f2;
      }
      const int64_t nxt = tile + gridDim.x;               // next tile's slab of q -> L2
      if (nxt < ntiles && threadIdx.x * 128 < 32 * n * (int)sizeof(T))
        asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(q + nxt * 32 * n) + threadIdx.x * 128));
    }
    __syncthreads();
    // ================================================================ stage 1a: rotations, root -> leaf
    for (int lvl = 0; lvl < tp.nflevel; ++lvl) {
      for (int it = tp.f_begin[lvl * nwarps + warp]; it < tp.f_begin[lvl * nwarps + warp + 1]; ++it) {
        const int c = tp.f_item[it];
        const int cb = tp.chain_begin[c], ce = tp.chain_end[c];
        T E[9];
        {
          const int par = m.parent[cb];
          if (par < 0) {
#pragma unroll
            for (int k = 0; k < 9; ++k) E[k] = (k % 4 == 0) ? T(1) : T(0);
          } else {
#pragma unroll
            for (int k = 0; k < 6; ++k) E[k] = TTAB(par, 4 + k);
            cross3(E, E + 3, E + 6);
          }
        }
#pragma unroll 1
        for (int i = cb; i < ce; ++i) {
          const T f1 = TTAB(i, 3), f2 = TTAB(i, 0);
          const T* mb = mdl + i * kTmMdl;
          T Ej[9], r[3];
#pragma unroll
          for (int k = 0; k < 9; ++k) Ej[k] = fma_t(mb[18 + k], f2, fma_t(mb[9 + k], f1, mb[k]));
#pragma unroll
          for (int k = 0; k < 3; ++k) r[k] = fma_t(mb[33 + k], f2, fma_t(mb[30 + k], f1, mb[27 + k]));
          // r_i = p_i - p_parent in world axes = E_parent^T r
#pragma unroll
          for (int cc = 0; cc < 3; ++cc) TTAB(i, 10 + cc) = E[cc] * r[0] + E[3 + cc] * r[1] + E[6 + cc] * r[2];
          // E_i = Ej E_parent, column by column
#pragma unroll
          for (int cc = 0; cc < 3; ++cc) {
            const T t0 = E[cc], t1 = E[3 + cc], t2 = E[6 + cc];
#pragma unroll
            for (int rr = 0; rr < 3; ++rr) E[3 * rr + cc] = fma_t(Ej[3 * rr + 2], t2, fma_t(Ej[3 * rr + 1], t1, Ej[3 * rr] * t0));
          }
#pragma unroll
          for (int cc = 0; cc < 3; ++cc)
            TTAB(i, cc) = E[cc] * mb[36] + E[3 + cc] * mb[37] + E[6 + cc] * mb[38];   // world joint axis
#pragma unroll
          for (int k = 0; k < 6; ++k) TTAB(i, 4 + k) = E[k];
        }
      }
      __syncthreads();
    }
    // ================================================================ stage 1b: articulated inertias, leaf -> root
    for (int lvl = 0; lvl < tp.nblevel; ++lvl) {
      for (int it = tp.b_begin[lvl * nwarps + warp]; it < tp.b_begin[lvl * nwarps + warp + 1]; ++it) {
        const int c = tp.b_item[it];
        const int cb = tp.chain_begin[c], ce = tp.chain_end[c];
        // IA = [[A, Bm], [Bm^T, C]] : A sym (0..5), Bm 3x3 row-major (6..14), C sym (15..20)
        T IA[21];
#pragma unroll 1
        for (int i = ce - 1; i >= cb; --i) {
          T E[9], w[3], rw[3];
#pragma unroll
          for (int k = 0; k < 6; ++k) E[k] = TTAB(i, 4 + k);
          cross3(E, E + 3, E + 6);
#pragma unroll
          for (int k = 0; k < 3; ++k) { w[k] = TTAB(i, k); rw[k] = TTAB(i, 10 + k); }
          const int kind = PRISM ? m.kind[i] : 0;
          // own rigid inertia about p_i, world-aligned axes
          {
            const T* mb = mdl + i * kTmMdl;
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
            T own[6];
            int idx = 0;
#pragma unroll
            for (int rr = 0; rr < 3; ++rr)
#pragma unroll
              for (int cc = rr; cc < 3; ++cc)
                own[idx++] = E[rr] * IbE[cc] + E[3 + rr] * IbE[3 + cc] + E[6 + rr] * IbE[6 + cc];
            if (i != ce - 1) {                            // the child i + 1 handed its inertia over in registers
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
          for (int s = tp.in_begin[i]; s < tp.in_begin[i + 1]; ++s) {      // chains hanging off body i
            const int sl = tp.in_slot[s];
#pragma unroll
            for (int k = 0; k < 21; ++k) IA[k] += TSLOT(sl, k);
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
          const T invD = T(1) / D;                                             // RBDReference.py:698-700
          TTAB(i, 3) = invD;
#pragma unroll
          for (int k = 0; k < 6; ++k) TTAB(i, 4 + k) = U[k];
          if (m.parent[i] >= 0) {
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
            if (i == cb) {                                 // head of the chain: the parent belongs to another chain
              const int sl = tp.out_slot[c];
#pragma unroll
              for (int k = 0; k < 21; ++k) TSLOT(sl, k) = IA[k];
            }
          }
        }
      }
      __syncthreads();
    }
    // ================================================================ stage 2: column groups (scratch aliases the slots)
    T* outk = Minv + (first + (lane < nk ? lane : 0)) * (int64_t)nn;
    const bool live = lane < nk;
    for (int gi = tp.g_begin[warp]; gi < tp.g_begin[warp + 1]; ++gi) {
      const int g = tp.g_item[gi];
      const int j0 = tp.g_first[g], nc = tp.g_ncols[g];
      const int ocol = tp.g_ocol[g];
      T V[GC][6];
#pragma unroll
      for (int c = 0; c < GC; ++c)
#pragma unroll
        for (int k = 0; k < 6; ++k) V[c][k] = T(0);
      // ---------------------------------------------------------------- phase B: leaf -> root (:700-726)
#pragma unroll 1
      for (int s = tp.g_sb[g]; s < tp.g_sc[g]; ++s) {
        const int st = tp.steps[s];
        const int a = st & 31, da = (st >> 5) & 15, mask = (st >> 9) & 15, self = (st >> 13) & 7;
        const T* row = tab + (size_t)a * kTmTab * 32 + lane;
        T w[3], U[6], r[3];
        const T invD = row[3 * 32];
#pragma unroll
        for (int k = 0; k < 3; ++k) { w[k] = row[k * 32]; r[k] = row[(10 + k) * 32]; }
#pragma unroll
        for (int k = 0; k < 6; ++k) U[k] = row[(4 + k) * 32];
        const bool pris = PRISM && ((st >> 25) & 1);
#pragma unroll
        for (int c = 0; c < GC; ++c) {
          if ((mask >> c) & 1) {                          // a is column c or an ancestor of it (warp-uniform)
            const T sF = pris ? dot3s(w, V[c] + 3) : dot3s(w, V[c]);
            const T mij = (self == c ? invD : T(0)) - invD * sF;
            TMB(da, c) = mij;
#pragma unroll
            for (int k = 0; k < 6; ++k) V[c][k] = fma_t(U[k], mij, V[c][k]);
            cross3_add(r, V[c] + 3, V[c]);                                     // moment about the parent's origin
          }
        }
      }
      // ---------------------------------------------------------------- phase C: every body of the component (:771-781)
#pragma unroll 1
      for (int s = tp.g_sc[g]; s < tp.g_sz[g]; ++s) {
        const int st = tp.steps[s];
        const int a = st & 31, da = (st >> 5) & 15, mask = (st >> 9) & 15;
        const int sl = ((st >> 16) & 15) - 1, psl = ((st >> 20) & 15) - 1;
        const T* row = tab + (size_t)a * kTmTab * 32 + lane;
        T w[3], U[6], r[3];
        const T invD = row[3 * 32];
#pragma unroll
        for (int k = 0; k < 3; ++k) { w[k] = row[k * 32]; r[k] = row[(10 + k) * 32]; }
#pragma unroll
        for (int k = 0; k < 6; ++k) U[k] = row[(4 + k) * 32];
        const bool pris = PRISM && ((st >> 25) & 1);
        T mij[GC];
#pragma unroll
        for (int c = 0; c < GC; ++c) mij[c] = ((mask >> c) & 1) ? TMB(da, c) : T(0);
        if ((st >> 24) & 1) {
          // the root of the component: no parent term (:778-781)
#pragma unroll
          for (int c = 0; c < GC; ++c)
#pragma unroll
            for (int k = 0; k < 6; ++k) V[c][k] = T(0);
        } else {
          if (psl >= 0) {                                 // parent is a branch point: its G was stashed
#pragma unroll
            for (int c = 0; c < GC; ++c)
#pragma unroll
              for (int k = 0; k < 6; ++k) V[c][k] = TGST(psl, c, k);
          }
#pragma unroll
          for (int c = 0; c < GC; ++c) {
            cross3_add(V[c], r, V[c] + 3);                                     // velocity at p_a: v += w x r
            mij[c] = fma_t(-invD, dot6s(U, V[c]), mij[c]);                     // :771-773
          }
        }
#pragma unroll
        for (int c = 0; c < GC; ++c) {
          if (pris) {
#pragma unroll
            for (int k = 0; k < 3; ++k) V[c][3 + k] = fma_t(w[k], mij[c], V[c][3 + k]);
          } else {
#pragma unroll
            for (int k = 0; k < 3; ++k) V[c][k] = fma_t(w[k], mij[c], V[c][k]);   // :774-781
          }
        }
        if (sl >= 0) {
#pragma unroll
          for (int c = 0; c < GC; ++c)
#pragma unroll
            for (int k = 0; k < 6; ++k) TGST(sl, c, k) = V[c][k];
        }
        // column j of Minv at row a == row j at column a (:799-804): a whole row segment per lane
        if (live) {
          const int orow = (st >> 26) & 31;
          if (ocol >= 0) tm_store_row<T, GC>(outk + orow * n + ocol, mij, nc, vec_ok ? ((orow * n + ocol) & 1) : 2);
          else {
#pragma unroll
            for (int c = 0; c < GC; ++c)
              if (c < nc) __stcs(outk + orow * n + plan.orig[j0 + c], mij[c]);
          }
        }
      }
      // ---------------------------------------------------------------- bodies of the other root components: zeros
      if (live) {
        T z[GC];
#pragma unroll
        for (int c = 0; c < GC; ++c) z[c] = T(0);
#pragma unroll 1
        for (int s = tp.g_sz[g]; s < tp.g_se[g]; ++s) {
          const int orow = (tp.steps[s] >> 26) & 31;
          if (ocol >= 0) tm_store_row<T, GC>(outk + orow * n + ocol, z, nc, vec_ok ? ((orow * n + ocol) & 1) : 2);
          else {
#pragma unroll
            for (int c = 0; c < GC; ++c)
              if (c < nc) __stcs(outk + orow * n + plan.orig[j0 + c], z[c]);
          }
        }
      }
    }
    __syncthreads();                                      // the table is rewritten by the next tile
  }
#undef TTAB
#undef TSLOT
#undef TMB
#undef TGST
}

}  // namespace rbd
