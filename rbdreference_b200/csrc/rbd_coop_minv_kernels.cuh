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

// table row stride: 16-byte aligned and = 16 (mod 128) bytes, so rows of different bodies start in
// different bank groups (lanes of one warp gather rows of up to eight different bodies at once)
template <typename T> __host__ __device__ constexpr int cm_tab_stride() { return sizeof(T) == 8 ? 14 : 20; }
#define kCmTabStride (cm_tab_stride<T>())
constexpr int kCmIaStride = 22;      // articulated inertia handed to the parent (21, eleven pairs)
constexpr int kCmMaxWarps = 8;       // warps per CTA is chosen at launch (blockDim.x / 32)

struct CoopMinvPlan {
  int maxcomp;                       // largest root component
  int depth[RBD_MAX_DOF];
  int comp_root[RBD_MAX_DOF];
};

// per-warp shared memory, in values of T:  tab | mb | big, where `big` holds the children's
// inertias in phase A and is re-used for the G stashes and the output tile in phase C
// fb: floating base - the matrix has n + 5 rows and the warp keeps the inverse of the base's articulated inertia
// (21 values per knot point, kCmDinvStride apart) after `big`
constexpr int kCmDinvStride = 24;     // (a multiple of four values: the next warp's table stays 16-byte aligned in FP32)
// bpass_pairs > 0: the minv_bpass helper (BPASS = true) - the warp also keeps every body's rotation (12 values per lane)
// and, per knot point, the seven results (Minv entry, F column) of each (ancestor, column) pair
__host__ __device__ inline int coop_minv_bpass_stage(int npairs) { return (7 * npairs + 3) & ~3; }
template <typename T>
__host__ __device__ inline int coop_minv_warp_vals(int n, int G, int maxdepth, int nslot, bool fb = false, int bpass_pairs = 0) {
  const int ipw = 32 / G;
  const int nvv = fb ? n + 5 : n;
  const int a = 32 * kCmIaStride;
  const int c = 6 * nslot * 32 + ((ipw * nvv * nvv + 3) & ~3);
  return 32 * kCmTabStride + (maxdepth + 1) * 32 + (a > c ? a : c) + (fb ? ipw * kCmDinvStride : 0) +
         (bpass_pairs > 0 ? 32 * 12 + ipw * coop_minv_bpass_stage(bpass_pairs) : 0);
}
// per-CTA tables of the minv_bpass helper: first pair of every column (n ints) and the maps from every position of
// the warp's Minv / F slabs to its staged value or -1 (16-bit)
__host__ __device__ inline size_t coop_minv_bpass_map_bytes(int n, int G) {
  const int ipw = 32 / G;
  return (size_t)((n + 3) & ~3) * sizeof(int) + (size_t)(((ipw * n * n + 7) & ~7) + ((ipw * 6 * n * n + 7) & ~7)) * sizeof(short);
}
template <typename T>
__host__ __device__ inline size_t coop_minv_smem_bytes(int n, int G, int maxdepth, int nslot, int warps, bool fb = false, int bpass_pairs = 0) {
  return (size_t)(((n * kCoopMdlStride + 3) & ~3) + warps * coop_minv_warp_vals<T>(n, G, maxdepth, nslot, fb, bpass_pairs)) * sizeof(T) +
         (size_t)n * 4 * sizeof(int) + (bpass_pairs > 0 ? coop_minv_bpass_map_bytes(n, G) : 0);
}

// index of entry (r, c) of a symmetric 6x6 stored as its upper triangle, row by row (21 values)
__host__ __device__ constexpr int sym6_idx(int r, int c) { return r <= c ? r * 6 - r * (r - 1) / 2 + (c - r) : c * 6 - c * (c - 1) / 2 + (r - c); }

// s <- inverse of the symmetric positive definite 6x6 s (upper triangle): six symmetric sweeps (each pivot is the
// diagonal of a Schur complement, positive without pivoting), which leave -inverse
template <typename T>
__device__ __forceinline__ void sym6_invert(T (&s)[21]) {
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const T d = T(1) / s[sym6_idx(k, k)];
    T bk[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) bk[i] = s[sym6_idx(i, k)] * d;
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
      for (int j = i; j < 6; ++j)
        if (i != k && j != k) s[sym6_idx(i, j)] = fma_t(-bk[i], s[sym6_idx(k, j)], s[sym6_idx(i, j)]);
#pragma unroll
    for (int i = 0; i < 6; ++i)
      if (i != k) s[sym6_idx(i, k)] = bk[i];
    s[sym6_idx(k, k)] = -d;
  }
#pragma unroll
  for (int k = 0; k < 21; ++k) s[k] = -s[k];
}

// ---- per-body table entry in shared memory: w(3) invD | U(6) | r(3), stride kCmTabStride ------
template <typename T> struct TabEntry { T w[3], invD, U[6], r[3]; };
__device__ __forceinline__ void tab_load(const double* p, TabEntry<double>& e) {
  const double2* v = reinterpret_cast<const double2*>(p);
  const double2 a = v[0], b = v[1], c = v[2], d = v[3], f = v[4], g = v[5], h = v[6];
  e.w[0] = a.x; e.w[1] = a.y; e.w[2] = b.x; e.invD = b.y;
  e.U[0] = c.x; e.U[1] = c.y; e.U[2] = d.x; e.U[3] = d.y; e.U[4] = f.x; e.U[5] = f.y;
  e.r[0] = g.x; e.r[1] = g.y; e.r[2] = h.x;
}
__device__ __forceinline__ void tab_load(const float* p, TabEntry<float>& e) {
  const float4* v = reinterpret_cast<const float4*>(p);
  const float4 a = v[0], b = v[1], c = v[2];
  const float d = p[12];
  e.w[0] = a.x; e.w[1] = a.y; e.w[2] = a.z; e.invD = a.w;
  e.U[0] = b.x; e.U[1] = b.y; e.U[2] = b.z; e.U[3] = b.w; e.U[4] = c.x; e.U[5] = c.y;
  e.r[0] = c.z; e.r[1] = c.w; e.r[2] = d;
}
// packed topology word of a body: depth | (slot_a + 1) << 8 | (parent's slot_a + 1) << 16 | kind << 24
__host__ __device__ inline int cm_pack(int depth, int slot, int pslot, int kind) {
  return depth | ((slot + 1) << 8) | ((pslot + 1) << 16) | (kind << 24);
}

// Phases B and C of the header comment plus the mirrored, coalesced store, for the IPW knot points
// whose per-body table is in `tab`.  Lane (g, i) is column i of knot point g.
// imdl[a] = {parent, sub_end, orig, cm_pack(...)}.  PRISM = false: every joint is revolute.
// ZERO = false: the caller zeroed the tile and nothing else touched it since (every entry inside a
// root component is rewritten on each call, entries between components are never written).
// FB = true (floating base, :652-691 / :761-779): body 0 is the base, whose "1 / D" is the 6x6 `dinv` of its knot point;
// lane i >= 1 is column i + 5, the base's rows of that column are -dinv F (:687-691), its columns come from the mirror
// (the reference fills both triangles with the same numbers up to rounding) and lanes 0..5 write dinv itself (:686).
template <typename T, int G, bool PRISM, bool ZERO = true, bool FB = false>
__device__ __forceinline__ void minv_column_phases(int n, int maxdepth, int maxcomp, int nslot, bool valid, int i, int gbase,
                                                   int lane, int oi, int comp_root, const int4* imdl, const T* tab, T* mbw,
                                                   T* big, T* __restrict__ dst, int nknots, const T* dinv = nullptr) {
  constexpr int IPW = 32 / G;
  typedef typename Vec2<T>::type V2;
  const int nvv = FB ? n + 5 : n;                         // rows of the matrix
  const int nn = nvv * nvv;
  const int g = lane / G;
  const int col = FB ? oi + 5 : oi;                       // this lane's column
  const bool colv = valid && !(FB && i == 0);             // the lane owns a joint's column
  T m0[6];                                                // FB: the base's rows of the column
  T* tile = big + 6 * nslot * 32;                         // [IPW][n][n]
  T* mytile = tile + g * nn;
  const int tile_vals = IPW * nn;
  const bool pair_ok = (tile_vals & 1) == 0;              // slab is a whole number of aligned pairs
  // zero the output tile (entries between different root components stay zero)
  if (ZERO) {
    if (pair_ok) {
      V2 z; z.x = T(0); z.y = T(0);
      for (int k = lane; k < (tile_vals >> 1); k += 32) reinterpret_cast<V2*>(tile)[k] = z;
    } else {
      for (int k = lane; k < tile_vals; k += 32) tile[k] = T(0);
    }
  }
  // ------------------------------------------------------------------ phase B: walk to the root
  {
    T F[6] = {T(0), T(0), T(0), T(0), T(0), T(0)};
    int a = colv ? i : -1;
    for (int t = 0; t <= maxdepth; ++t) {
      if (a >= (FB ? 1 : 0)) {
        TabEntry<T> e;
        tab_load(tab + (gbase + a) * kCmTabStride, e);
        const int4 ia = imdl[a];
        const bool pris = PRISM && ((ia.w >> 24) != 0);
        const T sF = pris ? dot3s(e.w, F + 3) : dot3s(e.w, F);
        const T mij = (a == i ? e.invD : T(0)) - e.invD * sF;                  // :700-708
        mbw[(ia.w & 0xff) * 32 + lane] = mij;
#pragma unroll
        for (int k = 0; k < 6; ++k) F[k] = fma_t(e.U[k], mij, F[k]);           // :721-726
        cross3_add(e.r, F + 3, F);                                             // moment about the parent's origin
        a = ia.x;
      }
    }
    if (FB) {
      const T* dg = dinv + g * kCmDinvStride;
#pragma unroll
      for (int r = 0; r < 6; ++r) {
        T acc = T(0);
#pragma unroll
        for (int k = 0; k < 6; ++k) acc = fma_t(dg[sym6_idx(r, k)], F[k], acc);
        m0[r] = -acc;                                                          // :687-691
      }
    }
  }
  __syncwarp();
  // ------------------------------------------------------------------ phase C: preorder sweep
  if (colv) {
    T Gv[6];
    const int oin = col * nvv;
    if (FB) {
      // the base: G_0 = S_0 Minv[0:6, j] with S_0 = eye(6) (:778-781)
      const int4 ia = imdl[0];
#pragma unroll
      for (int k = 0; k < 6; ++k) Gv[k] = m0[k];
      const int sl = ((ia.w >> 8) & 0xff) - 1;
      if (sl >= 0) {
        T* gs = big + (sl * 6) * 32 + lane;
#pragma unroll
        for (int k = 0; k < 6; ++k) gs[k * 32] = Gv[k];
      }
#pragma unroll
      for (int k = 0; k < 6; ++k) { mytile[k * nvv + col] = m0[k]; mytile[oin + k] = m0[k]; }
    } else {
      // the root of the component: no parent term (:778-781)
      const int a = comp_root;
      TabEntry<T> e;
      tab_load(tab + (gbase + a) * kCmTabStride, e);
      const int4 ia = imdl[a];
      const T mij = mbw[lane];                            // depth 0; the root is an ancestor of (or is) i
      const bool pris = PRISM && ((ia.w >> 24) != 0);
#pragma unroll
      for (int k = 0; k < 3; ++k) { Gv[k] = pris ? T(0) : e.w[k] * mij; Gv[3 + k] = pris ? e.w[k] * mij : T(0); }
      const int sl = ((ia.w >> 8) & 0xff) - 1;
      if (sl >= 0) {
        T* gs = big + (sl * 6) * 32 + lane;
#pragma unroll
        for (int k = 0; k < 6; ++k) gs[k * 32] = Gv[k];
      }
      mytile[ia.z * n + oi] = mij;
      mytile[oin + ia.z] = mij;
    }
    for (int a = comp_root + 1; a <= i; ++a) {
      const int ra = FB ? imdl[a].z + 5 : imdl[a].z;      // row of body a
      TabEntry<T> e;
      tab_load(tab + (gbase + a) * kCmTabStride, e);
      const int4 ia = imdl[a];
      T mij = (i < ia.y) ? mbw[(ia.w & 0xff) * 32 + lane] : T(0);
      if (ia.x != a - 1) {                                // parent is a branch point: its G was stashed
        const T* gs = big + ((((ia.w >> 16) & 0xff) - 1) * 6) * 32 + lane;
#pragma unroll
        for (int k = 0; k < 6; ++k) Gv[k] = gs[k * 32];
      }
      cross3_add(Gv, e.r, Gv + 3);                                             // velocity at p_a: v += w x r
      mij = fma_t(-e.invD, dot6s(e.U, Gv), mij);                               // :771-773
      if (PRISM && ((ia.w >> 24) != 0)) {
#pragma unroll
        for (int k = 0; k < 3; ++k) Gv[3 + k] = fma_t(e.w[k], mij, Gv[3 + k]);
      } else {
#pragma unroll
        for (int k = 0; k < 3; ++k) Gv[k] = fma_t(e.w[k], mij, Gv[k]);         // :774-781
      }
      const int sl = ((ia.w >> 8) & 0xff) - 1;
      if (sl >= 0) {
        T* gs = big + (sl * 6) * 32 + lane;
#pragma unroll
        for (int k = 0; k < 6; ++k) gs[k * 32] = Gv[k];
      }
      mytile[ra * nvv + col] = mij;
      mytile[oin + ra] = mij;                                                  // :799-804
    }
  }
  if (FB && i < 6) {
    const T* dg = dinv + g * kCmDinvStride;
#pragma unroll
    for (int r = 0; r < 6; ++r) mytile[r * nvv + i] = dg[sym6_idx(r, i)];      // :686: the base block is inv(IA_0)
  }
  __syncwarp();
  // ------------------------------------------------------------------ coalesced slab write
  {
    const int count = nknots * nn;
    if (warp_bulk_store(dst, tile, count, lane)) {
      warp_bulk_store_wait(lane);                         // the caller rewrites the tile right away
    } else if (pair_ok && count == tile_vals && (reinterpret_cast<uintptr_t>(dst) & (sizeof(V2) - 1)) == 0) {
      for (int k = lane; k < (tile_vals >> 1); k += 32) __stcs(reinterpret_cast<V2*>(dst) + k, reinterpret_cast<const V2*>(tile)[k]);
    } else {
      for (int k = lane; k < count; k += 32) __stcs(dst + k, tile[k]);
    }
  }
}

// Composite-rigid-body algorithm (RBDReference.py:1090-1124, fixed-base branch) on the same
// table: with the rank-one downdate of phase A switched off, IA is the composite inertia IC, the
// table row of body i holds U_i = IC_i S_i and D_i = S_i.U_i = H[i,i] (in the invD slot), and lane
// (g, i) carries fh = U_i up its root path (:1113-1122): H[i,a] = H[a,i] = S_a . fh for every
// ancestor a, fh re-expressed about each parent's origin on the way.
template <typename T, int G, bool PRISM>
__device__ __forceinline__ void crba_column_phase(int n, int maxdepth, bool valid, int i, int gbase, int lane, int oi,
                                                  const int4* imdl, const T* tab, T* tile, T* __restrict__ dst, int nknots) {
  constexpr int IPW = 32 / G;
  typedef typename Vec2<T>::type V2;
  const int nn = n * n;
  const int g = lane / G;
  T* mytile = tile + g * nn;
  const int tile_vals = IPW * nn;
  for (int k = lane; k < tile_vals; k += 32) tile[k] = T(0);   // unrelated pairs of bodies (:1106)
  __syncwarp();
  if (valid) {
    TabEntry<T> e;
    tab_load(tab + (gbase + i) * kCmTabStride, e);
    T F[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) F[k] = e.U[k];
    mytile[oi * n + oi] = e.invD;                                               // H[i,i] (:1111)
    int a = i;
    for (int t = 0; t < maxdepth; ++t) {
      const int par = imdl[a].x;
      if (par < 0) break;
      cross3_add(e.r, F + 3, F);                                               // X^T fh (:1117): about the parent's origin
      a = par;
      tab_load(tab + (gbase + a) * kCmTabStride, e);
      const int4 ia = imdl[a];
      const bool pris = PRISM && ((ia.w >> 24) != 0);
      const T h = pris ? dot3s(e.w, F + 3) : dot3s(e.w, F);                    // :1121
      mytile[oi * n + ia.z] = h;
      mytile[ia.z * n + oi] = h;                                               // :1122
    }
  }
  __syncwarp();
  {
    const int count = nknots * nn;
    if ((tile_vals & 1) == 0 && count == tile_vals && (reinterpret_cast<uintptr_t>(dst) & (sizeof(V2) - 1)) == 0) {
      for (int k = lane; k < (tile_vals >> 1); k += 32) __stcs(reinterpret_cast<V2*>(dst) + k, reinterpret_cast<const V2*>(tile)[k]);
    } else {
      for (int k = lane; k < count; k += 32) __stcs(dst + k, tile[k]);
    }
  }
}

// CRBA = false: minv (Minv out).  CRBA = true: the joint-space inertia matrix H (same launch
// geometry and shared-memory layout; phases B and C are replaced by crba_column_phase).
// BPASS = true: the minv_bpass helper (RBDReference.py:630-735) on the same phases 0 and A.  The pass returns the
// reference's arrays in BODY frames: U (n, 6) and Dinv (n; it holds D, :698) leave from phase A (lane = body: a rotation
// of the lane's own U), and lane j then walks from body j to its root once more, producing for every ancestor a the
// entry Minv[a, j] (:700-708) and the column F[a][:, j] as the reference leaves it (:721-723: U Minv[a, j] is added only
// when a has a parent), rotated into a's axes.  The results wait in shared memory and the warp's Minv and F slabs -
// mostly structural zeros: F[a][:, j] = 0 unless j is in subtree(a) - leave in one coalesced pass through a per-CTA
// map, every sector written once (the scheme of grad_fpass_level_kernel).
template <typename T, int G, bool PRISM, bool CRBA = false, bool FB = false, bool BPASS = false>
__global__ void __launch_bounds__(kCmMaxWarps * 32)
minv_coop_kernel(const __grid_constant__ FastModel<T> m, const __grid_constant__ DfsPlan plan,
                 const __grid_constant__ CoopPlan cp, const __grid_constant__ CoopMinvPlan mp, int64_t B,
                 const T* __restrict__ q, T* __restrict__ Minv, T* __restrict__ Fo, T* __restrict__ Uo,
                 T* __restrict__ Do, int npairs) {
  static_assert(!(FB && CRBA), "crba has no floating-base branch here");
  static_assert(!(BPASS && (CRBA || FB)), "the minv_bpass helper is a fixed-base, articulated-inertia mode");
  constexpr int IPW = 32 / G;
  typedef typename Vec2<T>::type V2;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = m.n;                                      // bodies (FB: body 0 is the base, identity transform)
  const int nvv = FB ? n + 5 : n;
  const int nq = FB ? n + 6 : n;
  const int nn = nvv * nvv;
  const int maxdepth = cp.maxdepth;
  const int nwarps = blockDim.x >> 5;
  const int warp_vals = coop_minv_warp_vals<T>(n, G, maxdepth, m.n_slot_a, FB, BPASS ? npairs : 0);
  int4* imdl = reinterpret_cast<int4*>(smem_raw);                            // [n]
  T* mdl = reinterpret_cast<T*>(smem_raw + (size_t)n * sizeof(int4));        // [n][51]
  T* warp_all = mdl + ((n * kCoopMdlStride + 3) & ~3);                       // [nwarps][warp_vals]

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
    // slot_a: this body's G is kept for a later (non-first) child
    imdl[i] = make_int4(p, plan.sub_end[i], plan.orig[i], cm_pack(mp.depth[i], m.slot_a[i], p >= 0 ? m.slot_a[p] : -1, m.kind[i]));
  }
  __syncthreads();
  // BPASS: pair tables behind the warps' regions
  const int stage = coop_minv_bpass_stage(npairs);          // staged values per knot point: [7][npairs]
  int* pfirst = reinterpret_cast<int*>(warp_all + (size_t)nwarps * warp_vals);                 // [n] first pair of column j
  short* pmapM = reinterpret_cast<short*>(pfirst + ((n + 3) & ~3));                            // [IPW][n * n]
  short* pmapF = pmapM + ((IPW * nn + 7) & ~7);                                                // [IPW][n][6][n]
  if (BPASS) {
    for (int k = threadIdx.x; k < IPW * nn; k += blockDim.x) pmapM[k] = (short)-1;
    for (int k = threadIdx.x; k < IPW * 6 * nn; k += blockDim.x) pmapF[k] = (short)-1;
    if (threadIdx.x == 0) {
      int first = 0;
      for (int j = 0; j < n; ++j) { pfirst[j] = first; first += mp.depth[j] + 1; }
    }
    __syncthreads();
    if (threadIdx.x < n) {                                  // pairs of column j: (j, j), (parent(j), j), ...
      const int j = threadIdx.x, oj = imdl[j].z;
      int idx = pfirst[j];
      for (int a = j; a >= 0; a = imdl[a].x) {
        const int oa = imdl[a].z;
        for (int kk = 0; kk < IPW; ++kk) {
          pmapM[kk * nn + oa * n + oj] = (short)(kk * stage + idx);
          for (int r = 0; r < 6; ++r) pmapF[kk * 6 * nn + (oa * 6 + r) * n + oj] = (short)(kk * stage + (1 + r) * npairs + idx);
        }
        ++idx;
      }
    }
    __syncthreads();
  }

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane / G, i = lane - g * G;
  const bool valid = i < n;
  const int ib = valid ? i : 0;
  const int gbase = g * G;
  const T* mb = mdl + ib * kCoopMdlStride;
  const int4 myA = imdl[ib];
  const int par = valid ? myA.x : -1;
  const int kind = PRISM ? (myA.w >> 24) : 0;
  const int sub_end = myA.y;
  const int oi = myA.z;
  const int depth = valid ? (myA.w & 0xff) : -1;
  const int comp_root = mp.comp_root[ib];
  int jump[5];
#pragma unroll
  for (int s = 0; s < 5; ++s) jump[s] = valid ? cp.jump[s][ib] : -1;
  T* tab = warp_all + warp * warp_vals;                   // [32][14]
  T* mbw = tab + 32 * kCmTabStride;                       // [maxdepth+1][32]
  T* big = mbw + (maxdepth + 1) * 32;                     // phase A: [32][22]; phase C: G stashes [slot][6][32] | tile
  T* mytab = tab + lane * kCmTabStride;
  T* dinv = tab + warp_vals - IPW * kCmDinvStride;        // FB: [IPW][22] inverse of the base's articulated inertia
  T* etab = tab + warp_vals - (32 * 12 + IPW * stage);    // BPASS: [32][12] rotation of every body | [IPW][7][npairs] results
  T* stg = etab + 32 * 12;
  const int nsteps = cp.nsteps;

  const int64_t ngroups = (B + IPW - 1) / IPW;
  const int64_t gstride = (int64_t)gridDim.x * nwarps;
  const int qoff = FB ? oi + 6 : oi;                      // (the base's lane reads a value it does not use: E = 1)
  // the joint position of the next knot point is requested one evaluation ahead
  T q_nx = T(0);
  {
    const int64_t grp0 = (int64_t)blockIdx.x * nwarps + warp;
    if (grp0 < ngroups) {
      int64_t b0 = grp0 * IPW + g;
      if (b0 >= B) b0 = B - 1;
      q_nx = q[b0 * nq + qoff];
    }
  }
  for (int64_t grp = (int64_t)blockIdx.x * nwarps + warp; grp < ngroups; grp += gstride) {
    int64_t b = grp * IPW + g;
    if (b >= B) b = B - 1;                                // duplicate work, never stored

    // ------------------------------------------------------------------ phase 0: rotations
    T E[9], rw[3], w[3];
    {
      const T qi = q_nx;
      if (grp + gstride < ngroups) {
        int64_t bn = (grp + gstride) * IPW + g;
        if (bn >= B) bn = B - 1;
        q_nx = q[bn * nq + qoff];
      }
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
    if (BPASS) {
#pragma unroll
      for (int k = 0; k < 9; ++k) etab[lane * 12 + k] = E[k];
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
        for (int c = i + 1; c < sub_end; c = imdl[c].y) {                   // children of i
          const V2* src = reinterpret_cast<const V2*>(big + (gbase + c) * kCmIaStride);
#pragma unroll
          for (int k = 0; k < 11; ++k) { const V2 t = src[k]; IA[2 * k] += t.x; IA[2 * k + 1] += t.y; }
        }
        if (FB && i == 0) {
          // the base: S = eye(6), D = IA_0, fb_Dinv = inv(IA_0) (:681-684)
          T s[21];
          s[sym6_idx(0, 0)] = IA[0]; s[sym6_idx(0, 1)] = IA[1]; s[sym6_idx(0, 2)] = IA[2];
          s[sym6_idx(1, 1)] = IA[3]; s[sym6_idx(1, 2)] = IA[4]; s[sym6_idx(2, 2)] = IA[5];
#pragma unroll
          for (int rr = 0; rr < 3; ++rr)
#pragma unroll
            for (int cc = 0; cc < 3; ++cc) s[sym6_idx(rr, 3 + cc)] = IA[6 + 3 * rr + cc];
          s[sym6_idx(3, 3)] = IA[15]; s[sym6_idx(3, 4)] = IA[16]; s[sym6_idx(3, 5)] = IA[17];
          s[sym6_idx(4, 4)] = IA[18]; s[sym6_idx(4, 5)] = IA[19]; s[sym6_idx(5, 5)] = IA[20];
          sym6_invert(s);
          T* dg = dinv + g * kCmDinvStride;
#pragma unroll
          for (int k = 0; k < 21; ++k) dg[k] = s[k];
        } else {
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
        if (BPASS && valid && (grp * IPW + g) < B) {
          // U (B, n, 6) in body i's axes (same origin: a rotation), Dinv (B, n) = D (:697-698)
          T* ub = Uo + (b * n + oi) * (int64_t)6;
#pragma unroll
          for (int rr = 0; rr < 3; ++rr) {
            ub[rr] = E[3 * rr] * U[0] + E[3 * rr + 1] * U[1] + E[3 * rr + 2] * U[2];
            ub[3 + rr] = E[3 * rr] * U[3] + E[3 * rr + 1] * U[4] + E[3 * rr + 2] * U[5];
          }
          Do[b * n + oi] = D;
        }
        {
          V2* dst = reinterpret_cast<V2*>(mytab);
          V2 t;
          t.x = w[0]; t.y = w[1]; dst[0] = t;
          t.x = w[2]; t.y = CRBA ? D : invD; dst[1] = t;
          t.x = U[0]; t.y = U[1]; dst[2] = t;
          t.x = U[2]; t.y = U[3]; dst[3] = t;
          t.x = U[4]; t.y = U[5]; dst[4] = t;
          t.x = rw[0]; t.y = rw[1]; dst[5] = t;
          t.x = rw[2]; t.y = T(0); dst[6] = t;
        }
        if (par >= 0) {
          // IA -= U U^T / D                                                   (:728-731)
          if (!CRBA) {
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
          }
          // translate to the parent's origin (:732-733 / :1100-1103 with X = [[1,0],[-r x,1]]):
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
      }
      __syncwarp();
    }

    if (BPASS) {
      // ---------------------------------------------------------------- lane j: column j up to its root
      if (valid) {
        T F[6] = {T(0), T(0), T(0), T(0), T(0), T(0)};
        T* sj = stg + g * stage + pfirst[i];
        int a = i;
        for (int t = 0; t <= maxdepth; ++t) {
          if (a >= 0) {
            TabEntry<T> e;
            tab_load(tab + (gbase + a) * kCmTabStride, e);
            const int4 ia = imdl[a];
            const bool pris = PRISM && ((ia.w >> 24) != 0);
            const T sF = pris ? dot3s(e.w, F + 3) : dot3s(e.w, F);
            const T mij = (a == i ? e.invD : T(0)) - e.invD * sF;              // :700-708
            if (ia.x >= 0) {
#pragma unroll
              for (int k = 0; k < 6; ++k) F[k] = fma_t(e.U[k], mij, F[k]);     // :721-723 (only with a parent)
            }
            const T* Ea = etab + (gbase + a) * 12;
            sj[t] = mij;
#pragma unroll
            for (int rr = 0; rr < 3; ++rr) {
              sj[(1 + rr) * npairs + t] = Ea[3 * rr] * F[0] + Ea[3 * rr + 1] * F[1] + Ea[3 * rr + 2] * F[2];
              sj[(4 + rr) * npairs + t] = Ea[3 * rr] * F[3] + Ea[3 * rr + 1] * F[4] + Ea[3 * rr + 2] * F[5];
            }
            cross3_add(e.r, F + 3, F);                                         // moment about the parent's origin (:724-726)
            a = ia.x;
          }
        }
      }
      __syncwarp();
      // ---------------------------------------------------------------- the warp's Minv and F slabs, every sector once
      const int64_t first = grp * IPW;
      const int nk = (int)((B - first) < IPW ? (B - first) : IPW);
#pragma unroll 1
      for (int w2 = 0; w2 < 2; ++w2) {
        const int per = w2 == 0 ? nn : 6 * nn;
        T* out = (w2 == 0 ? Minv : Fo) + first * per;
        const short* pm = w2 == 0 ? pmapM : pmapF;
        const int total = nk * per;
        if ((total & 1) == 0 && (reinterpret_cast<uintptr_t>(out) & (sizeof(V2) - 1)) == 0) {
          for (int f2 = lane; f2 < (total >> 1); f2 += 32) {
            const short2 pp = reinterpret_cast<const short2*>(pm)[f2];
            V2 x;
            x.x = pp.x >= 0 ? stg[pp.x] : T(0);
            x.y = pp.y >= 0 ? stg[pp.y] : T(0);
            __stcs(reinterpret_cast<V2*>(out) + f2, x);
          }
        } else {
          for (int f = lane; f < total; f += 32) {
            const int pp = pm[f];
            __stcs(out + f, pp >= 0 ? stg[pp] : T(0));
          }
        }
      }
    } else if (CRBA)
      crba_column_phase<T, G, PRISM>(n, maxdepth, valid, i, gbase, lane, oi, imdl, tab, big + 6 * m.n_slot_a * 32,
                                     Minv + grp * IPW * (int64_t)nn, (int)((B - grp * IPW) < IPW ? (B - grp * IPW) : IPW));
    else if (FB)   // every entry of the matrix is written (the base couples all bodies): no zero fill
      minv_column_phases<T, G, PRISM, false, true>(n, maxdepth, mp.maxcomp, m.n_slot_a, valid, i, gbase, lane, oi, comp_root, imdl, tab,
                                                   mbw, big, Minv + grp * IPW * (int64_t)nn,
                                                   (int)((B - grp * IPW) < IPW ? (B - grp * IPW) : IPW), dinv);
    else
      minv_column_phases<T, G, PRISM>(n, maxdepth, mp.maxcomp, m.n_slot_a, valid, i, gbase, lane, oi, comp_root, imdl, tab, mbw, big,
                                      Minv + grp * IPW * (int64_t)nn, (int)((B - grp * IPW) < IPW ? (B - grp * IPW) : IPW));
    __syncwarp();
  }
}

// =============================================================================================
// Hybrid mapping: phase A does not parallelise over the bodies of one knot point (it is a serial
// recursion along every root path), so a warp takes 32 knot points at a time and runs
//   stage 1  ONE KNOT POINT PER LANE: rotation sweep root -> leaf, articulated-inertia sweep
//            leaf -> root (the parent's rotation is re-derived from the child's, only branch points
//            and chain ends are stashed), writing each body's (w, 1/D, U, r) to a warp-private
//            scratch table in global memory (written and re-read within microseconds: L2 traffic);
//   stage 2  ONE COLUMN PER LANE (minv_column_phases), IPW knot points per pass, reading the table
//            back with coalesced loads.
// Every lane is busy in both stages; the model is read from the constant bank in stage 1
// (warp-uniform body index) and from a shared-memory index table in stage 2.
constexpr int kHyScrStride = 16;     // w(3) invD | U(6) | r(3) pad | f1 f2   (eight aligned pairs)

template <typename T>
__host__ __device__ inline int hybrid_minv_warp_vals(int n, int G, int maxdepth, int nslot_a, int nslot_b) {
  const int ipw = 32 / G;
  const int s2 = 32 * kCmTabStride + (maxdepth + 1) * 32 + 6 * nslot_a * 32 + ((ipw * n * n + 3) & ~3);
  const int s1 = (22 * nslot_a + 9 * nslot_b) * 32;
  return s1 > s2 ? s1 : s2;
}
template <typename T>
__host__ __device__ inline size_t hybrid_minv_smem_bytes(int n, int G, int maxdepth, int nslot_a, int nslot_b, int warps) {
  return (size_t)warps * hybrid_minv_warp_vals<T>(n, G, maxdepth, nslot_a, nslot_b) * sizeof(T) + (size_t)n * 4 * sizeof(int);
}

template <typename T, int G, bool PRISM>
__global__ void __launch_bounds__(kCmMaxWarps * 32)
minv_hybrid_kernel(const __grid_constant__ FastModel<T> m, const __grid_constant__ DfsPlan plan,
                   const __grid_constant__ CoopMinvPlan mp, int maxdepth, int64_t B, const T* __restrict__ q,
                   T* __restrict__ Minv, T* __restrict__ scratch) {
  constexpr int IPW = 32 / G;
  typedef typename Vec2<T>::type V2;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = m.n;
  const int nn = n * n;
  const int nwarps = blockDim.x >> 5;
  const int warp_vals = hybrid_minv_warp_vals<T>(n, G, maxdepth, m.n_slot_a, m.n_slot_b);
  int4* imdl = reinterpret_cast<int4*>(smem_raw);                            // [n]
  T* warp_all = reinterpret_cast<T*>(smem_raw + (size_t)n * sizeof(int4));
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int p = m.parent[i];
    imdl[i] = make_int4(p, plan.sub_end[i], plan.orig[i], cm_pack(mp.depth[i], m.slot_a[i], p >= 0 ? m.slot_a[p] : -1, m.kind[i]));
  }
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane / G, i2 = lane - g * G;
  const bool valid = i2 < n;
  const int ib = valid ? i2 : 0;
  const int gbase = g * G;
  const int oi = imdl[ib].z;
  const int comp_root = mp.comp_root[ib];
  T* ws = warp_all + warp * warp_vals;
  // stage-2 views
  T* tab = ws;                                            // [32][14]
  T* mbw = tab + 32 * kCmTabStride;                       // [maxdepth+1][32]
  T* big = mbw + (maxdepth + 1) * 32;                     // G stashes [slot][6][32] | tile
  // stage-1 views (same memory, different time)
  T* sta = ws;                                            // [slot_a][22][32]
  T* stb = sta + m.n_slot_a * 22 * 32;                    // [slot_b][9][32]
#define HSTA(s, k) sta[((s) * 22 + (k)) * 32 + lane]
#define HSTB(s, k) stb[((s) * 9 + (k)) * 32 + lane]
  T* myscr = scratch + (size_t)(blockIdx.x * nwarps + warp) * 32 * n * kHyScrStride;   // [32][n][16]
  T* scr = myscr + (size_t)lane * n * kHyScrStride;

  const int64_t ntasks = (B + 31) / 32;
  for (int64_t task = (int64_t)blockIdx.x * nwarps + warp; task < ntasks; task += (int64_t)gridDim.x * nwarps) {
    // ================================================================ stage 1: lane = knot point
    {
      int64_t b = task * 32 + lane;
      if (b >= B) b = B - 1;                              // duplicate work, never stored
      const T* qb = q + b * n;
      T E[9];
      // ---- rotations, root -> leaf; q is fetched four bodies ahead (the loads of a lane are 8 bytes
      // out of every n*8, so their latency has to be covered by independent work)
      T qpre[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) qpre[u] = qb[plan.orig[u < n ? u : 0]];
      if (lane < 2) {
        // next task's slab of q -> L2
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
          if (i + 4 < n) qpre[u] = qb[plan.orig[i + 4]];  // in flight while the next four bodies are processed
          if (!PRISM || m.kind[i] == 0) sincos_t(qi, &f2, &f1);
          else { f1 = qi; f2 = T(0); }
        }
        {
          V2 t; t.x = f1; t.y = f2;
          __stcg(reinterpret_cast<V2*>(scr + i * kHyScrStride + 14), t);
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
      // ---- articulated inertias, leaf -> root
      for (int s = 0; s < m.n_slot_a; ++s)
#pragma unroll
        for (int k = 0; k < 22; ++k) HSTA(s, k) = T(0);
      T IA[21];
      V2 ffnext = __ldcg(reinterpret_cast<const V2*>(scr + (n - 1) * kHyScrStride + 14));
      V2 ffnext2 = __ldcg(reinterpret_cast<const V2*>(scr + (n > 1 ? n - 2 : 0) * kHyScrStride + 14));
#pragma unroll 1
      for (int i = n - 1; i >= 0; --i) {
        const bool chained = (i != n - 1) && (m.parent[i + 1] == i);
        if (!chained && i != n - 1) {
          const int s = m.slot_b[i];
#pragma unroll
          for (int k = 0; k < 9; ++k) E[k] = HSTB(s, k);
        }
        const V2 ff = ffnext;
        ffnext = ffnext2;                                 // (f1, f2) arrive two bodies ahead
        if (i > 1) ffnext2 = __ldcg(reinterpret_cast<const V2*>(scr + (i - 2) * kHyScrStride + 14));
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
          for (int k = 0; k < 21; ++k) IA[k] += HSTA(sa, k);
        }
        T w[3], rw[3], Ej[9];
        {
          const T f1 = ff.x, f2 = ff.y;
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
          V2* dst = reinterpret_cast<V2*>(scr + i * kHyScrStride);
          V2 t;
          t.x = w[0]; t.y = w[1]; __stcg(dst + 0, t);
          t.x = w[2]; t.y = invD; __stcg(dst + 1, t);
          t.x = U[0]; t.y = U[1]; __stcg(dst + 2, t);
          t.x = U[2]; t.y = U[3]; __stcg(dst + 3, t);
          t.x = U[4]; t.y = U[5]; __stcg(dst + 4, t);
          t.x = rw[0]; t.y = rw[1]; __stcg(dst + 5, t);
          t.x = rw[2]; t.y = T(0); __stcg(dst + 6, t);
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
    // ================================================================ stage 2: lane = column
    {
      V2 pre[7];                                          // table rows of the next pass, in flight
      const V2* src0 = reinterpret_cast<const V2*>(myscr + ((size_t)g * n + ib) * kHyScrStride);
      const size_t pass_stride = (size_t)IPW * n * kHyScrStride / 2;
#pragma unroll
      for (int k = 0; k < 7; ++k) pre[k] = __ldcg(src0 + k);
      {
        // stage 1 used this memory for its stashes: clear the output tile once per task
        T* tile = big + 6 * m.n_slot_a * 32;
        const int tile_vals = IPW * nn;
        for (int k = lane; k < tile_vals; k += 32) tile[k] = T(0);
      }
      for (int pass = 0; pass < G; ++pass) {
        const int64_t first = task * 32 + pass * IPW;
        if (first >= B) break;
        {
          V2* dst = reinterpret_cast<V2*>(tab + lane * kCmTabStride);
#pragma unroll
          for (int k = 0; k < 7; ++k) dst[k] = pre[k];
        }
        __syncwarp();
        if (pass + 1 < G) {
          const V2* src = src0 + (size_t)(pass + 1) * pass_stride;
#pragma unroll
          for (int k = 0; k < 7; ++k) pre[k] = __ldcg(src + k);
        }
        minv_column_phases<T, G, PRISM, false>(n, maxdepth, mp.maxcomp, m.n_slot_a, valid, i2, gbase, lane, oi, comp_root, imdl,
                                               tab, mbw, big, Minv + first * (int64_t)nn,
                                               (int)((B - first) < IPW ? (B - first) : IPW));
        __syncwarp();
      }
    }
  }
#undef HSTA
#undef HSTB
}

}  // namespace rbd
