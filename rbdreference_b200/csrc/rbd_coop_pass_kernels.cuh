// rbd_coop_pass_kernels.cuh - the four gradient passes (RBDReference.py:1127-1343) with one
// DERIVATIVE COLUMN PER LANE: lane (g, c) of a warp owns column c of knot point g (G = 8 / 16 / 32
// lanes per knot point, 32 / G knot points per warp pass).
//
// The reference's recursions never mix columns: dv/da/df[:, c, i] only depend on [:, c, parent(i)]
// (forward passes) and df[:, c, parent] only receives X^T df[:, c, i] (backward passes).  So every
// lane runs the body loop on its own column, and the (6, n, NB) tensors of the warp's knot points
// live in a shared-memory tile that has the layout of the warp's contiguous slab in HBM (rows
// padded to an odd pitch): they are read / written with coalesced full-line accesses instead of
// the 8-byte accesses 6 n^2 * 8 bytes apart of a knot-point-per-thread mapping.
// Same arithmetic as the generic kernels in rbd_pass_kernels.cuh (body frame, dense 6x6 inertia,
// any S), so any model the reference accepts is served.
#pragma once
#include "rbd_common.cuh"

namespace rbd {

constexpr int kCpMaxWarps = 8;

__host__ __device__ inline int cp_pitch(int n) { return n | 1; }
// values of T per warp
__host__ __device__ inline int cp_fpass_warp_vals(int n, int G) {
  const int ipw = 32 / G;
  return (ipw * (3 * 6 * n * cp_pitch(n) + 12 * n + 3 * n) + 3) & ~3;      // dv da df tiles | v a rows | f1 f2 qd
}
__host__ __device__ inline int cp_bpass_warp_vals(int n, int G) {
  const int ipw = 32 / G;
  // df tile | dc tile | f rows | f1 f2 | mbarrier of the bulk load (last two values), rounded to 16 / 32 bytes
  return (ipw * (6 * n * cp_pitch(n) + n * n + 6 * n + 2 * n) + 2 + 3) & ~3;
}

__device__ __forceinline__ void cp_async_val(double* dst, const double* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_val(float* dst, const float* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}

// Tiles of an odd n are dense (pitch = n): they have the exact layout of the warp's slab in HBM and move with one
// cp.async.bulk per tensor (warp_bulk_store / warp_bulk_load, rbd_common.cuh); padded tiles (even n) and slabs the
// instruction cannot take (16-byte rule) go through cp_copy:
// copy `count` values of a [rows][n] slab between global memory and a tile of pitch cp_pitch(n).
// Global -> shared uses cp.async: all of a lane's copies are in flight at once (a load + store per
// iteration keeps one load per lane in flight and is latency-bound); the caller's __syncwarp follows
// the wait.
template <typename T, bool TO_SMEM>
__device__ __forceinline__ void cp_copy(T* tile, T* __restrict__ gmem, int n, int count, int total, int lane) {
  const int np = cp_pitch(n);
  int row = lane / n, col = lane - row * n;
  const int drow = 32 / n, dcol = 32 - drow * n;
  for (int e = lane; e < total; e += 32) {
    if (TO_SMEM) {
      if (e < count) cp_async_val(tile + row * np + col, gmem + e);
      else tile[row * np + col] = T(0);
    } else if (e < count) {
      __stcs(gmem + e, tile[row * np + col]);
    }
    row += drow; col += dcol;
    if (col >= n) { col -= n; row += 1; }
  }
  if (TO_SMEM) asm volatile("cp.async.wait_all;" ::: "memory");
}

// ---- rnea_grad_fpass_dq / _dqd (:1127-1187, :1189-1255) --------------------------------------
template <typename T, int G, bool DQ>
__global__ void __launch_bounds__(kCpMaxWarps * 32)
grad_fpass_coop_kernel(const __grid_constant__ DevModel<T> m, int64_t B, const T* __restrict__ q,
                       const T* __restrict__ qd, const T* __restrict__ v, const T* __restrict__ a, T gravity,
                       T* __restrict__ dv, T* __restrict__ da, T* __restrict__ df) {
  constexpr int IPW = 32 / G;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = m.n, np = cp_pitch(n);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int g = lane / G, c = lane - g * G;
  const bool valid = c < n;
  const int tvals = 6 * n * np;                           // one knot point's padded tile
  T* ws = reinterpret_cast<T*>(smem_raw) + (size_t)warp * cp_fpass_warp_vals(n, G);
  T* tv = ws;                                             // [IPW][6 n][np]
  T* ta = tv + IPW * tvals;
  T* tf = ta + IPW * tvals;
  T* sv = tf + IPW * tvals;                               // [IPW][6][n]
  T* sa = sv + IPW * 6 * n;
  T* sj = sa + IPW * 6 * n;                               // [IPW][n][3]: f1 f2 qd
  const int64_t slab = (int64_t)6 * n * n;
  const int64_t ngroups = (B + IPW - 1) / IPW;
  for (int64_t grp = (int64_t)blockIdx.x * nwarps + warp; grp < ngroups; grp += (int64_t)gridDim.x * nwarps) {
    const int64_t first = grp * IPW;
    const int nk = (int)((B - first) < IPW ? (B - first) : IPW);
    if (valid) {
      int64_t b = first + g;
      if (b >= B) b = B - 1;
      T f1, f2;
      joint_basis(m, c, q[b * n + c], f1, f2);
      T* d = sj + (g * n + c) * 3;
      d[0] = f1; d[1] = f2; d[2] = qd[b * n + c];
    }
    for (int e = lane; e < IPW * 6 * n; e += 32) {
      const bool ok = e < nk * 6 * n;
      sv[e] = ok ? v[first * 6 * n + e] : T(0);
      if (DQ) sa[e] = ok ? a[first * 6 * n + e] : T(0);
    }
    warp_bulk_store_wait(lane);                           // the previous pass's tiles have left (includes __syncwarp)
    if (valid) {
      T* mv = tv + g * tvals + c * np;                    // element (r, c, i) at (r n + c) np + i
      T* ma = ta + g * tvals + c * np;
      T* mf = tf + g * tvals + c * np;
      const int rstride = n * np;
      const T* vb = sv + g * 6 * n;
      const T* ab = sa + g * 6 * n;
      for (int i = 0; i < n; ++i) {
        const T* ji = sj + (g * n + i) * 3;
        T X[18];
        build_X(m, i, ji[0], ji[1], X);
        const T qdi = ji[2];
        const int p = m.parent[i];
        T S[6], vi[6], Iv[6];
#pragma unroll
        for (int r = 0; r < 6; ++r) { S[r] = m.S[i][r]; vi[r] = vb[r * n + i]; }
        mat6_apply(m.I[i], vi, Iv);                                            // :1180 / :1248
        T dvc[6], dac[6];
        if (p >= 0) {
          T pv[6], pa[6];
#pragma unroll
          for (int r = 0; r < 6; ++r) { pv[r] = mv[r * rstride + p]; pa[r] = ma[r * rstride + p]; }
          X_apply(X, pv, dvc);                                                 // :1158 / :1230
          X_apply(X, pa, dac);                                                 // :1163 / :1234
        } else {
#pragma unroll
          for (int r = 0; r < 6; ++r) { dvc[r] = T(0); dac[r] = T(0); }
        }
        T seed_a[6] = {T(0), T(0), T(0), T(0), T(0), T(0)};
        if (c == i) {
          // seed terms of the column of the body's own joint
          if (DQ) {
            T par[6], t[6], seed_v[6];
            if (p >= 0) {
#pragma unroll
              for (int r = 0; r < 6; ++r) par[r] = vb[r * n + p];
              X_apply(X, par, t);
              crm_mul(t, S, seed_v);                                           // :1159
#pragma unroll
              for (int r = 0; r < 6; ++r) { dvc[r] += seed_v[r]; par[r] = ab[r * n + p]; }
            } else {
#pragma unroll
              for (int r = 0; r < 6; ++r) par[r] = T(0);
              par[5] = -gravity;                                               // :1137
            }
            X_apply(X, par, t);
            crm_mul(t, S, seed_a);                                             // :1173 / :1175
          } else {
#pragma unroll
            for (int r = 0; r < 6; ++r) dvc[r] += S[r];                        // :1231
            crm_mul(vi, S, seed_a);                                            // :1243
          }
        }
        T t[6];
        crm_mul(dvc, S, t);                                                    // :1170 / :1240
#pragma unroll
        for (int r = 0; r < 6; ++r) dac[r] = fma_t(qdi, t[r], dac[r]) + seed_a[r];
        T Ida[6], Idv[6], t1[6], t2[6];
        mat6_apply(m.I[i], dac, Ida);                                          // :1179 / :1247
        mat6_apply(m.I[i], dvc, Idv);
        crf_mul(dvc, Iv, t1);                                                  // :1184 / :1251
        crf_mul(vi, Idv, t2);                                                  // :1185 / :1252
#pragma unroll
        for (int r = 0; r < 6; ++r) {
          mv[r * rstride + i] = dvc[r];
          ma[r * rstride + i] = dac[r];
          mf[r * rstride + i] = Ida[r] + t1[r] + t2[r];
        }
      }
    }
    __syncwarp();
    {
      const int cnt = nk * 6 * n * n;
      if (!(np == n && warp_bulk_store(dv + first * slab, tv, cnt, lane))) cp_copy<T, false>(tv, dv + first * slab, n, cnt, IPW * 6 * n * n, lane);
      if (!(np == n && warp_bulk_store(da + first * slab, ta, cnt, lane))) cp_copy<T, false>(ta, da + first * slab, n, cnt, IPW * 6 * n * n, lane);
      if (!(np == n && warp_bulk_store(df + first * slab, tf, cnt, lane))) cp_copy<T, false>(tf, df + first * slab, n, cnt, IPW * 6 * n * n, lane);
    }
    __syncwarp();
  }
  warp_bulk_store_wait(lane);                             // shared memory must outlive the copies
}

// ---- rnea_grad_fpass_dq / _dqd: one BODY per lane, one ancestor distance per round -------------------------------
// The three (6, n, NB) tensors of a knot point are mostly structural zeros: column c of body i is non-zero only when c
// is i or one of its ancestors (150 of Atlas' 900 pairs, 28 of iiwa14's 49).  A group of G = 8 / 16 / 32 lanes owns one
// knot point (32 / G knot points per warp):
//   * lane i of the group is body i.  In round d the lane works on the pair (c, i) with c = the ancestor of i at
//     distance d: the recursion dv[c, i] = X_i dv[c, parent(i)] (:1158-1175 / :1230-1243) needs the pair (c, parent(i)),
//     which is what the parent's lane produced in round d - 1 - twelve values by warp shuffle.  max depth + 1 rounds
//     (10 for Atlas) instead of n steps over all n columns, and no lane works on a zero;
//   * the 18 results of a pair wait in shared memory ([tensor row][pair]; 22 KB per knot point for Atlas, 4 KB for iiwa14);
//   * the warp's slabs (contiguous in HBM: consecutive knot points) then leave in ONE coalesced pass of 16-byte stores:
//     a per-CTA table maps every (c, i) of a tensor row to its pair or to "structural zero".  Every sector is written
//     once (DRAM traffic = the tensors' size; a first version that zero-filled the slabs and scattered the pairs
//     afterwards wrote 1.4x: with 130 KB per warp in flight the zeroed lines had left the L2 before their pairs arrived).
// Shared memory: the per-body constants (one copy per CTA, odd stride: lanes read different bodies), the pair map, and per
// warp the staged v / a rows and the pair results.
constexpr int kCpLvlMdl = 97;
__host__ __device__ inline int cp_level_warp_vals(int n, int npairs, int G) { return (32 / G) * ((12 * n + 18 * npairs + 3) & ~3); }   // per knot: v a rows | results
__host__ __device__ inline size_t cp_level_head_bytes(int n, size_t tsize, int G = 32) {
  return (((size_t)n * 4 * sizeof(int) + (size_t)(32 / G) * 6 * n * n * sizeof(short) + 15) & ~(size_t)15) + (((size_t)n * kCpLvlMdl + 3) & ~(size_t)3) * tsize;
}

constexpr int kCpLvlMaxWarps = 8;
template <typename T, int G, bool DQ>
__global__ void __launch_bounds__(kCpLvlMaxWarps * 32)
grad_fpass_level_kernel(const __grid_constant__ DevModel<T> m, int npairs, int64_t B, const T* __restrict__ q,
                        const T* __restrict__ qd, const T* __restrict__ v, const T* __restrict__ a, T gravity,
                        T* __restrict__ dv, T* __restrict__ da, T* __restrict__ df) {
  constexpr int IPW = 32 / G;
  typedef typename Vec2<T>::type V2;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = m.n;
  const int nn = n * n;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int g = lane / G, i = lane - g * G;
  const int gbase = g * G;
  const bool valid = i < n;
  int* topo = reinterpret_cast<int*>(smem_raw);                            // [n][4]: parent kind depth first-pair
  short* pmap = reinterpret_cast<short*>(topo + 4 * n);                     // [IPW][6][n * n]: where value f of the warp's slab run waits (offset from the
                                                                            // first knot point's results), or -1 = structural zero
  T* mdl = reinterpret_cast<T*>(smem_raw + cp_level_head_bytes(n, 0, G));  // [n][97]: XA XB XC S I
  const int knot_vals = (12 * n + 18 * npairs + 3) & ~3;
  T* ws = mdl + (((size_t)n * kCpLvlMdl + 3) & ~(size_t)3) + (size_t)warp * IPW * knot_vals;
  T* sv = ws + g * knot_vals;                              // [6][n] of this lane's knot point
  T* sa = sv + 6 * n;
  T* res = sa + 6 * n;                                     // [3][6][npairs]
  for (int k = threadIdx.x; k < n * kCpLvlMdl; k += blockDim.x) {
    const int b = k / kCpLvlMdl, w = k - b * kCpLvlMdl;
    mdl[k] = w < 18 ? m.XA[b][w] : w < 36 ? m.XB[b][w - 18] : w < 54 ? m.XC[b][w - 36] : w < 60 ? m.S[b][w - 54] : w < 96 ? m.I[b][w - 60] : T(0);
  }
  for (int k = threadIdx.x; k < IPW * 6 * nn; k += blockDim.x) pmap[k] = (short)-1;
  int maxdepth = 0;
  {
    int first = 0;
    for (int b = 0; b < n; ++b) {                          // (every thread walks the same entries of the constant bank)
      int d = 0;
      for (int p = m.parent[b]; p >= 0; p = m.parent[p]) ++d;
      if (threadIdx.x == 0) { topo[4 * b] = m.parent[b]; topo[4 * b + 1] = m.kind[b]; topo[4 * b + 2] = d; topo[4 * b + 3] = first; }
      first += d + 1;
      maxdepth = d > maxdepth ? d : maxdepth;
    }
  }
  __syncthreads();
  if (threadIdx.x < n) {                                   // pairs of body b: (b, b), (parent(b), b), ...
    const int b = threadIdx.x;
    int idx = topo[4 * b + 3];
    for (int c = b; c >= 0; c = topo[4 * c]) {
      for (int kk = 0; kk < IPW; ++kk)
        for (int r = 0; r < 6; ++r) pmap[(kk * 6 + r) * nn + c * n + b] = (short)(kk * knot_vals + r * npairs + idx);
      ++idx;
    }
  }
  __syncthreads();
  const int ib = valid ? i : 0;
  const int par = topo[4 * ib], kind = topo[4 * ib + 1];
  const int depth = valid ? topo[4 * ib + 2] : -1;
  const int pair0 = topo[4 * ib + 3];
  const T* mc = mdl + ib * kCpLvlMdl;
  // lane = body for the whole kernel: its motion subspace and spatial inertia stay in registers (read from shared
  // memory in every round, the 72 loads of the two inertia products saturated the memory-instruction queue)
  T S[6], Im[36];
#pragma unroll
  for (int r = 0; r < 6; ++r) S[r] = mc[54 + r];
#pragma unroll
  for (int k = 0; k < 36; ++k) Im[k] = mc[60 + k];
  const int slab = 6 * nn;                                  // values of one tensor of one knot point (even)
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(dv) | reinterpret_cast<uintptr_t>(da) | reinterpret_cast<uintptr_t>(df)) & (2 * sizeof(T) - 1)) == 0;
  const int64_t ngroups = (B + IPW - 1) / IPW;
  for (int64_t grp = (int64_t)blockIdx.x * nwarps + warp; grp < ngroups; grp += (int64_t)gridDim.x * nwarps) {
    const int64_t first = grp * IPW;
    const int nk = (int)((B - first) < IPW ? (B - first) : IPW);
    int64_t b = first + g;
    if (b >= B) b = B - 1;                                  // duplicate work, never stored
    // ---- inputs of the knot point (those of the warp's next knot points are requested into L2 meanwhile: the
    //      loads below were 28 % of the stall samples)
    int64_t bn = first + (int64_t)gridDim.x * nwarps * IPW + g;
    const bool more = bn < B;
    for (int e = i; e < 6 * n; e += G) {
      sv[e] = v[b * 6 * n + e];
      if (DQ) sa[e] = a[b * 6 * n + e];
      if (more) {
        prefetch_l2(v + bn * 6 * n + e);
        if (DQ) prefetch_l2(a + bn * 6 * n + e);
      }
    }
    T X[18], vi[6], Iv[6], qdi = T(0);
    {
      T f1 = T(0), f2 = T(0);
      if (valid) {
        const T qi = q[b * n + i];
        qdi = qd[b * n + i];
        if (more) { prefetch_l2(q + bn * n + i); prefetch_l2(qd + bn * n + i); }
        if (kind == 0) sincos_t(qi, &f2, &f1);
        else f1 = qi;
      }
#pragma unroll
      for (int k = 0; k < 18; ++k) X[k] = fma_t(mc[36 + k], f2, fma_t(mc[18 + k], f1, mc[k]));
    }
    __syncwarp();                                           // staged rows are in place; the previous slabs have been read
#pragma unroll
    for (int r = 0; r < 6; ++r) vi[r] = sv[r * n + ib];
    mat6_apply(Im, vi, Iv);                                                      // :1180 / :1248
    T cdv[6], cda[6];                                       // dv / da of this lane's pair of the previous round
#pragma unroll
    for (int r = 0; r < 6; ++r) { cdv[r] = T(0); cda[r] = T(0); }
#pragma unroll 1
    for (int d = 0; d <= maxdepth; ++d) {
      T pv[6], pa[6];
      const int src = gbase + (par >= 0 ? par : 0);
#pragma unroll
      for (int r = 0; r < 6; ++r) { pv[r] = __shfl_sync(0xffffffffu, cdv[r], src); pa[r] = __shfl_sync(0xffffffffu, cda[r], src); }
      if (depth >= d) {
        T dvc[6], dac[6], t[6];
        if (d == 0) {
          T seed_a[6];
          if (DQ) {
            T pr[6], xp[6];
            if (par >= 0) {
#pragma unroll
              for (int r = 0; r < 6; ++r) pr[r] = sv[r * n + par];
              X_apply(X, pr, xp);
              crm_mul(xp, S, dvc);                                               // :1159
#pragma unroll
              for (int r = 0; r < 6; ++r) pr[r] = sa[r * n + par];
            } else {
#pragma unroll
              for (int r = 0; r < 6; ++r) { dvc[r] = T(0); pr[r] = T(0); }
              pr[5] = -gravity;                                                  // :1137
            }
            X_apply(X, pr, xp);
            crm_mul(xp, S, seed_a);                                              // :1173 / :1175
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
        T Ida[6], Idv[6], t1[6], t2[6];
        mat6_apply(Im, dac, Ida);                                                // :1179 / :1247
        mat6_apply(Im, dvc, Idv);
        crf_mul(dvc, Iv, t1);                                                    // :1184 / :1251
        crf_mul(vi, Idv, t2);                                                    // :1185 / :1252
        T* rp = res + pair0 + d;
#pragma unroll
        for (int r = 0; r < 6; ++r) {
          rp[r * npairs] = dvc[r];
          rp[(6 + r) * npairs] = dac[r];
          rp[(12 + r) * npairs] = Ida[r] + t1[r] + t2[r];
          cdv[r] = dvc[r];
          cda[r] = dac[r];
        }
      }
    }
    __syncwarp();
    // ---- the warp's slabs of the three tensors (contiguous: consecutive knot points), every sector once: values
    //      2 f2, 2 f2 + 1 of the run come from the map (one 32-bit load for the two entries), zero where it says -1
    const int total = nk * slab;
#pragma unroll 1
    for (int w = 0; w < 3; ++w) {
      T* out = (w == 0 ? dv : (w == 1 ? da : df)) + first * slab;
      const T* rk = ws + 12 * n + (size_t)w * 6 * npairs;   // tensor w of the warp's first knot point
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

// ---- rnea_grad_bpass_dq / _dqd (:1257-1297, :1299-1343) --------------------------------------
// df arrives from the caller (arbitrary contents), is accumulated IN PLACE and written back.
template <typename T, int G, bool DQ>
__global__ void __launch_bounds__(kCpMaxWarps * 32)
grad_bpass_coop_kernel(const __grid_constant__ DevModel<T> m, int64_t B, const T* __restrict__ q,
                       const T* __restrict__ f, T* __restrict__ df, int use_damping, T* __restrict__ dc) {
  constexpr int IPW = 32 / G;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = m.n, np = cp_pitch(n), nn = n * n;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int g = lane / G, c = lane - g * G;
  const bool valid = c < n;
  const int tvals = 6 * n * np;
  T* ws = reinterpret_cast<T*>(smem_raw) + (size_t)warp * cp_bpass_warp_vals(n, G);
  T* td = ws;                                             // [IPW][6 n][np]
  T* tc = td + IPW * tvals;                               // [IPW][n][n]
  T* sf = tc + IPW * nn;                                  // [IPW][6][n]
  T* sj = sf + IPW * 6 * n;                               // [IPW][n][2]
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(ws + cp_bpass_warp_vals(n, G) - 2);
  unsigned phase = 0;
  warp_bulk_bar_init(bar, lane);
  const int64_t slab = (int64_t)6 * n * n;
  const int64_t ngroups = (B + IPW - 1) / IPW;
  for (int64_t grp = (int64_t)blockIdx.x * nwarps + warp; grp < ngroups; grp += (int64_t)gridDim.x * nwarps) {
    const int64_t first = grp * IPW;
    const int nk = (int)((B - first) < IPW ? (B - first) : IPW);
    warp_bulk_store_wait(lane);                           // the previous pass's tiles have left
    const bool bulk_in = np == n && warp_bulk_load(td, df + first * slab, nk * 6 * n * n, bar, lane);
    if (valid) {
      int64_t b = first + g;
      if (b >= B) b = B - 1;
      T f1, f2;
      joint_basis(m, c, q[b * n + c], f1, f2);
      sj[(g * n + c) * 2] = f1; sj[(g * n + c) * 2 + 1] = f2;
    }
    if (DQ) {
      for (int e = lane; e < IPW * 6 * n; e += 32) sf[e] = e < nk * 6 * n ? f[first * 6 * n + e] : T(0);
    }
    if (bulk_in) {
      warp_bulk_load_wait(bar, phase);
      phase ^= 1u;
    } else {
      cp_copy<T, true>(td, df + first * slab, n, nk * 6 * n * n, IPW * 6 * n * n, lane);
    }
    __syncwarp();
    if (valid) {
      T* md = td + g * tvals + c * np;
      T* mc = tc + g * nn;
      const int rstride = n * np;
      for (int i = n - 1; i >= 0; --i) {
        const int p = m.parent[i];
        T col[6];
#pragma unroll
        for (int r = 0; r < 6; ++r) col[r] = md[r * rstride + i];
        T val = dot6(m.S[i], col);                                             // :1284 / :1325
        if (!DQ && use_damping && c == i) val += m.damping[i];                 // :1341
        mc[i * n + c] = val;
        if (p >= 0) {
          T X[18], t[6];
          build_X(m, i, sj[(g * n + i) * 2], sj[(g * n + i) * 2 + 1], X);
          XT_apply(X, col, t);                                                 // :1291 / :1331
          if (DQ && c == i) {
            T fi[6], S[6], fxs[6], t2[6];
#pragma unroll
            for (int r = 0; r < 6; ++r) { fi[r] = sf[g * 6 * n + r * n + i]; S[r] = m.S[i][r]; }
            crm_mul(fi, S, fxs);
#pragma unroll
            for (int r = 0; r < 6; ++r) fxs[r] = -fxs[r];                      // fxS :166-168
            XT_apply(X, fxs, t2);                                              // :1292
#pragma unroll
            for (int r = 0; r < 6; ++r) t[r] += t2[r];                         // :1293-1294
          }
#pragma unroll
          for (int r = 0; r < 6; ++r) md[r * rstride + p] += t[r];
        }
      }
    }
    __syncwarp();
    if (!(np == n && warp_bulk_store(df + first * slab, td, nk * 6 * n * n, lane)))
      cp_copy<T, false>(td, df + first * slab, n, nk * 6 * n * n, IPW * 6 * n * n, lane);
    if (!warp_bulk_store(dc + first * nn, tc, nk * nn, lane))
      for (int e = lane; e < nk * nn; e += 32) __stcs(dc + first * nn + e, tc[e]);
    __syncwarp();
  }
  warp_bulk_store_wait(lane);                             // shared memory must outlive the copies
}

}  // namespace rbd
