// rbd_pass_kernels.cuh - the eight per-pass helper kernels.
//
// The reference keeps each sweep of rnea / rnea_grad / minv as its own public method so that
// accelerator implementations can be checked pass by pass (README.md:19).  These kernels are
// those entry points: one thread per knot point, tensors in the reference's shapes with the
// batch axis first, same in-place contracts.  They are bound by the HBM traffic of their
// (6,n,NB)-sized tensors, not by arithmetic; the fused drivers in rbd_fused_kernels.cuh are the
// throughput path.
#pragma once
#include "rbd_common.cuh"

namespace rbd {

constexpr int kPassThreads = 128;

// ---- rnea_fpass (RBDReference.py:559-598) --------------------------------------------------
// Generic version (large robots, non-rigid inertias): one knot point per thread, v / a / f kept per thread
// in memory order and stored through warp_flush_blocks (coalesced 256-byte rows).
template <typename T>
__global__ void __launch_bounds__(kPassThreads)
rnea_fpass_kernel(const __grid_constant__ DevModel<T> m, int64_t B, const T* __restrict__ q,
                  const T* __restrict__ qd, const T* __restrict__ qdd, T gravity,
                  T* __restrict__ v, T* __restrict__ a, T* __restrict__ f) {
  __shared__ T tiles[kPassThreads / 32][32 * (kFlushChunk + 1)];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t b0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) - lane;
  if (b0 >= B) return;                                     // whole warp past the end
  const int nlive = (int)(B - b0 < 32 ? B - b0 : 32);
  const int64_t b = b0 + (lane < nlive ? lane : 0);        // idle lanes shadow a valid knot point
  const int n = m.n;
  const T* qb = q + b * n;
  const T* qdb = qd + b * n;
  const T* qddb = qdd ? qdd + b * n : nullptr;
  T lv[6 * RBD_MAX_DOF], la[6 * RBD_MAX_DOF], lf[6 * RBD_MAX_DOF];     // (6, NB): element (r, i) at r*n + i
  for (int i = 0; i < n; ++i) {
    T X[18];
    build_X_from_q(m, i, qb[i], X);
    const int p = m.parent[i];
    T vp[6], ap[6], vi[6], ai[6];
    if (p < 0) {
#pragma unroll
      for (int r = 0; r < 6; ++r) { vi[r] = T(0); ap[r] = T(0); }
      ap[5] = -gravity;                                  // :566
      X_apply(X, ap, ai);                                // :578
    } else {
#pragma unroll
      for (int r = 0; r < 6; ++r) { vp[r] = lv[r * n + p]; ap[r] = la[r * n + p]; }
      X_apply(X, vp, vi);                                // :580
      X_apply(X, ap, ai);                                // :581
    }
    T vJ[6], t[6];
    const T qdi = qdb[i];
#pragma unroll
    for (int r = 0; r < 6; ++r) { vJ[r] = m.S[i][r] * qdi; vi[r] += vJ[r]; }   // :586-587
    crm_mul(vi, vJ, t);                                                        // :588
#pragma unroll
    for (int r = 0; r < 6; ++r) ai[r] += t[r];
    if (qddb) {
      const T qddi = qddb[i];
#pragma unroll
      for (int r = 0; r < 6; ++r) ai[r] = fma_t(m.S[i][r], qddi, ai[r]);      // :589-593
    }
    T Ia[6], Iv[6], vxIv[6];
    mat6_apply(m.I[i], ai, Ia);
    mat6_apply(m.I[i], vi, Iv);
    crf_mul(vi, Iv, vxIv);                                                     // :170-182
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      lv[r * n + i] = vi[r];
      la[r * n + i] = ai[r];
      lf[r * n + i] = Ia[r] + vxIv[r];                                         // :596
    }
  }
  warp_flush_blocks(lv, 6 * n, tiles[warp], v + b0 * 6 * n, nlive, lane);
  warp_flush_blocks(la, 6 * n, tiles[warp], a + b0 * 6 * n, nlive, lane);
  warp_flush_blocks(lf, 6 * n, tiles[warp], f + b0 * 6 * n, nlive, lane);
}

// ---- rnea_bpass (RBDReference.py:600-621) --------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kPassThreads)
rnea_bpass_kernel(const __grid_constant__ DevModel<T> m, int64_t B, const T* __restrict__ q,
                  T* __restrict__ f, T* __restrict__ c) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int n = m.n;
  const T* qb = q + b * n;
  T* fb = f + b * 6 * n;
  T* cb = c + b * n;
  for (int i = n - 1; i >= 0; --i) {
    T fi[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) fi[r] = fb[r * n + i];
    cb[i] = dot6(m.S[i], fi);                                                  // :612
    const int p = m.parent[i];
    if (p >= 0) {
      T X[18], t[6];
      build_X_from_q(m, i, qb[i], X);
      XT_apply(X, fi, t);                                                      // :618
#pragma unroll
      for (int r = 0; r < 6; ++r) fb[r * n + p] += t[r];                       // :619
    }
  }
}

// ---- rnea_grad_fpass_dq / _dqd (RBDReference.py:1127-1187, :1189-1255) ---------------------
// Tensors are (6, n, NB): element (r, c, i) at (r*n + c)*n + i.  Every entry is written
// (structural zeros included) so the caller may pass uninitialised output buffers.
template <typename T, bool DQ>
__global__ void __launch_bounds__(kPassThreads)
rnea_grad_fpass_kernel(const __grid_constant__ DevModel<T> m, int64_t B, const T* __restrict__ q,
                       const T* __restrict__ qd, const T* __restrict__ v, const T* __restrict__ a,
                       T gravity, T* __restrict__ dv, T* __restrict__ da, T* __restrict__ df) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int n = m.n;
  const T* qb = q + b * n;
  const T* qdb = qd + b * n;
  const T* vb = v + b * 6 * n;
  const T* ab = DQ ? a + b * 6 * n : nullptr;
  const int64_t slab = (int64_t)6 * n * n;
  T* dvb = dv + b * slab;
  T* dab = da + b * slab;
  T* dfb = df + b * slab;
  for (int i = 0; i < n; ++i) {
    T X[18];
    build_X_from_q(m, i, qb[i], X);
    const int p = m.parent[i];
    const T qdi = qdb[i];
    T S[6], vi[6], Iv[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) { S[r] = m.S[i][r]; vi[r] = vb[r * n + i]; }
    mat6_apply(m.I[i], vi, Iv);                                                // :1180 / :1248
    // seed terms for column c == i
    T seed_v[6], seed_a[6];
    if (DQ) {
      T par[6], t[6];
      if (p >= 0) {
#pragma unroll
        for (int r = 0; r < 6; ++r) par[r] = vb[r * n + p];
        X_apply(X, par, t);
        crm_mul(t, S, seed_v);                                                 // :1159
#pragma unroll
        for (int r = 0; r < 6; ++r) par[r] = ab[r * n + p];
      } else {
#pragma unroll
        for (int r = 0; r < 6; ++r) { seed_v[r] = T(0); par[r] = T(0); }
        par[5] = -gravity;                                                     // :1137
      }
      X_apply(X, par, t);
      crm_mul(t, S, seed_a);                                                   // :1173 / :1175
    } else {
#pragma unroll
      for (int r = 0; r < 6; ++r) seed_v[r] = S[r];                            // :1231
      crm_mul(vi, S, seed_a);                                                  // :1243
    }
    for (int c = 0; c < n; ++c) {
      T dvc[6], dac[6];
      if (p >= 0) {
        T pv[6], pa[6];
#pragma unroll
        for (int r = 0; r < 6; ++r) {
          pv[r] = dvb[(r * n + c) * n + p];
          pa[r] = dab[(r * n + c) * n + p];
        }
        X_apply(X, pv, dvc);                                                   // :1158 / :1230
        X_apply(X, pa, dac);                                                   // :1163 / :1234
      } else {
#pragma unroll
        for (int r = 0; r < 6; ++r) { dvc[r] = T(0); dac[r] = T(0); }
      }
      if (c == i) {
#pragma unroll
        for (int r = 0; r < 6; ++r) dvc[r] += seed_v[r];
      }
      T t[6];
      crm_mul(dvc, S, t);                                                      // :1170 / :1240
#pragma unroll
      for (int r = 0; r < 6; ++r) dac[r] = fma_t(qdi, t[r], dac[r]);
      if (c == i) {
#pragma unroll
        for (int r = 0; r < 6; ++r) dac[r] += seed_a[r];
      }
      T Ida[6], Idv[6], t1[6], t2[6];
      mat6_apply(m.I[i], dac, Ida);                                            // :1179 / :1247
      mat6_apply(m.I[i], dvc, Idv);
      crf_mul(dvc, Iv, t1);                                                    // :1184 / :1251
      crf_mul(vi, Idv, t2);                                                    // :1185 / :1252
#pragma unroll
      for (int r = 0; r < 6; ++r) {
        dvb[(r * n + c) * n + i] = dvc[r];
        dab[(r * n + c) * n + i] = dac[r];
        dfb[(r * n + c) * n + i] = Ida[r] + t1[r] + t2[r];
      }
    }
  }
}

// ---- rnea_grad_bpass_dq / _dqd (RBDReference.py:1257-1297, :1299-1343) ---------------------
// df is an arbitrary caller tensor, so all n columns of every body are processed.
template <typename T, bool DQ>
__global__ void __launch_bounds__(kPassThreads)
rnea_grad_bpass_kernel(const __grid_constant__ DevModel<T> m, int64_t B, const T* __restrict__ q,
                       const T* __restrict__ f, T* __restrict__ df, int use_damping,
                       T* __restrict__ dc) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int n = m.n;
  const T* qb = q + b * n;
  const T* fb = DQ ? f + b * 6 * n : nullptr;
  T* dfb = df + b * (int64_t)6 * n * n;
  T* dcb = dc + b * (int64_t)n * n;
  for (int i = n - 1; i >= 0; --i) {
    const int p = m.parent[i];
    T X[18];
    if (p >= 0) build_X_from_q(m, i, qb[i], X);
    for (int c = 0; c < n; ++c) {
      T col[6];
#pragma unroll
      for (int r = 0; r < 6; ++r) col[r] = dfb[(r * n + c) * n + i];
      T val = dot6(m.S[i], col);                                               // :1284 / :1325
      if (!DQ && use_damping && c == i) val += m.damping[i];                   // :1341
      dcb[i * n + c] = val;
      if (p >= 0) {
        T t[6];
        XT_apply(X, col, t);                                                   // :1291 / :1331
        if (DQ && c == i) {
          T fi[6], S[6], fxs[6], t2[6];
#pragma unroll
          for (int r = 0; r < 6; ++r) { fi[r] = fb[r * n + i]; S[r] = m.S[i][r]; }
          crm_mul(fi, S, fxs);
#pragma unroll
          for (int r = 0; r < 6; ++r) fxs[r] = -fxs[r];                        // fxS :166-168
          XT_apply(X, fxs, t2);                                                // :1292
#pragma unroll
          for (int r = 0; r < 6; ++r) t[r] += t2[r];                           // :1293-1294
        }
#pragma unroll
        for (int r = 0; r < 6; ++r) dfb[(r * n + c) * n + p] += t[r];
      }
    }
  }
}

// ---- minv_bpass (RBDReference.py:630-735) --------------------------------------------------
// Minv (n,n), F (n,6,n): element (i, r, j) at (i*6 + r)*n + j, U (n,6), Dinv (n) [= D].
// The articulated inertias IA are private per-thread state (the reference deep-copies the
// inertia dict at :662); they live in local memory here.
template <typename T>
__global__ void __launch_bounds__(kPassThreads)
minv_bpass_kernel(const __grid_constant__ DevModel<T> m, int64_t B, const T* __restrict__ q,
                  T* __restrict__ Minv, T* __restrict__ F, T* __restrict__ U, T* __restrict__ Dinv) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int n = m.n;
  const T* qb = q + b * n;
  T* Mb = Minv + b * (int64_t)n * n;
  T* Fb = F + b * (int64_t)6 * n * n;
  T* Ub = U + b * (int64_t)6 * n;
  T* Db = Dinv + b * n;
  T IA[RBD_MAX_DOF][36];
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int k = 0; k < 36; ++k) IA[i][k] = m.I[i][k];
    for (int j = 0; j < n; ++j) {
      Mb[i * n + j] = T(0);
#pragma unroll
      for (int r = 0; r < 6; ++r) Fb[(i * 6 + r) * n + j] = T(0);
    }
  }
  for (int i = n - 1; i >= 0; --i) {
    T S[6], Ui[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) S[r] = m.S[i][r];
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      T acc = T(0);
#pragma unroll
      for (int k = 0; k < 6; ++k) acc = fma_t(IA[i][6 * r + k], S[k], acc);
      Ui[r] = acc;                                                             // :697
      Ub[i * 6 + r] = acc;
    }
    const T D = dot6(S, Ui);                                                   // :698
    Db[i] = D;
    const T invD = T(1) / D;
    const unsigned sub = m.sub_mask[i];
    const int p = m.parent[i];
    T X[18];
    if (p >= 0) build_X_from_q(m, i, qb[i], X);
    for (int j = i; j < n; ++j) {
      if (!((sub >> j) & 1u)) continue;
      T Fij[6];
#pragma unroll
      for (int r = 0; r < 6; ++r) Fij[r] = Fb[(i * 6 + r) * n + j];
      T mij = (j == i ? invD : T(0)) - invD * dot6(S, Fij);                    // :700-708
      Mb[i * n + j] = mij;
      if (p >= 0) {
        T t[6];
#pragma unroll
        for (int r = 0; r < 6; ++r) {
          Fij[r] = fma_t(Ui[r], mij, Fij[r]);                                  // :721-723
          Fb[(i * 6 + r) * n + j] = Fij[r];
        }
        XT_apply(X, Fij, t);                                                   // :724-726
#pragma unroll
        for (int r = 0; r < 6; ++r) Fb[(p * 6 + r) * n + j] += t[r];
      }
    }
    if (p >= 0) {
      // Ia = IA_i - U U^T / D ; IA_p += X^T Ia X                              // :728-733
      T Xf[6][6];
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          Xf[r][c] = X[3 * r + c];
          Xf[r][3 + c] = T(0);
          Xf[3 + r][c] = X[9 + 3 * r + c];
          Xf[3 + r][3 + c] = X[3 * r + c];
        }
      T Ia[6][6], tmp[6][6];
#pragma unroll
      for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int c = 0; c < 6; ++c) Ia[r][c] = IA[i][6 * r + c] - Ui[r] * (invD * Ui[c]);
#pragma unroll
      for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int c = 0; c < 6; ++c) {
          T acc = T(0);
#pragma unroll
          for (int k = 0; k < 6; ++k) acc = fma_t(Ia[r][k], Xf[k][c], acc);
          tmp[r][c] = acc;
        }
#pragma unroll
      for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int c = 0; c < 6; ++c) {
          T acc = T(0);
#pragma unroll
          for (int k = 0; k < 6; ++k) acc = fma_t(Xf[k][r], tmp[k][c], acc);
          IA[p][6 * r + c] += acc;
        }
    }
  }
}

// ---- minv_fpass (RBDReference.py:737-783) --------------------------------------------------
// F[i] (6,n) is rewritten for every body; whole rows of Minv are updated (:771).
template <typename T>
__global__ void __launch_bounds__(kPassThreads)
minv_fpass_kernel(const __grid_constant__ DevModel<T> m, int64_t B, const T* __restrict__ q,
                  T* __restrict__ Minv, T* __restrict__ F, const T* __restrict__ U,
                  const T* __restrict__ Dinv) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int n = m.n;
  const T* qb = q + b * n;
  T* Mb = Minv + b * (int64_t)n * n;
  T* Fb = F + b * (int64_t)6 * n * n;
  const T* Ub = U + b * (int64_t)6 * n;
  const T* Db = Dinv + b * n;
  for (int i = 0; i < n; ++i) {
    const int p = m.parent[i];
    if (p >= 0) {
      T X[18], Ui[6], UX[6];
      build_X_from_q(m, i, qb[i], X);
#pragma unroll
      for (int r = 0; r < 6; ++r) Ui[r] = Ub[i * 6 + r];
      XT_apply(X, Ui, UX);                                                     // U^T X = (X^T U)^T
      const T invD = T(1) / Db[i];
      for (int j = 0; j < n; ++j) {
        T Fp[6], Fi[6];
#pragma unroll
        for (int r = 0; r < 6; ++r) Fp[r] = Fb[(p * 6 + r) * n + j];
        const T mij = Mb[i * n + j] - invD * dot6(UX, Fp);                     // :771-773
        Mb[i * n + j] = mij;
        X_apply(X, Fp, Fi);
#pragma unroll
        for (int r = 0; r < 6; ++r) Fb[(i * 6 + r) * n + j] = fma_t(m.S[i][r], mij, Fi[r]);   // :774-776
      }
    } else {
      for (int j = 0; j < n; ++j) {
        const T mij = Mb[i * n + j];
#pragma unroll
        for (int r = 0; r < 6; ++r) Fb[(i * 6 + r) * n + j] = m.S[i][r] * mij;               // :781
      }
    }
  }
}

// ---- minv_bpass, one column per lane ----------------------------------------------------------
// Lane (g, l) of a warp carries column l of knot point g (G = 8 | 16 | 32 lanes per knot point).
// Columns never mix in :700-726: column j becomes active at body j and walks to its root, its
// running F (six registers) is F[i][:, j] of the body it currently belongs to.  Rows of Minv and of
// F[i][r, :] are stored once, contiguously across the lanes (zeros outside subtree(i)); the
// knot-point-per-thread version used F in global memory as read-modify-write working storage, 8 bytes
// per 32-byte sector.  The articulated inertias (:728-733) are shared by the lanes of a knot point
// through shared memory (IA[body][36], U): lane c < 6 owns column c of the update - Ia X[:, c] from shared
// memory, then X^T of it straight from its registers (XT_apply), added to column c of IA_parent.
template <int G> __host__ __device__ constexpr int minv_bpass_col_warp_vals(int n) { return (32 / G) * (n * 36 + 8); }

template <typename T, int G>
__global__ void __launch_bounds__(kPassThreads)
minv_bpass_col_kernel(const __grid_constant__ DevModel<T> m, int64_t B, const T* __restrict__ q,
                      T* __restrict__ Minv, T* __restrict__ F, T* __restrict__ U, T* __restrict__ Dinv) {
  extern __shared__ __align__(16) unsigned char bp_smem_raw[];
  constexpr int IPW = 32 / G;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int g = lane / G, l = lane % G;
  const int n = m.n;
  const int knot_vals = n * 36 + 8;
  T* ws = reinterpret_cast<T*>(bp_smem_raw) + (size_t)warp * IPW * knot_vals + (size_t)g * knot_vals;
  T* IA = ws;                  // [n][36]
  T* Us = IA + n * 36;         // U(6)
  const int64_t ntask = (B + IPW - 1) / IPW;
  for (int64_t task = (int64_t)blockIdx.x * nwarps + warp; task < ntask; task += (int64_t)gridDim.x * nwarps) {
    const int64_t b = task * IPW + g;
    const bool live = b < B;
    const bool col = live && l < n;                         // this lane owns a column
    const int64_t bb = live ? b : task * IPW;               // idle groups shadow a valid knot point (no stores)
    const T* qb = q + bb * n;
    T* Mb = Minv + bb * (int64_t)n * n;
    T* Fb = F + bb * (int64_t)6 * n * n;
    T* Ub = U + bb * (int64_t)6 * n;
    T* Db = Dinv + bb * n;
    for (int e = l; e < n * 36; e += G) IA[e] = m.I[e / 36][e % 36];     // :662
    __syncwarp();
    T Frun[6] = {T(0), T(0), T(0), T(0), T(0), T(0)};
    for (int i = n - 1; i >= 0; --i) {
      const int p = m.parent[i];
      T X[18];
      if (p >= 0) build_X_from_q(m, i, qb[i], X);
      // U = IA_i S, D = S . U (:697-698): lane r < 6 computes U[r]
      if (l < 6) {
        T acc = T(0);
#pragma unroll
        for (int k = 0; k < 6; ++k) acc = fma_t(IA[i * 36 + 6 * l + k], m.S[i][k], acc);
        Us[l] = acc;
      }
      __syncwarp();
      T Ui[6];
#pragma unroll
      for (int r = 0; r < 6; ++r) Ui[r] = Us[r];
      const T D = dot6(m.S[i], Ui);
      const T invD = T(1) / D;
      if (live) {
        if (l < 6) Ub[i * 6 + l] = Ui[l];
        if (l == 6) Db[i] = D;                               // "Dinv" holds D (:698)
      }
      // column l at body i (:700-726)
      if (col) {
        const bool mine = (m.sub_mask[i] >> l) & 1u;
        T mij = T(0), Fout[6] = {T(0), T(0), T(0), T(0), T(0), T(0)};
        if (mine) {
          mij = (l == i ? invD : T(0)) - invD * dot6(m.S[i], Frun);
#pragma unroll
          for (int r = 0; r < 6; ++r) Fout[r] = p >= 0 ? fma_t(Ui[r], mij, Frun[r]) : Frun[r];
          if (p >= 0) XT_apply(X, Fout, Frun);               // now F[p][:, l] (:724-726)
        }
        Mb[i * n + l] = mij;
#pragma unroll
        for (int r = 0; r < 6; ++r) Fb[(i * 6 + r) * n + l] = Fout[r];
      }
      // IA_p += X^T (IA_i - U U^T / D) X (:728-733): lane c < 6 owns column c
      if (p >= 0 && l < 6) {
        // column c of X = [[E, 0], [L, E]] picked out of the lane's registers (c is a run-time value)
        const int cc = l < 3 ? l : l - 3;
        T xk[6];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const T e = cc == 0 ? X[3 * r] : (cc == 1 ? X[3 * r + 1] : X[3 * r + 2]);
          const T lo = cc == 0 ? X[9 + 3 * r] : (cc == 1 ? X[9 + 3 * r + 1] : X[9 + 3 * r + 2]);
          xk[r] = l < 3 ? e : T(0);
          xk[3 + r] = l < 3 ? lo : e;
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
        for (int r = 0; r < 6; ++r) IA[p * 36 + 6 * r + l] += t[r];
      }
      __syncwarp();
    }
  }
}

// ---- minv_fpass, one column per lane ----------------------------------------------------------
// The forward pass never mixes columns (:771-776 act on whole rows, entry by entry), so lane (g, j)
// carries column j of knot point g through the bodies: every access to Minv[i, :] and F[i][r, :] is
// one contiguous row per knot point (the knot-point-per-thread version touched 8 bytes per 32-byte
// sector), the parent's F stays in registers along chains and is re-read (own store) at branch
// points.  G = 8 | 16 | 32 lanes per knot point; lanes j >= n idle.
template <typename T, int G>
__global__ void __launch_bounds__(kPassThreads)
minv_fpass_col_kernel(const __grid_constant__ DevModel<T> m, int64_t B, const T* __restrict__ q,
                      T* __restrict__ Minv, T* __restrict__ F, const T* __restrict__ U,
                      const T* __restrict__ Dinv) {
  constexpr int IPW = 32 / G;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int g = lane / G, j = lane % G;
  const int n = m.n;
  const int64_t ntask = (B + IPW - 1) / IPW;
  for (int64_t task = (int64_t)blockIdx.x * nwarps + warp; task < ntask; task += (int64_t)gridDim.x * nwarps) {
    const int64_t b = task * IPW + g;
    if (b >= B || j >= n) continue;
    const T* qb = q + b * n;
    T* Mb = Minv + b * (int64_t)n * n;
    T* Fb = F + b * (int64_t)6 * n * n;
    const T* Ub = U + b * (int64_t)6 * n;
    const T* Db = Dinv + b * n;
    T Fprev[6];
    int prev = -1;
    for (int i = 0; i < n; ++i) {
      const int p = m.parent[i];
      T mij = Mb[i * n + j];
      if (p >= 0) {
        T X[18], Ui[6], UX[6], Fp[6], Fi[6];
        build_X_from_q(m, i, qb[i], X);
#pragma unroll
        for (int r = 0; r < 6; ++r) Ui[r] = Ub[i * 6 + r];
        XT_apply(X, Ui, UX);                                                   // U^T X = (X^T U)^T
        const T invD = T(1) / Db[i];
        if (p == prev) {
#pragma unroll
          for (int r = 0; r < 6; ++r) Fp[r] = Fprev[r];
        } else {
#pragma unroll
          for (int r = 0; r < 6; ++r) Fp[r] = Fb[(p * 6 + r) * n + j];
        }
        mij -= invD * dot6(UX, Fp);                                            // :771-773
        Mb[i * n + j] = mij;
        X_apply(X, Fp, Fi);
#pragma unroll
        for (int r = 0; r < 6; ++r) Fprev[r] = fma_t(m.S[i][r], mij, Fi[r]);   // :774-776
      } else {
#pragma unroll
        for (int r = 0; r < 6; ++r) Fprev[r] = m.S[i][r] * mij;                // :781
      }
#pragma unroll
      for (int r = 0; r < 6; ++r) Fb[(i * 6 + r) * n + j] = Fprev[r];
      prev = i;
    }
  }
}

// ---- minv epilogue (RBDReference.py:799-804): lower <- upper --------------------------------
template <typename T>
__global__ void __launch_bounds__(kPassThreads)
minv_mirror_kernel(int n, int64_t B, T* __restrict__ Minv) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  T* Mb = Minv + b * (int64_t)n * n;
  for (int r = 1; r < n; ++r)
    for (int c = 0; c < r; ++c) Mb[r * n + c] = Mb[c * n + r];
}

}  // namespace rbd
