// rbd_launch_fb.cu - part of librbd_b200.so: floating-base rnea / rnea_grad / minv (include/rbd_b200.h).
// Reference: the `floating_base` branches of RBDReference.py:559-806 and :1127-1368.
#include "rbd_internal.cuh"
#include "rbd_fb_kernels.cuh"
#include "rbd_fb_pass_kernels.cuh"

using namespace rbd;
using namespace rbd_host;

struct rbd_fb_model {
  FbModel<double> d;
  FbModel<float> f;
  // warp-cooperative kernels (rbd_coop_kernels.cuh with FB = true): the NB bodies in depth-first numbering, the base
  // as body 0 with an identity transform; valid when every inertia has rigid-body structure
  bool fast_ok;
  FastModel<double> fd_dfs;
  FastModel<float> ff_dfs;
  DfsPlan plan;
  CoopPlan coop;
  CoopMinvPlan coop_minv;
  FbBaseLayout layout;
  mutable std::atomic<int> variant{-1};   // per-handle kernel family (-1: follow rbd_set_kernel_variant)
};

namespace {

// kernel family of one call: the handle's own choice (rbd_fb_model_set_kernel_variant) or the process default
inline int fb_variant_of(const rbd_fb_model* m) {
  const int v = m ? m->variant.load(std::memory_order_relaxed) : -1;
  return v >= 0 ? v : g_variant.load(std::memory_order_relaxed);
}

template <typename T> const FastModel<T>& pick_fb_dfs(const rbd_fb_model* m);
template <> const FastModel<double>& pick_fb_dfs<double>(const rbd_fb_model* m) { return m->fd_dfs; }
template <> const FastModel<float>& pick_fb_dfs<float>(const rbd_fb_model* m) { return m->ff_dfs; }

// The robot as the cooperative kernels see it: body 0 becomes a joint with the identity transform (the kernels work in
// base coordinates and treat lane 0 as the base), bodies 1.. keep their 1-DoF joints.
void build_fb_coop(const RbdFbModelDesc* fd, rbd_fb_model* m) {
  const RbdModelDesc* d = &fd->bodies;
  const int n = d->n;
  double S[RBD_MAX_DOF * 6], XA[RBD_MAX_DOF * 18], XB[RBD_MAX_DOF * 18], XC[RBD_MAX_DOF * 18];
  int32_t kind[RBD_MAX_DOF];
  std::memcpy(S, d->S, sizeof(double) * 6 * n);
  std::memcpy(XA, d->XA, sizeof(double) * 18 * n);
  std::memcpy(XB, d->XB, sizeof(double) * 18 * n);
  std::memcpy(XC, d->XC, sizeof(double) * 18 * n);
  std::memcpy(kind, d->kind, sizeof(int32_t) * n);
  for (int k = 0; k < 6; ++k) S[k] = k == 2 ? 1.0 : 0.0;
  for (int k = 0; k < 18; ++k) { XA[k] = (k == 0 || k == 4 || k == 8) ? 1.0 : 0.0; XB[k] = 0.0; XC[k] = 0.0; }
  kind[0] = 0;
  RbdModelDesc pd = {n, d->parent, kind, S, XA, XB, XC, d->I, d->damping};
  m->fast_ok = build_dfs_model(&pd, m->fd_dfs, m->plan) && m->plan.orig[0] == 0;
  narrow_fast_model(m->fd_dfs, m->ff_dfs);
  build_coop_plans(m->fd_dfs, m->coop, m->coop_minv);
  m->layout.quat_off = fd->quat_off;
  m->layout.w_first = fd->w_first;
  m->layout.transpose = fd->transpose;
}

template <typename T> const FbModel<T>& pick_fb(const rbd_fb_model* m);
template <> const FbModel<double>& pick_fb<double>(const rbd_fb_model* m) { return m->d; }
template <> const FbModel<float>& pick_fb<float>(const rbd_fb_model* m) { return m->f; }

template <typename T>
void fill_fb(const RbdFbModelDesc* fd, FbModel<T>& out) {
  std::memset(&out, 0, sizeof(out));
  const RbdModelDesc* d = &fd->bodies;
  const int n = d->n;
  out.pos_off = fd->pos_off; out.quat_off = fd->quat_off; out.w_first = fd->w_first; out.transpose = fd->transpose;
  DevModel<T>& o = out.d;
  o.n = n;
  for (int i = 0; i < n; ++i) {
    o.parent[i] = d->parent[i];
    o.kind[i] = d->kind[i];
    o.damping[i] = (T)(d->damping ? d->damping[i] : 0.0);
    for (int k = 0; k < 36; ++k) o.I[i][k] = (T)d->I[i * 36 + k];
    if (i > 0) {
      for (int k = 0; k < 6; ++k) o.S[i][k] = (T)d->S[i * 6 + k];
      for (int k = 0; k < 18; ++k) {
        o.XA[i][k] = (T)d->XA[i * 18 + k]; o.XB[i][k] = (T)d->XB[i * 18 + k]; o.XC[i][k] = (T)d->XC[i * 18 + k];
      }
    }
    unsigned anc = 1u << i;
    if (d->parent[i] >= 0) anc |= o.anc_mask[d->parent[i]];
    o.anc_mask[i] = anc;
    if (i > 0 && d->parent[i] != i - 1) out.store_mask |= 1u << d->parent[i];
  }
  for (int i = n - 1; i >= 0; --i) {
    o.sub_mask[i] |= 1u << i;
    if (d->parent[i] >= 0) o.sub_mask[d->parent[i]] |= o.sub_mask[i];
  }
}

// dynamic shared memory of the rnea_grad and minv kernels: the per-body constants (rbd_fb_kernels.cuh: fb_stage_model)
template <typename T>
size_t fb_smem_bytes(const rbd_fb_model* m) { return (size_t)m->d.d.n * kFbSmStride * sizeof(T); }

template <typename T>
int launch_fb_rnea(const rbd_fb_model* m, int64_t B, const T* q, const T* qd, const T* qdd, T g, T* c, T* v, T* a, T* f,
                   void* stream) {
  RBD_CHECK_ARGS(m && q && qd && c && B >= 0, "rbd_fb_rnea: null model/q/qd/c or negative B");
  if (B == 0) return 0;
  fb_rnea_kernel<T><<<blocks_for(B, kFbThreads), kFbThreads, 0, (cudaStream_t)stream>>>(pick_fb<T>(m), B, q, qd, qdd, g, c, v, a, f);
  return cuda_status("rbd_fb_rnea");
}

template <typename T>
int launch_fb_rnea_grad(const rbd_fb_model* m, int64_t B, const T* q, const T* qd, const T* qdd, T g, int damp, T* dc_du,
                        T* c_out, void* stream) {
  RBD_CHECK_ARGS(m && q && qd && dc_du && B >= 0, "rbd_fb_rnea_grad: null model/q/qd/dc_du or negative B");
  if (B == 0) return 0;
  const int variant = fb_variant_of(m);
  if (m->fast_ok && variant != 1 && variant != 2) {
    // world-frame composites in base coordinates, one body per lane (rbd_coop_kernels.cuh, FB = true)
    const FastModel<T>& fm = pick_fb_dfs<T>(m);
    const int n = fm.n, nv = n + 5;
    const int G = n <= 8 ? 8 : (n <= 16 ? 16 : 32);
    const int ipw = 32 / G;
    const int tile_stride = coop_grad_tile_stride(nv, ipw, false);
    const size_t smem = (size_t)(((n * kCoopMdlStride + 1) & ~1) + kCoopWarps * 32 * kCoopVecStride + kCoopWarps * tile_stride) * sizeof(T) +
                        (size_t)n * kCoopIntStride * sizeof(int);
    if (smem <= kMaxDynSmem) {
      auto kern = G == 8 ? rnea_grad_coop_kernel<T, 8, false, false, true>
                         : (G == 16 ? rnea_grad_coop_kernel<T, 16, false, false, true> : rnea_grad_coop_kernel<T, 32, false, false, true>);
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDynSmem);
      if (e != cudaSuccess) return fail((int)e, cudaGetErrorString(e));
      const int64_t ngroups = (B + ipw - 1) / ipw;
      int64_t blocks = (ngroups + kCoopWarps - 1) / kCoopWarps;
      if (blocks > grid_cap()) blocks = grid_cap();
      kern<<<(unsigned)blocks, kCoopWarps * 32, smem, (cudaStream_t)stream>>>(fm, m->plan, m->coop, B, q, qd, qdd, g, damp, dc_du, c_out,
                                                                              m->layout);
      return cuda_status("rbd_fb_rnea_grad(coop)");
    }
  }
  fb_rnea_grad_kernel<T><<<blocks_for(B, kFbThreads), kFbThreads, fb_smem_bytes<T>(m), (cudaStream_t)stream>>>(pick_fb<T>(m), B, q, qd, qdd, g, damp,
                                                                                             dc_du, c_out);
  return cuda_status("rbd_fb_rnea_grad");
}

template <typename T>
int launch_fb_minv(const rbd_fb_model* m, int64_t B, const T* q, int dense, T* Minv, void* stream) {
  RBD_CHECK_ARGS(m && q && Minv && B >= 0, "rbd_fb_minv: null model/q/Minv or negative B");
  if (B == 0) return 0;
  const int variant = fb_variant_of(m);
  if (m->fast_ok && variant != 1 && variant != 2) {
    // warp-cooperative kernel (rbd_coop_minv_kernels.cuh, FB = true).  The reference fills both triangles of a
    // floating-base Minv (:761-781 works on whole rows; output_dense then copies the upper triangle of the leading
    // NB x NB block over the lower one, :799-804), so either setting of output_dense is the full symmetric matrix.
    (void)dense;
    const FastModel<T>& fm = pick_fb_dfs<T>(m);
    const int n = fm.n;
    const int G = n <= 8 ? 8 : (n <= 16 ? 16 : 32);
    auto kern = fm.has_prismatic
                    ? (G == 8 ? minv_coop_kernel<T, 8, true, false, true>
                              : (G == 16 ? minv_coop_kernel<T, 16, true, false, true> : minv_coop_kernel<T, 32, true, false, true>))
                    : (G == 8 ? minv_coop_kernel<T, 8, false, false, true>
                              : (G == 16 ? minv_coop_kernel<T, 16, false, false, true> : minv_coop_kernel<T, 32, false, false, true>));
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDynSmem);
    if (e != cudaSuccess) return fail((int)e, cudaGetErrorString(e));
    int warps = 0, best = 0;
    size_t smem = 0;
    for (int w = 4; w <= kCmMaxWarps; ++w) {
      const size_t sz = coop_minv_smem_bytes<T>(n, G, m->coop.maxdepth, fm.n_slot_a, w, true);
      if (sz > kMaxDynSmem) break;
      int nb = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, w * 32, sz) != cudaSuccess) { cudaGetLastError(); continue; }
      if (nb * w > best) { best = nb * w; warps = w; smem = sz; }
    }
    if (warps > 0) {
      const int ipw = 32 / G;
      const int64_t ngroups = (B + ipw - 1) / ipw;
      int64_t blocks = (ngroups + warps - 1) / warps;
      if (blocks > grid_cap()) blocks = grid_cap();
      kern<<<(unsigned)blocks, warps * 32, smem, (cudaStream_t)stream>>>(fm, m->plan, m->coop, m->coop_minv, B, q, Minv, nullptr, nullptr, nullptr, 0);
      return cuda_status("rbd_fb_minv(coop)");
    }
  }
  fb_minv_kernel<T><<<blocks_for(B, kFbThreads), kFbThreads, fb_smem_bytes<T>(m), (cudaStream_t)stream>>>(pick_fb<T>(m), B, q, dense, Minv);
  return cuda_status("rbd_fb_minv");
}


// ---- per-pass helpers (rbd_fb_pass_kernels.cuh): one knot point per thread, the passes' arrays are the working storage
template <typename T>
int launch_fb_rnea_fpass(const rbd_fb_model* m, int64_t B, const T* q, const T* qd, const T* qdd, T g, T* v, T* a, T* f, void* stream) {
  RBD_CHECK_ARGS(m && q && qd && v && a && f && B >= 0, "rbd_fb_rnea_fpass: null argument or negative B");
  if (B == 0) return 0;
  fbp_rnea_fpass_kernel<T><<<blocks_for(B, kFbPassThreads), kFbPassThreads, 0, (cudaStream_t)stream>>>(pick_fb<T>(m), B, q, qd, qdd, g, v, a, f);
  return cuda_status("rbd_fb_rnea_fpass");
}
template <typename T>
int launch_fb_rnea_bpass(const rbd_fb_model* m, int64_t B, const T* q, T* f, T* c, void* stream) {
  RBD_CHECK_ARGS(m && q && f && c && B >= 0, "rbd_fb_rnea_bpass: null argument or negative B");
  if (B == 0) return 0;
  fbp_rnea_bpass_kernel<T><<<blocks_for(B, kFbPassThreads), kFbPassThreads, 0, (cudaStream_t)stream>>>(pick_fb<T>(m), B, q, f, c);
  return cuda_status("rbd_fb_rnea_bpass");
}
template <typename T>
int launch_fb_minv_bpass(const rbd_fb_model* m, int64_t B, const T* q, T* Minv, T* F, T* U, T* Dinv, void* stream) {
  RBD_CHECK_ARGS(m && q && Minv && F && U && Dinv && B >= 0, "rbd_fb_minv_bpass: null argument or negative B");
  if (B == 0) return 0;
  if (fb_variant_of(m) != 1) {
    // articulated inertias shared through shared memory, then one column per lane (rbd_fb_pass_kernels.cuh)
    const int warps = kPassThreads / 32;
    const size_t smem = (size_t)warps * fbp_minv_bpass_warp_vals(m->d.d.n) * sizeof(T);
    auto kern = fbp_minv_bpass_col_kernel<T>;
    if (smem <= 48 * 1024 || cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess) {
      int64_t blocks = (B + warps - 1) / warps;
      if (blocks > grid_cap()) blocks = grid_cap();
      kern<<<(unsigned)blocks, kPassThreads, smem, (cudaStream_t)stream>>>(pick_fb<T>(m), B, q, Minv, F, U, Dinv);
      return cuda_status("rbd_fb_minv_bpass(col)");
    }
    cudaGetLastError();
  }
  fbp_minv_bpass_kernel<T><<<blocks_for(B, kFbPassThreads), kFbPassThreads, 0, (cudaStream_t)stream>>>(pick_fb<T>(m), B, q, Minv, F, U, Dinv);
  return cuda_status("rbd_fb_minv_bpass");
}
template <typename T>
int launch_fb_minv_fpass(const rbd_fb_model* m, int64_t B, const T* q, T* Minv, T* F, const T* U, const T* Dinv, void* stream) {
  RBD_CHECK_ARGS(m && q && Minv && F && U && Dinv && B >= 0, "rbd_fb_minv_fpass: null argument or negative B");
  if (B == 0) return 0;
  if (fb_variant_of(m) != 1) {
    // one column per lane, a warp per knot point: coalesced rows of Minv and F (rbd_fb_pass_kernels.cuh)
    const int warps = kPassThreads / 32;
    int64_t blocks = (B + warps - 1) / warps;
    if (blocks > grid_cap()) blocks = grid_cap();
    fbp_minv_fpass_col_kernel<T><<<(unsigned)blocks, kPassThreads, (size_t)warps * 2 * RBD_MAX_DOF * sizeof(T), (cudaStream_t)stream>>>(
        pick_fb<T>(m), B, q, Minv, F, U, Dinv);
    return cuda_status("rbd_fb_minv_fpass(col)");
  }
  fbp_minv_fpass_kernel<T><<<blocks_for(B, kFbPassThreads), kFbPassThreads, 0, (cudaStream_t)stream>>>(pick_fb<T>(m), B, q, Minv, F, U, Dinv);
  return cuda_status("rbd_fb_minv_fpass");
}
template <typename T, bool DQ>
int launch_fb_grad_fpass(const rbd_fb_model* m, int64_t B, const T* q, const T* qd, const T* v, const T* a, T g, T* dv, T* da, T* df,
                         void* stream) {
  RBD_CHECK_ARGS(m && q && qd && v && (a || !DQ) && dv && da && df && B >= 0, "rbd_fb_rnea_grad_fpass: null argument or negative B");
  if (B == 0) return 0;
  if (DQ && m->d.d.n < 6)        // RBDReference.py:1168 indexes bodies 0..5 (IndexError upstream)
    return fail(RBD_E_UNSUPPORTED, "rbd_fb_rnea_grad_fpass_dq: the reference needs at least 6 bodies on this path (:1168)");
  if (fb_variant_of(m) != 1) {
    // one body per lane, one ancestor distance per round, slabs written in one coalesced pass
    // (rbd_fb_pass_kernels.cuh: fbp_grad_fpass_level_kernel)
    const int NB = m->d.d.n;
    const int G = NB <= 8 ? 8 : (NB <= 16 ? 16 : 32);
    int npairs = 0;
    for (int i = 0; i < NB; ++i) {
      npairs += 6;
      for (int c = i; c >= 1; c = m->d.d.parent[c]) ++npairs;
    }
    auto kern = G == 8 ? fbp_grad_fpass_level_kernel<T, 8, DQ>
                       : (G == 16 ? fbp_grad_fpass_level_kernel<T, 16, DQ> : fbp_grad_fpass_level_kernel<T, 32, DQ>);
    int warps = 0, best = 0, ctas = 0;
    size_t smem = 0;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDynSmem) == cudaSuccess) {
      for (int w = 1; w <= kCpLvlMaxWarps; ++w) {
        const size_t sz = fbp_level_head_bytes(NB, G, sizeof(T)) + (size_t)(32 / G) * fbp_level_knot_vals(NB, npairs) * sizeof(T) * w;
        if (sz > kMaxDynSmem) break;
        int nb = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, w * 32, sz) != cudaSuccess) { cudaGetLastError(); continue; }
        if (nb * w > best) { best = nb * w; warps = w; smem = sz; ctas = nb; }
      }
    }
    if (warps > 0) {
      const int64_t ngroups = (B + 32 / G - 1) / (32 / G);
      int64_t blocks = (ngroups + warps - 1) / warps;
      const int64_t cap = (int64_t)sm_count() * ctas * 4;
      if (blocks > cap) blocks = cap;
      kern<<<(unsigned)blocks, warps * 32, smem, (cudaStream_t)stream>>>(pick_fb<T>(m), npairs, B, q, qd, v, a, g, dv, da, df);
      return cuda_status("rbd_fb_rnea_grad_fpass(level)");
    }
    cudaGetLastError();
  }
  fbp_grad_fpass_kernel<T, DQ><<<blocks_for(B, kFbPassThreads), kFbPassThreads, 0, (cudaStream_t)stream>>>(pick_fb<T>(m), B, q, qd, v, a, g, dv, da, df);
  return cuda_status("rbd_fb_rnea_grad_fpass");
}
template <typename T, bool DQ>
int launch_fb_grad_bpass(const rbd_fb_model* m, int64_t B, const T* q, const T* f, T* df, int damp, T* dc, void* stream) {
  RBD_CHECK_ARGS(m && q && (f || !DQ) && df && dc && B >= 0, "rbd_fb_rnea_grad_bpass: null argument or negative B");
  if (B == 0) return 0;
  if (fb_variant_of(m) != 1) {
    // one column per lane, the knot point's df slab in a shared-memory tile (rbd_fb_pass_kernels.cuh: fbp_grad_bpass_coop_kernel)
    const int NB = m->d.d.n;
    auto kern = fbp_grad_bpass_coop_kernel<T, DQ>;
    const size_t per_warp = (size_t)fbp_bpass_warp_vals(NB) * sizeof(T);
    int warps = 0, best = 0, ctas = 0;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDynSmem) == cudaSuccess) {
      for (int w = 1; w <= kCpMaxWarps; ++w) {
        if (per_warp * w > kMaxDynSmem) break;
        int nb = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, w * 32, per_warp * w) != cudaSuccess) { cudaGetLastError(); continue; }
        if (nb * w > best) { best = nb * w; warps = w; ctas = nb; }
      }
    }
    if (warps > 0) {
      int64_t blocks = (B + warps - 1) / warps;
      const int64_t cap = (int64_t)sm_count() * ctas * 4;
      if (blocks > cap) blocks = cap;
      kern<<<(unsigned)blocks, warps * 32, per_warp * warps, (cudaStream_t)stream>>>(pick_fb<T>(m), B, q, f, df, damp, dc);
      return cuda_status("rbd_fb_rnea_grad_bpass(coop)");
    }
    cudaGetLastError();
  }
  fbp_grad_bpass_kernel<T, DQ><<<blocks_for(B, kFbPassThreads), kFbPassThreads, 0, (cudaStream_t)stream>>>(pick_fb<T>(m), B, q, f, df, damp, dc);
  return cuda_status("rbd_fb_rnea_grad_bpass");
}

// forward_dynamics / forward_dynamics_grad (RBDReference.py:1369-1384) are robot-agnostic compositions; with a
// floating base they run the three kernels above plus the per-knot-point product of rbd_fd_kernels.cuh.
template <typename T>
int launch_fb_forward_dynamics(const rbd_fb_model* m, int64_t B, const T* q, const T* qd, const T* u, T* qdd, T* Minv_out,
                               void* stream) {
  RBD_CHECK_ARGS(m && q && qd && u && qdd && B >= 0, "rbd_fb_forward_dynamics: null argument or negative B");
  if (B == 0) return 0;
  const int nv = m->d.d.n + 5;
  PoolBuf c((cudaStream_t)stream), Mi((cudaStream_t)stream);
  int rc = c.alloc((size_t)B * nv * sizeof(T));
  if (rc) return rc;
  T* Minv = Minv_out;
  if (!Minv) {
    rc = Mi.alloc((size_t)B * nv * nv * sizeof(T));
    if (rc) return rc;
    Minv = (T*)Mi.p;
  }
  rc = launch_fb_rnea<T>(m, B, q, qd, nullptr, T(-9.81), (T*)c.p, nullptr, nullptr, nullptr, stream);     // :1370
  if (rc) return rc;
  rc = launch_fb_minv<T>(m, B, q, 1, Minv, stream);                                                       // :1371
  if (rc) return rc;
  return launch_fd_apply<T, false>(fb_variant_of(m), nv, 1, B, Minv, u, (const T*)c.p, T(1), qdd, nullptr, stream);         // :1372
}

template <typename T>
int launch_fb_forward_dynamics_grad(const rbd_fb_model* m, int64_t B, const T* q, const T* qd, const T* u, T* qdd_dq,
                                    T* qdd_dqd, T* qdd_out, void* stream) {
  RBD_CHECK_ARGS(m && q && qd && u && qdd_dq && qdd_dqd && B >= 0, "rbd_fb_forward_dynamics_grad: null argument or negative B");
  if (B == 0) return 0;
  const int nv = m->d.d.n + 5;
  PoolBuf Mi((cudaStream_t)stream), dd((cudaStream_t)stream), dc((cudaStream_t)stream);
  int rc = Mi.alloc((size_t)B * nv * nv * sizeof(T));
  if (rc) return rc;
  rc = dc.alloc((size_t)B * nv * 2 * nv * sizeof(T));
  if (rc) return rc;
  T* qdd = qdd_out;
  if (!qdd) {
    rc = dd.alloc((size_t)B * nv * sizeof(T));
    if (rc) return rc;
    qdd = (T*)dd.p;
  }
  rc = launch_fb_forward_dynamics<T>(m, B, q, qd, u, qdd, (T*)Mi.p, stream);                               // :1377
  if (rc) return rc;
  rc = launch_fb_rnea_grad<T>(m, B, q, qd, qdd, T(-9.81), 0, (T*)dc.p, nullptr, stream);                   // :1378
  if (rc) return rc;
  return launch_fd_apply<T, true>(fb_variant_of(m), nv, 2 * nv, B, (const T*)Mi.p, (const T*)dc.p, nullptr, T(-1), qdd_dq, qdd_dqd, stream);
}

}  // namespace

extern "C" {

int rbd_fb_model_create(const RbdFbModelDesc* fd, rbd_fb_model_t** out) {
  if (!fd || !out) return fail(RBD_E_INVALID_ARGUMENT, "rbd_fb_model_create: null argument");
  *out = nullptr;
  const RbdModelDesc* d = &fd->bodies;
  if (d->n < 1 || d->n > RBD_MAX_DOF) return fail(RBD_E_UNSUPPORTED, "rbd_fb_model_create: number of bodies outside 1..RBD_MAX_DOF");
  if (!d->parent || !d->kind || !d->S || !d->XA || !d->XB || !d->XC || !d->I)
    return fail(RBD_E_INVALID_ARGUMENT, "rbd_fb_model_create: null table pointer");
  if (d->parent[0] != -1) return fail(RBD_E_INVALID_ARGUMENT, "rbd_fb_model_create: body 0 must be the floating base (parent -1)");
  for (int i = 1; i < d->n; ++i) {
    if (d->parent[i] < 0 || d->parent[i] >= i)
      return fail(RBD_E_INVALID_ARGUMENT, "rbd_fb_model_create: parent[i] must satisfy 0 <= parent[i] < i for i >= 1");
    if (d->kind[i] != 0 && d->kind[i] != 1)
      return fail(RBD_E_INVALID_ARGUMENT, "rbd_fb_model_create: kind[i] must be 0 (revolute) or 1 (prismatic)");
  }
  const bool layout_ok = (fd->pos_off == 0 && fd->quat_off == 3) || (fd->pos_off == 4 && fd->quat_off == 0);
  if (!layout_ok || (fd->w_first | 1) != 1 || (fd->transpose | 1) != 1)
    return fail(RBD_E_INVALID_ARGUMENT, "rbd_fb_model_create: base layout must be position+quaternion or quaternion+position");
  rbd_fb_model* m = new (std::nothrow) rbd_fb_model;
  if (!m) return fail(RBD_E_INVALID_ARGUMENT, "rbd_fb_model_create: out of host memory");
  fill_fb<double>(fd, m->d);
  fill_fb<float>(fd, m->f);
  build_fb_coop(fd, m);
  *out = m;
  return 0;
}

int rbd_fb_model_destroy(rbd_fb_model_t* m) {
  delete m;
  return 0;
}

int rbd_fb_model_num_vel(const rbd_fb_model_t* m) { return m ? m->d.d.n + 5 : RBD_E_INVALID_ARGUMENT; }

int rbd_fb_model_set_kernel_variant(rbd_fb_model_t* m, int variant) {
  if (!m || variant < -1 || variant > 3)
    return fail(RBD_E_INVALID_ARGUMENT, "floating-base kernel family: -1 follow the process default, 0 automatic (cooperative), "
                                        "1 / 2 one knot point per thread, 3 cooperative");
  m->variant.store(variant, std::memory_order_relaxed);
  return 0;
}

#define RBD_FB_DEFINE(SUF, T)                                                                                        \
  int rbd_fb_rnea_##SUF(const rbd_fb_model_t* m, int64_t B, const T* q, const T* qd, const T* qdd, T gravity, T* c,  \
                        T* v, T* a, T* f, void* stream) {                                                            \
    RBD_NVTX(__func__); return launch_fb_rnea<T>(m, B, q, qd, qdd, gravity, c, v, a, f, stream);                                         \
  }                                                                                                                  \
  int rbd_fb_rnea_grad_##SUF(const rbd_fb_model_t* m, int64_t B, const T* q, const T* qd, const T* qdd, T gravity,   \
                             int use_velocity_damping, T* dc_du, T* c_out, void* stream) {                           \
    RBD_NVTX(__func__); return launch_fb_rnea_grad<T>(m, B, q, qd, qdd, gravity, use_velocity_damping, dc_du, c_out, stream);            \
  }                                                                                                                  \
  int rbd_fb_minv_##SUF(const rbd_fb_model_t* m, int64_t B, const T* q, int output_dense, T* Minv, void* stream) {   \
    RBD_NVTX(__func__); return launch_fb_minv<T>(m, B, q, output_dense, Minv, stream);                                                   \
  }                                                                                                                  \
  int rbd_fb_rnea_fpass_##SUF(const rbd_fb_model_t* m, int64_t B, const T* q, const T* qd, const T* qdd, T gravity,  \
                              T* v, T* a, T* f, void* stream) {                                                      \
    RBD_NVTX(__func__); return launch_fb_rnea_fpass<T>(m, B, q, qd, qdd, gravity, v, a, f, stream);                                      \
  }                                                                                                                  \
  int rbd_fb_rnea_bpass_##SUF(const rbd_fb_model_t* m, int64_t B, const T* q, T* f, T* c, void* stream) {            \
    RBD_NVTX(__func__); return launch_fb_rnea_bpass<T>(m, B, q, f, c, stream);                                                           \
  }                                                                                                                  \
  int rbd_fb_minv_bpass_##SUF(const rbd_fb_model_t* m, int64_t B, const T* q, T* Minv, T* F, T* U, T* Dinv,          \
                              void* stream) {                                                                        \
    RBD_NVTX(__func__); return launch_fb_minv_bpass<T>(m, B, q, Minv, F, U, Dinv, stream);                                               \
  }                                                                                                                  \
  int rbd_fb_minv_fpass_##SUF(const rbd_fb_model_t* m, int64_t B, const T* q, T* Minv, T* F, const T* U,             \
                              const T* Dinv, void* stream) {                                                         \
    RBD_NVTX(__func__); return launch_fb_minv_fpass<T>(m, B, q, Minv, F, U, Dinv, stream);                                               \
  }                                                                                                                  \
  int rbd_fb_rnea_grad_fpass_dq_##SUF(const rbd_fb_model_t* m, int64_t B, const T* q, const T* qd, const T* v,       \
                                      const T* a, T gravity, T* dv, T* da, T* df, void* stream) {                    \
    RBD_NVTX(__func__); return launch_fb_grad_fpass<T, true>(m, B, q, qd, v, a, gravity, dv, da, df, stream);                            \
  }                                                                                                                  \
  int rbd_fb_rnea_grad_fpass_dqd_##SUF(const rbd_fb_model_t* m, int64_t B, const T* q, const T* qd, const T* v,      \
                                       T* dv, T* da, T* df, void* stream) {                                          \
    RBD_NVTX(__func__); return launch_fb_grad_fpass<T, false>(m, B, q, qd, v, nullptr, T(0), dv, da, df, stream);                        \
  }                                                                                                                  \
  int rbd_fb_rnea_grad_bpass_dq_##SUF(const rbd_fb_model_t* m, int64_t B, const T* q, const T* f, T* df_dq,          \
                                      T* dc_dq, void* stream) {                                                      \
    RBD_NVTX(__func__); return launch_fb_grad_bpass<T, true>(m, B, q, f, df_dq, 0, dc_dq, stream);                                       \
  }                                                                                                                  \
  int rbd_fb_rnea_grad_bpass_dqd_##SUF(const rbd_fb_model_t* m, int64_t B, const T* q, T* df_dqd,                    \
                                       int use_velocity_damping, T* dc_dqd, void* stream) {                          \
    RBD_NVTX(__func__); return launch_fb_grad_bpass<T, false>(m, B, q, nullptr, df_dqd, use_velocity_damping, dc_dqd, stream);           \
  }                                                                                                                  \
  int rbd_fb_forward_dynamics_##SUF(const rbd_fb_model_t* m, int64_t B, const T* q, const T* qd, const T* u, T* qdd, \
                                    T* Minv_out, void* stream) {                                                     \
    RBD_NVTX(__func__); return launch_fb_forward_dynamics<T>(m, B, q, qd, u, qdd, Minv_out, stream);                                     \
  }                                                                                                                  \
  int rbd_fb_forward_dynamics_grad_##SUF(const rbd_fb_model_t* m, int64_t B, const T* q, const T* qd, const T* u,    \
                                         T* qdd_dq, T* qdd_dqd, T* qdd_out, void* stream) {                          \
    RBD_NVTX(__func__); return launch_fb_forward_dynamics_grad<T>(m, B, q, qd, u, qdd_dq, qdd_dqd, qdd_out, stream);                     \
  }

RBD_FB_DEFINE(f64, double)
RBD_FB_DEFINE(f32, float)

}  // extern "C"
