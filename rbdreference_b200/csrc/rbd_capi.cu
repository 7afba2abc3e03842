// rbd_capi.cu - C ABI of librbd_b200.so (declared in include/rbd_b200.h).
// Plain pointers and sizes only; no torch types; never synchronises; no CPU fallback.
#include "rbd_internal.cuh"
#include "rbd_fused_kernels.cuh"
#include "rbd_pass_kernels.cuh"
#include "rbd_fd_kernels.cuh"

using namespace rbd;

namespace {
thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};
}  // namespace

namespace rbd_host {

std::atomic<int> g_variant{0};

// shared-memory budget per warp beyond which the local-memory variants are used
// (RBD_SMEM_LIMIT_KB overrides it for experiments)
size_t smem_limit() {
  static size_t v = [] {
    const char* e = std::getenv("RBD_SMEM_LIMIT_KB");
    return (size_t)(e ? std::atoi(e) : 75) * 1024;
  }();
  return v;
}

// Stream-ordered scratch allocations come from one private pool per device that keeps its
// memory between calls (release threshold = max), so steady-state calls never reach the driver's
// allocator and the application's default pool is left alone.
cudaMemPool_t scratch_pool(int dev) {
  static std::mutex mu;
  static cudaMemPool_t pools[64] = {};
  if (dev < 0 || dev >= 64) return nullptr;
  std::lock_guard<std::mutex> lock(mu);
  if (!pools[dev]) {
    cudaMemPoolProps props = {};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = dev;
    cudaMemPool_t p = nullptr;
    if (cudaMemPoolCreate(&p, &props) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    uint64_t keep = UINT64_MAX;
    cudaMemPoolSetAttribute(p, cudaMemPoolAttrReleaseThreshold, &keep);
    pools[dev] = p;
  }
  return pools[dev];
}

int sm_count() {
  static std::atomic<int> cache[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { cudaGetLastError(); return 148; }
  int v = cache[dev].load(std::memory_order_relaxed);
  if (v <= 0) {
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) { cudaGetLastError(); return 148; }
    cache[dev].store(v, std::memory_order_relaxed);
  }
  return v;
}

int fail(int code, const char* msg) {
  std::snprintf(g_err, sizeof(g_err), "%s", msg);
  return code;
}

int cuda_status(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    std::snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

// ---- forward dynamics (RBDReference.py:1369-1384): compositions of the fused drivers plus the
// per-knot-point products of rbd_fd_kernels.cuh; temporaries are stream-ordered pool allocations.
template <typename T, bool SPLIT>
int launch_fd_apply(int variant, int n, int mcols, int64_t B, const T* A, const T* R1, const T* R2, T alpha, T* out0, T* out1,
                    void* stream) {
  const size_t per_knot = (size_t)(n * n + 2 * n * mcols) * sizeof(T);
  // small per-CTA batches: several CTAs per SM keep loads, products and stores of different batches overlapped
  int KB = (int)((size_t)(28 * 1024) / per_knot);
  if (KB < 1) KB = 1;
  if (KB > 32) KB = 32;
  const size_t smem = per_knot * KB;
  if (mcols > 1 && std::is_same<T, double>::value && !R2 && variant != 1) {
    // FP64 tensor-core product (forward_dynamics_grad)
    const FdMmaShape sh = fd_mma_shape(n, mcols);
    const size_t per_knot_m = (size_t)sh.vals * sizeof(double);
    int KM = (int)((size_t)(28 * 1024) / per_knot_m);
    if (KM < 1) KM = 1;
    if (KM > 32) KM = 32;
    auto mk = fd_apply_mma_kernel<SPLIT>;
    cudaError_t e = cudaFuncSetAttribute(mk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDynSmem);
    if (e != cudaSuccess) return fail((int)e, cudaGetErrorString(e));
    int64_t blocks = (B + KM - 1) / KM;
    if (blocks > grid_cap()) blocks = grid_cap();
    mk<<<(unsigned)blocks, kFdMmaThreads, per_knot_m * KM, (cudaStream_t)stream>>>(
        n, mcols, KM, B, (const double*)A, (const double*)R1, (double)alpha, (double*)out0, (double*)out1);
    return cuda_status("rbd_forward_dynamics(apply, mma)");
  }
  if (mcols > 1) {
    // register-tiled product (forward_dynamics_grad)
    const size_t per_knot_t = fd_tiled_vals_per_knot(n, mcols) * sizeof(T);
    int KT = (int)((size_t)(28 * 1024) / per_knot_t);
    if (KT < 1) KT = 1;
    if (KT > 32) KT = 32;
    const size_t smem_t = per_knot_t * KT + 64;
    auto tk = fd_apply_tiled_kernel<T, SPLIT>;
    cudaError_t e = cudaFuncSetAttribute(tk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDynSmem);
    if (e != cudaSuccess) return fail((int)e, cudaGetErrorString(e));
    int64_t blocks = (B + KT - 1) / KT;
    if (blocks > grid_cap()) blocks = grid_cap();
    tk<<<(unsigned)blocks, kFdTiledThreads, smem_t, (cudaStream_t)stream>>>(n, mcols, KT, B, A, R1, R2, alpha, out0, out1);
    return cuda_status("rbd_forward_dynamics(apply, tiled)");
  }
  auto kern = fd_apply_kernel<T, SPLIT>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDynSmem);
  if (e != cudaSuccess) return fail((int)e, cudaGetErrorString(e));
  int64_t blocks = (B + KB - 1) / KB;
  if (blocks > grid_cap()) blocks = grid_cap();
  kern<<<(unsigned)blocks, kFdThreads, smem, (cudaStream_t)stream>>>(n, mcols, KB, B, A, R1, R2, alpha, out0, out1);
  return cuda_status("rbd_forward_dynamics(apply)");
}

template int launch_fd_apply<double, false>(int, int, int, int64_t, const double*, const double*, const double*, double, double*, double*, void*);
template int launch_fd_apply<double, true>(int, int, int, int64_t, const double*, const double*, const double*, double, double*, double*, void*);
template int launch_fd_apply<float, false>(int, int, int, int64_t, const float*, const float*, const float*, float, float*, float*, void*);
template int launch_fd_apply<float, true>(int, int, int, int64_t, const float*, const float*, const float*, float, float*, float*, void*);

}  // namespace rbd_host

using namespace rbd_host;

namespace {

template <typename T>
int launch_minv_bpass(const rbd_model* m, int64_t B, const T* q, T* Minv, T* F, T* U, T* Dinv, void* stream) {
  RBD_CHECK_ARGS(m && q && Minv && F && U && Dinv && B >= 0, "rbd_minv_bpass: null argument or negative B");
  if (B == 0) return 0;
  if (variant_of(m) == 1) {          // knot point per thread (the first version)
    minv_bpass_kernel<T><<<blocks_for(B, kPassThreads), kPassThreads, 0, (cudaStream_t)stream>>>(
        pick<T>(m), B, q, Minv, F, U, Dinv);
    return cuda_status("rbd_minv_bpass");
  }
  if (m->fast_ok && variant_of(m) != 3) {
    // rigid-body inertias: phases 0 and A of the cooperative minv kernel (one body per lane, local world-aligned
    // frames), results rotated into body frames, Minv / F slabs written in one coalesced pass (BPASS = true)
    const FastModel<T>& fm = pick_dfs<T>(m);
    const int n = fm.n;
    const int G = n <= 8 ? 8 : (n <= 16 ? 16 : 32);
    int npairs = 0;
    for (int i = 0; i < n; ++i) npairs += m->coop_minv.depth[i] + 1;
    auto kern = fm.has_prismatic
                    ? (G == 8 ? minv_coop_kernel<T, 8, true, false, false, true>
                              : (G == 16 ? minv_coop_kernel<T, 16, true, false, false, true> : minv_coop_kernel<T, 32, true, false, false, true>))
                    : (G == 8 ? minv_coop_kernel<T, 8, false, false, false, true>
                              : (G == 16 ? minv_coop_kernel<T, 16, false, false, false, true> : minv_coop_kernel<T, 32, false, false, false, true>));
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDynSmem) == cudaSuccess) {
      int warps = 0, best = 0, ctas = 0;
      size_t smem = 0;
      for (int w = 1; w <= kCmMaxWarps; ++w) {
        const size_t sz = coop_minv_smem_bytes<T>(n, G, m->coop.maxdepth, fm.n_slot_a, w, false, npairs);
        if (sz > kMaxDynSmem) break;
        int nb = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, w * 32, sz) != cudaSuccess) { cudaGetLastError(); continue; }
        if (nb * w > best) { best = nb * w; warps = w; smem = sz; ctas = nb; }
      }
      if (warps > 0) {
        const int ipw = 32 / G;
        const int64_t ngroups = (B + ipw - 1) / ipw;
        int64_t blocks = (ngroups + warps - 1) / warps;
        const int64_t cap = (int64_t)sm_count() * ctas * 4;
        if (blocks > cap) blocks = cap;
        kern<<<(unsigned)blocks, warps * 32, smem, (cudaStream_t)stream>>>(fm, m->plan, m->coop, m->coop_minv, B, q, Minv, F, U, Dinv, npairs);
        return cuda_status("rbd_minv_bpass(coop)");
      }
    }
    cudaGetLastError();
  }
  // one column per lane, articulated inertias shared through shared memory
  const int n = m->d.n;
  const int G = n <= 8 ? 8 : (n <= 16 ? 16 : 32);
  const int vals = G == 8 ? minv_bpass_col_warp_vals<8>(n) : (G == 16 ? minv_bpass_col_warp_vals<16>(n) : minv_bpass_col_warp_vals<32>(n));
  const size_t smem = (size_t)(kPassThreads / 32) * vals * sizeof(T);
  auto kern = G == 8 ? minv_bpass_col_kernel<T, 8> : (G == 16 ? minv_bpass_col_kernel<T, 16> : minv_bpass_col_kernel<T, 32>);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail((int)e, cudaGetErrorString(e));
  }
  const int64_t ntask = (B + 32 / G - 1) / (32 / G);
  int64_t blocks = (ntask + kPassThreads / 32 - 1) / (kPassThreads / 32);
  if (blocks > grid_cap()) blocks = grid_cap();
  kern<<<(unsigned)blocks, kPassThreads, smem, (cudaStream_t)stream>>>(pick<T>(m), B, q, Minv, F, U, Dinv);
  return cuda_status("rbd_minv_bpass");
}

template <typename T>
int launch_minv_fpass(const rbd_model* m, int64_t B, const T* q, T* Minv, T* F, const T* U, const T* Dinv,
                      void* stream) {
  RBD_CHECK_ARGS(m && q && Minv && F && U && Dinv && B >= 0, "rbd_minv_fpass: null argument or negative B");
  if (B == 0) return 0;
  if (variant_of(m) == 1) {          // knot point per thread (the first version)
    minv_fpass_kernel<T><<<blocks_for(B, kPassThreads), kPassThreads, 0, (cudaStream_t)stream>>>(
        pick<T>(m), B, q, Minv, F, U, Dinv);
    return cuda_status("rbd_minv_fpass");
  }
  // one column per lane: rows of Minv and F are read and written contiguously
  const int n = m->d.n;
  const int G = n <= 8 ? 8 : (n <= 16 ? 16 : 32);
  const int64_t ntask = (B + 32 / G - 1) / (32 / G);
  int64_t blocks = (ntask + kPassThreads / 32 - 1) / (kPassThreads / 32);
  if (blocks > grid_cap()) blocks = grid_cap();
  auto kern = G == 8 ? minv_fpass_col_kernel<T, 8> : (G == 16 ? minv_fpass_col_kernel<T, 16> : minv_fpass_col_kernel<T, 32>);
  kern<<<(unsigned)blocks, kPassThreads, 0, (cudaStream_t)stream>>>(pick<T>(m), B, q, Minv, F, U, Dinv);
  return cuda_status("rbd_minv_fpass");
}

template <typename T>
int launch_forward_dynamics(const rbd_model* m, int64_t B, const T* q, const T* qd, const T* u, T* qdd, T* Minv_out,
                            void* stream) {
  RBD_CHECK_ARGS(m && q && qd && u && qdd && B >= 0, "rbd_forward_dynamics: null argument or negative B");
  if (B == 0) return 0;
  const int n = m->d.n;
  PoolBuf c((cudaStream_t)stream), Mi((cudaStream_t)stream);
  int rc = c.alloc((size_t)B * n * sizeof(T));
  if (rc) return rc;
  T* Minv = Minv_out;
  if (!Minv) {
    rc = Mi.alloc((size_t)B * n * n * sizeof(T));
    if (rc) return rc;
    Minv = (T*)Mi.p;
  }
  // c = rnea(q, qd) with the S*qdd term skipped (:1370), Minv = minv(q) (:1371), qdd = Minv (u - c) (:1372)
  rc = launch_rnea<T>(m, B, q, qd, nullptr, T(-9.81), (T*)c.p, nullptr, nullptr, nullptr, stream);
  if (rc) return rc;
  rc = launch_minv<T>(m, B, q, 1, Minv, stream);
  if (rc) return rc;
  return launch_fd_apply<T, false>(variant_of(m), n, 1, B, Minv, u, (const T*)c.p, T(1), qdd, nullptr, stream);
}

template <typename T>
int launch_forward_dynamics_grad(const rbd_model* m, int64_t B, const T* q, const T* qd, const T* u, T* qdd_dq,
                                 T* qdd_dqd, T* qdd_out, void* stream) {
  RBD_CHECK_ARGS(m && q && qd && u && qdd_dq && qdd_dqd && B >= 0, "rbd_forward_dynamics_grad: null argument or negative B");
  if (B == 0) return 0;
  const int n = m->d.n;
  PoolBuf Mi((cudaStream_t)stream), dd((cudaStream_t)stream), dc((cudaStream_t)stream);
  int rc = Mi.alloc((size_t)B * n * n * sizeof(T));
  if (rc) return rc;
  rc = dc.alloc((size_t)B * n * 2 * n * sizeof(T));
  if (rc) return rc;
  T* qdd = qdd_out;
  if (!qdd) {
    rc = dd.alloc((size_t)B * n * sizeof(T));
    if (rc) return rc;
    qdd = (T*)dd.p;
  }
  rc = launch_forward_dynamics<T>(m, B, q, qd, u, qdd, (T*)Mi.p, stream);                       // :1377
  if (rc) return rc;
  rc = launch_rnea_grad<T>(m, B, q, qd, qdd, T(-9.81), 0, (T*)dc.p, nullptr, stream);           // :1378
  if (rc) return rc;
  // qdd_dq = -Minv dc_dq, qdd_dqd = -Minv dc_dqd (:1381-1383); the reference recomputes minv (:1381)
  return launch_fd_apply<T, true>(variant_of(m), n, 2 * n, B, (const T*)Mi.p, (const T*)dc.p, nullptr, T(-1), qdd_dq, qdd_dqd, stream);
}

template <typename T>
void fill_model(const RbdModelDesc* d, DevModel<T>& out) {
  std::memset(&out, 0, sizeof(out));
  const int n = d->n;
  out.n = n;
  for (int i = 0; i < n; ++i) {
    out.parent[i] = d->parent[i];
    out.kind[i] = d->kind[i];
    out.damping[i] = (T)(d->damping ? d->damping[i] : 0.0);
    for (int k = 0; k < 6; ++k) out.S[i][k] = (T)d->S[i * 6 + k];
    for (int k = 0; k < 18; ++k) {
      out.XA[i][k] = (T)d->XA[i * 18 + k];
      out.XB[i][k] = (T)d->XB[i * 18 + k];
      out.XC[i][k] = (T)d->XC[i * 18 + k];
    }
    for (int k = 0; k < 36; ++k) out.I[i][k] = (T)d->I[i * 36 + k];
    unsigned anc = 1u << i;
    if (d->parent[i] >= 0) anc |= out.anc_mask[d->parent[i]];
    out.anc_mask[i] = anc;
  }
  for (int i = n - 1; i >= 0; --i) {
    out.sub_mask[i] |= 1u << i;
    if (d->parent[i] >= 0) out.sub_mask[d->parent[i]] |= out.sub_mask[i];
  }
}

}  // namespace

namespace rbd_host {

// ---- FastModel: rigid-body parameters, r(q) coefficients, stash-slot plan ---------------------
bool build_fast_model(const RbdModelDesc* d, FastModel<double>& out) {
  std::memset(&out, 0, sizeof(out));
  const int n = d->n;
  out.n = n;
  out.rigid = 1;
  for (int i = 0; i < n; ++i) {
    out.parent[i] = d->parent[i];
    out.kind[i] = d->kind[i];
    out.slot_a[i] = out.slot_b[i] = -1;
    out.damping[i] = d->damping ? d->damping[i] : 0.0;
    if (d->kind[i] == 1) out.has_prismatic = 1;
    const double* S = d->S + 6 * i;
    const bool ang = S[0] != 0 || S[1] != 0 || S[2] != 0, lin = S[3] != 0 || S[4] != 0 || S[5] != 0;
    if ((ang && lin) || (!ang && !lin) || (ang != (d->kind[i] == 0))) return false;
    for (int k = 0; k < 3; ++k) out.axis[i][k] = ang ? S[k] : S[3 + k];
    const double* XA = d->XA + 18 * i; const double* XB = d->XB + 18 * i; const double* XC = d->XC + 18 * i;
    for (int k = 0; k < 9; ++k) { out.EA[i][k] = XA[k]; out.EB[i][k] = XB[k]; out.EC[i][k] = XC[k]; }
    // r(q) from r x = -E^T L, sampled where (f1, f2) is a valid basis value
    auto r_of = [&](double f1, double f2, double* r) {
      double E[9], L[9], R[9];
      for (int k = 0; k < 9; ++k) { E[k] = XA[k] + f1 * XB[k] + f2 * XC[k]; L[k] = XA[9 + k] + f1 * XB[9 + k] + f2 * XC[9 + k]; }
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) R[3 * a + b] = -(E[a] * L[b] + E[3 + a] * L[3 + b] + E[6 + a] * L[6 + b]);
      r[0] = 0.5 * (R[7] - R[5]); r[1] = 0.5 * (R[2] - R[6]); r[2] = 0.5 * (R[3] - R[1]);
    };
    double r0[3], r1[3], r2[3];
    if (d->kind[i] == 0) {
      r_of(1, 0, r0); r_of(-1, 0, r1); r_of(0, 1, r2);
      for (int k = 0; k < 3; ++k) {
        out.rA[i][k] = 0.5 * (r0[k] + r1[k]); out.rB[i][k] = 0.5 * (r0[k] - r1[k]); out.rC[i][k] = r2[k] - out.rA[i][k];
      }
    } else {
      r_of(0, 0, r0); r_of(1, 0, r1);
      for (int k = 0; k < 3; ++k) { out.rA[i][k] = r0[k]; out.rB[i][k] = r1[k] - r0[k]; out.rC[i][k] = 0; }
    }
    // rigid-body structure of the spatial inertia: [[Ibar, hx], [hx^T, m 1]]
    const double* I = d->I + 36 * i;
    const double mss = I[21];
    const double h[3] = {I[2 * 6 + 4], I[0 * 6 + 5], I[1 * 6 + 3]};
    double ref[36] = {0};
    const double Ib[6] = {I[0], I[1], I[2], I[7], I[8], I[14]};
    ref[0] = Ib[0]; ref[1] = Ib[1]; ref[2] = Ib[2]; ref[6] = Ib[1]; ref[7] = Ib[3]; ref[8] = Ib[4];
    ref[12] = Ib[2]; ref[13] = Ib[4]; ref[14] = Ib[5];
    const double hx[9] = {0, -h[2], h[1], h[2], 0, -h[0], -h[1], h[0], 0};
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) { ref[6 * a + 3 + b] = hx[3 * a + b]; ref[6 * (3 + b) + a] = hx[3 * a + b]; }
    ref[21] = ref[28] = ref[35] = mss;
    double scale = 0, err = 0;
    for (int k = 0; k < 36; ++k) { scale = fmax(scale, fabs(I[k])); err = fmax(err, fabs(I[k] - ref[k])); }
    if (err > 1e-12 * fmax(scale, 1e-300)) out.rigid = 0;
    out.mass[i] = mss;
    for (int k = 0; k < 3; ++k) out.h[i][k] = h[k];
    for (int k = 0; k < 6; ++k) out.Ib[i][k] = Ib[k];
    unsigned anc = 1u << i;
    if (d->parent[i] >= 0) anc |= out.anc_mask[d->parent[i]];
    out.anc_mask[i] = anc;
  }
  for (int i = n - 1; i >= 0; --i) {
    out.sub_mask[i] |= 1u << i;
    if (d->parent[i] >= 0) out.sub_mask[d->parent[i]] |= out.sub_mask[i];
  }
  for (int c = 1; c < n; ++c) {
    const int p = d->parent[c];
    if (p >= 0 && p != c - 1 && out.slot_a[p] < 0) out.slot_a[p] = out.n_slot_a++;
  }
  for (int i = 0; i + 1 < n; ++i)
    if (d->parent[i + 1] != i) out.slot_b[i] = out.n_slot_b++;
  return out.rigid != 0;
}

// Renumber the robot in depth-first preorder and build its FastModel + index tables.
bool build_dfs_model(const RbdModelDesc* d, FastModel<double>& out, DfsPlan& plan) {
  const int n = d->n;
  std::memset(&plan, 0, sizeof(plan));
  int order[RBD_MAX_DOF], cnt = 0, stack[RBD_MAX_DOF], sp = 0;
  for (int r = n - 1; r >= 0; --r)
    if (d->parent[r] < 0) stack[sp++] = r;              // roots, smallest id on top
  while (sp > 0) {
    const int i = stack[--sp];
    order[cnt++] = i;
    for (int c = n - 1; c > i; --c)
      if (d->parent[c] == i) stack[sp++] = c;           // children, smallest id on top
  }
  if (cnt != n) return false;
  int32_t parent[RBD_MAX_DOF], kind[RBD_MAX_DOF];
  double S[RBD_MAX_DOF * 6], XA[RBD_MAX_DOF * 18], XB[RBD_MAX_DOF * 18], XC[RBD_MAX_DOF * 18], I[RBD_MAX_DOF * 36],
      damping[RBD_MAX_DOF];
  for (int k = 0; k < n; ++k) { plan.orig[k] = order[k]; plan.pos[order[k]] = k; }
  for (int k = 0; k < n; ++k) {
    const int o = order[k];
    parent[k] = d->parent[o] < 0 ? -1 : plan.pos[d->parent[o]];
    kind[k] = d->kind[o];
    damping[k] = d->damping ? d->damping[o] : 0.0;
    std::memcpy(S + 6 * k, d->S + 6 * o, 6 * sizeof(double));
    std::memcpy(XA + 18 * k, d->XA + 18 * o, 18 * sizeof(double));
    std::memcpy(XB + 18 * k, d->XB + 18 * o, 18 * sizeof(double));
    std::memcpy(XC + 18 * k, d->XC + 18 * o, 18 * sizeof(double));
    std::memcpy(I + 36 * k, d->I + 36 * o, 36 * sizeof(double));
  }
  RbdModelDesc pd = {n, parent, kind, S, XA, XB, XC, I, damping};
  const bool ok = build_fast_model(&pd, out);
  // subtree / component ranges (preorder => contiguous)
  int size[RBD_MAX_DOF];
  for (int k = 0; k < n; ++k) size[k] = 1;
  for (int k = n - 1; k >= 0; --k)
    if (parent[k] >= 0) size[parent[k]] += size[k];
  for (int k = 0; k < n; ++k) {
    plan.sub_end[k] = k + size[k];
    plan.comp_end[k] = parent[k] < 0 ? k + size[k] : plan.comp_end[parent[k]];
  }
  return ok;
}

// Index tables of the warp-cooperative kernels for a robot in depth-first numbering: pointer-jumping ancestor
// table, root components, depths.
void build_coop_plans(const FastModel<double>& dfs, CoopPlan& cp, CoopMinvPlan& mp) {
  std::memset(&cp, 0, sizeof(cp));
  std::memset(&mp, 0, sizeof(mp));
  const int n = dfs.n;
  int depth[RBD_MAX_DOF];
  for (int i = 0; i < n; ++i) {
    const int p = dfs.parent[i];
    cp.jump[0][i] = p;
    if (p < 0) cp.comp_begin[cp.ncomp++] = i;
    depth[i] = p < 0 ? 0 : depth[p] + 1;
    if (depth[i] > cp.maxdepth) cp.maxdepth = depth[i];
  }
  for (int s = 1; s < 5; ++s)
    for (int i = 0; i < n; ++i) cp.jump[s][i] = cp.jump[s - 1][i] < 0 ? -1 : cp.jump[s - 1][cp.jump[s - 1][i]];
  while ((1 << cp.nsteps) < cp.maxdepth + 1) ++cp.nsteps;
  cp.comp_begin[cp.ncomp] = n;
  for (int i = 0; i < n; ++i) {
    const int p = dfs.parent[i];
    mp.depth[i] = depth[i];
    mp.comp_root[i] = p < 0 ? i : mp.comp_root[p];
    if (i - mp.comp_root[i] + 1 > mp.maxcomp) mp.maxcomp = i - mp.comp_root[i] + 1;
  }
}

void narrow_fast_model(const FastModel<double>& a, FastModel<float>& b) {
  std::memset(&b, 0, sizeof(b));
  b.n = a.n; b.n_slot_a = a.n_slot_a; b.n_slot_b = a.n_slot_b; b.rigid = a.rigid; b.has_prismatic = a.has_prismatic;
  for (int i = 0; i < RBD_MAX_DOF; ++i) {
    b.parent[i] = a.parent[i]; b.kind[i] = a.kind[i]; b.slot_a[i] = a.slot_a[i]; b.slot_b[i] = a.slot_b[i];
    b.anc_mask[i] = a.anc_mask[i]; b.sub_mask[i] = a.sub_mask[i];
    b.damping[i] = (float)a.damping[i]; b.mass[i] = (float)a.mass[i];
    for (int k = 0; k < 9; ++k) { b.EA[i][k] = (float)a.EA[i][k]; b.EB[i][k] = (float)a.EB[i][k]; b.EC[i][k] = (float)a.EC[i][k]; }
    for (int k = 0; k < 3; ++k) {
      b.rA[i][k] = (float)a.rA[i][k]; b.rB[i][k] = (float)a.rB[i][k]; b.rC[i][k] = (float)a.rC[i][k];
      b.axis[i][k] = (float)a.axis[i][k]; b.h[i][k] = (float)a.h[i][k];
    }
    for (int k = 0; k < 6; ++k) b.Ib[i][k] = (float)a.Ib[i][k];
  }
}

}  // namespace rbd_host

namespace {


// Schedule of the tile minv kernel (rbd_tile_minv_kernels.cuh): chains of the depth-first numbering with their
// dependency levels and hand-off slots, per-warp work lists, and the step tables of the column groups.
void build_tile_plan(const FastModel<double>& fm, const DfsPlan& plan, const CoopMinvPlan& mp, int maxdepth, int gc, int maxwarps,
                     TilePlan& tp) {
  std::memset(&tp, 0, sizeof(tp));
  tp.gc = gc;
  const int n = fm.n;
  int chain_of[RBD_MAX_DOF], flevel[RBD_MAX_DOF] = {0}, blevel[RBD_MAX_DOF] = {0};
  for (int i = 0; i < n; ++i) {
    if (i == 0 || fm.parent[i] != i - 1) { tp.chain_begin[tp.nchain] = i; ++tp.nchain; }
    chain_of[i] = tp.nchain - 1;
    tp.chain_end[tp.nchain - 1] = i + 1;
  }
  const int nch = tp.nchain;
  for (int c = 0; c < nch; ++c) {
    const int p = fm.parent[tp.chain_begin[c]];
    flevel[c] = p < 0 ? 0 : flevel[chain_of[p]] + 1;
    if (flevel[c] + 1 > tp.nflevel) tp.nflevel = flevel[c] + 1;
    tp.out_slot[c] = p < 0 ? -1 : tp.nslot++;
  }
  // forward sweep as late as possible (a root chain nobody waits for runs next to the deepest level)
  for (int c = nch - 1; c >= 0; --c) {
    int lv = tp.nflevel - 1;
    for (int d = c + 1; d < nch; ++d) {
      const int p = fm.parent[tp.chain_begin[d]];
      if (p >= 0 && chain_of[p] == c && flevel[d] - 1 < lv) lv = flevel[d] - 1;
    }
    flevel[c] = lv;
  }
  for (int c = nch - 1; c >= 0; --c) {                    // chains hanging off c have larger indices: final before use
    const int p = fm.parent[tp.chain_begin[c]];
    if (p >= 0 && blevel[chain_of[p]] < blevel[c] + 1) blevel[chain_of[p]] = blevel[c] + 1;
    if (blevel[c] + 1 > tp.nblevel) tp.nblevel = blevel[c] + 1;
  }
  int cnt = 0;
  for (int i = 0; i < n; ++i) {
    tp.in_begin[i] = cnt;
    for (int c = 0; c < nch; ++c)
      if (fm.parent[tp.chain_begin[c]] == i) tp.in_slot[cnt++] = tp.out_slot[c];
  }
  tp.in_begin[n] = cnt;
  tp.maxdepth = maxdepth;
  tp.nslot_g = fm.n_slot_a;
  // warps per CTA: as many as the FP64 shared-memory budget allows (the FP32 kernel uses the same schedule)
  int w = maxwarps;
  while (w > 1 && tile_minv_smem_vals(n, tp.nslot, maxdepth, tp.nslot_g, w, gc) * sizeof(double) > rbd_host::kMaxDynSmem - 1024) --w;
  tp.nwarps = w;
  tp.ok = maxdepth < 16 && fm.n_slot_a < 15 &&
          tile_minv_smem_vals(n, tp.nslot, maxdepth, tp.nslot_g, w, gc) * sizeof(double) <= rbd_host::kMaxDynSmem - 1024;
  // per-(level, warp) chain lists: longest chains first, each to the least loaded warp of its level
  auto fill = [&](const int* level, int nlevel, int* begin, int* item) {
    int pos = 0;
    for (int lv = 0; lv < nlevel; ++lv) {
      int load[kTmMaxWarps] = {0}, owner[RBD_MAX_DOF];
      bool used[RBD_MAX_DOF] = {false};
      for (int c = 0; c < nch; ++c) owner[c] = -1;
      for (;;) {
        int best = -1;
        for (int c = 0; c < nch; ++c)
          if (level[c] == lv && !used[c] && (best < 0 || tp.chain_end[c] - tp.chain_begin[c] > tp.chain_end[best] - tp.chain_begin[best])) best = c;
        if (best < 0) break;
        int ww = 0;
        for (int x = 1; x < w; ++x)
          if (load[x] < load[ww]) ww = x;
        used[best] = true;
        owner[best] = ww;
        load[ww] += tp.chain_end[best] - tp.chain_begin[best];
      }
      for (int x = 0; x < w; ++x) {
        begin[lv * w + x] = pos;
        for (int c = 0; c < nch; ++c)
          if (owner[c] == x) item[pos++] = c;
      }
    }
    begin[nlevel * w] = pos;
  };
  fill(flevel, tp.nflevel, tp.f_begin, tp.f_item);
  fill(blevel, tp.nblevel, tp.b_begin, tp.b_item);
  // column groups: up to kTmGC consecutive columns of one root component, with their step tables
  int cost[kTmMaxGroups], ns = 0;
  for (int r = 0; r < n && tp.ok; r = plan.comp_end[r]) {
    const int cend = plan.comp_end[r];
    for (int j0 = r; j0 < cend; j0 += gc) {
      const int g = tp.ngroup++;
      const int nc = cend - j0 < gc ? cend - j0 : gc;
      tp.g_first[g] = j0;
      tp.g_ncols[g] = nc;
      tp.g_ocol[g] = plan.orig[j0];
      for (int c = 1; c < nc; ++c)
        if (plan.orig[j0 + c] != plan.orig[j0] + c) tp.g_ocol[g] = -1;
      if (ns + 3 * n > kTmMaxSteps) { tp.ok = 0; break; }
      tp.g_sb[g] = ns;
      for (int a = j0 + nc - 1; a >= r; --a) {            // phase B: bodies with a column of the group below them
        if (plan.sub_end[a] <= j0) continue;
        int mask = 0, self = 7;
        for (int c = 0; c < nc; ++c) {
          const int j = j0 + c;
          if (j >= a && j < plan.sub_end[a]) mask |= 1 << c;
          if (j == a) self = c;
        }
        tp.steps[ns++] = tm_pack_step(a, mp.depth[a], mask, self, -1, -1, 0, fm.kind[a], plan.orig[a]);
      }
      tp.g_sc[g] = ns;
      for (int a = r; a < cend; ++a) {                    // phase C: the whole component in preorder
        int mask = 0;
        for (int c = 0; c < nc; ++c) {
          const int j = j0 + c;
          if (j >= a && j < plan.sub_end[a]) mask |= 1 << c;
        }
        const int par = fm.parent[a];
        const int psl = (par >= 0 && par != a - 1) ? fm.slot_a[par] : -1;
        tp.steps[ns++] = tm_pack_step(a, mp.depth[a], mask, 7, fm.slot_a[a], psl, par < 0 ? 1 : 0, fm.kind[a], plan.orig[a]);
      }
      tp.g_sz[g] = ns;
      for (int a = 0; a < n; ++a)                         // rows of the other components: zeros
        if (a < r || a >= cend) tp.steps[ns++] = tm_pack_step(a, 0, 0, 7, -1, -1, 0, 0, plan.orig[a]);
      tp.g_se[g] = ns;
      cost[g] = 8 * (tp.g_sz[g] - tp.g_sb[g]) + (tp.g_se[g] - tp.g_sz[g]);
    }
  }
  // longest-processing-time assignment of the groups to the warps
  int load[kTmMaxWarps] = {0}, owner[kTmMaxGroups];
  bool done[kTmMaxGroups] = {false};
  for (int k = 0; k < tp.ngroup; ++k) {
    int best = -1;
    for (int g = 0; g < tp.ngroup; ++g)
      if (!done[g] && (best < 0 || cost[g] > cost[best])) best = g;
    int ww = 0;
    for (int x = 1; x < w; ++x)
      if (load[x] < load[ww]) ww = x;
    done[best] = true;
    owner[best] = ww;
    load[ww] += cost[best];
  }
  int pos = 0;
  for (int x = 0; x < kTmMaxWarps; ++x) {
    tp.g_begin[x] = pos;
    if (x < w)
      for (int g = 0; g < tp.ngroup; ++g)
        if (owner[g] == x) tp.g_item[pos++] = g;
  }
  tp.g_begin[kTmMaxWarps] = pos;
}

template <typename T>
void fill_chain_model(const FastModel<double>& fm, ChainModel<T>& cm) {
  std::memset(&cm, 0, sizeof(cm));
  cm.n = fm.n;
  cm.r_const = 1;
  for (int i = 0; i < fm.n && i < kChainMaxN; ++i) {
    cm.kind[i] = fm.kind[i];
    typename ChainModel<T>::Body& b = cm.b[i];
    for (int k = 0; k < 9; ++k) { b.EA[k] = (T)fm.EA[i][k]; b.EB[k] = (T)fm.EB[i][k]; b.EC[k] = (T)fm.EC[i][k]; }
    for (int k = 0; k < 3; ++k) {
      b.rA[k] = (T)fm.rA[i][k]; b.rB[k] = (T)fm.rB[i][k]; b.rC[k] = (T)fm.rC[i][k];
      b.axis[k] = (T)fm.axis[i][k]; b.h[k] = (T)fm.h[i][k];
      if (fm.rB[i][k] != 0.0 || fm.rC[i][k] != 0.0) cm.r_const = 0;
    }
    for (int k = 0; k < 6; ++k) b.Ib[k] = (T)fm.Ib[i][k];
    b.mass = (T)fm.mass[i];
    b.damping = (T)fm.damping[i];
  }
}

}  // namespace

extern "C" {

int rbd_abi_version(void) { return RBD_ABI_VERSION; }
const char* rbd_last_error_string(void) { return g_err; }
int64_t rbd_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int rbd_model_create(const RbdModelDesc* desc, rbd_model_t** out) {
  if (!desc || !out) return fail(RBD_E_INVALID_ARGUMENT, "rbd_model_create: null argument");
  *out = nullptr;
  if (desc->n < 1 || desc->n > RBD_MAX_DOF) return fail(RBD_E_UNSUPPORTED, "rbd_model_create: n outside 1..RBD_MAX_DOF");
  if (!desc->parent || !desc->kind || !desc->S || !desc->XA || !desc->XB || !desc->XC || !desc->I)
    return fail(RBD_E_INVALID_ARGUMENT, "rbd_model_create: null table pointer");
  for (int i = 0; i < desc->n; ++i) {
    if (desc->parent[i] < -1 || desc->parent[i] >= i)
      return fail(RBD_E_INVALID_ARGUMENT, "rbd_model_create: parent[i] must satisfy -1 <= parent[i] < i");
    if (desc->kind[i] != 0 && desc->kind[i] != 1)
      return fail(RBD_E_INVALID_ARGUMENT, "rbd_model_create: kind[i] must be 0 (revolute) or 1 (prismatic)");
  }
  rbd_model* m = new (std::nothrow) rbd_model;
  if (!m) return fail(RBD_E_INVALID_ARGUMENT, "rbd_model_create: out of host memory");
  fill_model<double>(desc, m->d);
  fill_model<float>(desc, m->f);
  m->fast_ok = build_fast_model(desc, m->fd);
  m->is_chain = true;
  for (int i = 0; i < desc->n; ++i) m->is_chain = m->is_chain && desc->parent[i] == i - 1;
  narrow_fast_model(m->fd, m->ff);
  m->fast_ok = build_dfs_model(desc, m->fd_dfs, m->plan) && m->fast_ok;
  narrow_fast_model(m->fd_dfs, m->ff_dfs);
  build_coop_plans(m->fd_dfs, m->coop, m->coop_minv);
  build_tile_plan(m->fd_dfs, m->plan, m->coop_minv, m->coop.maxdepth, 4, 8, m->tile);
  build_tile_plan(m->fd_dfs, m->plan, m->coop_minv, m->coop.maxdepth, 2, 16, m->tile2);
  fill_chain_model<double>(m->fd, m->chain_d);
  fill_chain_model<float>(m->fd, m->chain_f);
  {
    // create the current device's scratch pool now, so that no call made later under CUDA-graph
    // capture has to create it (pool creation is not allowed while a global-mode capture is open)
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) scratch_pool(dev);
    cudaGetLastError();
  }
  *out = m;
  return 0;
}

int rbd_model_destroy(rbd_model_t* m) {
  delete m;
  return 0;
}

int rbd_model_num_dof(const rbd_model_t* m) { return m ? m->d.n : RBD_E_INVALID_ARGUMENT; }

static const char* kVariantHelp =
    "kernel variant: 0 auto, 1 generic, 2 world (thread per knot point), 3 cooperative, 4 hybrid (minv), 5 lane (minv), "
    "7 chain (rnea_grad, serial chains), 8 tile (minv, large trees), 9 tile with 2-column groups / 16 warps";
int rbd_set_kernel_variant(int variant) {
  if (variant < 0 || variant > 9 || variant == 6) return fail(RBD_E_INVALID_ARGUMENT, kVariantHelp);
  g_variant.store(variant, std::memory_order_relaxed);
  return 0;
}
int rbd_model_set_kernel_variant(rbd_model_t* m, int variant) {
  if (!m || variant < -1 || variant > 9 || variant == 6) return fail(RBD_E_INVALID_ARGUMENT, kVariantHelp);
  m->variant.store(variant, std::memory_order_relaxed);
  return 0;
}
int rbd_model_uses_world_kernels(const rbd_model_t* m) { return (m && m->fast_ok) ? 1 : 0; }

int rbd_prepare_device(int device) {
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) {
    cudaGetLastError();
    return fail(RBD_E_NO_DEVICE, "rbd_prepare_device: no such CUDA device");
  }
  return scratch_pool(device) ? 0 : fail(RBD_E_NO_DEVICE, "rbd_prepare_device: cannot create the scratch memory pool");
}

int rbd_trim_scratch(int64_t keep_bytes) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return fail(RBD_E_NO_DEVICE, "rbd_trim_scratch: no CUDA device"); }
  cudaMemPool_t pool = scratch_pool(dev);
  if (!pool) return fail(RBD_E_NO_DEVICE, "rbd_trim_scratch: cannot create the scratch memory pool");
  cudaError_t e = cudaMemPoolTrimTo(pool, keep_bytes < 0 ? 0 : (size_t)keep_bytes);
  if (e != cudaSuccess) return fail((int)e, cudaGetErrorString(e));
  return 0;
}

#define RBD_DEFINE(SUF, T)                                                                                           \
  int rbd_rnea_##SUF(const rbd_model_t* m, int64_t B, const T* q, const T* qd, const T* qdd, T gravity, T* c, T* v,  \
                     T* a, T* f, void* stream) {                                                                     \
    RBD_NVTX(__func__); return launch_rnea<T>(m, B, q, qd, qdd, gravity, c, v, a, f, stream);                                            \
  }                                                                                                                  \
  int rbd_rnea_grad_##SUF(const rbd_model_t* m, int64_t B, const T* q, const T* qd, const T* qdd, T gravity,         \
                          int use_velocity_damping, T* dc_du, T* c_out, void* stream) {                              \
    RBD_NVTX(__func__); return launch_rnea_grad<T>(m, B, q, qd, qdd, gravity, use_velocity_damping, dc_du, c_out, stream);               \
  }                                                                                                                  \
  int rbd_minv_##SUF(const rbd_model_t* m, int64_t B, const T* q, int output_dense, T* Minv, void* stream) {         \
    RBD_NVTX(__func__); return launch_minv<T>(m, B, q, output_dense, Minv, stream);                                                      \
  }                                                                                                                  \
  int rbd_crba_##SUF(const rbd_model_t* m, int64_t B, const T* q, T* H, void* stream) {                              \
    RBD_NVTX(__func__); return launch_crba<T>(m, B, q, H, stream);                                                                       \
  }                                                                                                                  \
  int rbd_aba_##SUF(const rbd_model_t* m, int64_t B, const T* q, const T* qd, const T* tau, T gravity, T* qdd,       \
                    void* stream) {                                                                                  \
    RBD_NVTX(__func__); return launch_aba<T>(m, B, q, qd, tau, gravity, qdd, stream);                                                    \
  }                                                                                                                  \
  int rbd_rnea_fpass_##SUF(const rbd_model_t* m, int64_t B, const T* q, const T* qd, const T* qdd, T gravity, T* v,  \
                           T* a, T* f, void* stream) {                                                               \
    RBD_NVTX(__func__); return launch_rnea_fpass<T>(m, B, q, qd, qdd, gravity, v, a, f, stream);                                         \
  }                                                                                                                  \
  int rbd_rnea_bpass_##SUF(const rbd_model_t* m, int64_t B, const T* q, T* f, T* c, void* stream) {                  \
    RBD_NVTX(__func__); return launch_rnea_bpass<T>(m, B, q, f, c, stream);                                                              \
  }                                                                                                                  \
  int rbd_rnea_grad_fpass_dq_##SUF(const rbd_model_t* m, int64_t B, const T* q, const T* qd, const T* v,             \
                                   const T* a, T gravity, T* dv, T* da, T* df, void* stream) {                       \
    RBD_NVTX(__func__); return launch_grad_fpass<T, true>(m, B, q, qd, v, a, gravity, dv, da, df, stream);                               \
  }                                                                                                                  \
  int rbd_rnea_grad_fpass_dqd_##SUF(const rbd_model_t* m, int64_t B, const T* q, const T* qd, const T* v, T* dv,     \
                                    T* da, T* df, void* stream) {                                                    \
    RBD_NVTX(__func__); return launch_grad_fpass<T, false>(m, B, q, qd, v, nullptr, T(0), dv, da, df, stream);                           \
  }                                                                                                                  \
  int rbd_rnea_grad_bpass_dq_##SUF(const rbd_model_t* m, int64_t B, const T* q, const T* f, T* df_dq, T* dc_dq,      \
                                   void* stream) {                                                                   \
    RBD_NVTX(__func__); return launch_grad_bpass<T, true>(m, B, q, f, df_dq, 0, dc_dq, stream);                                          \
  }                                                                                                                  \
  int rbd_rnea_grad_bpass_dqd_##SUF(const rbd_model_t* m, int64_t B, const T* q, T* df_dqd,                          \
                                    int use_velocity_damping, T* dc_dqd, void* stream) {                             \
    RBD_NVTX(__func__); return launch_grad_bpass<T, false>(m, B, q, nullptr, df_dqd, use_velocity_damping, dc_dqd, stream);              \
  }                                                                                                                  \
  int rbd_minv_bpass_##SUF(const rbd_model_t* m, int64_t B, const T* q, T* Minv, T* F, T* U, T* Dinv,                \
                           void* stream) {                                                                           \
    RBD_NVTX(__func__); return launch_minv_bpass<T>(m, B, q, Minv, F, U, Dinv, stream);                                                  \
  }                                                                                                                  \
  int rbd_minv_fpass_##SUF(const rbd_model_t* m, int64_t B, const T* q, T* Minv, T* F, const T* U, const T* Dinv,    \
                           void* stream) {                                                                           \
    RBD_NVTX(__func__); return launch_minv_fpass<T>(m, B, q, Minv, F, U, Dinv, stream);                                                  \
  }                                                                                                                  \
  int rbd_forward_dynamics_##SUF(const rbd_model_t* m, int64_t B, const T* q, const T* qd, const T* u, T* qdd,       \
                                 T* Minv_out, void* stream) {                                                        \
    RBD_NVTX(__func__); return launch_forward_dynamics<T>(m, B, q, qd, u, qdd, Minv_out, stream);                                        \
  }                                                                                                                  \
  int rbd_forward_dynamics_grad_##SUF(const rbd_model_t* m, int64_t B, const T* q, const T* qd, const T* u,          \
                                      T* qdd_dq, T* qdd_dqd, T* qdd_out, void* stream) {                             \
    RBD_NVTX(__func__); return launch_forward_dynamics_grad<T>(m, B, q, qd, u, qdd_dq, qdd_dqd, qdd_out, stream);                        \
  }

RBD_DEFINE(f64, double)
RBD_DEFINE(f32, float)

int rbd_measure_fma_peak(int is_f64, double* flops_per_s, double* elapsed_ms, void* stream) {
  if (!flops_per_s) return fail(RBD_E_INVALID_ARGUMENT, "rbd_measure_fma_peak: null output");
  cudaStream_t s = (cudaStream_t)stream;
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
    return fail(RBD_E_NO_DEVICE, "rbd_measure_fma_peak: no CUDA device");
  const int threads = 256, blocks = sms * 8, iters = is_f64 ? 4096 : 8192;
  void* buf = nullptr;
  cudaError_t e = cudaMalloc(&buf, (size_t)threads * blocks * sizeof(double));
  if (e != cudaSuccess) return fail((int)e, cudaGetErrorString(e));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0, s);
    if (is_f64) fma_peak_kernel<double><<<blocks, threads, 0, s>>>((double*)buf, iters, 1.0);
    else fma_peak_kernel<float><<<blocks, threads, 0, s>>>((float*)buf, iters, 1.0f);
    cudaEventRecord(e1, s);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  g_launches.fetch_add(4, std::memory_order_relaxed);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(buf);
  e = cudaGetLastError();
  if (e != cudaSuccess) return fail((int)e, cudaGetErrorString(e));
  const double fmas = (double)threads * blocks * (double)iters * 64.0;
  *flops_per_s = 2.0 * fmas / (best * 1e-3);
  if (elapsed_ms) *elapsed_ms = best;
  return 0;
}

}  // extern "C"
