// rbd_capi.cu - C ABI of librbd_b200.so (declared in include/rbd_b200.h).
// Plain pointers and sizes only; no torch types; never synchronises; no CPU fallback.
#include <cstdio>
#include <cstring>
#include <atomic>
#include <new>

#include "../../include/rbd_b200.h"
#include "rbd_common.cuh"
#include "rbd_fused_kernels.cuh"
#include "rbd_pass_kernels.cuh"

using namespace rbd;

struct rbd_model {
  DevModel<double> d;
  DevModel<float> f;
};

namespace {

thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

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

template <typename T> const DevModel<T>& pick(const rbd_model* m);
template <> const DevModel<double>& pick<double>(const rbd_model* m) { return m->d; }
template <> const DevModel<float>& pick<float>(const rbd_model* m) { return m->f; }

inline unsigned blocks_for(int64_t B, int threads) { return (unsigned)((B + threads - 1) / threads); }

#define RBD_CHECK_ARGS(cond, msg) \
  do { if (!(cond)) return fail(RBD_E_INVALID_ARGUMENT, msg); } while (0)

template <typename T>
int launch_rnea(const rbd_model* m, int64_t B, const T* q, const T* qd, const T* qdd, T g, T* c, T* v, T* a,
                T* f, void* stream) {
  RBD_CHECK_ARGS(m && q && qd && c && B >= 0, "rbd_rnea: null model/q/qd/c or negative B");
  if (B == 0) return 0;
  rnea_fused_kernel<T><<<blocks_for(B, kFusedThreads), kFusedThreads, 0, (cudaStream_t)stream>>>(
      pick<T>(m), B, q, qd, qdd, g, c, v, a, f);
  return cuda_status("rbd_rnea");
}

template <typename T>
int launch_rnea_grad(const rbd_model* m, int64_t B, const T* q, const T* qd, const T* qdd, T g, int damp,
                     T* dc_du, T* c_out, void* stream) {
  RBD_CHECK_ARGS(m && q && qd && dc_du && B >= 0, "rbd_rnea_grad: null model/q/qd/dc_du or negative B");
  if (B == 0) return 0;
  rnea_grad_fused_kernel<T><<<blocks_for(B, kFusedThreads), kFusedThreads, 0, (cudaStream_t)stream>>>(
      pick<T>(m), B, q, qd, qdd, g, damp, dc_du, c_out);
  return cuda_status("rbd_rnea_grad");
}

template <typename T>
int launch_minv(const rbd_model* m, int64_t B, const T* q, int dense, T* Minv, void* stream) {
  RBD_CHECK_ARGS(m && q && Minv && B >= 0, "rbd_minv: null model/q/Minv or negative B");
  if (B == 0) return 0;
  minv_fused_kernel<T><<<blocks_for(B, kFusedThreads), kFusedThreads, 0, (cudaStream_t)stream>>>(
      pick<T>(m), B, q, dense, Minv);
  return cuda_status("rbd_minv");
}

template <typename T>
int launch_rnea_fpass(const rbd_model* m, int64_t B, const T* q, const T* qd, const T* qdd, T g, T* v, T* a,
                      T* f, void* stream) {
  RBD_CHECK_ARGS(m && q && qd && v && a && f && B >= 0, "rbd_rnea_fpass: null argument or negative B");
  if (B == 0) return 0;
  rnea_fpass_kernel<T><<<blocks_for(B, kPassThreads), kPassThreads, 0, (cudaStream_t)stream>>>(
      pick<T>(m), B, q, qd, qdd, g, v, a, f);
  return cuda_status("rbd_rnea_fpass");
}

template <typename T>
int launch_rnea_bpass(const rbd_model* m, int64_t B, const T* q, T* f, T* c, void* stream) {
  RBD_CHECK_ARGS(m && q && f && c && B >= 0, "rbd_rnea_bpass: null argument or negative B");
  if (B == 0) return 0;
  rnea_bpass_kernel<T><<<blocks_for(B, kPassThreads), kPassThreads, 0, (cudaStream_t)stream>>>(
      pick<T>(m), B, q, f, c);
  return cuda_status("rbd_rnea_bpass");
}

template <typename T, bool DQ>
int launch_grad_fpass(const rbd_model* m, int64_t B, const T* q, const T* qd, const T* v, const T* a, T g,
                      T* dv, T* da, T* df, void* stream) {
  RBD_CHECK_ARGS(m && q && qd && v && (a || !DQ) && dv && da && df && B >= 0,
                 "rbd_rnea_grad_fpass: null argument or negative B");
  if (B == 0) return 0;
  rnea_grad_fpass_kernel<T, DQ><<<blocks_for(B, kPassThreads), kPassThreads, 0, (cudaStream_t)stream>>>(
      pick<T>(m), B, q, qd, v, a, g, dv, da, df);
  return cuda_status("rbd_rnea_grad_fpass");
}

template <typename T, bool DQ>
int launch_grad_bpass(const rbd_model* m, int64_t B, const T* q, const T* f, T* df, int damp, T* dc,
                      void* stream) {
  RBD_CHECK_ARGS(m && q && (f || !DQ) && df && dc && B >= 0, "rbd_rnea_grad_bpass: null argument or negative B");
  if (B == 0) return 0;
  rnea_grad_bpass_kernel<T, DQ><<<blocks_for(B, kPassThreads), kPassThreads, 0, (cudaStream_t)stream>>>(
      pick<T>(m), B, q, f, df, damp, dc);
  return cuda_status("rbd_rnea_grad_bpass");
}

template <typename T>
int launch_minv_bpass(const rbd_model* m, int64_t B, const T* q, T* Minv, T* F, T* U, T* Dinv, void* stream) {
  RBD_CHECK_ARGS(m && q && Minv && F && U && Dinv && B >= 0, "rbd_minv_bpass: null argument or negative B");
  if (B == 0) return 0;
  minv_bpass_kernel<T><<<blocks_for(B, kPassThreads), kPassThreads, 0, (cudaStream_t)stream>>>(
      pick<T>(m), B, q, Minv, F, U, Dinv);
  return cuda_status("rbd_minv_bpass");
}

template <typename T>
int launch_minv_fpass(const rbd_model* m, int64_t B, const T* q, T* Minv, T* F, const T* U, const T* Dinv,
                      void* stream) {
  RBD_CHECK_ARGS(m && q && Minv && F && U && Dinv && B >= 0, "rbd_minv_fpass: null argument or negative B");
  if (B == 0) return 0;
  minv_fpass_kernel<T><<<blocks_for(B, kPassThreads), kPassThreads, 0, (cudaStream_t)stream>>>(
      pick<T>(m), B, q, Minv, F, U, Dinv);
  return cuda_status("rbd_minv_fpass");
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
  *out = m;
  return 0;
}

int rbd_model_destroy(rbd_model_t* m) {
  delete m;
  return 0;
}

int rbd_model_num_dof(const rbd_model_t* m) { return m ? m->d.n : RBD_E_INVALID_ARGUMENT; }

#define RBD_DEFINE(SUF, T)                                                                                           \
  int rbd_rnea_##SUF(const rbd_model_t* m, int64_t B, const T* q, const T* qd, const T* qdd, T gravity, T* c, T* v,  \
                     T* a, T* f, void* stream) {                                                                     \
    return launch_rnea<T>(m, B, q, qd, qdd, gravity, c, v, a, f, stream);                                            \
  }                                                                                                                  \
  int rbd_rnea_grad_##SUF(const rbd_model_t* m, int64_t B, const T* q, const T* qd, const T* qdd, T gravity,         \
                          int use_velocity_damping, T* dc_du, T* c_out, void* stream) {                              \
    return launch_rnea_grad<T>(m, B, q, qd, qdd, gravity, use_velocity_damping, dc_du, c_out, stream);               \
  }                                                                                                                  \
  int rbd_minv_##SUF(const rbd_model_t* m, int64_t B, const T* q, int output_dense, T* Minv, void* stream) {         \
    return launch_minv<T>(m, B, q, output_dense, Minv, stream);                                                      \
  }                                                                                                                  \
  int rbd_rnea_fpass_##SUF(const rbd_model_t* m, int64_t B, const T* q, const T* qd, const T* qdd, T gravity, T* v,  \
                           T* a, T* f, void* stream) {                                                               \
    return launch_rnea_fpass<T>(m, B, q, qd, qdd, gravity, v, a, f, stream);                                         \
  }                                                                                                                  \
  int rbd_rnea_bpass_##SUF(const rbd_model_t* m, int64_t B, const T* q, T* f, T* c, void* stream) {                  \
    return launch_rnea_bpass<T>(m, B, q, f, c, stream);                                                              \
  }                                                                                                                  \
  int rbd_rnea_grad_fpass_dq_##SUF(const rbd_model_t* m, int64_t B, const T* q, const T* qd, const T* v,             \
                                   const T* a, T gravity, T* dv, T* da, T* df, void* stream) {                       \
    return launch_grad_fpass<T, true>(m, B, q, qd, v, a, gravity, dv, da, df, stream);                               \
  }                                                                                                                  \
  int rbd_rnea_grad_fpass_dqd_##SUF(const rbd_model_t* m, int64_t B, const T* q, const T* qd, const T* v, T* dv,     \
                                    T* da, T* df, void* stream) {                                                    \
    return launch_grad_fpass<T, false>(m, B, q, qd, v, nullptr, T(0), dv, da, df, stream);                           \
  }                                                                                                                  \
  int rbd_rnea_grad_bpass_dq_##SUF(const rbd_model_t* m, int64_t B, const T* q, const T* f, T* df_dq, T* dc_dq,      \
                                   void* stream) {                                                                   \
    return launch_grad_bpass<T, true>(m, B, q, f, df_dq, 0, dc_dq, stream);                                          \
  }                                                                                                                  \
  int rbd_rnea_grad_bpass_dqd_##SUF(const rbd_model_t* m, int64_t B, const T* q, T* df_dqd,                          \
                                    int use_velocity_damping, T* dc_dqd, void* stream) {                             \
    return launch_grad_bpass<T, false>(m, B, q, nullptr, df_dqd, use_velocity_damping, dc_dqd, stream);              \
  }                                                                                                                  \
  int rbd_minv_bpass_##SUF(const rbd_model_t* m, int64_t B, const T* q, T* Minv, T* F, T* U, T* Dinv,                \
                           void* stream) {                                                                           \
    return launch_minv_bpass<T>(m, B, q, Minv, F, U, Dinv, stream);                                                  \
  }                                                                                                                  \
  int rbd_minv_fpass_##SUF(const rbd_model_t* m, int64_t B, const T* q, T* Minv, T* F, const T* U, const T* Dinv,    \
                           void* stream) {                                                                           \
    return launch_minv_fpass<T>(m, B, q, Minv, F, U, Dinv, stream);                                                  \
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
