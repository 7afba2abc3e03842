// rbd_launch_ee.cu - part of librbd_b200.so: end-effector pose / pose gradient (include/rbd_b200.h).
// Reference: end_effector_pose RBDReference.py:220-283, end_effector_pose_gradient :295-386.
#include "rbd_internal.cuh"
#include "rbd_ee_kernels.cuh"

using namespace rbd;
using namespace rbd_host;

struct rbd_ee_model {
  EeModel<double> d;
  EeModel<float> f;
};

namespace {

template <typename T> const EeModel<T>& pick_ee(const rbd_ee_model* m);
template <> const EeModel<double>& pick_ee<double>(const rbd_ee_model* m) { return m->d; }
template <> const EeModel<float>& pick_ee<float>(const rbd_ee_model* m) { return m->f; }

template <typename T>
void fill_ee(const RbdEeDesc* d, EeModel<T>& out) {
  std::memset(&out, 0, sizeof(out));
  out.n = d->n;
  out.n_ee = d->n_ee;
  for (int k = 0; k < 4; ++k) out.off[k] = (T)d->offset[k];
  for (int i = 0; i < d->n; ++i) {
    out.kind[i] = d->kind[i];
    for (int k = 0; k < 12; ++k) {
      out.TA[i][k] = (T)d->TA[12 * i + k]; out.TB[i][k] = (T)d->TB[12 * i + k]; out.TC[i][k] = (T)d->TC[12 * i + k];
      out.DA[i][k] = (T)d->DA[12 * i + k]; out.DB[i][k] = (T)d->DB[12 * i + k]; out.DC[i][k] = (T)d->DC[12 * i + k];
    }
  }
  for (int e = 0; e < d->n_ee; ++e) {
    int tmp[RBD_MAX_DOF], len = 0;
    for (int j = d->ee_joint[e]; j >= 0; j = d->parent[j]) tmp[len++] = j;     // leaf -> base (:238-242)
    out.chain_len[e] = len;
    for (int j = 0; j < RBD_MAX_DOF; ++j) out.chain_pos[e][j] = -1;
    for (int t = 0; t < len; ++t) {
      out.chain[e][t] = (unsigned char)tmp[len - 1 - t];
      out.chain_pos[e][tmp[len - 1 - t]] = (signed char)t;
    }
    for (int k = 0; k < 12; ++k) out.fin[e][k] = (T)d->ee_final[12 * e + k];
  }
}

template <typename T, bool GRAD>
int launch_ee(const rbd_ee_model* m, int64_t B, const T* q, T* pose, T* grad, void* stream, const char* what) {
  RBD_CHECK_ARGS(m && q && B >= 0 && (GRAD ? grad != nullptr : pose != nullptr), "end_effector_pose: null argument or negative B");
  if (B == 0) return 0;
  const EeModel<T>& em = pick_ee<T>(m);
  const int n = em.n, n_ee = em.n_ee;
  // several end effectors: the tile holds only the chain's columns and the copy-out expands them (Atlas: 46 KB
  // -> 16 KB per warp); one end effector: dense tile, the warp's slab leaves in one linear pass
  const int compact = n_ee > 1 ? 1 : 0;
  int maxlen = 1;
  for (int e = 0; e < n_ee; ++e) maxlen = em.chain_len[e] > maxlen ? em.chain_len[e] : maxlen;
  const int pitch = (6 * (compact ? maxlen : n)) | 1;             // odd pitch: lanes hit distinct banks
  const size_t per_warp = ((size_t)32 * n_ee * 6 + (GRAD ? (size_t)32 * pitch : 0)) * sizeof(T);
  // joint coefficients in shared memory for small robots (iiwa14: +13 % FP64, +52 % FP32); for large ones the
  // 72 n values per CTA cost more occupancy than the constant-bank stalls they remove (Atlas FP64: -33 %)
  const bool coef_smem = (size_t)n * 72 * sizeof(T) <= 9 * 1024;
  const size_t coef_bytes = coef_smem ? (size_t)((n * 72 + 1) & ~1) * sizeof(T) : 0;
  if (per_warp + coef_bytes > kMaxDynSmem) return fail(RBD_E_UNSUPPORTED, "end_effector_pose: too many end effectors for the staging tile");
  // warps per CTA that put the most warps on an SM (227 KB of shared memory, 1 KB reserved per CTA)
  int warps = 1, best = 0, resident_ctas = 1;
  for (int w = 1; w <= kEeMaxWarps; ++w) {
    const size_t cta = coef_bytes + (size_t)w * per_warp + 2048;          // static chain table + the per-CTA reserve
    if (cta > kMaxDynSmem) break;
    const int resident = (int)((size_t)(228 * 1024) / cta) * w;
    if (resident > best) { best = resident; warps = w; resident_ctas = resident / w; }
  }
  auto kern = coef_smem ? ee_pose_kernel<T, GRAD, true> : ee_pose_kernel<T, GRAD, false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(coef_bytes + per_warp * warps));   // + 1 KB static
  if (e != cudaSuccess) return fail((int)e, cudaGetErrorString(e));
  const int64_t ntask = (B + 31) / 32;
  int64_t blocks = (ntask + warps - 1) / warps;
  const int sms = rbd_host::sm_count();
  // a few waves of CTAs: exactly one wave (fully persistent) measured 20 % slower on iiwa14 (tail imbalance)
  if (blocks > (int64_t)sms * resident_ctas * 8) blocks = (int64_t)sms * resident_ctas * 8;
  kern<<<(unsigned)blocks, warps * 32, coef_bytes + per_warp * warps, (cudaStream_t)stream>>>(em, B, q, pose, grad, pitch, compact);
  return cuda_status(what);
}

}  // namespace

extern "C" {

int rbd_ee_model_create(const RbdEeDesc* d, rbd_ee_model_t** out) {
  if (!d || !out) return fail(RBD_E_INVALID_ARGUMENT, "rbd_ee_model_create: null argument");
  *out = nullptr;
  if (d->n < 1 || d->n > RBD_MAX_DOF) return fail(RBD_E_UNSUPPORTED, "rbd_ee_model_create: n outside 1..RBD_MAX_DOF");
  if (d->n_ee < 1 || d->n_ee > RBD_MAX_EE) return fail(RBD_E_UNSUPPORTED, "rbd_ee_model_create: n_ee outside 1..RBD_MAX_EE");
  if (!d->parent || !d->kind || !d->TA || !d->TB || !d->TC || !d->DA || !d->DB || !d->DC || !d->ee_joint || !d->ee_final)
    return fail(RBD_E_INVALID_ARGUMENT, "rbd_ee_model_create: null table pointer");
  for (int i = 0; i < d->n; ++i)
    if (d->parent[i] < -1 || d->parent[i] >= i)
      return fail(RBD_E_INVALID_ARGUMENT, "rbd_ee_model_create: parent[i] must satisfy -1 <= parent[i] < i");
  for (int e = 0; e < d->n_ee; ++e)
    if (d->ee_joint[e] < 0 || d->ee_joint[e] >= d->n)
      return fail(RBD_E_INVALID_ARGUMENT, "rbd_ee_model_create: ee_joint out of range");
  rbd_ee_model* m = new (std::nothrow) rbd_ee_model;
  if (!m) return fail(RBD_E_INVALID_ARGUMENT, "rbd_ee_model_create: out of host memory");
  fill_ee<double>(d, m->d);
  fill_ee<float>(d, m->f);
  *out = m;
  return 0;
}

int rbd_ee_model_destroy(rbd_ee_model_t* m) {
  delete m;
  return 0;
}

int rbd_ee_model_num_ee(const rbd_ee_model_t* m) { return m ? m->d.n_ee : RBD_E_INVALID_ARGUMENT; }

int rbd_end_effector_pose_f64(const rbd_ee_model_t* m, int64_t B, const double* q, double* pose, void* stream) {
  RBD_NVTX(__func__); return launch_ee<double, false>(m, B, q, pose, nullptr, stream, "rbd_end_effector_pose");
}
int rbd_end_effector_pose_f32(const rbd_ee_model_t* m, int64_t B, const float* q, float* pose, void* stream) {
  RBD_NVTX(__func__); return launch_ee<float, false>(m, B, q, pose, nullptr, stream, "rbd_end_effector_pose");
}
int rbd_end_effector_pose_gradient_f64(const rbd_ee_model_t* m, int64_t B, const double* q, double* dpose,
                                       double* pose, void* stream) {
  RBD_NVTX(__func__); return launch_ee<double, true>(m, B, q, pose, dpose, stream, "rbd_end_effector_pose_gradient");
}
int rbd_end_effector_pose_gradient_f32(const rbd_ee_model_t* m, int64_t B, const float* q, float* dpose, float* pose,
                                       void* stream) {
  RBD_NVTX(__func__); return launch_ee<float, true>(m, B, q, pose, dpose, stream, "rbd_end_effector_pose_gradient");
}

}  // extern "C"
