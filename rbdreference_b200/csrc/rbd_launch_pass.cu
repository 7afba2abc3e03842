// rbd_launch_pass.cu - part of librbd_b200.so (see rbd_internal.cuh); compiled with -DRBD_LAUNCH_T=double|float.
// Launchers of the four gradient passes (RBDReference.py:1127-1343).
#include "rbd_internal.cuh"
#include "rbd_pass_kernels.cuh"
#include "rbd_coop_pass_kernels.cuh"

#ifndef RBD_LAUNCH_T
#error "compile with -DRBD_LAUNCH_T=double or -DRBD_LAUNCH_T=float"
#endif

using namespace rbd;

namespace rbd_host {

// warps per CTA that keep the most warps resident; 0 if one warp's tile does not fit
template <typename K>
static int cp_geometry(K kern, size_t per_warp, size_t* smem_out, int* ctas_out) {
  if (per_warp > kMaxDynSmem) return 0;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDynSmem) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  int warps = 0, best = 0;
  for (int w = 1; w <= kCpMaxWarps; ++w) {
    const size_t sz = per_warp * w;
    if (sz > kMaxDynSmem) break;
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, w * 32, sz) != cudaSuccess) { cudaGetLastError(); continue; }
    if (nb * w >= best) { best = nb * w; warps = w; *smem_out = sz; *ctas_out = nb; }
  }
  return warps;
}

static int64_t cp_blocks(int64_t ngroups, int warps, int ctas) {
  const int sms = rbd_host::sm_count();
  int64_t blocks = (ngroups + warps - 1) / warps;
  const int64_t cap = (int64_t)sms * ctas;
  return blocks > cap ? cap : blocks;
}

template <typename T, bool DQ>
int launch_grad_fpass(const rbd_model* m, int64_t B, const T* q, const T* qd, const T* v, const T* a, T g,
                      T* dv, T* da, T* df, void* stream) {
  RBD_CHECK_ARGS(m && q && qd && v && (a || !DQ) && dv && da && df && B >= 0,
                 "rbd_rnea_grad_fpass: null argument or negative B");
  if (B == 0) return 0;
  if (variant_of(m) != 1 && variant_of(m) != 3) {
    // one body per lane, one ancestor distance per round, slabs written in one coalesced pass
    // (rbd_coop_pass_kernels.cuh: grad_fpass_level_kernel)
    const int n = m->d.n;
    const int G = n <= 8 ? 8 : (n <= 16 ? 16 : 32);
    int npairs = 0;
    for (int i = 0; i < n; ++i)
      for (int c = i; c >= 0; c = m->d.parent[c]) ++npairs;
    auto kern = G == 8 ? grad_fpass_level_kernel<T, 8, DQ> : (G == 16 ? grad_fpass_level_kernel<T, 16, DQ> : grad_fpass_level_kernel<T, 32, DQ>);
    // warps per CTA: the choice that keeps the most warps resident per SM (the constants and the slab map are per CTA,
    // the staged inputs and pair results per warp: 25 KB for Atlas, 75 KB for a 32-link chain)
    int warps = 0, best = 0, ctas = 0;
    size_t smem = 0;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDynSmem) == cudaSuccess) {
      for (int w = 1; w <= kCpLvlMaxWarps; ++w) {
        const size_t sz = cp_level_head_bytes(n, sizeof(T), G) + (size_t)cp_level_warp_vals(n, npairs, G) * sizeof(T) * w;
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
      kern<<<(unsigned)blocks, warps * 32, smem, (cudaStream_t)stream>>>(pick<T>(m), npairs, B, q, qd, v, a, g, dv, da, df);
      return cuda_status("rbd_rnea_grad_fpass(level)");
    }
    cudaGetLastError();
  }
  if (variant_of(m) != 1) {
    // one derivative column per lane, tensors staged through a shared-memory tile
    const int n = m->d.n;
    const int G = n <= 8 ? 8 : (n <= 16 ? 16 : 32);
    auto kern = G == 8 ? grad_fpass_coop_kernel<T, 8, DQ> : (G == 16 ? grad_fpass_coop_kernel<T, 16, DQ> : grad_fpass_coop_kernel<T, 32, DQ>);
    size_t smem = 0;
    int ctas = 0;
    const int warps = cp_geometry(kern, (size_t)cp_fpass_warp_vals(n, G) * sizeof(T), &smem, &ctas);
    if (warps > 0) {
      const int64_t ngroups = (B + 32 / G - 1) / (32 / G);
      kern<<<(unsigned)cp_blocks(ngroups, warps, ctas), warps * 32, smem, (cudaStream_t)stream>>>(
          pick<T>(m), B, q, qd, v, a, g, dv, da, df);
      return cuda_status("rbd_rnea_grad_fpass(coop)");
    }
  }
  rnea_grad_fpass_kernel<T, DQ><<<blocks_for(B, kPassThreads), kPassThreads, 0, (cudaStream_t)stream>>>(
      pick<T>(m), B, q, qd, v, a, g, dv, da, df);
  return cuda_status("rbd_rnea_grad_fpass");
}

template <typename T, bool DQ>
int launch_grad_bpass(const rbd_model* m, int64_t B, const T* q, const T* f, T* df, int damp, T* dc,
                      void* stream) {
  RBD_CHECK_ARGS(m && q && (f || !DQ) && df && dc && B >= 0, "rbd_rnea_grad_bpass: null argument or negative B");
  if (B == 0) return 0;
  if (variant_of(m) != 1) {
    const int n = m->d.n;
    const int G = n <= 8 ? 8 : (n <= 16 ? 16 : 32);
    auto kern = G == 8 ? grad_bpass_coop_kernel<T, 8, DQ> : (G == 16 ? grad_bpass_coop_kernel<T, 16, DQ> : grad_bpass_coop_kernel<T, 32, DQ>);
    size_t smem = 0;
    int ctas = 0;
    const int warps = cp_geometry(kern, (size_t)cp_bpass_warp_vals(n, G) * sizeof(T), &smem, &ctas);
    if (warps > 0) {
      const int64_t ngroups = (B + 32 / G - 1) / (32 / G);
      kern<<<(unsigned)cp_blocks(ngroups, warps, ctas), warps * 32, smem, (cudaStream_t)stream>>>(
          pick<T>(m), B, q, f, df, damp, dc);
      return cuda_status("rbd_rnea_grad_bpass(coop)");
    }
  }
  rnea_grad_bpass_kernel<T, DQ><<<blocks_for(B, kPassThreads), kPassThreads, 0, (cudaStream_t)stream>>>(
      pick<T>(m), B, q, f, df, damp, dc);
  return cuda_status("rbd_rnea_grad_bpass");
}

#define RBD_INST(DQ)                                                                                                  \
  template int launch_grad_fpass<RBD_LAUNCH_T, DQ>(const rbd_model*, int64_t, const RBD_LAUNCH_T*, const RBD_LAUNCH_T*, \
                                                   const RBD_LAUNCH_T*, const RBD_LAUNCH_T*, RBD_LAUNCH_T, RBD_LAUNCH_T*, \
                                                   RBD_LAUNCH_T*, RBD_LAUNCH_T*, void*);                               \
  template int launch_grad_bpass<RBD_LAUNCH_T, DQ>(const rbd_model*, int64_t, const RBD_LAUNCH_T*, const RBD_LAUNCH_T*, \
                                                   RBD_LAUNCH_T*, int, RBD_LAUNCH_T*, void*);
RBD_INST(true)
RBD_INST(false)
#undef RBD_INST

}  // namespace rbd_host
