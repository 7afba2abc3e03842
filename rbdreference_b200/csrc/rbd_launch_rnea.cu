// rbd_launch_rnea.cu - part of librbd_b200.so (see rbd_internal.cuh); compiled with -DRBD_LAUNCH_T=double|float.
#include "rbd_internal.cuh"
#include "rbd_fused_kernels.cuh"
#include "rbd_coop_kernels.cuh"
#include "rbd_pass_kernels.cuh"
#include "rbd_lane_rnea_kernels.cuh"

#ifndef RBD_LAUNCH_T
#error "compile with -DRBD_LAUNCH_T=double or -DRBD_LAUNCH_T=float"
#endif

using namespace rbd;

namespace rbd_host {

template <typename T>
int launch_rnea(const rbd_model* m, int64_t B, const T* q, const T* qd, const T* qdd, T g, T* c, T* v, T* a,
                T* f, void* stream) {
  RBD_CHECK_ARGS(m && q && qd && c && B >= 0, "rbd_rnea: null model/q/qd/c or negative B");
  if (B == 0) return 0;
  const int variant = variant_of(m);
  const int nd = m->d.n;
  // measured crossovers against the lane kernel: iiwa14 ~16k, Atlas ~4k knot points
  const int64_t small = nd <= 8 ? 16384 : (nd <= 16 ? 8192 : 4096);
  if (m->fast_ok && !(v || a || f) && ((variant == 0 && B < small) || variant == 3)) {
    // Small batches, c only: the warp-cooperative rnea_grad kernel in its rnea-only mode (one body per
    // lane).  The lane kernel below works through 32 knot points per warp, so below one wave of tasks
    // its time is the latency of one task (Atlas: 79 us, flat from 1k to 16k knot points; the
    // cooperative kernel needs 50 us for 1k and 88 us for 4k Atlas knot points).
    const FastModel<T>& fm = pick_dfs<T>(m);
    const int n = fm.n;
    const int G = n <= 8 ? 8 : (n <= 16 ? 16 : 32);
    const int ipw = 32 / G;
    const int tile_stride = coop_grad_tile_stride(n, ipw, false);
    const size_t smem = (size_t)(((n * kCoopMdlStride + 1) & ~1) + kCoopWarps * 32 * kCoopVecStride + kCoopWarps * tile_stride) * sizeof(T) +
                        (size_t)n * kCoopIntStride * sizeof(int);
    if (smem <= kMaxDynSmem) {
      auto kern = G == 8 ? rnea_grad_coop_kernel<T, 8, false, true>
                         : (G == 16 ? rnea_grad_coop_kernel<T, 16, false, true> : rnea_grad_coop_kernel<T, 32, false, true>);
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDynSmem);
      if (e != cudaSuccess) return fail((int)e, cudaGetErrorString(e));
      const int64_t ngroups = (B + ipw - 1) / ipw;
      int64_t blocks = (ngroups + kCoopWarps - 1) / kCoopWarps;
      if (blocks > grid_cap()) blocks = grid_cap();
      kern<<<(unsigned)blocks, kCoopWarps * 32, smem, (cudaStream_t)stream>>>(fm, m->plan, m->coop, B, q, qd, qdd, g, 0, nullptr, c, FbBaseLayout{});
      return cuda_status("rbd_rnea(coop)");
    }
  }
  if (m->fast_ok && variant != 1) {
    // knot point per lane, depth-first chains in registers, coalesced staging (rbd_lane_rnea_kernels.cuh)
    const FastModel<T>& fm = pick_dfs<T>(m);
    const int n = fm.n;
    const bool vaf = v || a || f;
    const bool localf = !vaf && (size_t)lane_rnea_warp_vals(n, fm.n_slot_a, false, false) * sizeof(T) > 32 * 1024;
    const size_t per_warp = (size_t)lane_rnea_warp_vals(n, fm.n_slot_a, localf, vaf) * sizeof(T);
    if (per_warp <= 72 * 1024) {
      void (*kern)(const FastModel<T>, const DfsPlan, int64_t, const T*, const T*, const T*, T, T*, T*, T*, T*);
      if (vaf)
        kern = n <= 8 ? rnea_lane_kernel<T, false, true, 8> : (n <= 16 ? rnea_lane_kernel<T, false, true, 16> : rnea_lane_kernel<T, false, true, 0>);
      else if (localf)
        kern = rnea_lane_kernel<T, true, false, 0>;
      else
        kern = n <= 8 ? rnea_lane_kernel<T, false, false, 8> : (n <= 16 ? rnea_lane_kernel<T, false, false, 16> : rnea_lane_kernel<T, false, false, 0>);
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDynSmem);
      if (e != cudaSuccess) return fail((int)e, cudaGetErrorString(e));
      int warps = 0, best = 0;
      for (int w = 1; w <= 4; ++w) {
        const size_t sz = per_warp * w;
        if (sz > kMaxDynSmem) break;
        int nb = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, w * 32, sz) != cudaSuccess) { cudaGetLastError(); continue; }
        if (nb * w >= best) { best = nb * w; warps = w; }
      }
      if (warps > 0) {
        const int64_t ntasks = (B + 31) / 32;
        kern<<<(unsigned)((ntasks + warps - 1) / warps), warps * 32, per_warp * warps, (cudaStream_t)stream>>>(
            fm, m->plan, B, q, qd, qdd, g, c, v, a, f);
        return cuda_status("rbd_rnea(lane)");
      }
    }
  }
  rnea_fused_kernel<T><<<blocks_for(B, kFusedThreads), kFusedThreads, 0, (cudaStream_t)stream>>>(
      pick<T>(m), B, q, qd, qdd, g, c, v, a, f);
  return cuda_status("rbd_rnea");
}

// Launch geometry of rnea_lane_kernel: warps per CTA that keep the most warps resident.
template <typename K>
static bool lane_rnea_geometry(K kern, size_t per_warp, int* warps_out) {
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDynSmem) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  int warps = 0, best = 0;
  for (int w = 1; w <= 4; ++w) {
    const size_t sz = per_warp * w;
    if (sz > kMaxDynSmem) break;
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, w * 32, sz) != cudaSuccess) { cudaGetLastError(); continue; }
    if (nb * w >= best) { best = nb * w; warps = w; }
  }
  *warps_out = warps;
  return warps > 0;
}

// rnea_fpass (RBDReference.py:559-598): the forward half of the lane kernel when the robot has
// rigid-body inertias and its staging rows fit in shared memory, else the generic pass kernel.
template <typename T>
int launch_rnea_fpass(const rbd_model* m, int64_t B, const T* q, const T* qd, const T* qdd, T g, T* v, T* a,
                      T* f, void* stream) {
  RBD_CHECK_ARGS(m && q && qd && v && a && f && B >= 0, "rbd_rnea_fpass: null argument or negative B");
  if (B == 0) return 0;
  if (m->fast_ok && variant_of(m) != 1) {
    const FastModel<T>& fm = pick_dfs<T>(m);
    const int n = fm.n;
    const size_t per_warp = (size_t)lane_rnea_warp_vals(n, fm.n_slot_a, false, true) * sizeof(T);
    auto kern = n <= 8 ? rnea_lane_kernel<T, false, true, 8, 1> : (n <= 16 ? rnea_lane_kernel<T, false, true, 16, 1> : rnea_lane_kernel<T, false, true, 0, 1>);
    int warps = 0;
    if (per_warp <= 72 * 1024 && lane_rnea_geometry(kern, per_warp, &warps)) {
      const int64_t ntasks = (B + 31) / 32;
      kern<<<(unsigned)((ntasks + warps - 1) / warps), warps * 32, per_warp * warps, (cudaStream_t)stream>>>(
          fm, m->plan, B, q, qd, qdd, g, nullptr, v, a, f);
      return cuda_status("rbd_rnea_fpass(lane)");
    }
  }
  rnea_fpass_kernel<T><<<blocks_for(B, kPassThreads), kPassThreads, 0, (cudaStream_t)stream>>>(
      pick<T>(m), B, q, qd, qdd, g, v, a, f);
  return cuda_status("rbd_rnea_fpass");
}

// rnea_bpass (RBDReference.py:600-621): f is accumulated in place.
template <typename T>
int launch_rnea_bpass(const rbd_model* m, int64_t B, const T* q, T* f, T* c, void* stream) {
  RBD_CHECK_ARGS(m && q && f && c && B >= 0, "rbd_rnea_bpass: null argument or negative B");
  if (B == 0) return 0;
  if (m->fast_ok && variant_of(m) != 1) {
    const FastModel<T>& fm = pick_dfs<T>(m);
    const int n = fm.n;
    const size_t per_warp = (size_t)lane_rnea_warp_vals(n, fm.n_slot_a, false, false) * sizeof(T);
    auto kern = n <= 8 ? rnea_lane_kernel<T, false, false, 8, 2> : (n <= 16 ? rnea_lane_kernel<T, false, false, 16, 2> : rnea_lane_kernel<T, false, false, 0, 2>);
    int warps = 0;
    if (per_warp <= 48 * 1024 && lane_rnea_geometry(kern, per_warp, &warps)) {
      const int64_t ntasks = (B + 31) / 32;
      kern<<<(unsigned)((ntasks + warps - 1) / warps), warps * 32, per_warp * warps, (cudaStream_t)stream>>>(
          fm, m->plan, B, q, nullptr, nullptr, T(0), c, nullptr, nullptr, f);
      return cuda_status("rbd_rnea_bpass(lane)");
    }
  }
  rnea_bpass_kernel<T><<<blocks_for(B, kPassThreads), kPassThreads, 0, (cudaStream_t)stream>>>(
      pick<T>(m), B, q, f, c);
  return cuda_status("rbd_rnea_bpass");
}

template int launch_rnea_fpass<RBD_LAUNCH_T>(const rbd_model*, int64_t, const RBD_LAUNCH_T*, const RBD_LAUNCH_T*,
                                             const RBD_LAUNCH_T*, RBD_LAUNCH_T, RBD_LAUNCH_T*, RBD_LAUNCH_T*, RBD_LAUNCH_T*, void*);
template int launch_rnea_bpass<RBD_LAUNCH_T>(const rbd_model*, int64_t, const RBD_LAUNCH_T*, RBD_LAUNCH_T*, RBD_LAUNCH_T*, void*);
template int launch_rnea<RBD_LAUNCH_T>(const rbd_model*, int64_t, const RBD_LAUNCH_T*, const RBD_LAUNCH_T*, const RBD_LAUNCH_T*,
                                       RBD_LAUNCH_T, RBD_LAUNCH_T*, RBD_LAUNCH_T*, RBD_LAUNCH_T*, RBD_LAUNCH_T*, void*);

}  // namespace rbd_host
