// rbd_launch_minv.cu - part of librbd_b200.so (see rbd_internal.cuh); compiled with -DRBD_LAUNCH_T=double|float.
#include "rbd_internal.cuh"
#include "rbd_fused_kernels.cuh"
#include "rbd_lane_minv_kernels.cuh"

#ifndef RBD_LAUNCH_T
#error "compile with -DRBD_LAUNCH_T=double or -DRBD_LAUNCH_T=float"
#endif

using namespace rbd;

namespace rbd_host {

template <typename T>
int launch_minv(const rbd_model* m, int64_t B, const T* q, int dense, T* Minv, void* stream) {
  RBD_CHECK_ARGS(m && q && Minv && B >= 0, "rbd_minv: null model/q/Minv or negative B");
  if (B == 0) return 0;
  int variant = variant_of(m);
  if (variant == 0 && m->fast_ok && dense) {
    // automatic choice, measured on B200 (evals/s FP64 | FP32 at the BASELINE batch sizes):
    //   Atlas  n=30: hybrid 1.21e8 | 2.0e8   cooperative 1.19e8 | 1.8e8   thread 5.8e7 | 7.1e7
    //   HyQ    n=12: hybrid 6.1e8 | 8.5e8    cooperative 8.8e8 | 1.2e9    thread 5.2e8
    //   iiwa14 n=7 : hybrid 8.5e8 | 1.38e9   cooperative 7.8e8 | 1.2e9    thread 1.13e9 | 1.88e9 (body frame)
    const int n = m->d.n;
    //   iiwa14 n=7 : lane 1.49e9 | 2.60e9 (knot point per lane, table + tile in shared memory)
    // n > 16, measured on B200 (Atlas, 2^18 knot points, FP64 | FP32 ms): tile with 2-column groups and 16 warps
    // (variant 9) 1.88 | 1.23, tile with 4-column groups and 8 warps (variant 8) 1.98 | 1.34, hybrid 2.02 | 1.25; the
    // tile kernels move 2.5 GB of DRAM traffic where the hybrid kernel's scratch hand-off moves 5.1 GB
    variant = n > 16 ? (m->tile2.ok ? 9 : (m->tile.ok ? 8 : 4)) : (n > 8 ? 3 : 5);
    // Small batches (MPC-sized): the knot-point-per-lane kernels process 32 knot points per warp one
    // after the other, so below one wave of tasks their time is the latency of a single task
    // (iiwa14 48 us, Atlas 243 us, flat from 1k to 16k knot points); the cooperative kernel spreads
    // the same batch over 8x - 32x more warps (measured crossovers: iiwa14 ~8k, Atlas ~24k knot points).
    if (B < (n <= 8 ? 8192 : 24576)) variant = 3;
  }
  if (m->fast_ok && dense && variant == 5) {
    // knot point per lane in every phase, per-body table + output tile in shared memory
    const FastModel<T>& fm = pick_dfs<T>(m);
    const int n = fm.n;
    constexpr int GC = 4;
    auto kern = n <= 8 ? (fm.has_prismatic ? minv_lane_kernel<T, GC, true, 8, 0> : minv_lane_kernel<T, GC, false, 8, 0>)
                       : (fm.has_prismatic ? minv_lane_kernel<T, GC, true, 0, 0> : minv_lane_kernel<T, GC, false, 0, 0>);
    int warps = 0, best = 0, ctas = 0;
    size_t smem = 0;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDynSmem);
    if (e != cudaSuccess) return fail((int)e, cudaGetErrorString(e));
    for (int w = 1; w <= kLmMaxWarps; ++w) {
      const size_t sz = lane_minv_smem_bytes<T, GC>(n, fm.n_slot_a, fm.n_slot_b, w);
      if (sz > kMaxDynSmem) break;
      int nb = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, w * 32, sz) != cudaSuccess) { cudaGetLastError(); continue; }
      if (nb * w >= best) { best = nb * w; warps = w; smem = sz; ctas = nb; }
    }
    int dev = 0, sms = 0;
    if (warps > 0 && cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess) {
      const int64_t ntasks = (B + 31) / 32;
      int64_t blocks = (ntasks + warps - 1) / warps;
      if (blocks > (int64_t)sms * ctas) blocks = (int64_t)sms * ctas;
      kern<<<(unsigned)blocks, warps * 32, smem, (cudaStream_t)stream>>>(fm, m->plan, m->coop_minv, B, q, Minv);
      return cuda_status("rbd_minv(lane)");
    }
  }
  if (m->fast_ok && dense && (variant == 8 || variant == 9) && (variant == 8 ? m->tile.ok : m->tile2.ok)) {
    // one CTA per tile of 32 knot points: branch-parallel articulated inertias, column groups per warp,
    // per-body table in shared memory, rows stored straight from registers (rbd_tile_minv_kernels.cuh)
    const FastModel<T>& fm = pick_dfs<T>(m);
    const TilePlan& tp = variant == 8 ? m->tile : m->tile2;
    void (*kern)(const FastModel<T>, const DfsPlan, const TilePlan, int64_t, const T*, T*);
    if (variant == 8) kern = fm.has_prismatic ? minv_tile_kernel<T, true, 4> : minv_tile_kernel<T, false, 4>;
    else kern = fm.has_prismatic ? minv_tile_kernel<T, true, 2> : minv_tile_kernel<T, false, 2>;
    const int warps = tp.nwarps;
    const size_t smem = tile_minv_smem_vals(fm.n, tp.nslot, tp.maxdepth, tp.nslot_g, warps, tp.gc) * sizeof(T);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return fail((int)e, cudaGetErrorString(e));
    int nb = 0;
    cudaError_t eo = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, warps * 32, smem);
    if (eo == cudaSuccess && nb > 0) {
      const int64_t ntiles = (B + 31) / 32;
      int64_t blocks = (int64_t)sm_count() * nb;
      if (blocks > ntiles) blocks = ntiles;
      kern<<<(unsigned)blocks, warps * 32, smem, (cudaStream_t)stream>>>(fm, m->plan, tp, B, q, Minv);
      return cuda_status("rbd_minv(tile)");
    }
    cudaGetLastError();
  }
  if (variant == 8 || variant == 9) variant = 4;
  if (m->fast_ok && dense && variant == 4) {
    // hybrid kernel: knot point per lane for the articulated inertias, column per lane for the
    // rows of Minv, per-body table handed over through an L2-resident scratch buffer
    const FastModel<T>& fm = pick_dfs<T>(m);
    const int n = fm.n;
    const int G = n <= 8 ? 8 : (n <= 16 ? 16 : 32);
    auto kern = fm.has_prismatic
                    ? (G == 8 ? minv_hybrid_kernel<T, 8, true> : (G == 16 ? minv_hybrid_kernel<T, 16, true> : minv_hybrid_kernel<T, 32, true>))
                    : (G == 8 ? minv_hybrid_kernel<T, 8, false> : (G == 16 ? minv_hybrid_kernel<T, 16, false> : minv_hybrid_kernel<T, 32, false>));
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDynSmem);
    if (e != cudaSuccess) return fail((int)e, cudaGetErrorString(e));
    int warps = 0, best = 0, ctas = 0;
    size_t smem = 0;
    for (int w = 4; w <= kCmMaxWarps; ++w) {
      const size_t sz = hybrid_minv_smem_bytes<T>(n, G, m->coop.maxdepth, fm.n_slot_a, fm.n_slot_b, w);
      if (sz > kMaxDynSmem) break;
      int nb = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, w * 32, sz) != cudaSuccess) { cudaGetLastError(); continue; }
      if (nb * w > best) { best = nb * w; warps = w; smem = sz; ctas = nb; }
    }
    int dev = 0, sms = 0;
    if (warps > 0 && cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess) {
      const int64_t ntasks = (B + 31) / 32;
      int64_t blocks = (ntasks + warps - 1) / warps;
      if (blocks > (int64_t)sms * ctas) blocks = (int64_t)sms * ctas;
      T* scratch = nullptr;
      const size_t scratch_bytes = (size_t)blocks * warps * 32 * n * kHyScrStride * sizeof(T);
      cudaMemPool_t pool = scratch_pool(dev);
      if (!pool) return fail(RBD_E_NO_DEVICE, "rbd_minv: cannot create the scratch memory pool");
      e = cudaMallocFromPoolAsync((void**)&scratch, scratch_bytes, pool, (cudaStream_t)stream);
      if (e != cudaSuccess) return fail((int)e, cudaGetErrorString(e));
      kern<<<(unsigned)blocks, warps * 32, smem, (cudaStream_t)stream>>>(fm, m->plan, m->coop_minv, m->coop.maxdepth, B, q, Minv,
                                                                         scratch);
      const int rc = cuda_status("rbd_minv(hybrid)");
      cudaFreeAsync(scratch, (cudaStream_t)stream);
      return rc;
    }
  }
  if (m->fast_ok && dense && variant == 3) {
    // warp-cooperative kernel (local world-aligned frames: accurate in FP32 as well)
    const FastModel<T>& fm = pick_dfs<T>(m);
    const int n = fm.n;
    const int G = n <= 8 ? 8 : (n <= 16 ? 16 : 32);
    auto kern = fm.has_prismatic
                    ? (G == 8 ? minv_coop_kernel<T, 8, true> : (G == 16 ? minv_coop_kernel<T, 16, true> : minv_coop_kernel<T, 32, true>))
                    : (G == 8 ? minv_coop_kernel<T, 8, false> : (G == 16 ? minv_coop_kernel<T, 16, false> : minv_coop_kernel<T, 32, false>));
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDynSmem);
    if (e != cudaSuccess) return fail((int)e, cudaGetErrorString(e));
    // warps per CTA: the choice that keeps the most warps resident per SM (the model copy is
    // shared by the CTA, the rest of the shared memory is per warp)
    int warps = 0, best = 0;
    size_t smem = 0;
    for (int w = 4; w <= kCmMaxWarps; ++w) {
      const size_t sz = coop_minv_smem_bytes<T>(n, G, m->coop.maxdepth, fm.n_slot_a, w);
      if (sz > kMaxDynSmem) break;
      int nb = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, w * 32, sz) != cudaSuccess) { cudaGetLastError(); continue; }
      if (nb * w > best) { best = nb * w; warps = w; smem = sz; }
    }
    if (warps > 0) {
      const int ipw = 32 / G;
      const int64_t ngroups = (B + ipw - 1) / ipw;
      int64_t blocks = (ngroups + warps - 1) / warps;
      const int64_t cap = grid_cap();
      if (blocks > cap) blocks = cap;
      kern<<<(unsigned)blocks, warps * 32, smem, (cudaStream_t)stream>>>(fm, m->plan, m->coop, m->coop_minv, B, q, Minv, nullptr, nullptr, nullptr, 0);
      return cuda_status("rbd_minv(coop)");
    }
  }
  // thread-per-knot-point world-frame kernel.  FP32: coordinates about the world origin lose
  // digits on light distal links far from the base (m |p|^2 cancellation, measured 1.2e-4 on
  // Atlas), so in single precision this family is never selected.
  if (m->fast_ok && dense && std::is_same<T, double>::value && variant == 2) {
    const FastModel<T>& fm = pick_dfs<T>(m);
    const int n = fm.n;
    const size_t stash = (size_t)(fm.n_slot_a * kMinvSlotA + fm.n_slot_b * kMinvSlotB) * 32 * sizeof(T);
    size_t smem = (size_t)(n * (kMinvPerBody + 6) + n * (n + 1) / 2) * 32 * sizeof(T) + stash;
    // measured on B200 (iiwa14, 1M points): shared-memory variant 8.8e8 evals/s (4 warps/SM),
    // local-memory variant 1.12e9 evals/s (12 warps/SM) -> local memory is the default here
    auto kern = minv_world_kernel<T, 1, 12>;
    static const bool force_smem = std::getenv("RBD_MINV_SMEM") != nullptr;
    if (force_smem && smem <= kMaxDynSmem) kern = minv_world_kernel<T, 0, 1>;
    else smem = stash;
    if (smem <= kMaxDynSmem) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDynSmem);
      if (e != cudaSuccess) return fail((int)e, cudaGetErrorString(e));
      kern<<<blocks_for(B, 32), 32, smem, (cudaStream_t)stream>>>(fm, m->plan, B, q, Minv);
      return cuda_status("rbd_minv(world)");
    }
  }
  minv_fused_kernel<T><<<blocks_for(B, kFusedThreads), kFusedThreads, 0, (cudaStream_t)stream>>>(
      pick<T>(m), B, q, dense, Minv);
  return cuda_status("rbd_minv");
}

// crba (RBDReference.py:1026-1124, fixed base): warp-cooperative kernel (the minv kernel with the
// articulated downdate switched off) for rigid-body inertias, generic dense kernel otherwise.
template <typename T>
int launch_crba(const rbd_model* m, int64_t B, const T* q, T* H, void* stream) {
  RBD_CHECK_ARGS(m && q && H && B >= 0, "rbd_crba: null model/q/H or negative B");
  if (B == 0) return 0;
  const int variant = variant_of(m);
  if (m->fast_ok && variant != 1) {
    const FastModel<T>& fm = pick_dfs<T>(m);
    const int n = fm.n;
    const int G = n <= 8 ? 8 : (n <= 16 ? 16 : 32);
    auto kern = fm.has_prismatic
                    ? (G == 8 ? minv_coop_kernel<T, 8, true, true> : (G == 16 ? minv_coop_kernel<T, 16, true, true> : minv_coop_kernel<T, 32, true, true>))
                    : (G == 8 ? minv_coop_kernel<T, 8, false, true> : (G == 16 ? minv_coop_kernel<T, 16, false, true> : minv_coop_kernel<T, 32, false, true>));
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDynSmem);
    if (e != cudaSuccess) return fail((int)e, cudaGetErrorString(e));
    int warps = 0, best = 0;
    size_t smem = 0;
    for (int w = 4; w <= kCmMaxWarps; ++w) {
      const size_t sz = coop_minv_smem_bytes<T>(n, G, m->coop.maxdepth, fm.n_slot_a, w);
      if (sz > kMaxDynSmem) break;
      int nb = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, w * 32, sz) != cudaSuccess) { cudaGetLastError(); continue; }
      if (nb * w > best) { best = nb * w; warps = w; smem = sz; }
    }
    if (warps > 0) {
      const int ipw = 32 / G;
      const int64_t ngroups = (B + ipw - 1) / ipw;
      int64_t blocks = (ngroups + warps - 1) / warps;
      const int64_t cap = grid_cap();
      if (blocks > cap) blocks = cap;
      kern<<<(unsigned)blocks, warps * 32, smem, (cudaStream_t)stream>>>(fm, m->plan, m->coop, m->coop_minv, B, q, H, nullptr, nullptr, nullptr, 0);
      return cuda_status("rbd_crba(coop)");
    }
  }
  crba_fused_kernel<T><<<blocks_for(B, kFusedThreads), kFusedThreads, 0, (cudaStream_t)stream>>>(pick<T>(m), B, q, H);
  return cuda_status("rbd_crba");
}

// aba (RBDReference.py:817, fixed-base branch): one launch, knot point per thread.
template <typename T>
int launch_aba(const rbd_model* m, int64_t B, const T* q, const T* qd, const T* tau, T g, T* qdd, void* stream) {
  RBD_CHECK_ARGS(m && q && qd && tau && qdd && B >= 0, "rbd_aba: null argument or negative B");
  if (B == 0) return 0;
  aba_fused_kernel<T><<<blocks_for(B, kFusedThreads), kFusedThreads, 0, (cudaStream_t)stream>>>(pick<T>(m), B, q, qd, tau, g, qdd);
  return cuda_status("rbd_aba");
}
template int launch_aba<RBD_LAUNCH_T>(const rbd_model*, int64_t, const RBD_LAUNCH_T*, const RBD_LAUNCH_T*, const RBD_LAUNCH_T*,
                                      RBD_LAUNCH_T, RBD_LAUNCH_T*, void*);

template int launch_crba<RBD_LAUNCH_T>(const rbd_model*, int64_t, const RBD_LAUNCH_T*, RBD_LAUNCH_T*, void*);
template int launch_minv<RBD_LAUNCH_T>(const rbd_model*, int64_t, const RBD_LAUNCH_T*, int, RBD_LAUNCH_T*, void*);

}  // namespace rbd_host
