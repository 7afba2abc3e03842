// rbd_launch_grad.cu - part of librbd_b200.so (see rbd_internal.cuh); compiled with -DRBD_LAUNCH_T=double|float.
#include "rbd_internal.cuh"
#include "rbd_fused_kernels.cuh"

#ifndef RBD_LAUNCH_T
#error "compile with -DRBD_LAUNCH_T=double or -DRBD_LAUNCH_T=float"
#endif

using namespace rbd;

namespace rbd_host {

template <typename T>
int launch_rnea_grad(const rbd_model* m, int64_t B, const T* q, const T* qd, const T* qdd, T g, int damp,
                     T* dc_du, T* c_out, void* stream) {
  RBD_CHECK_ARGS(m && q && qd && dc_du && B >= 0, "rbd_rnea_grad: null model/q/qd/dc_du or negative B");
  if (B == 0) return 0;
  int variant = variant_of(m);
  if (variant >= 4 && variant != 7) variant = 0;       // 4, 5, 8, 9 only select among the minv kernels
  if (m->fast_ok && m->is_chain && (variant == 0 || variant == 7) && B >= chain_min_batch(m->d.n) &&
      (reinterpret_cast<uintptr_t>(dc_du) & (2 * sizeof(T) - 1)) == 0) {
    // serial chain, one knot point per lane (rbd_chain_grad_kernels.cuh)
    void (*kern)(const ChainModel<T>, int64_t, const T*, const T*, const T*, T, int, T*, T*, int) = nullptr;
    switch (m->d.n) {
      case 4: kern = rnea_grad_chain_kernel<T, 4>; break;
      case 5: kern = rnea_grad_chain_kernel<T, 5>; break;
      case 6: kern = rnea_grad_chain_kernel<T, 6>; break;
      case 7: kern = rnea_grad_chain_kernel<T, 7>; break;
      case 8: kern = rnea_grad_chain_kernel<T, 8>; break;
      default: break;
    }
    if (kern) {
      const size_t smem = chain_grad_smem_pairs(m->d.n) * 32 * 2 * sizeof(T);
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
      if (e != cudaSuccess) return fail((int)e, cudaGetErrorString(e));
      int nb = 8;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, 32, smem) != cudaSuccess || nb < 1) { cudaGetLastError(); nb = 8; }
      kern<<<blocks_for(B, 32), 32, smem, (cudaStream_t)stream>>>(pick_chain<T>(m), B, q, qd, qdd, g, damp, dc_du, c_out,
                                                                  sm_count() * nb);
      return cuda_status("rbd_rnea_grad(chain)");
    }
  }
  if (variant == 7) variant = 0;
  if (m->fast_ok && (variant == 0 || variant == 3)) {
    // warp-cooperative kernel: one body per lane, 32/G knot points per warp
    const FastModel<T>& fm = pick_dfs<T>(m);
    const int n = fm.n;
    const int G = n <= 8 ? 8 : (n <= 16 ? 16 : 32);
    const int ipw = 32 / G;
    // large robots: the output tile is produced in two halves (dc_dq, dc_dqd) to halve its shared memory
    const bool split = false;   // measured on B200 (Atlas): halving the tile does not pay (FP64 1.19e8 vs 1.2e8+, FP32 1.6e8 vs 2.4e8)
    const int tile_stride = coop_grad_tile_stride(n, ipw, split);
    const size_t smem = (size_t)(((n * kCoopMdlStride + 1) & ~1) + kCoopWarps * 32 * kCoopVecStride + kCoopWarps * tile_stride) * sizeof(T) +
                        (size_t)n * kCoopIntStride * sizeof(int);
    if (smem <= kMaxDynSmem) {
      auto kern = G == 8 ? rnea_grad_coop_kernel<T, 8, false>
                         : (G == 16 ? rnea_grad_coop_kernel<T, 16, false>
                                    : (split ? rnea_grad_coop_kernel<T, 32, true> : rnea_grad_coop_kernel<T, 32, false>));
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDynSmem);
      if (e != cudaSuccess) return fail((int)e, cudaGetErrorString(e));
      const int64_t ngroups = (B + ipw - 1) / ipw;
      int64_t blocks = (ngroups + kCoopWarps - 1) / kCoopWarps;
      const int64_t cap = grid_cap();
      if (blocks > cap) blocks = cap;
      kern<<<(unsigned)blocks, kCoopWarps * 32, smem, (cudaStream_t)stream>>>(fm, m->plan, m->coop, B, q, qd, qdd, g, damp,
                                                                              dc_du, c_out, FbBaseLayout{});
      return cuda_status("rbd_rnea_grad(coop)");
    }
  }
  const FastModel<T>& fm = pick_fast<T>(m);
  if (m->fast_ok && (variant == 0 || variant == 2)) {
    // (a fully unrolled compile-time-n instantiation was measured 28 % slower on B200: the
    //  straight-line code no longer fits the instruction cache with ~5 resident warps per SM)
    const size_t stash = (size_t)(fm.n_slot_a * 28 + fm.n_slot_b * 24) * 32 * sizeof(T);
    size_t smem = (size_t)fm.n * kVecPerBody * 32 * sizeof(T) + stash;
    auto kern = rnea_grad_world_kernel<T, 0, 1>;
    if (smem > kGradSmemLimit) {          // large trees: per-body vectors go to local memory
      kern = rnea_grad_world_kernel<T, 1, 8>;   // (tighter register caps measured slower: spills)
      smem = stash;
    }
    if (smem <= kMaxDynSmem) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDynSmem);
      if (e != cudaSuccess) return fail((int)e, cudaGetErrorString(e));
      kern<<<blocks_for(B, 32), 32, smem, (cudaStream_t)stream>>>(fm, B, q, qd, qdd, g, damp, dc_du, c_out);
      return cuda_status("rbd_rnea_grad(world)");
    }
  }
  rnea_grad_fused_kernel<T><<<blocks_for(B, kFusedThreads), kFusedThreads, 0, (cudaStream_t)stream>>>(
      pick<T>(m), B, q, qd, qdd, g, damp, dc_du, c_out);
  return cuda_status("rbd_rnea_grad");
}

template int launch_rnea_grad<RBD_LAUNCH_T>(const rbd_model*, int64_t, const RBD_LAUNCH_T*, const RBD_LAUNCH_T*,
                                            const RBD_LAUNCH_T*, RBD_LAUNCH_T, int, RBD_LAUNCH_T*, RBD_LAUNCH_T*, void*);

}  // namespace rbd_host
