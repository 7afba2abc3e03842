// rbd_internal.cuh - host-side declarations shared by the translation units of librbd_b200.so.
// The library is built from several .cu files compiled in parallel (rbdreference_b200/build.py):
//   rbd_capi.cu         C ABI, model compilation, per-pass and forward-dynamics launchers
//   rbd_launch_rnea.cu  launch_rnea<T>       (rnea kernels)
//   rbd_launch_grad.cu  launch_rnea_grad<T>  (rnea_grad kernels)
//   rbd_launch_minv.cu  launch_minv<T>, launch_crba<T>
//   rbd_launch_pass.cu  launch_grad_fpass / launch_grad_bpass (the four gradient passes)
// The launcher files are compiled once per precision (-DRBD_LAUNCH_T=double / float).
#pragma once
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <type_traits>
#include <atomic>
#include <mutex>
#include <new>

#include <nvtx3/nvToolsExt.h>

#include "../../include/rbd_b200.h"
#include "rbd_common.cuh"
#include "rbd_grad_kernels.cuh"
#include "rbd_minv_kernels.cuh"
#include "rbd_coop_kernels.cuh"
#include "rbd_coop_minv_kernels.cuh"
#include "rbd_tile_minv_kernels.cuh"
#include "rbd_chain_grad_kernels.cuh"

struct rbd_model {
  rbd::DevModel<double> d;
  rbd::DevModel<float> f;
  rbd::FastModel<double> fd;     // world-frame kernels (rigid-body inertias only)
  rbd::FastModel<float> ff;
  bool fast_ok;                  // FastModel valid (rigid inertias, 1-DoF revolute/prismatic joints)
  bool is_chain;                 // parent[i] == i - 1 for every body (serial chain)
  rbd::ChainModel<double> chain_d;   // per-body records of the chain rnea_grad kernel (is_chain && n <= kChainMaxN)
  rbd::ChainModel<float> chain_f;
  rbd::FastModel<double> fd_dfs; // the same robot renumbered in depth-first preorder
  rbd::FastModel<float> ff_dfs;
  rbd::DfsPlan plan;
  rbd::CoopPlan coop;
  rbd::CoopMinvPlan coop_minv;
  rbd::TilePlan tile;            // chain / column-group schedule of the tile minv kernel (4-column groups, 8 warps)
  rbd::TilePlan tile2;           // the same with 2-column groups and 16 warps per CTA
  mutable std::atomic<int> variant{-1};   // per-handle kernel family (-1: follow rbd_set_kernel_variant)
};

namespace rbd_host {

// NVTX range around one C-ABI call (header-only nvtx3: a no-op unless a profiler is attached)
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};
#define RBD_NVTX(name) ::rbd_host::NvtxRange rbd_nvtx_range_(name)

constexpr size_t kMaxDynSmem = 227 * 1024;
extern std::atomic<int> g_variant;      // 0 auto, 1 force the generic body-frame kernels, 2.. see rbd_b200.h
// kernel family of one call: the handle's own choice (rbd_model_set_kernel_variant) or the process default
inline int variant_of(const rbd_model* m) {
  const int v = m ? m->variant.load(std::memory_order_relaxed) : -1;
  return v >= 0 ? v : g_variant.load(std::memory_order_relaxed);
}
int sm_count();                         // SMs of the current device (cudaDevAttrMultiProcessorCount, cached per device)
inline int64_t grid_cap(int ctas_per_sm = 16) { return (int64_t)sm_count() * ctas_per_sm; }
size_t smem_limit();                    // shared-memory budget per warp of the world-frame grad kernel
cudaMemPool_t scratch_pool(int dev);    // private stream-ordered pool for scratch buffers
// model compilation shared by the fixed-base and the floating-base handles (rbd_capi.cu)
bool build_fast_model(const RbdModelDesc* d, rbd::FastModel<double>& out);                        // false: not rigid / not 1-DoF
bool build_dfs_model(const RbdModelDesc* d, rbd::FastModel<double>& out, rbd::DfsPlan& plan);      // depth-first renumbering
void build_coop_plans(const rbd::FastModel<double>& dfs, rbd::CoopPlan& cp, rbd::CoopMinvPlan& mp);
void narrow_fast_model(const rbd::FastModel<double>& a, rbd::FastModel<float>& b);
int fail(int code, const char* msg);    // records the message for rbd_last_error_string()
int cuda_status(const char* what);      // cudaGetLastError -> status code; counts the launch

template <typename T> inline const rbd::DevModel<T>& pick(const rbd_model* m);
template <> inline const rbd::DevModel<double>& pick<double>(const rbd_model* m) { return m->d; }
template <> inline const rbd::DevModel<float>& pick<float>(const rbd_model* m) { return m->f; }
template <typename T> inline const rbd::FastModel<T>& pick_dfs(const rbd_model* m);
template <> inline const rbd::FastModel<double>& pick_dfs<double>(const rbd_model* m) { return m->fd_dfs; }
template <> inline const rbd::FastModel<float>& pick_dfs<float>(const rbd_model* m) { return m->ff_dfs; }
template <typename T> inline const rbd::ChainModel<T>& pick_chain(const rbd_model* m);
template <> inline const rbd::ChainModel<double>& pick_chain<double>(const rbd_model* m) { return m->chain_d; }
template <> inline const rbd::ChainModel<float>& pick_chain<float>(const rbd_model* m) { return m->chain_f; }
template <typename T> inline const rbd::FastModel<T>& pick_fast(const rbd_model* m);
template <> inline const rbd::FastModel<double>& pick_fast<double>(const rbd_model* m) { return m->fd; }
template <> inline const rbd::FastModel<float>& pick_fast<float>(const rbd_model* m) { return m->ff; }

// below this batch the one-knot-point-per-lane chain kernel leaves SMs idle (32 knot points per warp)
inline int64_t chain_min_batch(int n) { (void)n; return 16384; }
inline unsigned blocks_for(int64_t B, int threads) { return (unsigned)((B + threads - 1) / threads); }

#define RBD_CHECK_ARGS(cond, msg) \
  do { if (!(cond)) return fail(RBD_E_INVALID_ARGUMENT, msg); } while (0)
#define kGradSmemLimit smem_limit()

// stream-ordered temporary from the private scratch pool, freed (stream-ordered) on scope exit
struct PoolBuf {
  void* p = nullptr;
  cudaStream_t s;
  explicit PoolBuf(cudaStream_t st) : s(st) {}
  int alloc(size_t bytes) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return fail(RBD_E_NO_DEVICE, "no CUDA device");
    cudaMemPool_t pool = scratch_pool(dev);
    if (!pool) return fail(RBD_E_NO_DEVICE, "cannot create the scratch memory pool");
    cudaError_t e = cudaMallocFromPoolAsync(&p, bytes, pool, s);
    if (e != cudaSuccess) { p = nullptr; return fail((int)e, cudaGetErrorString(e)); }
    return 0;
  }
  ~PoolBuf() { if (p) cudaFreeAsync(p, s); }
};

// Y = alpha A (R1 - R2) per knot point, the product of forward_dynamics(_grad) (rbd_fd_kernels.cuh); defined and
// explicitly instantiated in rbd_capi.cu
template <typename T, bool SPLIT>
int launch_fd_apply(int variant, int n, int mcols, int64_t B, const T* A, const T* R1, const T* R2, T alpha, T* out0, T* out1,
                    void* stream);

// fused drivers (defined in rbd_launch_*.cu, explicitly instantiated for double and float)
template <typename T>
int launch_rnea(const rbd_model* m, int64_t B, const T* q, const T* qd, const T* qdd, T g, T* c, T* v, T* a,
                T* f, void* stream);
template <typename T>
int launch_rnea_grad(const rbd_model* m, int64_t B, const T* q, const T* qd, const T* qdd, T g, int damp,
                     T* dc_du, T* c_out, void* stream);
template <typename T>
int launch_minv(const rbd_model* m, int64_t B, const T* q, int dense, T* Minv, void* stream);
template <typename T>
int launch_rnea_fpass(const rbd_model* m, int64_t B, const T* q, const T* qd, const T* qdd, T g, T* v, T* a,
                      T* f, void* stream);
template <typename T>
int launch_rnea_bpass(const rbd_model* m, int64_t B, const T* q, T* f, T* c, void* stream);
template <typename T, bool DQ>
int launch_grad_fpass(const rbd_model* m, int64_t B, const T* q, const T* qd, const T* v, const T* a, T g,
                      T* dv, T* da, T* df, void* stream);
template <typename T, bool DQ>
int launch_grad_bpass(const rbd_model* m, int64_t B, const T* q, const T* f, T* df, int damp, T* dc,
                      void* stream);
template <typename T>
int launch_crba(const rbd_model* m, int64_t B, const T* q, T* H, void* stream);
template <typename T>
int launch_aba(const rbd_model* m, int64_t B, const T* q, const T* qd, const T* tau, T g, T* qdd, void* stream);

}  // namespace rbd_host
