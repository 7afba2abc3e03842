// rbd_fd_kernels.cuh - the small per-knot-point products of forward dynamics
// (RBDReference.py:1369-1384):   qdd = Minv (u - c),   [qdd_dq | qdd_dqd] = -Minv [dc_dq | dc_dqd].
//
// One n x n by n x m product per knot point (m = 1 or 2n): far too small for tensor cores and
// bound by the HBM traffic of its operands, so a CTA stages the contiguous slabs of KB knot points
// in shared memory with coalesced loads, every thread accumulates output elements from shared
// memory, and the results leave through shared memory as coalesced slabs again.
#pragma once
#include "rbd_common.cuh"

namespace rbd {

constexpr int kFdThreads = 128;

// Y[b] = alpha * A[b] * (R1[b] - R2[b])      A: (B,n,n)   R1, R2: (B,n,m)   (R2 may be null)
// SPLIT = false: Y -> out0 (B,n,m).   SPLIT = true: Y[:, :, :m/2] -> out0, Y[:, :, m/2:] -> out1, each (B,n,m/2).
template <typename T, bool SPLIT>
__global__ void __launch_bounds__(kFdThreads)
fd_apply_kernel(int n, int m, int KB, int64_t B, const T* __restrict__ A, const T* __restrict__ R1,
                const T* __restrict__ R2, T alpha, T* __restrict__ out0, T* __restrict__ out1) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* sA = reinterpret_cast<T*>(smem_raw);          // [KB][n*n]
  T* sR = sA + (size_t)KB * n * n;                 // [KB][n*m]
  T* sY = sR + (size_t)KB * n * m;                 // [KB][n*m]
  const int nn = n * n, nm = n * m;
  const int64_t ngroups = (B + KB - 1) / KB;
  for (int64_t grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
    const int64_t b0 = grp * KB;
    const int kb = (int)((B - b0) < KB ? (B - b0) : KB);
    const T* gA = A + b0 * nn;
    const T* gR1 = R1 + b0 * nm;
    for (int e = threadIdx.x; e < kb * nn; e += kFdThreads) sA[e] = gA[e];
    if (R2) {
      const T* gR2 = R2 + b0 * nm;
      for (int e = threadIdx.x; e < kb * nm; e += kFdThreads) sR[e] = gR1[e] - gR2[e];
    } else {
      for (int e = threadIdx.x; e < kb * nm; e += kFdThreads) sR[e] = gR1[e];
    }
    __syncthreads();
    for (int e = threadIdx.x; e < kb * nm; e += kFdThreads) {
      const int k = e / nm, rem = e - k * nm;
      const int r = rem / m, c = rem - r * m;
      const T* a = sA + k * nn + r * n;
      const T* x = sR + k * nm + c;
      T acc0 = T(0), acc1 = T(0);
      int t = 0;
      for (; t + 1 < n; t += 2) {
        acc0 = fma_t(a[t], x[t * m], acc0);
        acc1 = fma_t(a[t + 1], x[(t + 1) * m], acc1);
      }
      if (t < n) acc0 = fma_t(a[t], x[t * m], acc0);
      sY[e] = alpha * (acc0 + acc1);
    }
    __syncthreads();
    if (SPLIT) {
      const int h = m >> 1, nh = n * h;
      T* g0 = out0 + b0 * nh;
      T* g1 = out1 + b0 * nh;
      for (int e = threadIdx.x; e < kb * nh; e += kFdThreads) {
        const int k = e / nh, rem = e - k * nh;
        const int r = rem / h, c = rem - r * h;
        const T* y = sY + k * nm + r * m + c;
        g0[e] = y[0];
        g1[e] = y[h];
      }
    } else {
      T* g0 = out0 + b0 * nm;
      for (int e = threadIdx.x; e < kb * nm; e += kFdThreads) g0[e] = sY[e];
    }
    __syncthreads();
  }
}

}  // namespace rbd
