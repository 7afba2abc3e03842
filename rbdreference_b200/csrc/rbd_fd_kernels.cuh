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

// Asynchronous global -> shared copy of one value.  The staging loops issue ALL of a thread's copies
// before waiting: with a plain load + store per iteration only one load per thread is in flight and
// the kernel is bound by memory latency (measured: 10 % of the HBM bandwidth), not by bandwidth.
__device__ __forceinline__ void fd_cp_async(double* dst, const double* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void fd_cp_async(float* dst, const float* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void fd_cp_async_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Walks the elements start, start + step, ... of a dense [knot][r < n][c < w] slab without per-element
// divisions: (k, r, c) and the flat row index row = k * n + r are advanced incrementally.
struct FdWalk {
  int k, r, c, row, n, w, drow, dcol;
  __device__ __forceinline__ FdWalk(int start, int step, int n_, int w_) : n(n_), w(w_) {
    row = start / w; c = start - row * w;
    k = row / n; r = row - k * n;
    drow = step / w; dcol = step - drow * w;
  }
  __device__ __forceinline__ void next() {
    int dr = drow;
    c += dcol;
    if (c >= w) { c -= w; ++dr; }
    row += dr; r += dr;
    while (r >= n) { r -= n; ++k; }
  }
};

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
    for (int e = threadIdx.x; e < kb * nn; e += kFdThreads) fd_cp_async(sA + e, gA + e);
    if (R2) {
      const T* gR2 = R2 + b0 * nm;
      for (int e = threadIdx.x; e < kb * nm; e += kFdThreads) sR[e] = gR1[e] - gR2[e];
    } else {
      for (int e = threadIdx.x; e < kb * nm; e += kFdThreads) sR[e] = gR1[e];
    }
    fd_cp_async_wait();
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

// Register-tiled version for m > 1 (forward_dynamics_grad: m = 2n): every thread accumulates an
// 8 x 4 block of Y for one knot point, so twelve shared-memory values feed 32 FMAs (the
// one-output-per-thread kernel above needs two shared-memory loads per FMA and is LSU-bound:
// measured 10 ms for Atlas, 2^18 knot points).  Consecutive threads take consecutive column blocks
// of the same rows: their A operands are broadcast reads, their X operands one contiguous,
// 16-byte-aligned run (rows of X and Y are zero-padded to a multiple of four columns).
constexpr int kFdTiledThreads = 128;
constexpr int kFdTR = 8, kFdTC = 4;

__device__ __forceinline__ void fd_load4(const double* p, double (&v)[4]) {
  const double2 a = reinterpret_cast<const double2*>(p)[0], b = reinterpret_cast<const double2*>(p)[1];
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
__device__ __forceinline__ void fd_load4(const float* p, float (&v)[4]) {
  const float4 a = reinterpret_cast<const float4*>(p)[0];
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
}

__host__ __device__ inline size_t fd_tiled_vals_per_knot(int n, int m) { return (size_t)n * n + 2 * (size_t)n * ((m + 3) & ~3); }

template <typename T, bool SPLIT>
__global__ void __launch_bounds__(kFdTiledThreads)
fd_apply_tiled_kernel(int n, int m, int KB, int64_t B, const T* __restrict__ A, const T* __restrict__ R1,
                      const T* __restrict__ R2, T alpha, T* __restrict__ out0, T* __restrict__ out1) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int nn = n * n, nm = n * m;
  const int mp = (m + 3) & ~3, nmp = n * mp;           // padded row pitch of X and Y
  const int nna = (nn + 3) & ~3;                       // keep sR 16-byte aligned after sA
  T* sA = reinterpret_cast<T*>(smem_raw);              // [KB][n*n]
  T* sR = sA + (((size_t)KB * nn + 3) & ~(size_t)3);   // [KB][n][mp]
  T* sY = sR + (size_t)KB * nmp;                       // [KB][n][mp]
  (void)nna;
  const int RT = (n + kFdTR - 1) / kFdTR, CT = mp / kFdTC;
  const int64_t ngroups = (B + KB - 1) / KB;
  for (int64_t grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
    const int64_t b0 = grp * KB;
    const int kb = (int)((B - b0) < KB ? (B - b0) : KB);
    const T* gA = A + b0 * nn;
    const T* gR1 = R1 + b0 * nm;
    const T* gR2 = R2 ? R2 + b0 * nm : nullptr;
    for (int e = threadIdx.x; e < kb * nn; e += kFdTiledThreads) fd_cp_async(sA + e, gA + e);
    {
      FdWalk wk(threadIdx.x, kFdTiledThreads, n, mp);      // padding columns are zero
      for (int e = threadIdx.x; e < kb * nmp; e += kFdTiledThreads, wk.next()) {
        if (wk.c < m) {
          if (gR2) sR[e] = gR1[wk.row * m + wk.c] - gR2[wk.row * m + wk.c];
          else fd_cp_async(sR + e, gR1 + wk.row * m + wk.c);
        } else {
          sR[e] = T(0);
        }
      }
    }
    fd_cp_async_wait();
    __syncthreads();
    const int ntiles = kb * RT * CT;
    for (int tile = threadIdx.x; tile < ntiles; tile += kFdTiledThreads) {
      const int k = tile / (RT * CT), rem = tile - k * (RT * CT);
      const int rt = rem / CT, ct = rem - rt * CT;
      const int r0 = rt * kFdTR, c0 = ct * kFdTC;
      // clamped row pointers: out-of-range rows compute duplicates that are not stored
      const T* a[kFdTR];
#pragma unroll
      for (int ii = 0; ii < kFdTR; ++ii) a[ii] = sA + k * nn + (r0 + ii < n ? r0 + ii : n - 1) * n;
      const T* x = sR + k * nmp + c0;
      T acc[kFdTR][kFdTC];
#pragma unroll
      for (int ii = 0; ii < kFdTR; ++ii)
#pragma unroll
        for (int jj = 0; jj < kFdTC; ++jj) acc[ii][jj] = T(0);
      for (int t = 0; t < n; ++t) {
        T xv[4], av[kFdTR];
        fd_load4(x + t * mp, xv);
#pragma unroll
        for (int ii = 0; ii < kFdTR; ++ii) av[ii] = a[ii][t];
#pragma unroll
        for (int ii = 0; ii < kFdTR; ++ii)
#pragma unroll
          for (int jj = 0; jj < kFdTC; ++jj) acc[ii][jj] = fma_t(av[ii], xv[jj], acc[ii][jj]);
      }
      T* y = sY + k * nmp + c0;
#pragma unroll
      for (int ii = 0; ii < kFdTR; ++ii)
        if (r0 + ii < n) {
#pragma unroll
          for (int jj = 0; jj < kFdTC; ++jj) y[(r0 + ii) * mp + jj] = alpha * acc[ii][jj];
        }
    }
    __syncthreads();
    if (SPLIT) {
      const int h = m >> 1, nh = n * h;
      T* g0 = out0 + b0 * nh;
      T* g1 = out1 + b0 * nh;
      FdWalk wk(threadIdx.x, kFdTiledThreads, n, h);
      for (int e = threadIdx.x; e < kb * nh; e += kFdTiledThreads, wk.next()) {
        const T* y = sY + wk.row * mp + wk.c;
        __stcs(g0 + e, y[0]);
        __stcs(g1 + e, y[h]);
      }
    } else {
      T* g0 = out0 + b0 * nm;
      FdWalk wk(threadIdx.x, kFdTiledThreads, n, m);
      for (int e = threadIdx.x; e < kb * nm; e += kFdTiledThreads, wk.next()) __stcs(g0 + e, sY[wk.row * mp + wk.c]);
    }
    __syncthreads();
  }
}

// FP64 tensor-core version of the same product (forward_dynamics_grad, m = 2n): the only dense
// contraction on the path that is worth tensor cores.  mma.sync.m8n8k4.f64 accumulates an 8 x 8 block
// of Y per warp instruction from ONE shared-memory value of A and one of X per thread (256 FMAs per
// two loads; the register-tiled kernel above needs twelve loads per 32 FMAs and is LSU-bound).
// A, X, Y of KB knot points are staged zero-padded to multiples of the fragment sizes, with row
// pitches = 4 (mod 8) values so that the fragment loads of a half-warp hit 16 different bank pairs.
constexpr int kFdMmaThreads = 128;

__host__ __device__ inline int fd_mma_pitch(int cols) {      // cols is a multiple of 4
  return (cols % 8 == 4) ? cols : cols + 4;
}
struct FdMmaShape {
  int R8, K4, C8, pa, px, vals;                           // padded sizes, pitches, values per knot point
};
__host__ __device__ inline FdMmaShape fd_mma_shape(int n, int m) {
  FdMmaShape s;
  s.R8 = (n + 7) & ~7; s.K4 = (n + 3) & ~3; s.C8 = (m + 7) & ~7;
  s.pa = fd_mma_pitch(s.K4); s.px = fd_mma_pitch(s.C8);
  s.vals = s.R8 * s.pa + s.K4 * s.px + s.R8 * s.px;
  return s;
}

template <bool SPLIT>
__global__ void __launch_bounds__(kFdMmaThreads)
fd_apply_mma_kernel(int n, int m, int KB, int64_t B, const double* __restrict__ A, const double* __restrict__ R1,
                    double alpha, double* __restrict__ out0, double* __restrict__ out1) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const FdMmaShape sh = fd_mma_shape(n, m);
  double* sA = reinterpret_cast<double*>(smem_raw);       // [KB][R8][pa]
  double* sX = sA + (size_t)KB * sh.R8 * sh.pa;           // [KB][K4][px]
  double* sY = sX + (size_t)KB * sh.K4 * sh.px;           // [KB][R8][px]
  const int nn = n * n, nm = n * m;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int RT = sh.R8 >> 3, CTs = sh.C8 >> 3, CP = (CTs + 1) >> 1;
  const int64_t ngroups = (B + KB - 1) / KB;
  // zero the padding once: the staging loops below only write the n x n / n x m interiors
  for (int e = threadIdx.x; e < KB * (sh.R8 * sh.pa + sh.K4 * sh.px); e += kFdMmaThreads) sA[e] = 0.0;
  __syncthreads();
  for (int64_t grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
    const int64_t b0 = grp * KB;
    const int kb = (int)((B - b0) < KB ? (B - b0) : KB);
    const double* gA = A + b0 * nn;
    const double* gX = R1 + b0 * nm;
    {
      FdWalk wa(threadIdx.x, kFdMmaThreads, n, n);
      for (int e = threadIdx.x; e < kb * nn; e += kFdMmaThreads, wa.next()) fd_cp_async(sA + (wa.k * sh.R8 + wa.r) * sh.pa + wa.c, gA + e);
      FdWalk wx(threadIdx.x, kFdMmaThreads, n, m);
      for (int e = threadIdx.x; e < kb * nm; e += kFdMmaThreads, wx.next()) fd_cp_async(sX + (wx.k * sh.K4 + wx.r) * sh.px + wx.c, gX + e);
    }
    fd_cp_async_wait();
    __syncthreads();
    // warp task = 16 x 16 block of Y of one knot point: two A fragments x two X fragments per k-step,
    // four independent accumulator chains
    const int RP = (RT + 1) >> 1;
    const int ntasks = kb * RP * CP;
    for (int task = warp; task < ntasks; task += kFdMmaThreads / 32) {
      const int k = task / (RP * CP), rem = task - k * (RP * CP);
      const int rp = rem / CP, cp = rem - rp * CP;
      const bool r2 = (2 * rp + 1) < RT, c2 = (2 * cp + 1) < CTs;
      const double* a = sA + (k * sh.R8 + rp * 16 + g) * sh.pa + t;
      const double* a8 = r2 ? a + 8 * sh.pa : a;            // second row block (or a harmless duplicate)
      const double* x = sX + (k * sh.K4 + t) * sh.px + cp * 16 + g;
      const int xo = c2 ? 8 : 0;
      double acc[4][2] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
#define RBD_DMMA(C, AF, BF)                                                                  \
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" \
               : "+d"(C[0]), "+d"(C[1]) : "d"(AF), "d"(BF))
      for (int k0 = 0; k0 < sh.K4; k0 += 4) {
        const double af0 = a[k0], af1 = a8[k0];
        const double bf0 = x[k0 * sh.px], bf1 = x[k0 * sh.px + xo];
        RBD_DMMA(acc[0], af0, bf0);
        RBD_DMMA(acc[1], af0, bf1);
        RBD_DMMA(acc[2], af1, bf0);
        RBD_DMMA(acc[3], af1, bf1);
      }
#undef RBD_DMMA
      double* y = sY + (k * sh.R8 + rp * 16 + g) * sh.px + cp * 16 + 2 * t;
      y[0] = alpha * acc[0][0]; y[1] = alpha * acc[0][1];
      if (c2) { y[8] = alpha * acc[1][0]; y[9] = alpha * acc[1][1]; }
      if (r2) {
        double* y8 = y + 8 * sh.px;
        y8[0] = alpha * acc[2][0]; y8[1] = alpha * acc[2][1];
        if (c2) { y8[8] = alpha * acc[3][0]; y8[9] = alpha * acc[3][1]; }
      }
    }
    __syncthreads();
    if (SPLIT) {
      const int h = m >> 1, nh = n * h;
      double* g0 = out0 + b0 * nh;
      double* g1 = out1 + b0 * nh;
      FdWalk wk(threadIdx.x, kFdMmaThreads, n, h);
      for (int e = threadIdx.x; e < kb * nh; e += kFdMmaThreads, wk.next()) {
        const double* y = sY + (wk.k * sh.R8 + wk.r) * sh.px + wk.c;
        __stcs(g0 + e, y[0]);
        __stcs(g1 + e, y[h]);
      }
    } else {
      double* g0 = out0 + b0 * nm;
      FdWalk wk(threadIdx.x, kFdMmaThreads, n, m);
      for (int e = threadIdx.x; e < kb * nm; e += kFdMmaThreads, wk.next()) __stcs(g0 + e, sY[(wk.k * sh.R8 + wk.r) * sh.px + wk.c]);
    }
    __syncthreads();
  }
}

}  // namespace rbd
