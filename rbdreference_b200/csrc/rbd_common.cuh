// rbd_common.cuh - device model layout and spatial-algebra primitives (sm_100a).
//
// The robot is compiled once on the host (rbdreference_b200/model.py) and travels to every
// kernel as a __grid_constant__ parameter, i.e. it lives in the constant bank: warp-uniform
// reads cost no shared memory or L1 bandwidth and there is no per-device upload to race on.
//
// Spatial vectors are [angular(3); linear(3)] (RBDReference.py:566).  A joint transform is the
// Pluecker matrix X = [[E,0],[L,E]]; only E and L (9 + 9 values) are ever formed.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define RBD_MAX_DOF 32

namespace rbd {

template <typename T>
struct DevModel {
  int n;
  int parent[RBD_MAX_DOF];
  int kind[RBD_MAX_DOF];            // 0 revolute (cos/sin), 1 prismatic (affine)
  unsigned anc_mask[RBD_MAX_DOF];   // bit c set iff c is an ancestor of i or c == i
  unsigned sub_mask[RBD_MAX_DOF];   // bit j set iff j is in subtree(i) (including i)
  T damping[RBD_MAX_DOF];
  T S[RBD_MAX_DOF][6];
  T XA[RBD_MAX_DOF][18];
  T XB[RBD_MAX_DOF][18];
  T XC[RBD_MAX_DOF][18];
  T I[RBD_MAX_DOF][36];
};

// ---- scalar helpers ---------------------------------------------------------------------
__device__ __forceinline__ void sincos_t(double x, double* s, double* c) { sincos(x, s, c); }
__device__ __forceinline__ void sincos_t(float x, float* s, float* c) { sincosf(x, s, c); }
__device__ __forceinline__ double fma_t(double a, double b, double c) { return fma(a, b, c); }
__device__ __forceinline__ float fma_t(float a, float b, float c) { return fmaf(a, b, c); }

// sin / cos of NV angles at once: Cody-Waite reduction by pi/2 (exact products through FMA) and the
// fdlibm minimax polynomials on [-pi/4, pi/4]; every step loops over the NV values so that NV
// independent dependency chains are in flight.  |x| > 1e5 (never the case for joint angles) takes
// the library path.  Max error ~1 ulp.
static __device__ __noinline__ void sincos_far(const double* x, double* s, double* c, int n) {
  for (int k = 0; k < n; ++k) sincos(x[k], &s[k], &c[k]);
}
template <int NV>
__device__ __forceinline__ void sincos_batch(const double* x, double* s, double* c) {
  const double MAGIC = 6755399441055744.0;           // 1.5 * 2^52: round-to-nearest integer in the low word
  double kd[NV], r[NV], z[NV], ps[NV], pc[NV];
  int ki[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const double t = fma(x[k], 6.36619772367581382433e-01, MAGIC);
    ki[k] = __double2loint(t);
    kd[k] = t - MAGIC;
  }
#pragma unroll
  for (int k = 0; k < NV; ++k) r[k] = fma(-kd[k], 1.57079632679489655800e+00, x[k]);
#pragma unroll
  for (int k = 0; k < NV; ++k) r[k] = fma(-kd[k], 6.12323399573676603587e-17, r[k]);
#pragma unroll
  for (int k = 0; k < NV; ++k) r[k] = fma(-kd[k], -1.49738490485916983e-33, r[k]);
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    z[k] = r[k] * r[k];
    ps[k] = fma(z[k], 1.58969099521155010221e-10, -2.50507602534068634195e-08);
    pc[k] = fma(z[k], -1.13596475577881948265e-11, 2.08757232129817482790e-09);
  }
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    ps[k] = fma(z[k], ps[k], 2.75573137070700676789e-06);
    pc[k] = fma(z[k], pc[k], -2.75573143513906633035e-07);
  }
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    ps[k] = fma(z[k], ps[k], -1.98412698298579493134e-04);
    pc[k] = fma(z[k], pc[k], 2.48015872894767294178e-05);
  }
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    ps[k] = fma(z[k], ps[k], 8.33333333332248946124e-03);
    pc[k] = fma(z[k], pc[k], -1.38888888888741095749e-03);
  }
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    ps[k] = fma(z[k], ps[k], -1.66666666666666324348e-01);
    pc[k] = fma(z[k], pc[k], 4.16666666666666019037e-02);
  }
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const double sv = fma(z[k] * r[k], ps[k], r[k]);
    const double cv = fma(z[k] * z[k], pc[k], fma(-0.5, z[k], 1.0));
    const bool swap = ki[k] & 1;
    double so = swap ? cv : sv, co = swap ? sv : cv;
    if (ki[k] & 2) so = -so;
    if ((ki[k] + 1) & 2) co = -co;
    s[k] = so;
    c[k] = co;
  }
  bool far = false;
#pragma unroll
  for (int k = 0; k < NV; ++k) far = far || !(fabs(x[k]) <= 1.0e5);   // also NaN / Inf
  if (far) {                                                          // cold: arrays live on the stack only here
    double xs[NV], ss[NV], cs[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) xs[k] = x[k];
    sincos_far(xs, ss, cs, NV);
#pragma unroll
    for (int k = 0; k < NV; ++k) { s[k] = ss[k]; c[k] = cs[k]; }
  }
}
template <int NV>
__device__ __forceinline__ void sincos_batch(const float* x, float* s, float* c) {
#pragma unroll
  for (int k = 0; k < NV; ++k) sincosf(x[k], &s[k], &c[k]);
}

// ---- joint transform X(q) = A + B*f1 + C*f2 ------------------------------------------------
// X[0..8] = E row-major, X[9..17] = L row-major.
template <typename T>
__device__ __forceinline__ void joint_basis(const DevModel<T>& m, int i, T q, T& f1, T& f2) {
  if (m.kind[i] == 0) {
    sincos_t(q, &f2, &f1);
  } else {
    f1 = q;
    f2 = T(0);
  }
}

template <typename T>
__device__ __forceinline__ void build_X(const DevModel<T>& m, int i, T f1, T f2, T (&X)[18]) {
#pragma unroll
  for (int k = 0; k < 18; ++k) X[k] = fma_t(m.XC[i][k], f2, fma_t(m.XB[i][k], f1, m.XA[i][k]));
}

template <typename T>
__device__ __forceinline__ void build_X_from_q(const DevModel<T>& m, int i, T q, T (&X)[18]) {
  T f1, f2;
  joint_basis(m, i, q, f1, f2);
  build_X(m, i, f1, f2, X);
}

// y = X x   (motion vector parent -> child)
template <typename T>
__device__ __forceinline__ void X_apply(const T (&X)[18], const T (&x)[6], T (&y)[6]) {
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    y[r] = X[3 * r] * x[0] + X[3 * r + 1] * x[1] + X[3 * r + 2] * x[2];
    y[3 + r] = X[9 + 3 * r] * x[0] + X[9 + 3 * r + 1] * x[1] + X[9 + 3 * r + 2] * x[2] +
               X[3 * r] * x[3] + X[3 * r + 1] * x[4] + X[3 * r + 2] * x[5];
  }
}

// y = X^T f   (force vector child -> parent)
template <typename T>
__device__ __forceinline__ void XT_apply(const T (&X)[18], const T (&f)[6], T (&y)[6]) {
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    y[c] = X[c] * f[0] + X[3 + c] * f[1] + X[6 + c] * f[2] +
           X[9 + c] * f[3] + X[12 + c] * f[4] + X[15 + c] * f[5];
    y[3 + c] = X[c] * f[3] + X[3 + c] * f[4] + X[6 + c] * f[5];
  }
}

// y += X^T f
template <typename T>
__device__ __forceinline__ void XT_apply_add(const T (&X)[18], const T (&f)[6], T (&y)[6]) {
  T t[6];
  XT_apply(X, f, t);
#pragma unroll
  for (int k = 0; k < 6; ++k) y[k] += t[k];
}

// y = M x for a dense row-major 6x6 held in the model (constant bank)
template <typename T>
__device__ __forceinline__ void mat6_apply(const T* __restrict__ M, const T (&x)[6], T (&y)[6]) {
#pragma unroll
  for (int r = 0; r < 6; ++r) {
    T acc = M[6 * r] * x[0];
#pragma unroll
    for (int k = 1; k < 6; ++k) acc = fma_t(M[6 * r + k], x[k], acc);
    y[r] = acc;
  }
}

// y = crm(v) s  : motion cross product  v x s            (RBDReference.py:9-21, :56-59)
template <typename T>
__device__ __forceinline__ void crm_mul(const T* v, const T* s, T* y) {
  y[0] = v[1] * s[2] - v[2] * s[1];
  y[1] = v[2] * s[0] - v[0] * s[2];
  y[2] = v[0] * s[1] - v[1] * s[0];
  y[3] = v[4] * s[2] - v[5] * s[1] + v[1] * s[5] - v[2] * s[4];
  y[4] = v[5] * s[0] - v[3] * s[2] + v[2] * s[3] - v[0] * s[5];
  y[5] = v[3] * s[1] - v[4] * s[0] + v[0] * s[4] - v[1] * s[3];
}

// y = crf(v) f  : force cross product  v x* f = -crm(v)^T f   (RBDReference.py:149-164)
template <typename T>
__device__ __forceinline__ void crf_mul(const T* v, const T* f, T* y) {
  y[0] = v[1] * f[2] - v[2] * f[1] + v[4] * f[5] - v[5] * f[4];
  y[1] = v[2] * f[0] - v[0] * f[2] + v[5] * f[3] - v[3] * f[5];
  y[2] = v[0] * f[1] - v[1] * f[0] + v[3] * f[4] - v[4] * f[3];
  y[3] = v[1] * f[5] - v[2] * f[4];
  y[4] = v[2] * f[3] - v[0] * f[5];
  y[5] = v[0] * f[4] - v[1] * f[3];
}

template <typename T>
__device__ __forceinline__ T dot6(const T* s, const T* x) {
  T acc = s[0] * x[0];
#pragma unroll
  for (int k = 1; k < 6; ++k) acc = fma_t(s[k], x[k], acc);
  return acc;
}

// request the cache line holding p into L2 (no register, no wait)
__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// ---- 1-D bulk copies (TMA engine, cp.async.bulk) between a warp's shared-memory tile and its contiguous slab in HBM ----
// A tile that has the exact layout of the slab moves with ONE instruction issued by one lane instead of count / 32
// load + store pairs per lane: the LSU issue slots stay free for the arithmetic and the copy engine works in the
// background.  Requirements of the instruction: both addresses 16-byte aligned, size a multiple of 16 bytes - checked
// at run time (warp-uniform); `false` tells the caller to fall back to its element loop.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// shared -> global.  Every lane calls (the fence orders each lane's own generic-proxy writes to the tile before the
// async proxy reads them); completion is awaited with warp_bulk_store_wait before the tile is written again.
template <typename T>
__device__ __forceinline__ bool warp_bulk_store(T* __restrict__ gdst, const T* ssrc, int count, int lane) {
  const unsigned bytes = (unsigned)count * (unsigned)sizeof(T);
  if (((bytes | (unsigned)(uintptr_t)gdst | smem_u32(ssrc)) & 15u) != 0u || bytes == 0u) return false;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  if (lane == 0) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
  return true;
}
// the bulk stores this warp issued have finished READING shared memory (the tile may be rewritten)
__device__ __forceinline__ void warp_bulk_store_wait(int lane) {
  if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  __syncwarp();
}

// global -> shared with mbarrier completion.  `bar` is an 8-byte shared-memory word owned by the warp, initialised
// once with warp_bulk_bar_init; `phase` is the warp's phase bit (flipped by the caller after every wait).
__device__ __forceinline__ void warp_bulk_bar_init(unsigned long long* bar, int lane) {
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
}
template <typename T>
__device__ __forceinline__ bool warp_bulk_load(T* sdst, const T* __restrict__ gsrc, int count, unsigned long long* bar, int lane) {
  const unsigned bytes = (unsigned)count * (unsigned)sizeof(T);
  if (((bytes | (unsigned)(uintptr_t)gsrc | smem_u32(sdst)) & 15u) != 0u || bytes == 0u) return false;
  __syncwarp();                                          // nobody still reads the tile
  if (lane == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(sdst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
  }
  return true;
}
__device__ __forceinline__ void warp_bulk_load_wait(unsigned long long* bar, unsigned phase) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(phase) : "memory");
}

// Cooperative store of per-lane result blocks: lane k of the warp holds the NV values of knot point b0 + k
// (in memory order) in `vals` (registers / local memory); dst + (b0 + k) * NV is where they go.  32 values
// per lane are transposed through `tile` ([32][33]) at a time, so every store instruction writes 256
// contiguous bytes of one knot point instead of 8 bytes per lane NV * sizeof(T) apart.  All 32 lanes call.
constexpr int kFlushChunk = 32;
template <typename T>
__device__ __forceinline__ void warp_flush_blocks(const T* vals, int NV, T* tile, T* __restrict__ dst, int nlive, int lane) {
  for (int c0 = 0; c0 < NV; c0 += kFlushChunk) {
    const int cnt = NV - c0 < kFlushChunk ? NV - c0 : kFlushChunk;
    for (int k = 0; k < cnt; ++k) tile[lane * (kFlushChunk + 1) + k] = vals[c0 + k];
    __syncwarp();
    if (lane < cnt)
      for (int r = 0; r < nlive; ++r) __stcs(dst + (int64_t)r * NV + c0 + lane, tile[r * (kFlushChunk + 1) + lane]);
    __syncwarp();
  }
}

}  // namespace rbd
