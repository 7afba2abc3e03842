// rbd_lane_rnea_kernels.cuh - fused rnea (RBDReference.py:623-628 = rnea_fpass :559-598 +
// rnea_bpass :600-621) for robots with rigid-body inertias and 1-DoF revolute / prismatic joints.
//
// Body-frame recursion exactly as the reference (v_i = X v_p + S qd, a_i = X a_p + crm(v_i) S qd +
// S qdd, f_i = I a_i + crf(v_i) I v_i, c_i = S.f_i, f_p += X^T f_i) with
//   * one knot point per lane, bodies in depth-first preorder: the parent's (v, a) and the child's
//     X^T f stay in registers along chains, only branch points go through a shared-memory stash;
//   * X = [[E, 0], [-E rx, E]] applied as two 3x3 products and one cross product (no 6x6), the
//     spatial inertia as (m, h, Ibar) - 10 numbers instead of 36;
//   * the warp's contiguous slabs of q / qd / qdd are staged with coalesced loads, and c (and
//     v, a, f when requested) leave through the same staging buffers with coalesced stores;
//   * per-body state kept for the backward sweep: f (6) and (cos q, sin q) only.
#pragma once
#include "rbd_common.cuh"
#include "rbd_grad_kernels.cuh"
#include "rbd_minv_kernels.cuh"

namespace rbd {

constexpr int kLrMaxWarps = 8;        // warps per CTA are chosen at launch (blockDim.x / 32)
constexpr int kLrStride = 33;        // lane stride of the staging rows: conflict-free both ways

// values of T per warp: io [n][3] rows | f [n][6] rows (LOCALF: in local memory) | v, a rows (VAF) | stash
__host__ __device__ inline int lane_rnea_warp_vals(int n, int nslot_a, bool localf, bool vaf) {
  return (3 * n + (localf ? 0 : 6 * n) + (vaf ? 12 * n : 0) + 12 * nslot_a) * kLrStride;
}

// NMAX >= n > 0: both sweeps are fully unrolled over the body index (guarded by i < n), which turns
// every model constant into an immediate constant-bank operand of the FP instruction that uses it
// (measured on B200: iiwa14 +8 % FP64 / +31 % FP32).  NMAX = 0: rolled loops with indexed constant
// loads, for large robots (unrolling 32 bodies costs more in instruction fetch than it saves).
// MODE 0: rnea.  MODE 1: rnea_fpass only (:559-598; VAF = true; f_out = the bodies' own forces, c is
// not touched).  MODE 2: rnea_bpass only (:600-621; f_out holds f on entry - staged with coalesced
// loads - and the accumulated forces on exit, c is written; qd / qdd / v_out / a_out are unused).
template <typename T, bool LOCALF, bool VAF, int NMAX, int MODE = 0>
__global__ void __launch_bounds__(kLrMaxWarps * 32)
rnea_lane_kernel(const __grid_constant__ FastModel<T> m, const __grid_constant__ DfsPlan plan, int64_t B,
                 const T* __restrict__ q, const T* __restrict__ qd, const T* __restrict__ qdd, T gravity,
                 T* __restrict__ c, T* __restrict__ v_out, T* __restrict__ a_out, T* __restrict__ f_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = m.n;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t task = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
  const int64_t first = task * 32;
  if (first >= B) return;
  const int nk = (int)((B - first) < 32 ? (B - first) : 32);
  const int warp_vals = lane_rnea_warp_vals(n, m.n_slot_a, LOCALF, VAF);
  T* ws = reinterpret_cast<T*>(smem_raw) + (size_t)warp * warp_vals;
  T* io = ws;                                             // [n][3] rows: q|f1, qd|f2, qdd|c
  T* fb = io + 3 * n * kLrStride;                         // [n][6] rows
  T* vb = fb + (LOCALF ? 0 : 6 * n * kLrStride);          // [n][6] rows (VAF)
  T* ab = vb + (VAF ? 6 * n * kLrStride : 0);
  T* st = ab + (VAF ? 6 * n * kLrStride : 0);             // [slot_a][12] rows
  T lf[LOCALF ? RBD_MAX_DOF * 6 : 1];
#define IO(i, k) io[((i) * 3 + (k)) * kLrStride + lane]
#define FB(i, k) (*(LOCALF ? &lf[(i) * 6 + (k)] : &fb[((i) * 6 + (k)) * kLrStride + lane]))
#define VB(i, k) vb[((i) * 6 + (k)) * kLrStride + lane]
#define AB(i, k) ab[((i) * 6 + (k)) * kLrStride + lane]
#define ST(s, k) st[((s) * 12 + (k)) * kLrStride + lane]

  // ---------------------------------------------------------------- stage q, qd, qdd (coalesced)
  {
    const int count = nk * n;
    const int64_t base = first * n;
    int kn = lane / n, jn = lane - kn * n;                // element e = kn * n + jn of the slab
    const int dk = 32 / n, dj = 32 - dk * n;
    for (int e = lane; e < 32 * n; e += 32) {
      const bool ok = e < count;
      const int row = plan.pos[jn] * 3;
      io[(row + 0) * kLrStride + kn] = ok ? __ldg(q + base + e) : T(0);
      if (MODE != 2) {
        io[(row + 1) * kLrStride + kn] = ok ? __ldg(qd + base + e) : T(0);
        io[(row + 2) * kLrStride + kn] = (ok && qdd) ? __ldg(qdd + base + e) : T(0);
      }
      kn += dk; jn += dj;
      if (jn >= n) { jn -= n; kn += 1; }
    }
  }
  if (MODE == 2 && !LOCALF) {
    // stage f (B, 6, NB): element e of the slab = knot e / 6n, row r = (e % 6n) / n, body i = e % n
    const int n6 = 6 * n;
    const int count = nk * n6;
    const int64_t base = first * n6;
    int kn = lane / n6, rem = lane - kn * n6;
    int rr = rem / n, jn = rem - rr * n;
    for (int e = lane; e < 32 * n6; e += 32) {
      fb[(plan.pos[jn] * 6 + rr) * kLrStride + kn] = e < count ? f_out[base + e] : T(0);
      jn += 32;
      while (jn >= n) { jn -= n; rr += 1; }
      while (rr >= 6) { rr -= 6; kn += 1; }
    }
  }
  __syncwarp();

  // ---------------------------------------------------------------- forward sweep (:559-598)
  T vc[6], ac[6];                                         // (v, a) of the body just processed
#pragma unroll(NMAX > 0 ? NMAX : 1)
  for (int i = 0; i < (NMAX > 0 ? NMAX : n); ++i) {
    if (MODE == 2) break;
    if (NMAX > 0 && i >= n) break;
    const int par = m.parent[i];
    const int kind = m.kind[i];
    const T qi = IO(i, 0), qdi = IO(i, 1), qddi = IO(i, 2);
    T f1, f2;
    if (kind == 0) sincos_t(qi, &f2, &f1);
    else { f1 = qi; f2 = T(0); }
    IO(i, 0) = f1; IO(i, 1) = f2;
    T vp[6], ap[6];
    if (par < 0) {
#pragma unroll
      for (int k = 0; k < 6; ++k) { vp[k] = T(0); ap[k] = T(0); }
      ap[5] = -gravity;                                                        // :566
    } else if (par != i - 1) {
      const int s = m.slot_a[par];
#pragma unroll
      for (int k = 0; k < 6; ++k) { vp[k] = ST(s, k); ap[k] = ST(s, 6 + k); }
    } else {
#pragma unroll
      for (int k = 0; k < 6; ++k) { vp[k] = vc[k]; ap[k] = ac[k]; }
    }
    T E[9], r[3];
#pragma unroll
    for (int k = 0; k < 9; ++k) E[k] = fma_t(m.EC[i][k], f2, fma_t(m.EB[i][k], f1, m.EA[i][k]));
#pragma unroll
    for (int k = 0; k < 3; ++k) r[k] = fma_t(m.rC[i][k], f2, fma_t(m.rB[i][k], f1, m.rA[i][k]));
    // X x = [E w; E (u + w x r)]                                              (:580, :583)
    T tv[3] = {vp[3], vp[4], vp[5]}, ta[3] = {ap[3], ap[4], ap[5]};
    cross3_add(vp, r, tv);
    cross3_add(ap, r, ta);
#pragma unroll
    for (int rr = 0; rr < 3; ++rr) {
      vc[rr] = E[3 * rr] * vp[0] + E[3 * rr + 1] * vp[1] + E[3 * rr + 2] * vp[2];
      vc[3 + rr] = E[3 * rr] * tv[0] + E[3 * rr + 1] * tv[1] + E[3 * rr + 2] * tv[2];
      ac[rr] = E[3 * rr] * ap[0] + E[3 * rr + 1] * ap[1] + E[3 * rr + 2] * ap[2];
      ac[3 + rr] = E[3 * rr] * ta[0] + E[3 * rr + 1] * ta[1] + E[3 * rr + 2] * ta[2];
    }
    // v += S qd ; a += crm(v) (S qd) + S qdd                                  (:586-590)
    const T ax[3] = {m.axis[i][0], m.axis[i][1], m.axis[i][2]};
    T vJ[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) vJ[k] = ax[k] * qdi;
    if (kind == 0) {
#pragma unroll
      for (int k = 0; k < 3; ++k) vc[k] += vJ[k];
      cross3_add(vc, vJ, ac);                             // w x vJ (vJ x vJ = 0)
      cross3_add(vc + 3, vJ, ac + 3);                     // u x vJ
#pragma unroll
      for (int k = 0; k < 3; ++k) ac[k] = fma_t(ax[k], qddi, ac[k]);
    } else {
#pragma unroll
      for (int k = 0; k < 3; ++k) vc[3 + k] += vJ[k];
      cross3_add(vc, vJ, ac + 3);                         // w x vJ
#pragma unroll
      for (int k = 0; k < 3; ++k) ac[3 + k] = fma_t(ax[k], qddi, ac[3 + k]);
    }
    // f = I a + crf(v) I v                                                    (:595-597)
    {
      const T mi = m.mass[i];
      const T h[3] = {m.h[i][0], m.h[i][1], m.h[i][2]};
      const T Ib[6] = {m.Ib[i][0], m.Ib[i][1], m.Ib[i][2], m.Ib[i][3], m.Ib[i][4], m.Ib[i][5]};
      T Ia[6], Iv[6], x[6];
      rigid_mul(mi, h, Ib, ac, Ia);
      rigid_mul(mi, h, Ib, vc, Iv);
      crf_mul(vc, Iv, x);
#pragma unroll
      for (int k = 0; k < 6; ++k) FB(i, k) = Ia[k] + x[k];
    }
    const int sa = m.slot_a[i];
    if (sa >= 0) {
#pragma unroll
      for (int k = 0; k < 6; ++k) { ST(sa, k) = vc[k]; ST(sa, 6 + k) = ac[k]; }
    }
    if (VAF) {
#pragma unroll
      for (int k = 0; k < 6; ++k) { VB(i, k) = vc[k]; AB(i, k) = ac[k]; }
    }
  }

  // ---------------------------------------------------------------- backward sweep (:600-621)
  T carry[6] = {T(0), T(0), T(0), T(0), T(0), T(0)};      // X^T f of body i + 1 when its parent is i
#pragma unroll(NMAX > 0 ? NMAX : 1)
  for (int i = (NMAX > 0 ? NMAX : n) - 1; i >= 0; --i) {
    if (MODE == 1) break;
    if (NMAX > 0 && i >= n) continue;
    const bool chained = (i != n - 1) && (m.parent[i + 1 < RBD_MAX_DOF ? i + 1 : i] == i);
    T f[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) f[k] = FB(i, k) + (chained ? carry[k] : T(0));
    const int kind = m.kind[i];
    const T ax[3] = {m.axis[i][0], m.axis[i][1], m.axis[i][2]};
    const T ci = kind == 0 ? dot3s(ax, f) : dot3s(ax, f + 3);                  // :613
    T f1 = IO(i, 0), f2 = IO(i, 1);
    if (MODE == 2) {                                      // the slot still holds q
      const T qi = f1;
      if (kind == 0) sincos_t(qi, &f2, &f1);
      else { f1 = qi; f2 = T(0); }
    }
    IO(i, 2) = ci;
    if (VAF || MODE == 2) {
#pragma unroll
      for (int k = 0; k < 6; ++k) FB(i, k) = f[k];                             // the accumulated force (:619, :628)
    }
    const int par = m.parent[i];
    if (par >= 0) {
      T E[9], r[3];
#pragma unroll
      for (int k = 0; k < 9; ++k) E[k] = fma_t(m.EC[i][k], f2, fma_t(m.EB[i][k], f1, m.EA[i][k]));
#pragma unroll
      for (int k = 0; k < 3; ++k) r[k] = fma_t(m.rC[i][k], f2, fma_t(m.rB[i][k], f1, m.rA[i][k]));
      // X^T f = [E^T n + r x (E^T l); E^T l]                                  (:618)
      T t[6];
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) {
        t[cc] = E[cc] * f[0] + E[3 + cc] * f[1] + E[6 + cc] * f[2];
        t[3 + cc] = E[cc] * f[3] + E[3 + cc] * f[4] + E[6 + cc] * f[5];
      }
      cross3_add(r, t + 3, t);
      if (par == i - 1) {
#pragma unroll
        for (int k = 0; k < 6; ++k) carry[k] = t[k];
      } else {
#pragma unroll
        for (int k = 0; k < 6; ++k) FB(par, k) += t[k];
      }
    }
  }
  __syncwarp();

  // ---------------------------------------------------------------- coalesced stores
  if (MODE != 1) {
    const int count = nk * n;
    T* dst = c + first * n;
    int kn = lane / n, jn = lane - kn * n;
    const int dk = 32 / n, dj = 32 - dk * n;
    for (int e = lane; e < count; e += 32) {
      __stcs(dst + e, io[(plan.pos[jn] * 3 + 2) * kLrStride + kn]);
      kn += dk; jn += dj;
      if (jn >= n) { jn -= n; kn += 1; }
    }
  }
  if (VAF || MODE == 2) {
    if (LOCALF) {
      // not instantiated: these modes need the shared-memory f rows
    } else {
      // (B, 6, NB): element e of the slab = knot e / 6n, row r = (e % 6n) / n, body i = e % n
      const int n6 = 6 * n;
      const int count = nk * n6;
      const int64_t base = first * n6;
      int kn = lane / n6, rem = lane - kn * n6;
      int rr = rem / n, jn = rem - rr * n;
      for (int e = lane; e < count; e += 32) {
        const int row = (plan.pos[jn] * 6 + rr) * kLrStride + kn;
        if (VAF && v_out) __stcs(v_out + base + e, vb[row]);
        if (VAF && a_out) __stcs(a_out + base + e, ab[row]);
        if (f_out) __stcs(f_out + base + e, fb[row]);
        // advance e by 32
        jn += 32;
        while (jn >= n) { jn -= n; rr += 1; }
        while (rr >= 6) { rr -= 6; kn += 1; }
      }
    }
  }
#undef IO
#undef FB
#undef VB
#undef AB
#undef ST
}

}  // namespace rbd
