// rbd_chain_grad_kernels.cuh - fused rnea_grad for serial chains (iiwa14 and other arms): one knot
// point per lane, no shuffles, nothing idle.
//
// Same world-frame composite formulation as rbd_grad_kernels.cuh (RBDReference.py:1345-1368 is the
// object computed), specialised at compile time for parent[i] = i - 1 and N bodies:
//
//   * stage 0: cos / sin of all N joint angles at once (N independent polynomial chains in flight);
//   * forward sweep (rolled over the bodies): pose, v, a in registers; S_i, Psi_dot_i, Psi_ddot_i of
//     bodies 1 .. N-2 and S_0 go to a [row][pair][lane] shared-memory table (16-byte accesses,
//     conflict free).  Psi_dot_0 = 0 and Psi_ddot_0 = a_base x S_0 (the base does not move) are
//     never stored; the leaf's vectors stay in registers;
//   * backward sweep (rolled): the pose / velocity / acceleration of the parent are re-derived from
//     the child's (nothing but the table is stored), composites are a running sum in registers,
//     F1..F4 of body i meet the table rows of its ancestors in a pair loop unrolled over the
//     ancestor index (table offsets and result registers are compile-time);
//   * results: row i of dc_du is complete at body i and leaves straight from registers as 16-byte
//     stores; the entries body i produces for the rows of its ancestors ([j, i], j < i) wait in the
//     table row of body i, which nobody reads any more (the leaf has no row: its entries go to
//     global memory as 8-byte stores into sectors the row stores complete later).
//     Shared memory per knot point: 10 (N - 2) + 4 pairs of values = 27 KB per warp for N = 7 in
//     FP64, i.e. eight resident warps per SM (two per scheduler) within 255 registers.
//
// The body loops stay rolled on purpose: the straight-line version of the per-knot-point kernel was
// measured 28 % slower (instruction cache, rbd_launch_grad.cu).
#pragma once
#include "rbd_common.cuh"
#include "rbd_grad_kernels.cuh"

namespace rbd {

constexpr int kChainRowPairs = 10;      // S(3) Psi_dot(3) Psi_ddot(3) (cos, sin)(1) pairs per table row
constexpr int kChainRow0Pairs = 4;      // row 0: S(3) (cos, sin)(1)

// rows 1 .. n-2 first, row 0 last
__host__ __device__ inline size_t chain_grad_smem_pairs(int n) {
  return (size_t)(n > 2 ? n - 2 : 0) * kChainRowPairs + kChainRow0Pairs;
}

constexpr int kChainMaxN = 8;

// The robot as this kernel reads it: one contiguous record per body (the kernel walks the bodies with a run-time
// index, so its constant-bank reads are indexed loads; a body's 50 values in 4 consecutive cache lines instead of
// ten tables 2.3 KB apart keep them in the first-level constant cache - ncu: short-scoreboard stalls on the FMAs
// that consume them).
template <typename T>
struct ChainModel {
  int n;
  int r_const;                    // rB = rC = 0 for every body: r(q) = rA (revolute joints through the child's origin)
  int kind[kChainMaxN];
  struct Body {
    T EA[9], EB[9], EC[9];        // E_J(q) = EA + EB f1 + EC f2
    T rA[3], rB[3], rC[3];
    T axis[3];
    T mass, h[3], Ib[6], damping;
  } b[kChainMaxN];
};

template <typename T> __device__ __forceinline__ void stcs2(T* p, T x, T y);
template <> __device__ __forceinline__ void stcs2<double>(double* p, double x, double y) {
  __stcs(reinterpret_cast<double2*>(p), make_double2(x, y));
}
template <> __device__ __forceinline__ void stcs2<float>(float* p, float x, float y) {
  __stcs(reinterpret_cast<float2*>(p), make_float2(x, y));
}

template <typename T, int N>
__global__ void __launch_bounds__(32, 8)
rnea_grad_chain_kernel(const __grid_constant__ ChainModel<T> m, int64_t B, const T* __restrict__ q,
                       const T* __restrict__ qd, const T* __restrict__ qdd, T gravity, int use_damping,
                       T* __restrict__ dc_du, T* __restrict__ c_out, int ahead) {
  static_assert(N >= 3, "chain kernel needs at least three bodies");
  typedef typename Vec2<T>::type V2;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  V2* tab = reinterpret_cast<V2*>(smem_raw) + threadIdx.x;       // pair k of row r: tab[(rowbase(r) + k) * 32]
  constexpr int RS = kChainRowPairs * 32;                          // row stride in pairs (rows 1 .. N-2)
  constexpr int ROW0 = (N - 2) * kChainRowPairs * 32;              // row 0: S pairs 0..2, (cos, sin) pair 3
  // row r >= 1 starts at (r - 1) * RS; its (cos, sin) pair is pair 9
  int64_t b = (int64_t)blockIdx.x * 32 + threadIdx.x;
  const bool active = b < B;
  if (!active) b = B - 1;                                          // keep the warp convergent; stores are masked
  const T* qb = q + b * N;
  const T* qdb = qd + b * N;
  const T* qddb = qdd ? qdd + b * N : nullptr;

  // the inputs of the CTA that will run in this slot next (`ahead` = resident CTAs of the grid) -> L2
  {
    const int64_t nxt = (int64_t)blockIdx.x + ahead;
    constexpr int kBytes = 32 * N * (int)sizeof(T);
    if ((nxt + 1) * 32 <= B && (int)threadIdx.x * 128 < kBytes + 127) {
      const size_t off = (size_t)nxt * kBytes + threadIdx.x * 128;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(q) + off));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(qd) + off));
      if (qdd) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(qdd) + off));
    }
  }
  // ------------------------------------------------------------------ stage 0: cos / sin of every joint angle
  T f1 = T(1), f2 = T(0);                                          // the leaf's (cos, sin) / prismatic (q, 0)
  {
    T qv[N], sv[N], cv[N];
#pragma unroll
    for (int i = 0; i < N; ++i) qv[i] = __ldg(qb + i);
    // FP64: N polynomial chains in flight.  FP32: sincosf is evaluated inside the (rolled) forward sweep instead -
    // N inlined copies of it made the straight-line stage larger than the instruction cache likes (0.35 -> 0.39 ms)
    if (sizeof(T) == 8) sincos_batch<N>(qv, sv, cv);
#pragma unroll
    for (int i = 0; i < N; ++i) {
      V2 t;
      if (sizeof(T) == 8 && m.kind[i] == 0) { t.x = cv[i]; t.y = sv[i]; }
      else { t.x = qv[i]; t.y = T(0); }
      if (i == 0) tab[ROW0 + 3 * 32] = t;
      else if (i < N - 1) tab[(i - 1) * RS + 9 * 32] = t;
      else { f1 = t.x; f2 = t.y; }
    }
  }

  T E[9], p[3], v[6], a[6];
  T S[6], Pd[6], Pdd[6];
#pragma unroll
  for (int k = 0; k < 9; ++k) E[k] = (k % 4 == 0) ? T(1) : T(0);
#pragma unroll
  for (int k = 0; k < 3; ++k) p[k] = T(0);
#pragma unroll
  for (int k = 0; k < 6; ++k) { v[k] = T(0); a[k] = T(0); }
  a[5] = -gravity;                                                 // RBDReference.py:566

  // ------------------------------------------------------------------ forward sweep
  T qd_nx = __ldg(qdb), qdd_nx = qddb ? __ldg(qddb) : T(0);
#pragma unroll 1
  for (int i = 0; i < N; ++i) {
    const typename ChainModel<T>::Body& mb = m.b[i];
    const T qdi = qd_nx, qddi = qdd_nx;
    if (i + 1 < N) {
      qd_nx = __ldg(qdb + i + 1);
      if (qddb) qdd_nx = __ldg(qddb + i + 1);
    }
    V2* row = tab + (i == 0 ? ROW0 : (i - 1) * RS);
    T c1 = f1, c2 = f2;
    if (i < N - 1) {
      const V2 t = row[(i == 0 ? 3 : 9) * 32];
      c1 = t.x; c2 = t.y;
    }
    const int kind = m.kind[i];
    if (sizeof(T) == 4 && kind == 0) {                    // FP32: (q, 0) was staged; the pair is rewritten for the way back
      sincos_t(c1, &c2, &c1);
      if (i < N - 1) { V2 t; t.x = c1; t.y = c2; row[(i == 0 ? 3 : 9) * 32] = t; }
      else { f1 = c1; f2 = c2; }
    }
    {
      T r[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) r[k] = m.r_const ? mb.rA[k] : fma_t(mb.rC[k], c2, fma_t(mb.rB[k], c1, mb.rA[k]));
      // p_i = p_parent + E_parent^T r
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) p[cc] = fma_t(E[6 + cc], r[2], fma_t(E[3 + cc], r[1], fma_t(E[cc], r[0], p[cc])));
      // E_i = Ej E_parent, column by column
      T Ej[9];
#pragma unroll
      for (int k = 0; k < 9; ++k) Ej[k] = fma_t(mb.EC[k], c2, fma_t(mb.EB[k], c1, mb.EA[k]));
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) {
        const T t0 = E[cc], t1 = E[3 + cc], t2 = E[6 + cc];
#pragma unroll
        for (int rr = 0; rr < 3; ++rr) E[3 * rr + cc] = fma_t(Ej[3 * rr + 2], t2, fma_t(Ej[3 * rr + 1], t1, Ej[3 * rr] * t0));
      }
    }
    {
      T w[3];
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) w[cc] = E[cc] * mb.axis[0] + E[3 + cc] * mb.axis[1] + E[6 + cc] * mb.axis[2];
      if (kind == 0) {
        S[0] = w[0]; S[1] = w[1]; S[2] = w[2];
        cross3(p, w, S + 3);
      } else {
        S[0] = S[1] = S[2] = T(0);
        S[3] = w[0]; S[4] = w[1]; S[5] = w[2];
      }
    }
    crm_mul(v, S, Pd);                        // Psi_dot  = v_parent x S
    {
      T t6[6];
      crm_mul(a, S, Pdd);                     // Psi_ddot = a_parent x S + v_parent x Psi_dot
      crm_mul(v, Pd, t6);
#pragma unroll
      for (int k = 0; k < 6; ++k) Pdd[k] += t6[k];
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      v[k] = fma_t(S[k], qdi, v[k]);
      a[k] = fma_t(Pd[k], qdi, fma_t(S[k], qddi, a[k]));
    }
    if (i < N - 1) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        V2 t;
        t.x = S[2 * k]; t.y = S[2 * k + 1]; row[k * 32] = t;
      }
      if (i > 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          V2 t;
          t.x = Pd[2 * k]; t.y = Pd[2 * k + 1]; row[(3 + k) * 32] = t;
          t.x = Pdd[2 * k]; t.y = Pdd[2 * k + 1]; row[(6 + k) * 32] = t;
        }
      }
    }
  }

  // ------------------------------------------------------------------ backward sweep
  // composites: 0 m | 1..3 h | 4..9 Ibar | 10..15 Sym | 16..18 n | 19..21 l | 22..27 f
  T acc[28];
#pragma unroll
  for (int k = 0; k < 28; ++k) acc[k] = T(0);
  T* out = dc_du + b * (int64_t)(2 * N * N);
  const T ag = -gravity;                      // a_base = [0 0 0 0 0 ag]: Psi_ddot_0 = [0; (-ag S0y, ag S0x, 0)]

#pragma unroll 1
  for (int i = N - 1; i >= 0; --i) {
    const typename ChainModel<T>::Body& mb = m.b[i];
    V2* rowi = tab + (i == 0 ? ROW0 : (i - 1) * RS);   // own table row; then the pending entries [j, i], j < i
    if (i < N - 1) {
#pragma unroll
      for (int k = 0; k < 3; ++k) { const V2 t = rowi[k * 32]; S[2 * k] = t.x; S[2 * k + 1] = t.y; }
      if (i > 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          V2 t = rowi[(3 + k) * 32]; Pd[2 * k] = t.x; Pd[2 * k + 1] = t.y;
          t = rowi[(6 + k) * 32]; Pdd[2 * k] = t.x; Pdd[2 * k + 1] = t.y;
        }
        const V2 t = rowi[9 * 32];
        f1 = t.x; f2 = t.y;
      } else {
#pragma unroll
        for (int k = 0; k < 6; ++k) { Pd[k] = T(0); Pdd[k] = T(0); }
        Pdd[3] = -ag * S[1];
        Pdd[4] = ag * S[0];
      }
    }
    T qdi = T(0), qddi = T(0);
    if (i > 0) {
      qdi = __ldg(qdb + i);
      if (qddb) qddi = __ldg(qddb + i);
    }
    // ---- own rigid-body terms in world coordinates, added to the running composites
    {
      const T mi = mb.mass;
      T hr[3], hw[3];
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) {
        hr[cc] = E[cc] * mb.h[0] + E[3 + cc] * mb.h[1] + E[6 + cc] * mb.h[2];
        hw[cc] = fma_t(mi, p[cc], hr[cc]);
      }
      T IbE[9];
      {
        const T xx = mb.Ib[0], xy = mb.Ib[1], xz = mb.Ib[2], yy = mb.Ib[3], yz = mb.Ib[4], zz = mb.Ib[5];
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
          IbE[cc] = xx * E[cc] + xy * E[3 + cc] + xz * E[6 + cc];
          IbE[3 + cc] = xy * E[cc] + yy * E[3 + cc] + yz * E[6 + cc];
          IbE[6 + cc] = xz * E[cc] + yz * E[3 + cc] + zz * E[6 + cc];
        }
      }
      T Iw[6];
      {
        const T tr = (hr[0] + hw[0]) * p[0] + (hr[1] + hw[1]) * p[1] + (hr[2] + hw[2]) * p[2];
        int idx = 0;
#pragma unroll
        for (int rr = 0; rr < 3; ++rr)
#pragma unroll
          for (int cc = rr; cc < 3; ++cc) {
            T val = E[rr] * IbE[cc] + E[3 + rr] * IbE[3 + cc] + E[6 + rr] * IbE[6 + cc];
            val -= hr[rr] * p[cc] + p[rr] * hw[cc];
            if (rr == cc) val += tr;
            Iw[idx++] = val;
          }
      }
      T mom[6], fo[6], t6[6];
      rigid_mul(mi, hw, Iw, v, mom);
      rigid_mul(mi, hw, Iw, a, fo);
      crf_mul(v, mom, t6);
      const T* wv = v;
      const T* uv = v + 3;
      T M[9];
      {
        const T Im[9] = {Iw[0], Iw[1], Iw[2], Iw[1], Iw[3], Iw[4], Iw[2], Iw[4], Iw[5]};
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
          M[cc] = wv[1] * Im[6 + cc] - wv[2] * Im[3 + cc];
          M[3 + cc] = wv[2] * Im[cc] - wv[0] * Im[6 + cc];
          M[6 + cc] = wv[0] * Im[3 + cc] - wv[1] * Im[cc];
        }
      }
      const T uh2 = T(2) * (uv[0] * hw[0] + uv[1] * hw[1] + uv[2] * hw[2]);
      acc[0] += mi;
      acc[1] += hw[0]; acc[2] += hw[1]; acc[3] += hw[2];
#pragma unroll
      for (int k = 0; k < 6; ++k) acc[4 + k] += Iw[k];
      acc[10] += T(2) * M[0] - T(2) * hw[0] * uv[0] + uh2;
      acc[11] += M[1] + M[3] - (hw[0] * uv[1] + uv[0] * hw[1]);
      acc[12] += M[2] + M[6] - (hw[0] * uv[2] + uv[0] * hw[2]);
      acc[13] += T(2) * M[4] - T(2) * hw[1] * uv[1] + uh2;
      acc[14] += M[5] + M[7] - (hw[1] * uv[2] + uv[1] * hw[2]);
      acc[15] += T(2) * M[8] - T(2) * hw[2] * uv[2] + uh2;
#pragma unroll
      for (int k = 0; k < 6; ++k) { acc[16 + k] += mom[k]; acc[22 + k] += fo[k] + t6[k]; }
    }
    const T mC = acc[0];
    const T* hC = acc + 1;
    const T* IC = acc + 4;
    const T* SyC = acc + 10;
    const T* nC = acc + 16;
    const T* lC = acc + 19;
    const T* fC = acc + 22;
    const int kind = m.kind[i];
    // ---- F vectors of body i
    T F1[6], F2[6], F3[3], F4[6];
    rigid_mul(mC, hC, IC, S, F4);
    {
      T t3[3];
      sym3_mul(SyC, S, F3);
      cross3_add(nC, S, F3);
      cross3(lC, S + 3, t3);
#pragma unroll
      for (int k = 0; k < 3; ++k) F3[k] = fma_t(T(0.5), F3[k], t3[k]);
    }
    {
      T t6[6], tb[3], tl[3];
      rigid_mul(mC, hC, IC, Pdd, F1);
      crf_mul(S, fC, t6);
      sym3_mul(SyC, Pd, tb);
      cross3(nC, Pd, tl);
#pragma unroll
      for (int k = 0; k < 3; ++k) F1[k] += t6[k] + tb[k] - tl[k];
      cross3(lC, Pd, tl);
#pragma unroll
      for (int k = 0; k < 3; ++k) F1[3 + k] += t6[3 + k] - T(2) * tl[k];
      rigid_mul(mC, hC, IC, Pd, F2);
      sym3_mul(SyC, S, tb);
      cross3(nC, S, tl);
#pragma unroll
      for (int k = 0; k < 3; ++k) F2[k] = T(2) * F2[k] + tb[k] - tl[k];
      cross3(lC, S, tl);
#pragma unroll
      for (int k = 0; k < 3; ++k) F2[3 + k] = T(2) * (F2[3 + k] - tl[k]);
    }
    if (c_out && active) c_out[b * N + i] = dot6s(S, fC);                       // :613
    T dqq = dot6s(S, F1);
    T ddd = dot6s(S, F2);
    if (use_damping) ddd += mb.damping;                                        // :1341
    if (kind == 1) {
      // reference quirk for prismatic joints: X^T(-crm(f)S) instead of X^T(S x* f) (:1292); needs p_i
      T nrot[3], dl[3], da[3], t3[3];
      cross3(p, fC + 3, t3);
#pragma unroll
      for (int k = 0; k < 3; ++k) nrot[k] = fC[k] - t3[k];
      cross3(S + 3, nrot, dl);
      cross3(S + 3, fC + 3, da);
      cross3(p, dl, t3);
#pragma unroll
      for (int k = 0; k < 3; ++k) { F1[k] += t3[k] - da[k]; F1[3 + k] += dl[k]; }
    }
    // ---- walk the running state back to the parent
    if (i > 0) {
      T r[3], Ej[9];
#pragma unroll
      for (int k = 0; k < 3; ++k) r[k] = m.r_const ? mb.rA[k] : fma_t(mb.rC[k], f2, fma_t(mb.rB[k], f1, mb.rA[k]));
#pragma unroll
      for (int k = 0; k < 9; ++k) Ej[k] = fma_t(mb.EC[k], f2, fma_t(mb.EB[k], f1, mb.EA[k]));
      // E_parent = Ej^T E_i, column by column; p_parent = p_i - E_parent^T r
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) {
        const T t0 = E[cc], t1 = E[3 + cc], t2 = E[6 + cc];
#pragma unroll
        for (int rr = 0; rr < 3; ++rr) E[3 * rr + cc] = fma_t(Ej[6 + rr], t2, fma_t(Ej[3 + rr], t1, Ej[rr] * t0));
        p[cc] -= E[cc] * r[0] + E[3 + cc] * r[1] + E[6 + cc] * r[2];
      }
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        v[k] = fma_t(-S[k], qdi, v[k]);
        a[k] = fma_t(-Pd[k], qdi, fma_t(-S[k], qddi, a[k]));
      }
    }
    // ---- row i of [dc_dq | dc_dqd]: column j < i from the pair (i, j), column i the diagonal,
    //      column k > i from the entries body k left in its table row (the leaf's are in HBM already)
    T rowv[2 * N];
#pragma unroll
    for (int j = 0; j < N; ++j) {
      if (j < i) {
        T Sj[6];
        V2 pend;
        if (j == 0) {
          const V2* rowj = tab + ROW0;
#pragma unroll
          for (int k = 0; k < 3; ++k) { const V2 t = rowj[k * 32]; Sj[2 * k] = t.x; Sj[2 * k + 1] = t.y; }
          // Psi_dot_0 = 0, Psi_ddot_0 = [0; (-ag S0y, ag S0x, 0)]
          rowv[j] = ag * (F4[4] * Sj[0] - F4[3] * Sj[1]);                        // dc_dq [i, 0]
          rowv[N + j] = T(2) * dot3s(F3, Sj);                                    // dc_dqd[i, 0]
        } else {
          const V2* rowj = tab + (j - 1) * RS;
          T Pdj[6], Pddj[6];
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            V2 t = rowj[k * 32]; Sj[2 * k] = t.x; Sj[2 * k + 1] = t.y;
            t = rowj[(3 + k) * 32]; Pdj[2 * k] = t.x; Pdj[2 * k + 1] = t.y;
            t = rowj[(6 + k) * 32]; Pddj[2 * k] = t.x; Pddj[2 * k + 1] = t.y;
          }
          rowv[j] = fma_t(T(2), dot3s(F3, Pdj), dot6s(F4, Pddj));                // dc_dq [i, j]
          rowv[N + j] = T(2) * (dot6s(F4, Pdj) + dot3s(F3, Sj));                 // dc_dqd[i, j]
        }
        pend.x = dot6s(Sj, F1);                                                  // dc_dq [j, i]
        pend.y = dot6s(Sj, F2);                                                  // dc_dqd[j, i]
        if (i == N - 1) {
          if (active) { __stcs(out + j * (2 * N) + (N - 1), pend.x); __stcs(out + j * (2 * N) + (2 * N - 1), pend.y); }
        } else {
          rowi[j * 32] = pend;
        }
      } else if (j == i) {
        rowv[j] = dqq;
        rowv[N + j] = ddd;
      } else if (j < N - 1) {
        const V2 pend = tab[(j - 1) * RS + i * 32];
        rowv[j] = pend.x;
        rowv[N + j] = pend.y;
      } else {
        rowv[j] = T(0);                       // leaf column of an inner row: already stored by the leaf
        rowv[N + j] = T(0);
      }
    }
    if (active) {
      T* orow = out + i * (2 * N);
      if (i == N - 1) {
#pragma unroll
        for (int k = 0; k < N; ++k) stcs2<T>(orow + 2 * k, rowv[2 * k], rowv[2 * k + 1]);
      } else {
#pragma unroll
        for (int k = 0; k < N; ++k) {
          const bool leaf0 = (2 * k == N - 1) || (2 * k == 2 * N - 1);
          const bool leaf1 = (2 * k + 1 == N - 1) || (2 * k + 1 == 2 * N - 1);
          if (!leaf0 && !leaf1) stcs2<T>(orow + 2 * k, rowv[2 * k], rowv[2 * k + 1]);
          else {
            if (!leaf0) __stcs(orow + 2 * k, rowv[2 * k]);
            if (!leaf1) __stcs(orow + 2 * k + 1, rowv[2 * k + 1]);
          }
        }
      }
    }
  }
}

}  // namespace rbd
