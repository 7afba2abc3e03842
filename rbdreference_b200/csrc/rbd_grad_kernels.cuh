// rbd_grad_kernels.cuh - fast fused rnea_grad: world-frame composite formulation.
//
// Same mathematical object as RBDReference.py:1345-1368 (dc_du = [dc/dq | dc/dqd], the exact
// partial derivatives of RNEA), evaluated with O(n) six-vectors of state per knot point instead
// of the reference's O(n^2) body-frame derivative columns:
//
//   forward  (root -> leaf): world pose (E_i, p_i), world joint axis S_i, v_i, a_i and
//            Psi_dot_i = v_parent x S_i,   Psi_ddot_i = a_parent x S_i + v_parent x Psi_dot_i
//   backward (leaf -> root): composite rigid-body inertia I^C, composite momentum h^C,
//            composite "Coriolis" matrix B^C and composite force f^C of every subtree - all sums
//            of world-frame quantities, so no 6x6 transform is ever applied - then
//              F1 = I^C Psi_ddot_i + S_i x* f^C + 2 B^C Psi_dot_i
//              F2 = 2 I^C Psi_dot_i + 2 B^C S_i        F3 = B^C^T S_i        F4 = I^C S_i
//            and for every ancestor j of i (and j = i on the diagonal):
//              dc_dq [j,i] = S_j.F1                 dc_dq [i,j] = 2 F3.Psi_dot_j + F4.Psi_ddot_j
//              dc_dqd[j,i] = S_j.F2                 dc_dqd[i,j] = 2 (F4.Psi_dot_j + F3.S_j)
//   B(I,v) = 1/2 (v x* I - I v x + (I v) xbar*) has the block form [[1/2(Sym - n x), 0], [-l x, 0]]
//   for a rigid-body inertia, with (n,l) = I v the momentum, so B^C is 12 numbers (6 + 6).
//
// Mapping: one knot point per thread, one warp (32 knot points) per CTA.  The per-body vectors
// S, Psi_dot, Psi_ddot live in shared memory in [slot][lane] order (bank-conflict free); pose,
// velocity, acceleration and the composites are registers.  The backward sweep re-derives each
// parent's pose/velocity from its child's (reverse walk) instead of storing them; only
// branch points and leaves keep a stash.  The robot is a __grid_constant__ parameter.
//
// The reference's extra term in rnea_grad_bpass_dq uses -crm(f)S (RBDReference.py:1292, :166-168),
// which equals S x* f only for revolute joints.  For prismatic joints the same (non-covariant)
// expression is reproduced here in body coordinates so results match the reference.
#pragma once
#include "rbd_common.cuh"

namespace rbd {

template <typename T>
struct FastModel {
  int n;
  int n_slot_a;                 // stash slots of 28 values (forward parent state / composites)
  int n_slot_b;                 // stash slots of 24 values (states of reverse-chain starts)
  int rigid;                    // 1 iff every spatial inertia has rigid-body structure
  int has_prismatic;
  int parent[RBD_MAX_DOF];
  int kind[RBD_MAX_DOF];
  int slot_a[RBD_MAX_DOF];      // >= 0 for bodies with a child c != i + 1
  int slot_b[RBD_MAX_DOF];      // >= 0 for bodies i != n-1 with parent[i+1] != i
  unsigned anc_mask[RBD_MAX_DOF];
  unsigned sub_mask[RBD_MAX_DOF];
  T damping[RBD_MAX_DOF];
  T EA[RBD_MAX_DOF][9], EB[RBD_MAX_DOF][9], EC[RBD_MAX_DOF][9];   // E_J(q) = EA + EB f1 + EC f2
  T rA[RBD_MAX_DOF][3], rB[RBD_MAX_DOF][3], rC[RBD_MAX_DOF][3];   // r(q), X = [[E,0],[-E rx,E]]
  T axis[RBD_MAX_DOF][3];       // joint axis in body coordinates
  T mass[RBD_MAX_DOF];
  T h[RBD_MAX_DOF][3];          // m * com
  T Ib[RBD_MAX_DOF][6];         // rotational inertia about the body origin: xx xy xz yy yz zz
};

constexpr int kVecPerBody = 22;   // S(6) Psi_dot(6) Psi_ddot(6) f1 f2 qd qdd

template <typename T> struct Vec2;
template <> struct Vec2<double> { typedef double2 type; };
template <> struct Vec2<float> { typedef float2 type; };

// two 3-term chains instead of one 6-term chain: halves the dependent-FMA depth
template <typename T>
__device__ __forceinline__ T dot6s(const T* s, const T* x) {
  const T lo = fma_t(s[2], x[2], fma_t(s[1], x[1], s[0] * x[0]));
  const T hi = fma_t(s[5], x[5], fma_t(s[4], x[4], s[3] * x[3]));
  return lo + hi;
}
template <typename T>
__device__ __forceinline__ T dot3s(const T* s, const T* x) {
  return fma_t(s[2], x[2], fma_t(s[1], x[1], s[0] * x[0]));
}

// ---- small helpers -------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void cross3(const T* a, const T* b, T* y) {
  y[0] = a[1] * b[2] - a[2] * b[1];
  y[1] = a[2] * b[0] - a[0] * b[2];
  y[2] = a[0] * b[1] - a[1] * b[0];
}
template <typename T>
__device__ __forceinline__ void cross3_add(const T* a, const T* b, T* y) {
  y[0] += a[1] * b[2] - a[2] * b[1];
  y[1] += a[2] * b[0] - a[0] * b[2];
  y[2] += a[0] * b[1] - a[1] * b[0];
}
// y = Sym x, Sym = xx xy xz yy yz zz
template <typename T>
__device__ __forceinline__ void sym3_mul(const T* s, const T* x, T* y) {
  y[0] = s[0] * x[0] + s[1] * x[1] + s[2] * x[2];
  y[1] = s[1] * x[0] + s[3] * x[1] + s[4] * x[2];
  y[2] = s[2] * x[0] + s[4] * x[1] + s[5] * x[2];
}
// y = I x for a rigid-body inertia (m, h, Ibar): [Ibar w + h x u ; m u - h x w]
template <typename T>
__device__ __forceinline__ void rigid_mul(T m, const T* h, const T* Ib, const T* x, T* y) {
  sym3_mul(Ib, x, y);
  cross3_add(h, x + 3, y);
  T t[3];
  cross3(h, x, t);
  y[3] = m * x[3] - t[0];
  y[4] = m * x[4] - t[1];
  y[5] = m * x[5] - t[2];
}

// =============================================================================================
// LOCAL == 0: per-body vectors in shared memory (robots whose working set fits: a warp's 32 knot
// points need (22 n + stashes) * 32 values).  LOCAL == 1: the same vectors in per-thread local
// memory (L1/L2-backed) for large trees, trading on-chip latency for 8 resident warps per SM.
template <typename T, int LOCAL, int MINB>
__global__ void __launch_bounds__(32, MINB)
rnea_grad_world_kernel(const __grid_constant__ FastModel<T> m, int64_t B, const T* __restrict__ q,
                       const T* __restrict__ qd, const T* __restrict__ qdd, T gravity,
                       int use_damping, T* __restrict__ dc_du, T* __restrict__ c_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* sm = reinterpret_cast<T*>(smem_raw);
  const int lane = threadIdx.x;
  const int n = m.n;
  int64_t b = (int64_t)blockIdx.x * 32 + lane;
  const bool active = b < B;
  if (!active) b = B - 1;                      // keep the warp convergent; stores are masked
  // element k of body i lives at ((i*11 + k/2)*32 + lane)*2 + (k&1): each lane owns aligned
  // pairs, so a 6-vector is three 16-byte shared-memory accesses (LDS.128 / STS.128 in FP64)
  typedef typename Vec2<T>::type V2;
  V2 lvec[LOCAL ? RBD_MAX_DOF * (kVecPerBody / 2) : 1];
#define VEC2(i, k2) (*(LOCAL ? &lvec[(i) * (kVecPerBody / 2) + (k2)] \
                             : &reinterpret_cast<V2*>(sm)[(((i) * (kVecPerBody / 2)) + (k2)) * 32 + lane]))
#define VEC(i, k) (reinterpret_cast<T*>(&VEC2(i, (k) >> 1))[(k) & 1])
  T* stash_a = sm + (LOCAL ? 0 : (size_t)n * kVecPerBody * 32);   // [slot][28][32]
  T* stash_b = stash_a + (size_t)m.n_slot_a * 28 * 32;            // [slot][24][32]
#define STA(s, k) stash_a[((s) * 28 + (k)) * 32 + lane]
#define STB(s, k) stash_b[((s) * 24 + (k)) * 32 + lane]

  // stage this warp's contiguous slab of (q, qd, qdd) with coalesced loads: element e of the
  // slab belongs to knot point e / n, joint e % n and goes to that lane's slots 18 / 20 / 21
  if (LOCAL) {
    for (int i = 0; i < n; ++i) {
      VEC(i, 18) = q[b * n + i];
      VEC(i, 20) = qd[b * n + i];
      VEC(i, 21) = qdd ? qdd[b * n + i] : T(0);
    }
  } else {
    const int64_t base = (int64_t)blockIdx.x * 32 * n;
    const int64_t limit = B * (int64_t)n;
    int inst = lane / n, jnt = lane - inst * n;
    const int dinst = 32 / n, djnt = 32 - dinst * n;
    for (int e = lane; e < 32 * n; e += 32) {
      const int64_t g = base + e;
      const bool ok = g < limit;
      const T vq = ok ? q[g] : T(0);
      const T vqd = ok ? qd[g] : T(0);
      const T vqdd = (ok && qdd) ? qdd[g] : T(0);
      T* dst = sm + ((((jnt * (kVecPerBody / 2)) + 9) * 32 + inst) << 1);
      dst[0] = vq;                      // slot 18 (converted to f1 in the forward sweep)
      dst[64] = vqd;                    // slot 20
      dst[65] = vqdd;                   // slot 21
      inst += dinst; jnt += djnt;
      if (jnt >= n) { jnt -= n; inst += 1; }
    }
    __syncwarp();
  }

  // running state of the body processed last: E (world -> body), p (origin in world), v, a
  T E[9], p[3], v[6], a[6];

  // ------------------------------------------------------------------ forward sweep
#pragma unroll 1
  for (int i = 0; i < n; ++i) {
    T f1, f2;
    {
      const T qi = VEC(i, 18);
      if (m.kind[i] == 0) sincos_t(qi, &f2, &f1);
      else { f1 = qi; f2 = T(0); }
    }
    const T qdi = VEC(i, 20);
    const T qddi = VEC(i, 21);
    const int par = m.parent[i];
    T Ep[9], pp[3], vp[6], ap[6];
    if (par < 0) {
#pragma unroll
      for (int k = 0; k < 9; ++k) Ep[k] = (k % 4 == 0) ? T(1) : T(0);
#pragma unroll
      for (int k = 0; k < 3; ++k) pp[k] = T(0);
#pragma unroll
      for (int k = 0; k < 6; ++k) { vp[k] = T(0); ap[k] = T(0); }
      ap[5] = -gravity;                                                      // RBDReference.py:566
    } else if (par != i - 1) {
      const int s = m.slot_a[par];
#pragma unroll
      for (int k = 0; k < 9; ++k) Ep[k] = STA(s, k);
#pragma unroll
      for (int k = 0; k < 3; ++k) pp[k] = STA(s, 9 + k);
#pragma unroll
      for (int k = 0; k < 6; ++k) { vp[k] = STA(s, 12 + k); ap[k] = STA(s, 18 + k); }
    } else {
#pragma unroll
      for (int k = 0; k < 9; ++k) Ep[k] = E[k];
#pragma unroll
      for (int k = 0; k < 3; ++k) pp[k] = p[k];
#pragma unroll
      for (int k = 0; k < 6; ++k) { vp[k] = v[k]; ap[k] = a[k]; }
    }
    // joint transform pieces
    T Ej[9], r[3];
#pragma unroll
    for (int k = 0; k < 9; ++k) Ej[k] = fma_t(m.EC[i][k], f2, fma_t(m.EB[i][k], f1, m.EA[i][k]));
#pragma unroll
    for (int k = 0; k < 3; ++k) r[k] = fma_t(m.rC[i][k], f2, fma_t(m.rB[i][k], f1, m.rA[i][k]));
    // pose: E_i = Ej * Ep ; p_i = pp + Ep^T r
#pragma unroll
    for (int rr = 0; rr < 3; ++rr)
#pragma unroll
      for (int cc = 0; cc < 3; ++cc)
        E[3 * rr + cc] = Ej[3 * rr] * Ep[cc] + Ej[3 * rr + 1] * Ep[3 + cc] + Ej[3 * rr + 2] * Ep[6 + cc];
#pragma unroll
    for (int cc = 0; cc < 3; ++cc) p[cc] = pp[cc] + Ep[cc] * r[0] + Ep[3 + cc] * r[1] + Ep[6 + cc] * r[2];
    // world joint axis S
    T S[6], w[3];
#pragma unroll
    for (int cc = 0; cc < 3; ++cc)
      w[cc] = E[cc] * m.axis[i][0] + E[3 + cc] * m.axis[i][1] + E[6 + cc] * m.axis[i][2];
    if (m.kind[i] == 0) {
      S[0] = w[0]; S[1] = w[1]; S[2] = w[2];
      cross3(p, w, S + 3);
    } else {
      S[0] = S[1] = S[2] = T(0);
      S[3] = w[0]; S[4] = w[1]; S[5] = w[2];
    }
    T Pd[6], Pdd[6], t6[6];
    crm_mul(vp, S, Pd);                       // Psi_dot  = v_parent x S
    crm_mul(ap, S, Pdd);                      // Psi_ddot = a_parent x S + v_parent x Psi_dot
    crm_mul(vp, Pd, t6);
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      Pdd[k] += t6[k];
      v[k] = fma_t(S[k], qdi, vp[k]);
      a[k] = fma_t(Pd[k], qdi, fma_t(S[k], qddi, ap[k]));
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      V2 t;
      t.x = S[2 * k]; t.y = S[2 * k + 1]; VEC2(i, k) = t;
      t.x = Pd[2 * k]; t.y = Pd[2 * k + 1]; VEC2(i, 3 + k) = t;
      t.x = Pdd[2 * k]; t.y = Pdd[2 * k + 1]; VEC2(i, 6 + k) = t;
    }
    {
      V2 t;
      t.x = f1; t.y = f2; VEC2(i, 9) = t;
    }
    const int sa = m.slot_a[i], sb = m.slot_b[i];
    if (sa >= 0) {
#pragma unroll
      for (int k = 0; k < 9; ++k) STA(sa, k) = E[k];
#pragma unroll
      for (int k = 0; k < 3; ++k) STA(sa, 9 + k) = p[k];
#pragma unroll
      for (int k = 0; k < 6; ++k) { STA(sa, 12 + k) = v[k]; STA(sa, 18 + k) = a[k]; }
    }
    if (sb >= 0) {
#pragma unroll
      for (int k = 0; k < 9; ++k) STB(sb, k) = E[k];
#pragma unroll
      for (int k = 0; k < 3; ++k) STB(sb, 9 + k) = p[k];
#pragma unroll
      for (int k = 0; k < 6; ++k) { STB(sb, 12 + k) = v[k]; STB(sb, 18 + k) = a[k]; }
    }
  }

  // ------------------------------------------------------------------ backward sweep
  // composites: 0 m | 1..3 h | 4..9 Ibar | 10..15 Sym | 16..18 n | 19..21 l | 22..27 f
  T acc[28];
  T* out = dc_du + b * (int64_t)2 * n * n;
  const int n2 = 2 * n;
  for (int s = 0; s < m.n_slot_a; ++s)
#pragma unroll
    for (int k = 0; k < 28; ++k) STA(s, k) = T(0);

#pragma unroll 1
  for (int i = n - 1; i >= 0; --i) {
    const bool chained = (i != n - 1) && (m.parent[i + 1] == i);   // running state/composites valid
    if (!chained && i != n - 1) {
      const int s = m.slot_b[i];
#pragma unroll
      for (int k = 0; k < 9; ++k) E[k] = STB(s, k);
#pragma unroll
      for (int k = 0; k < 3; ++k) p[k] = STB(s, 9 + k);
#pragma unroll
      for (int k = 0; k < 6; ++k) { v[k] = STB(s, 12 + k); a[k] = STB(s, 18 + k); }
    }
    // ---- own rigid-body terms in world coordinates
    T own[28];
    {
      const T mi = m.mass[i];
      T hr[3], hw[3];
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) {
        hr[cc] = E[cc] * m.h[i][0] + E[3 + cc] * m.h[i][1] + E[6 + cc] * m.h[i][2];
        hw[cc] = fma_t(mi, p[cc], hr[cc]);
      }
      // Ir = E^T Ib E (symmetric)
      T IbE[9];   // Ib * E
      {
        const T xx = m.Ib[i][0], xy = m.Ib[i][1], xz = m.Ib[i][2], yy = m.Ib[i][3], yz = m.Ib[i][4], zz = m.Ib[i][5];
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
      // Sym = (w x Ibar) + (w x Ibar)^T - (h u^T + u h^T) + 2 (u.h) 1
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
      own[0] = mi;
      own[1] = hw[0]; own[2] = hw[1]; own[3] = hw[2];
#pragma unroll
      for (int k = 0; k < 6; ++k) own[4 + k] = Iw[k];
      own[10] = T(2) * M[0] - T(2) * hw[0] * uv[0] + uh2;
      own[11] = M[1] + M[3] - (hw[0] * uv[1] + uv[0] * hw[1]);
      own[12] = M[2] + M[6] - (hw[0] * uv[2] + uv[0] * hw[2]);
      own[13] = T(2) * M[4] - T(2) * hw[1] * uv[1] + uh2;
      own[14] = M[5] + M[7] - (hw[1] * uv[2] + uv[1] * hw[2]);
      own[15] = T(2) * M[8] - T(2) * hw[2] * uv[2] + uh2;
#pragma unroll
      for (int k = 0; k < 6; ++k) { own[16 + k] = mom[k]; own[22 + k] = fo[k] + t6[k]; }
    }
    if (chained) {
#pragma unroll
      for (int k = 0; k < 28; ++k) acc[k] += own[k];
    } else {
#pragma unroll
      for (int k = 0; k < 28; ++k) acc[k] = own[k];
    }
    const int sa = m.slot_a[i];
    if (sa >= 0) {
#pragma unroll
      for (int k = 0; k < 28; ++k) acc[k] += STA(sa, k);
    }
    // ---- per-body vectors and F vectors
    T S[6], Pd[6], Pdd[6];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      V2 t = VEC2(i, k); S[2 * k] = t.x; S[2 * k + 1] = t.y;
      t = VEC2(i, 3 + k); Pd[2 * k] = t.x; Pd[2 * k + 1] = t.y;
      t = VEC2(i, 6 + k); Pdd[2 * k] = t.x; Pdd[2 * k + 1] = t.y;
    }
    const T mC = acc[0];
    const T* hC = acc + 1;
    const T* IC = acc + 4;
    const T* SyC = acc + 10;
    const T* nC = acc + 16;
    const T* lC = acc + 19;
    const T* fC = acc + 22;
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
    if (c_out && active) c_out[b * n + i] = dot6(S, fC);
    {
      T dqq = dot6(S, F1);
      T ddd = dot6(S, F2);
      if (use_damping) ddd += m.damping[i];                                   // RBDReference.py:1341
      if (active) { out[i * n2 + i] = dqq; out[i * n2 + n + i] = ddd; }
    }
    if (m.kind[i] == 1) {
      // reference quirk for prismatic joints: X^T(-crm(f)S) instead of X^T(S x* f)
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
    // ---- ancestors of i
    for (int j = m.parent[i]; j >= 0; j = m.parent[j]) {
      T Sj[6], Pdj[6], Pddj[6];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        V2 t = VEC2(j, k); Sj[2 * k] = t.x; Sj[2 * k + 1] = t.y;
        t = VEC2(j, 3 + k); Pdj[2 * k] = t.x; Pdj[2 * k + 1] = t.y;
        t = VEC2(j, 6 + k); Pddj[2 * k] = t.x; Pddj[2 * k + 1] = t.y;
      }
      const T dq_ji = dot6s(Sj, F1);
      const T dd_ji = dot6s(Sj, F2);
      const T dq_ij = fma_t(T(2), dot3s(F3, Pdj), dot6s(F4, Pddj));
      const T dd_ij = T(2) * (dot6s(F4, Pdj) + dot3s(F3, Sj));
      if (active) {
        out[j * n2 + i] = dq_ji;
        out[j * n2 + n + i] = dd_ji;
        out[i * n2 + j] = dq_ij;
        out[i * n2 + n + j] = dd_ij;
      }
    }
    // ---- structural zeros of row i (bodies on other branches)
    {
      const unsigned touched = m.anc_mask[i] | m.sub_mask[i];
      if (active)
        for (int j = 0; j < n; ++j)
          if (!((touched >> j) & 1u)) { out[i * n2 + j] = T(0); out[i * n2 + n + j] = T(0); }
    }
    // ---- hand the composites to the parent / walk the state back up
    const int par = m.parent[i];
    if (par >= 0 && par != i - 1) {
      const int s = m.slot_a[par];
#pragma unroll
      for (int k = 0; k < 28; ++k) STA(s, k) += acc[k];
    }
    if (par >= 0 && par == i - 1) {
      const V2 bas = VEC2(i, 9), dd = VEC2(i, 10);
      const T f1 = bas.x, f2 = bas.y, qdi = dd.x, qddi = dd.y;
      T Ej[9], r[3], Ep[9];
#pragma unroll
      for (int k = 0; k < 9; ++k) Ej[k] = fma_t(m.EC[i][k], f2, fma_t(m.EB[i][k], f1, m.EA[i][k]));
#pragma unroll
      for (int k = 0; k < 3; ++k) r[k] = fma_t(m.rC[i][k], f2, fma_t(m.rB[i][k], f1, m.rA[i][k]));
      // Ep = Ej^T E ; pp = p - Ep^T r
#pragma unroll
      for (int rr = 0; rr < 3; ++rr)
#pragma unroll
        for (int cc = 0; cc < 3; ++cc)
          Ep[3 * rr + cc] = Ej[rr] * E[cc] + Ej[3 + rr] * E[3 + cc] + Ej[6 + rr] * E[6 + cc];
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) p[cc] -= Ep[cc] * r[0] + Ep[3 + cc] * r[1] + Ep[6 + cc] * r[2];
#pragma unroll
      for (int k = 0; k < 9; ++k) E[k] = Ep[k];
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        v[k] = fma_t(-S[k], qdi, v[k]);
        a[k] = fma_t(-Pd[k], qdi, fma_t(-S[k], qddi, a[k]));
      }
    }
  }
#undef VEC
#undef VEC2
#undef STA
#undef STB
}

}  // namespace rbd
