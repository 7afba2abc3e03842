// rbd_minv_kernels.cuh - fast fused minv: analytical inverse of the joint-space inertia matrix
// (Carpentier; RBDReference.py:785-806 = minv_bpass :630-735 + minv_fpass :737-783 + mirror)
// evaluated in WORLD coordinates.
//
// In world coordinates every parent<-child transfer of the reference's recursion is a plain sum:
//   backward (leaf -> root):  U_i = IA_i S_i, D_i = S_i.U_i,
//                             Minv[i,j] = delta_ij / D_i - (S_i . F_j) / D_i     for j in subtree(i)
//                             F_j += U_i Minv[i,j]                               (one 6-vector per column)
//                             IA_parent += IA_i - U_i U_i^T / D_i               (no X^T . X congruence)
//   forward  (root -> leaf):  Minv[i,j] -= (U_i . G_j) / D_i,   G_j += S_i Minv[i,j]
// where IA is the articulated inertia (symmetric 6x6, 21 numbers) seeded with each body's
// rigid inertia moved to the world frame, and G_j is the running sum of S_k Minv[k,j] over the
// current root path (a depth-first walk pushes on the way down and pops when it leaves a branch).
//
// Bodies are handled in depth-first preorder (`orig` maps back to the caller's numbering), so a
// subtree and a root component are contiguous index ranges.  One knot point per thread, one warp
// per CTA; per-body (S, U, 1/D, f1, f2), per-column (F / G) and the upper triangle of Minv live in
// shared memory (LOCAL = 0) or in per-thread local memory for large trees (LOCAL = 1).
// Only the dense (mirrored) result is produced here; output_dense = 0 runs the generic kernel.
#pragma once
#include "rbd_common.cuh"
#include "rbd_grad_kernels.cuh"

namespace rbd {

// topology tables of the depth-first renumbering (all indices are DFS positions)
struct DfsPlan {
  int orig[RBD_MAX_DOF];       // DFS position -> caller's body id
  int pos[RBD_MAX_DOF];        // caller's body id -> DFS position
  int sub_end[RBD_MAX_DOF];    // subtree(i) = [i, sub_end[i])
  int comp_end[RBD_MAX_DOF];   // root component of i = [root, comp_end[i])
};

constexpr int kMinvPerBody = 16;   // S(6) U(6) invD f1 f2 pad
constexpr int kMinvSlotA = 22;     // IA (21) rounded to a pair boundary; forward use: E(9) p(3)
constexpr int kMinvSlotB = 12;     // E(9) p(3)

template <typename T, int LOCAL, int MINB>
__global__ void __launch_bounds__(32, MINB)
minv_world_kernel(const __grid_constant__ FastModel<T> m, const __grid_constant__ DfsPlan plan, int64_t B,
                  const T* __restrict__ q, T* __restrict__ Minv) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* sm = reinterpret_cast<T*>(smem_raw);
  const int lane = threadIdx.x;
  const int n = m.n;
  int64_t b = (int64_t)blockIdx.x * 32 + lane;
  const bool active = b < B;
  if (!active) b = B - 1;
  const int ntri = n * (n + 1) / 2;
  constexpr int kLocalVals = LOCAL ? RBD_MAX_DOF * (kMinvPerBody + 6) + RBD_MAX_DOF * (RBD_MAX_DOF + 1) / 2 : 1;
  T lmem[kLocalVals];
  // region offsets (in values per lane): bodies | columns | triangle | stash A | stash B
  const int off_col = n * kMinvPerBody;
  const int off_tri = off_col + 6 * n;
  const int off_sta = LOCAL ? 0 : off_tri + ntri;
  const int off_stb = off_sta + m.n_slot_a * kMinvSlotA;
#define LV(idx) (*(LOCAL ? &lmem[(idx)] : &sm[(idx) * 32 + lane]))
#define BODY(i, k) LV((i) * kMinvPerBody + (k))
#define COL(j, k) LV(off_col + (j) * 6 + (k))
#define TRI(i, j) LV(off_tri + (i) * n - (((i) * ((i) - 1)) >> 1) + ((j) - (i)))
#define STA(s, k) sm[(off_sta + (s) * kMinvSlotA + (k)) * 32 + lane]
#define STB(s, k) sm[(off_stb + (s) * kMinvSlotB + (k)) * 32 + lane]

  // ---- stage q (coalesced for the shared-memory variant)
  if (LOCAL) {
    for (int i = 0; i < n; ++i) BODY(i, 13) = q[b * n + plan.orig[i]];
  } else {
    const int64_t base = (int64_t)blockIdx.x * 32 * n;
    const int64_t limit = B * (int64_t)n;
    int inst = lane / n, jnt = lane - inst * n;
    const int dinst = 32 / n, djnt = 32 - dinst * n;
    for (int e = lane; e < 32 * n; e += 32) {
      const int64_t g = base + e;
      sm[(plan.pos[jnt] * kMinvPerBody + 13) * 32 + inst] = (g < limit) ? q[g] : T(0);
      inst += dinst; jnt += djnt;
      if (jnt >= n) { jnt -= n; inst += 1; }
    }
    __syncwarp();
  }

  T E[9], p[3];
  // ------------------------------------------------------------------ forward: poses and axes
#pragma unroll 1
  for (int i = 0; i < n; ++i) {
    T f1, f2;
    {
      const T qi = BODY(i, 13);
      if (m.kind[i] == 0) sincos_t(qi, &f2, &f1);
      else { f1 = qi; f2 = T(0); }
    }
    const int par = m.parent[i];
    T Ep[9], pp[3];
    if (par < 0) {
#pragma unroll
      for (int k = 0; k < 9; ++k) Ep[k] = (k % 4 == 0) ? T(1) : T(0);
#pragma unroll
      for (int k = 0; k < 3; ++k) pp[k] = T(0);
    } else if (par != i - 1) {
      const int s = m.slot_a[par];
#pragma unroll
      for (int k = 0; k < 9; ++k) Ep[k] = STA(s, k);
#pragma unroll
      for (int k = 0; k < 3; ++k) pp[k] = STA(s, 9 + k);
    } else {
#pragma unroll
      for (int k = 0; k < 9; ++k) Ep[k] = E[k];
#pragma unroll
      for (int k = 0; k < 3; ++k) pp[k] = p[k];
    }
    T Ej[9], r[3];
#pragma unroll
    for (int k = 0; k < 9; ++k) Ej[k] = fma_t(m.EC[i][k], f2, fma_t(m.EB[i][k], f1, m.EA[i][k]));
#pragma unroll
    for (int k = 0; k < 3; ++k) r[k] = fma_t(m.rC[i][k], f2, fma_t(m.rB[i][k], f1, m.rA[i][k]));
#pragma unroll
    for (int rr = 0; rr < 3; ++rr)
#pragma unroll
      for (int cc = 0; cc < 3; ++cc)
        E[3 * rr + cc] = Ej[3 * rr] * Ep[cc] + Ej[3 * rr + 1] * Ep[3 + cc] + Ej[3 * rr + 2] * Ep[6 + cc];
#pragma unroll
    for (int cc = 0; cc < 3; ++cc) p[cc] = pp[cc] + Ep[cc] * r[0] + Ep[3 + cc] * r[1] + Ep[6 + cc] * r[2];
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
#pragma unroll
    for (int k = 0; k < 6; ++k) BODY(i, k) = S[k];
    BODY(i, 13) = f1; BODY(i, 14) = f2;
    const int sa = m.slot_a[i], sb = m.slot_b[i];
    if (sa >= 0) {
#pragma unroll
      for (int k = 0; k < 9; ++k) STA(sa, k) = E[k];
#pragma unroll
      for (int k = 0; k < 3; ++k) STA(sa, 9 + k) = p[k];
    }
    if (sb >= 0) {
#pragma unroll
      for (int k = 0; k < 9; ++k) STB(sb, k) = E[k];
#pragma unroll
      for (int k = 0; k < 3; ++k) STB(sb, 9 + k) = p[k];
    }
  }

  // ------------------------------------------------------------------ backward: IA, U, D, rows of Minv
  // IA = [[A, Bm], [Bm^T, C]] : A sym (0..5: xx xy xz yy yz zz), Bm 3x3 row-major (6..14), C sym (15..20)
  T IA[21];
  for (int j = 0; j < n; ++j)
#pragma unroll
    for (int k = 0; k < 6; ++k) COL(j, k) = T(0);
  for (int s = 0; s < m.n_slot_a; ++s)
#pragma unroll
    for (int k = 0; k < 21; ++k) STA(s, k) = T(0);

#pragma unroll 1
  for (int i = n - 1; i >= 0; --i) {
    const bool chained = (i != n - 1) && (m.parent[i + 1] == i);
    if (!chained && i != n - 1) {
      const int s = m.slot_b[i];
#pragma unroll
      for (int k = 0; k < 9; ++k) E[k] = STB(s, k);
#pragma unroll
      for (int k = 0; k < 3; ++k) p[k] = STB(s, 9 + k);
    }
    // own rigid inertia about the world origin
    {
      const T mi = m.mass[i];
      T hr[3], hw[3];
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) {
        hr[cc] = E[cc] * m.h[i][0] + E[3 + cc] * m.h[i][1] + E[6 + cc] * m.h[i][2];
        hw[cc] = fma_t(mi, p[cc], hr[cc]);
      }
      T IbE[9];
      {
        const T xx = m.Ib[i][0], xy = m.Ib[i][1], xz = m.Ib[i][2], yy = m.Ib[i][3], yz = m.Ib[i][4], zz = m.Ib[i][5];
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
          IbE[cc] = xx * E[cc] + xy * E[3 + cc] + xz * E[6 + cc];
          IbE[3 + cc] = xy * E[cc] + yy * E[3 + cc] + yz * E[6 + cc];
          IbE[6 + cc] = xz * E[cc] + yz * E[3 + cc] + zz * E[6 + cc];
        }
      }
      const T tr = (hr[0] + hw[0]) * p[0] + (hr[1] + hw[1]) * p[1] + (hr[2] + hw[2]) * p[2];
      T own[21];
      int idx = 0;
#pragma unroll
      for (int rr = 0; rr < 3; ++rr)
#pragma unroll
        for (int cc = rr; cc < 3; ++cc) {
          T val = E[rr] * IbE[cc] + E[3 + rr] * IbE[3 + cc] + E[6 + rr] * IbE[6 + cc];
          val -= hr[rr] * p[cc] + p[rr] * hw[cc];
          if (rr == cc) val += tr;
          own[idx++] = val;
        }
      // Bm = h x : [0 -hz hy; hz 0 -hx; -hy hx 0]
      own[6] = T(0); own[7] = -hw[2]; own[8] = hw[1];
      own[9] = hw[2]; own[10] = T(0); own[11] = -hw[0];
      own[12] = -hw[1]; own[13] = hw[0]; own[14] = T(0);
      own[15] = mi; own[16] = T(0); own[17] = T(0); own[18] = mi; own[19] = T(0); own[20] = mi;
      if (chained) {
#pragma unroll
        for (int k = 0; k < 21; ++k) IA[k] += own[k];
      } else {
#pragma unroll
        for (int k = 0; k < 21; ++k) IA[k] = own[k];
      }
    }
    const int sa = m.slot_a[i];
    if (sa >= 0) {
#pragma unroll
      for (int k = 0; k < 21; ++k) IA[k] += STA(sa, k);
    }
    T S[6], U[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) S[k] = BODY(i, k);
    // U = IA S
    sym3_mul(IA, S, U);
    sym3_mul(IA + 15, S + 3, U + 3);
#pragma unroll
    for (int rr = 0; rr < 3; ++rr) {
      U[rr] += IA[6 + 3 * rr] * S[3] + IA[6 + 3 * rr + 1] * S[4] + IA[6 + 3 * rr + 2] * S[5];
      U[3 + rr] += IA[6 + rr] * S[0] + IA[9 + rr] * S[1] + IA[12 + rr] * S[2];
    }
    const T invD = T(1) / dot6s(S, U);                                        // RBDReference.py:698-700
#pragma unroll
    for (int k = 0; k < 6; ++k) BODY(i, 6 + k) = U[k];
    BODY(i, 12) = invD;
    const int par = m.parent[i];
    const int send = plan.sub_end[i];
    for (int j = i; j < send; ++j) {
      T F[6];
#pragma unroll
      for (int k = 0; k < 6; ++k) F[k] = COL(j, k);
      const T mij = (j == i ? invD : T(0)) - invD * dot6s(S, F);              // :700-708
      TRI(i, j) = mij;
      if (par >= 0) {
#pragma unroll
        for (int k = 0; k < 6; ++k) COL(j, k) = fma_t(U[k], mij, F[k]);       // :721-726 (world frame: no X^T)
      }
    }
    if (par >= 0) {
      // IA -= U U^T / D                                                     // :728-731
      T Us[6];
#pragma unroll
      for (int k = 0; k < 6; ++k) Us[k] = U[k] * invD;
      IA[0] -= U[0] * Us[0]; IA[1] -= U[0] * Us[1]; IA[2] -= U[0] * Us[2];
      IA[3] -= U[1] * Us[1]; IA[4] -= U[1] * Us[2]; IA[5] -= U[2] * Us[2];
#pragma unroll
      for (int rr = 0; rr < 3; ++rr)
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) IA[6 + 3 * rr + cc] -= U[rr] * Us[3 + cc];
      IA[15] -= U[3] * Us[3]; IA[16] -= U[3] * Us[4]; IA[17] -= U[3] * Us[5];
      IA[18] -= U[4] * Us[4]; IA[19] -= U[4] * Us[5]; IA[20] -= U[5] * Us[5];
      if (par != i - 1) {
        const int s = m.slot_a[par];
#pragma unroll
        for (int k = 0; k < 21; ++k) STA(s, k) += IA[k];                      // :732-733 (world frame)
      } else {
        // reverse walk of the pose to the parent
        const T f1 = BODY(i, 13), f2 = BODY(i, 14);
        T Ej[9], r[3], Ep[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) Ej[k] = fma_t(m.EC[i][k], f2, fma_t(m.EB[i][k], f1, m.EA[i][k]));
#pragma unroll
        for (int k = 0; k < 3; ++k) r[k] = fma_t(m.rC[i][k], f2, fma_t(m.rB[i][k], f1, m.rA[i][k]));
#pragma unroll
        for (int rr = 0; rr < 3; ++rr)
#pragma unroll
          for (int cc = 0; cc < 3; ++cc)
            Ep[3 * rr + cc] = Ej[rr] * E[cc] + Ej[3 + rr] * E[3 + cc] + Ej[6 + rr] * E[6 + cc];
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) p[cc] -= Ep[cc] * r[0] + Ep[3 + cc] * r[1] + Ep[6 + cc] * r[2];
#pragma unroll
        for (int k = 0; k < 9; ++k) E[k] = Ep[k];
      }
    }
  }

  // ------------------------------------------------------------------ forward: finish the rows
  for (int j = 0; j < n; ++j)
#pragma unroll
    for (int k = 0; k < 6; ++k) COL(j, k) = T(0);
  T* out = Minv + b * (int64_t)n * n;
  int top = -1;                                  // deepest body whose S Minv row is in G
#pragma unroll 1
  for (int i = 0; i < n; ++i) {
    const int par = m.parent[i];
    const int cend = plan.comp_end[i];
    // leave finished branches: G -= S_top Minv[top, i..]   (a new root starts from G = 0)
    if (par < 0) top = -1;
    while (top != par) {
      T St[6];
#pragma unroll
      for (int k = 0; k < 6; ++k) St[k] = BODY(top, k);
      for (int j = i; j < cend; ++j) {
        const T mtj = TRI(top, j);
#pragma unroll
        for (int k = 0; k < 6; ++k) COL(j, k) = fma_t(-St[k], mtj, COL(j, k));
      }
      top = m.parent[top];
    }
    T S[6], U[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) { S[k] = BODY(i, k); U[k] = BODY(i, 6 + k); }
    const T invD = BODY(i, 12);
    const int send = plan.sub_end[i];
    const int oi = plan.orig[i];
    for (int j = i; j < cend; ++j) {
      T G[6];
#pragma unroll
      for (int k = 0; k < 6; ++k) G[k] = COL(j, k);
      T mij = (j < send) ? TRI(i, j) : T(0);
      if (par >= 0) mij -= invD * dot6s(U, G);                                // :771-773 (world frame)
      TRI(i, j) = mij;
#pragma unroll
      for (int k = 0; k < 6; ++k) COL(j, k) = fma_t(S[k], mij, G[k]);         // :774-776 / :781
      if (active) {
        const int oj = plan.orig[j];
        out[oi * n + oj] = mij;
        out[oj * n + oi] = mij;                                               // :799-804
      }
    }
    // other root components do not couple with this body
    if (active) {
      for (int j = 0; j < n; ++j) {
        if (plan.comp_end[j] != cend) out[oi * n + plan.orig[j]] = T(0);
      }
    }
    top = i;
  }
#undef LV
#undef BODY
#undef COL
#undef TRI
#undef STA
#undef STB
}

}  // namespace rbd
