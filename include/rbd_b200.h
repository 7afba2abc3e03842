/* rbd_b200.h - C ABI of the B200-native batched rigid-body-dynamics library (librbd_b200.so).
 *
 * Drop-in boundary for the hot path of A2R-Lab/RBDReference: rnea, rnea_grad, minv and their
 * eight per-pass helpers, evaluated over B knot points per call on one CUDA device.
 * The reference has no FFI of its own (one pure-Python class); each entry point below names
 * the reference method (RBDReference.py:line) it replaces.  INTEGRATION.md shows the ctypes
 * binding a reference maintainer would add.
 *
 * Conventions
 *  - All tensors are dense, contiguous, row-major DEVICE pointers with the batch axis first;
 *    out[k] equals the reference's result for (q[k], qd[k], qdd[k]).
 *      q, qd, qdd, c        (B, n)
 *      v, a, f              (B, 6, NB)          spatial vectors are [angular; linear]
 *      dv, da, df           (B, 6, n, NB)
 *      dc_dq, dc_dqd        (B, n, n)           dc_du (B, n, 2n) = [dc_dq | dc_dqd]
 *      Minv                 (B, n, n)           F (B, n, 6, n)   U (B, n, 6)   Dinv (B, n)
 *    n = NB = number of 1-DoF joints (fixed base), 1 <= n <= RBD_MAX_DOF.
 *  - `_f64` entry points take double*, `_f32` take float*; both compute in that type.
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Calls only
 *    enqueue work; they never synchronise.  A model handle is immutable after creation, so any
 *    number of host threads may call with the same handle on different streams.
 *  - Every function returns 0 on success, a positive cudaError_t value for CUDA failures, or a
 *    negative RBD_E_* code for argument errors; rbd_last_error_string() describes the last
 *    failure on the calling thread.  No exceptions cross the ABI, nothing is printed.
 *  - There is no CPU fallback: without a CUDA device the compute entry points fail.
 */
#ifndef RBD_B200_H_
#define RBD_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RBD_MAX_DOF 32
#define RBD_ABI_VERSION 1

#define RBD_E_INVALID_ARGUMENT (-1)
#define RBD_E_UNSUPPORTED      (-2)
#define RBD_E_NO_DEVICE        (-3)

/* Host-side description of a compiled robot (rbdreference_b200/model.py builds it by probing
 * the robot object handed to RBDReference.__init__, RBDReference.py:6-7).  All pointers are
 * HOST pointers, copied during rbd_model_create. */
typedef struct RbdModelDesc {
  int32_t n;                /* bodies == DoF                                                  */
  const int32_t* parent;    /* [n]   get_parent_id, -1 = fixed base; parent[i] < i            */
  const int32_t* kind;      /* [n]   0: X(q)=A+B cos q+C sin q (revolute), 1: X(q)=A+B q      */
  const double* S;          /* [n*6] get_S_by_id                                              */
  const double* XA;         /* [n*18] E (3x3 row-major) then L (3x3), X = [[E,0],[L,E]]       */
  const double* XB;         /* [n*18]                                                         */
  const double* XC;         /* [n*18]                                                         */
  const double* I;          /* [n*36] get_Imat_by_id, row-major 6x6                           */
  const double* damping;    /* [n]   get_damping_by_id                                        */
} RbdModelDesc;

typedef struct rbd_model rbd_model_t;

int rbd_abi_version(void);
const char* rbd_last_error_string(void);

/* RBDReference.__init__ (RBDReference.py:6-7): compile-once model handle (host memory only;
 * the model travels to the device as a kernel parameter, so one handle serves every device). */
int rbd_model_create(const RbdModelDesc* desc, rbd_model_t** out);
int rbd_model_destroy(rbd_model_t* m);
int rbd_model_num_dof(const rbd_model_t* m);
/* 1 if the fused drivers run the world-frame kernels for this model (every spatial inertia has
 * rigid-body structure), 0 if they run the generic body-frame kernels. */
int rbd_model_uses_world_kernels(const rbd_model_t* m);
/* Process-wide kernel selection for the fused drivers: 0 = automatic (default), 1 = always the
 * generic body-frame kernels (the reference's own recursion), 2 = world-frame kernels with one
 * knot point per thread, 3 = warp-cooperative kernels (one body per lane), 4 = hybrid minv kernel
 * (knot point per lane for the articulated inertias, column per lane for the rows of Minv; other
 * operations behave as 0), 5 = lane minv kernel (knot point per lane in every phase, per-body table
 * and output tile in shared memory; robots too large for it run the generic kernel; other
 * operations behave as 0), 7 = chain rnea_grad kernel (serial chains, knot point per lane; other robots and
 * operations behave as 0), 8 / 9 = tile minv kernel (one CTA per 32 knot points, branch-parallel, table in shared
 * memory) with 4-column groups and 8 warps / 2-column groups and 16 warps per CTA (9 is the automatic choice for
 * n > 16; other operations behave as 0).  6 is retired.  Used by the tests and the benchmark to cross-check / compare. */
int rbd_set_kernel_variant(int variant);
/* The same choice for one handle only (-1 = follow the process-wide setting, the default): calls on
 * different handles never see each other's choice, so handles stay independent across threads. */
int rbd_model_set_kernel_variant(rbd_model_t* m, int variant);

/* ---- fused drivers ------------------------------------------------------------------------ */
/* rnea (RBDReference.py:623-628).  qdd may be NULL (skips the S*qdd term, :589).  v, a, f may
 * each be NULL when the caller only wants c; f is the ACCUMULATED force (:619,:628). */
int rbd_rnea_f64(const rbd_model_t* m, int64_t B, const double* q, const double* qd, const double* qdd,
                 double gravity, double* c, double* v, double* a, double* f, void* stream);
int rbd_rnea_f32(const rbd_model_t* m, int64_t B, const float* q, const float* qd, const float* qdd,
                 float gravity, float* c, float* v, float* a, float* f, void* stream);

/* rnea_grad (RBDReference.py:1345-1368): dc_du (B,n,2n).  c_out (B,n) optional. */
int rbd_rnea_grad_f64(const rbd_model_t* m, int64_t B, const double* q, const double* qd, const double* qdd,
                      double gravity, int use_velocity_damping, double* dc_du, double* c_out, void* stream);
int rbd_rnea_grad_f32(const rbd_model_t* m, int64_t B, const float* q, const float* qd, const float* qdd,
                      float gravity, int use_velocity_damping, float* dc_du, float* c_out, void* stream);

/* minv (RBDReference.py:785-806).  output_dense=0 keeps the rows as the forward pass leaves them
 * (the reference updates whole rows at :771, so the matrix is still full). */
int rbd_minv_f64(const rbd_model_t* m, int64_t B, const double* q, int output_dense, double* Minv, void* stream);
int rbd_minv_f32(const rbd_model_t* m, int64_t B, const float* q, int output_dense, float* Minv, void* stream);

/* ---- per-pass helpers (same in-place contracts as the reference) ----------------------------- */
/* rnea_fpass (RBDReference.py:559-598) */
int rbd_rnea_fpass_f64(const rbd_model_t* m, int64_t B, const double* q, const double* qd, const double* qdd,
                       double gravity, double* v, double* a, double* f, void* stream);
int rbd_rnea_fpass_f32(const rbd_model_t* m, int64_t B, const float* q, const float* qd, const float* qdd,
                       float gravity, float* v, float* a, float* f, void* stream);
/* rnea_bpass (RBDReference.py:600-621): f is accumulated IN PLACE. */
int rbd_rnea_bpass_f64(const rbd_model_t* m, int64_t B, const double* q, double* f, double* c, void* stream);
int rbd_rnea_bpass_f32(const rbd_model_t* m, int64_t B, const float* q, float* f, float* c, void* stream);
/* rnea_grad_fpass_dq (RBDReference.py:1127-1187) */
int rbd_rnea_grad_fpass_dq_f64(const rbd_model_t* m, int64_t B, const double* q, const double* qd, const double* v,
                               const double* a, double gravity, double* dv, double* da, double* df, void* stream);
int rbd_rnea_grad_fpass_dq_f32(const rbd_model_t* m, int64_t B, const float* q, const float* qd, const float* v,
                               const float* a, float gravity, float* dv, float* da, float* df, void* stream);
/* rnea_grad_fpass_dqd (RBDReference.py:1189-1255) */
int rbd_rnea_grad_fpass_dqd_f64(const rbd_model_t* m, int64_t B, const double* q, const double* qd, const double* v,
                                double* dv, double* da, double* df, void* stream);
int rbd_rnea_grad_fpass_dqd_f32(const rbd_model_t* m, int64_t B, const float* q, const float* qd, const float* v,
                                float* dv, float* da, float* df, void* stream);
/* rnea_grad_bpass_dq (RBDReference.py:1257-1297): df_dq is accumulated IN PLACE. */
int rbd_rnea_grad_bpass_dq_f64(const rbd_model_t* m, int64_t B, const double* q, const double* f, double* df_dq,
                               double* dc_dq, void* stream);
int rbd_rnea_grad_bpass_dq_f32(const rbd_model_t* m, int64_t B, const float* q, const float* f, float* df_dq,
                               float* dc_dq, void* stream);
/* rnea_grad_bpass_dqd (RBDReference.py:1299-1343): df_dqd is accumulated IN PLACE. */
int rbd_rnea_grad_bpass_dqd_f64(const rbd_model_t* m, int64_t B, const double* q, double* df_dqd,
                                int use_velocity_damping, double* dc_dqd, void* stream);
int rbd_rnea_grad_bpass_dqd_f32(const rbd_model_t* m, int64_t B, const float* q, float* df_dqd,
                                int use_velocity_damping, float* dc_dqd, void* stream);
/* minv_bpass (RBDReference.py:630-735): Dinv receives D = S^T U (not its reciprocal, :698). */
int rbd_minv_bpass_f64(const rbd_model_t* m, int64_t B, const double* q, double* Minv, double* F, double* U,
                       double* Dinv, void* stream);
int rbd_minv_bpass_f32(const rbd_model_t* m, int64_t B, const float* q, float* Minv, float* F, float* U,
                       float* Dinv, void* stream);
/* minv_fpass (RBDReference.py:737-783): Minv and F are updated IN PLACE. */
int rbd_minv_fpass_f64(const rbd_model_t* m, int64_t B, const double* q, double* Minv, double* F, const double* U,
                       const double* Dinv, void* stream);
int rbd_minv_fpass_f32(const rbd_model_t* m, int64_t B, const float* q, float* Minv, float* F, const float* U,
                       const float* Dinv, void* stream);

/* ---- forward dynamics ---------------------------------------------------- */
/* forward_dynamics (RBDReference.py:1369-1372): qdd = Minv (u - c), c = rnea(q, qd) with the S*qdd
 * term skipped and GRAVITY = -9.81 (the reference's defaults).  u, qdd: (B, n).  Minv_out (B, n, n)
 * may be NULL; when given it receives minv(q).  Three launches (rnea, minv, product); temporaries
 * are stream-ordered allocations from a private memory pool. */
int rbd_forward_dynamics_f64(const rbd_model_t* m, int64_t B, const double* q, const double* qd, const double* u,
                             double* qdd, double* Minv_out, void* stream);
int rbd_forward_dynamics_f32(const rbd_model_t* m, int64_t B, const float* q, const float* qd, const float* u,
                             float* qdd, float* Minv_out, void* stream);
/* forward_dynamics_grad (RBDReference.py:1374-1384): qdd_dq = -Minv dc_dq, qdd_dqd = -Minv dc_dqd with
 * dc_du = rnea_grad(q, qd, qdd), qdd = forward_dynamics(q, qd, u).  qdd_dq, qdd_dqd: (B, n, n);
 * qdd_out (B, n) may be NULL.  minv is evaluated once (the reference evaluates it twice, :1371/:1381). */
int rbd_forward_dynamics_grad_f64(const rbd_model_t* m, int64_t B, const double* q, const double* qd,
                                  const double* u, double* qdd_dq, double* qdd_dqd, double* qdd_out, void* stream);
int rbd_forward_dynamics_grad_f32(const rbd_model_t* m, int64_t B, const float* q, const float* qd, const float* u,
                                  float* qdd_dq, float* qdd_dqd, float* qdd_out, void* stream);

/* ---- joint-space inertia matrix ---------------------------------------- */
/* crba, fixed-base branch (RBDReference.py:1026-1124, the `else` at :1090): H (B, n, n), symmetric,
 * H[i,j] = 0 for bodies on different branches.  One launch. */
int rbd_crba_f64(const rbd_model_t* m, int64_t B, const double* q, double* H, void* stream);
int rbd_crba_f32(const rbd_model_t* m, int64_t B, const float* q, float* H, void* stream);

/* ---- articulated-body algorithm ---------------------------------------- */
/* aba, fixed-base branch (RBDReference.py:817, :940-1024): qdd (B, n) from q, qd, tau (B, n).  Follows
 * the reference to the letter, INCLUDING :984, where every body's bias force pA is set to element 0
 * of crf(v) I v broadcast over its six entries; results therefore equal the reference's aba(), not
 * forward_dynamics().  f_ext is ignored by this branch upstream and has no argument here. */
int rbd_aba_f64(const rbd_model_t* m, int64_t B, const double* q, const double* qd, const double* tau,
                double gravity, double* qdd, void* stream);
int rbd_aba_f32(const rbd_model_t* m, int64_t B, const float* q, const float* qd, const float* tau,
                float gravity, float* qdd, void* stream);

/* ---- end-effector kinematics --------------------------------------------- */
/* end_effector_pose (RBDReference.py:220-283) and end_effector_pose_gradient (:295-386).
 * The reference takes ee_joint_names / ee_offsets per call and walks the robot's 4x4 homogeneous
 * joint transforms (get_Xmat_hom_Func_by_id, get_dXmat_hom_Func_by_id) from each end effector to
 * the base.  Here that selection is compiled once into a handle (rbdreference_b200/model.py probes
 * the same getters): T_i(q) = TA + TB f1 + TC f2 and dT_i/dq = DA + DB f1 + DC f2 with
 * (f1, f2) = (cos q, sin q) for kind 0 and (q, 0) for kind 1; only rows 0..2 are passed (row 3 is
 * (0,0,0,1) for T and zero for dT).  ee_joint[e] is the moving joint the chain of end effector e
 * starts from, ee_final[e] the rows 0..2 of the fixed-joint transform closing it (identity when a
 * moving joint was named, get_transformation_matrix_hom() of the fixed joint otherwise, :277-280);
 * offset is ee_offsets[0] = (x, y, z, w), the only offset the reference uses (:248, :335). */
#define RBD_MAX_EE 32
typedef struct RbdEeDesc {
  int32_t n;                 /* joints                                                        */
  const int32_t* parent;     /* [n]                                                           */
  const int32_t* kind;       /* [n]                                                           */
  const double* TA;          /* [n*12] row-major 3x4                                          */
  const double* TB;
  const double* TC;
  const double* DA;          /* [n*12]                                                        */
  const double* DB;
  const double* DC;
  int32_t n_ee;              /* 1..RBD_MAX_EE, in the reference's output order                */
  const int32_t* ee_joint;   /* [n_ee]                                                        */
  const double* ee_final;    /* [n_ee*12]                                                     */
  double offset[4];
} RbdEeDesc;
typedef struct rbd_ee_model rbd_ee_model_t;
int rbd_ee_model_create(const RbdEeDesc* desc, rbd_ee_model_t** out);
int rbd_ee_model_destroy(rbd_ee_model_t* m);
int rbd_ee_model_num_ee(const rbd_ee_model_t* m);
/* pose (B, n_ee, 6) = [x y z roll pitch yaw] per end effector (the reference returns a list of
 * (6,1) matrices).  One launch. */
int rbd_end_effector_pose_f64(const rbd_ee_model_t* m, int64_t B, const double* q, double* pose, void* stream);
int rbd_end_effector_pose_f32(const rbd_ee_model_t* m, int64_t B, const float* q, float* pose, void* stream);
/* dpose (B, n_ee, 6, n): column j = d pose / d q_j, zero for joints off the chain (:359-361).
 * pose (B, n_ee, 6) may be NULL; when given it receives end_effector_pose of the same launch. */
int rbd_end_effector_pose_gradient_f64(const rbd_ee_model_t* m, int64_t B, const double* q, double* dpose,
                                       double* pose, void* stream);
int rbd_end_effector_pose_gradient_f32(const rbd_ee_model_t* m, int64_t B, const float* q, float* dpose,
                                       float* pose, void* stream);

/* ---- floating base ----------------------------------------------------- */
/* The `self.robot.floating_base` branches of rnea (RBDReference.py:585, :591), minv (:652-691,
 * :761-779) and rnea_grad (:1141-1168, :1212-1238, :1267-1282, :1309-1341).  bodies.n = NB counts
 * the base as body 0 (parent -1, S = eye(6) upstream; only I and damping of entry 0 are read);
 * body i >= 1 is a 1-DoF joint with parent[i] in 0..i-1.  The base transform is
 * X0 = xrot(E) xlt(p) with p = q[pos_off..+3] and E = R(quat)^T (transpose = 0) or R(quat)
 * (transpose = 1), quat = q[quat_off..+4] a unit quaternion stored (x,y,z,w) (w_first = 0) or
 * (w,x,y,z) (w_first = 1); (pos_off, quat_off) is (0, 3) or (4, 0).  rbdreference_b200/model.py
 * finds the layout by probing the robot's own get_Xmat_Func_by_id(0).
 * Shapes: q (B, NB+6) with joint i at q[i+6]; qd, qdd, c (B, NB+5) with the base twist
 * [angular; linear] in base coordinates at [0:6] and joint i at [i+5]; v, a, f (B, 6, NB);
 * dc_du (B, NB+5, 2(NB+5)); Minv (B, NB+5, NB+5).  Reference behaviour kept: velocity damping lands on
 * [i, i] (body index) and on the block [0:5, 0:5] for the base (:1336-1341).  The reference fills BOTH triangles of a
 * floating-base Minv (:761-781 work on whole rows) and output_dense only copies the upper triangle of the leading
 * NB x NB block over the lower one (:799-804): either setting is the full symmetric matrix up to rounding.  The
 * automatic (cooperative) kernel returns that matrix for both; kernel family 1 reproduces the two code paths. */
typedef struct RbdFbModelDesc {
  RbdModelDesc bodies;
  int32_t pos_off, quat_off, w_first, transpose;
} RbdFbModelDesc;
typedef struct rbd_fb_model rbd_fb_model_t;
int rbd_fb_model_create(const RbdFbModelDesc* desc, rbd_fb_model_t** out);
int rbd_fb_model_destroy(rbd_fb_model_t* m);
int rbd_fb_model_num_vel(const rbd_fb_model_t* m);
/* Kernel family of the fused floating-base drivers for THIS handle: -1 follow rbd_set_kernel_variant, 0 automatic
 * (warp-cooperative kernels in base coordinates when every inertia has rigid-body structure), 1 / 2 one knot point per
 * thread (the reference's body-frame recursion), 3 cooperative.  Thread-safe; calls in flight keep their choice. */
int rbd_fb_model_set_kernel_variant(rbd_fb_model_t* m, int variant);
int rbd_fb_rnea_f64(const rbd_fb_model_t* m, int64_t B, const double* q, const double* qd, const double* qdd,
                    double gravity, double* c, double* v, double* a, double* f, void* stream);
int rbd_fb_rnea_f32(const rbd_fb_model_t* m, int64_t B, const float* q, const float* qd, const float* qdd,
                    float gravity, float* c, float* v, float* a, float* f, void* stream);
int rbd_fb_rnea_grad_f64(const rbd_fb_model_t* m, int64_t B, const double* q, const double* qd, const double* qdd,
                         double gravity, int use_velocity_damping, double* dc_du, double* c_out, void* stream);
int rbd_fb_rnea_grad_f32(const rbd_fb_model_t* m, int64_t B, const float* q, const float* qd, const float* qdd,
                         float gravity, int use_velocity_damping, float* dc_du, float* c_out, void* stream);
int rbd_fb_minv_f64(const rbd_fb_model_t* m, int64_t B, const double* q, int output_dense, double* Minv, void* stream);
int rbd_fb_minv_f32(const rbd_fb_model_t* m, int64_t B, const float* q, int output_dense, float* Minv, void* stream);
/* The eight per-pass helpers for a floating-base robot - the `floating_base` branches of rnea_fpass
 * (RBDReference.py:559-598), rnea_bpass (:600-621), minv_bpass (:630-735), minv_fpass (:737-783),
 * rnea_grad_fpass_dq (:1127-1187), rnea_grad_fpass_dqd (:1189-1255), rnea_grad_bpass_dq (:1257-1297) and
 * rnea_grad_bpass_dqd (:1299-1343), reference shapes with a leading batch axis and n = NB + 5:
 * v, a, f (B, 6, NB); Minv (B, n, n); F (B, n, 6, n); U (B, n, 6); Dinv (B, n) [= D, rows 0..5 unused];
 * dv, da, df (B, 6, n, NB); dc (B, n, n).  Same in-place contracts as the fixed-base helpers (f, Minv and F,
 * df are updated through the caller's pointers).  rnea_grad_fpass_dq needs NB >= 6 (the reference raises
 * IndexError at :1168 otherwise): RBD_E_UNSUPPORTED. */
int rbd_fb_rnea_fpass_f64(const rbd_fb_model_t* m, int64_t B, const double* q, const double* qd, const double* qdd, double gravity,
                          double* v, double* a, double* f, void* stream);
int rbd_fb_rnea_bpass_f64(const rbd_fb_model_t* m, int64_t B, const double* q, double* f, double* c, void* stream);
int rbd_fb_minv_bpass_f64(const rbd_fb_model_t* m, int64_t B, const double* q, double* Minv, double* F, double* U, double* Dinv, void* stream);
int rbd_fb_minv_fpass_f64(const rbd_fb_model_t* m, int64_t B, const double* q, double* Minv, double* F, const double* U, const double* Dinv,
                          void* stream);
int rbd_fb_rnea_grad_fpass_dq_f64(const rbd_fb_model_t* m, int64_t B, const double* q, const double* qd, const double* v, const double* a,
                                  double gravity, double* dv, double* da, double* df, void* stream);
int rbd_fb_rnea_grad_fpass_dqd_f64(const rbd_fb_model_t* m, int64_t B, const double* q, const double* qd, const double* v, double* dv,
                                   double* da, double* df, void* stream);
int rbd_fb_rnea_grad_bpass_dq_f64(const rbd_fb_model_t* m, int64_t B, const double* q, const double* f, double* df_dq, double* dc_dq,
                                  void* stream);
int rbd_fb_rnea_grad_bpass_dqd_f64(const rbd_fb_model_t* m, int64_t B, const double* q, double* df_dqd, int use_velocity_damping,
                                   double* dc_dqd, void* stream);
int rbd_fb_rnea_fpass_f32(const rbd_fb_model_t* m, int64_t B, const float* q, const float* qd, const float* qdd, float gravity,
                          float* v, float* a, float* f, void* stream);
int rbd_fb_rnea_bpass_f32(const rbd_fb_model_t* m, int64_t B, const float* q, float* f, float* c, void* stream);
int rbd_fb_minv_bpass_f32(const rbd_fb_model_t* m, int64_t B, const float* q, float* Minv, float* F, float* U, float* Dinv, void* stream);
int rbd_fb_minv_fpass_f32(const rbd_fb_model_t* m, int64_t B, const float* q, float* Minv, float* F, const float* U, const float* Dinv,
                          void* stream);
int rbd_fb_rnea_grad_fpass_dq_f32(const rbd_fb_model_t* m, int64_t B, const float* q, const float* qd, const float* v, const float* a,
                                  float gravity, float* dv, float* da, float* df, void* stream);
int rbd_fb_rnea_grad_fpass_dqd_f32(const rbd_fb_model_t* m, int64_t B, const float* q, const float* qd, const float* v, float* dv,
                                   float* da, float* df, void* stream);
int rbd_fb_rnea_grad_bpass_dq_f32(const rbd_fb_model_t* m, int64_t B, const float* q, const float* f, float* df_dq, float* dc_dq,
                                  void* stream);
int rbd_fb_rnea_grad_bpass_dqd_f32(const rbd_fb_model_t* m, int64_t B, const float* q, float* df_dqd, int use_velocity_damping,
                                   float* dc_dqd, void* stream);
/* forward_dynamics / forward_dynamics_grad (RBDReference.py:1369-1384) of a floating-base robot: the
 * same compositions as rbd_forward_dynamics*, u / qdd (B, NB+5), qdd_dq / qdd_dqd (B, NB+5, NB+5). */
int rbd_fb_forward_dynamics_f64(const rbd_fb_model_t* m, int64_t B, const double* q, const double* qd, const double* u,
                                double* qdd, double* Minv_out, void* stream);
int rbd_fb_forward_dynamics_f32(const rbd_fb_model_t* m, int64_t B, const float* q, const float* qd, const float* u,
                                float* qdd, float* Minv_out, void* stream);
int rbd_fb_forward_dynamics_grad_f64(const rbd_fb_model_t* m, int64_t B, const double* q, const double* qd,
                                     const double* u, double* qdd_dq, double* qdd_dqd, double* qdd_out, void* stream);
int rbd_fb_forward_dynamics_grad_f32(const rbd_fb_model_t* m, int64_t B, const float* q, const float* qd, const float* u,
                                     float* qdd_dq, float* qdd_dqd, float* qdd_out, void* stream);

/* ---- scratch memory ------------------------------------------------------------------------------ */
/* forward_dynamics(_grad) and the hybrid minv kernel take stream-ordered temporaries from one private memory pool per
 * device that KEEPS what it has allocated (no driver call on the steady-state path; the memory is invisible to
 * the caller's own allocator: Atlas forward_dynamics_grad on 2^20 knot points holds 23 GB).
 * rbd_trim_scratch returns everything above keep_bytes of the CURRENT device's pool to the driver (call it after a
 * large one-off batch; synchronise the streams that used the library first).
 * rbd_prepare_device creates the pool of `device` ahead of time: the first call that needs it must not happen
 * under CUDA-graph capture (pool creation is not capturable); rbd_model_create prepares the device that is
 * current at creation. */
int rbd_trim_scratch(int64_t keep_bytes);
int rbd_prepare_device(int device);

/* ---- measurement helpers (bench.py) ---------------------------------------------------------- */
/* Runs a dependent-chain FMA micro-benchmark on `stream`'s device and returns the achieved
 * FLOP/s (2 per FMA) in *flops_per_s; is_f64 selects DFMA or FFMA.  Used only to put a measured
 * denominator under the FP64/FP32 roofline fraction. */
int rbd_measure_fma_peak(int is_f64, double* flops_per_s, double* elapsed_ms, void* stream);
/* Number of kernel launches issued through this library by the calling process so far. */
int64_t rbd_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* RBD_B200_H_ */
