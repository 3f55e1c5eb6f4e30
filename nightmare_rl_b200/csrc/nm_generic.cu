// nm_generic.cu — physics step for general small legged robots whose model needs MuJoCo's Newton solver: models/anymal_c
// of the reference (BASELINE configs[3]; `anymal_c.xml:4` cone="elliptic" impratio="100", solver left at Newton, `:9` joint
// damping + friction loss, `:20-21` condim-6 sphere feet with priority, `:26` position actuators with a force range, joint
// limits, box / cylinder / sphere geoms against the plane, Euler integration with implicit joint damping).
//
// Replaces `mj.mj_step(model, data[i], nstep)` (reference call sites envs/nightmare_v3_env.py:200, simple_test.py:39) for
// such a model.  Scope: a free-floating base plus hinge joints (one joint per body, anchored at the body
// origin), collisions with ONE static plane; the model's geom-geom self collisions are not generated.
//
// Decomposition: ONE WARP PER ENVIRONMENT, four environments per CTA.  Everything of an environment lives in shared memory
// for the whole launch (state, kinematics, dense M / J / H for nv <= 24, up to 72 constraint rows: ~21 KB); every stage is a
// loop over bodies / dofs / rows / matrix entries strided over the 32 lanes with __syncwarp() between dependent stages;
// tree recursions go level by level.  The constraint solve is Newton's method on MuJoCo's convex primal problem with the
// exact cone Hessian and an exact line search (regula falsi on the directional derivative), dense Cholesky in shared memory.
#include <cuda_runtime.h>
#include <math_constants.h>

#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/nightmare_b200.h"

int nm_fail(int code, const std::string& msg);   // nm_abi.cu

#define GM_MAXB 16     // bodies, world included
#define GM_MAXV 24
#define GM_MAXQ 25
#define GM_MAXU 18
#define GM_MAXG 48     // collision geoms (plane excluded)
#define GM_MAXCON 16   // first tier: contacts / constraint rows per environment held in shared memory
#define GM_MAXROW 72
#ifndef GM_WARPS
#define GM_WARPS 4     // environments (warps) per CTA, in lockstep.  Measured per 4096-env x 4-substep launch: unsynchronised 12.4 ms; in lockstep with the first code 9.6 / 6.9 / 5.9 ms for 2 / 4 / 8 warps; with the final (4x smaller) code 3.53 ms for 2 CTAs x 4 warps, 3.80 for 1 x 8, 4.35 for 1 x 6
#endif
#define GM_MAXCHAIN 16  // dofs on the chain from the world to any body
#define GM_BIGCON 64   // second tier (same cap as the oracle's NMO_MAXCON); beyond it info[3] = 1 and the step is truncated
#define GM_BIGROW 240

enum { G_SPHERE = 2, G_CYLINDER = 5, G_BOX = 6 };
enum { R_FRICTION = 0, R_LIMIT = 1, R_CONTACT = 3 };

struct GenModel {
  int nq, nv, nu, nbody, ngeom, maxdepth, nfloss, cone, iterations, integrator, eulerdamp, pad0;
  float timestep, gravity[3], tolerance, impratio, solver_scale, pad1;
  float qpos0[GM_MAXQ];
  // bodies (one joint each; body 0 = world, body 1 = the free-floating base)
  int body_parent[GM_MAXB], body_depth[GM_MAXB], body_dofadr[GM_MAXB], body_qadr[GM_MAXB], body_dofmask[GM_MAXB];
  int body_ndof[GM_MAXB], body_dofs[GM_MAXB][GM_MAXCHAIN];   // dofs that move the body (its chain), ascending
  float body_pos[GM_MAXB][3], body_quat[GM_MAXB][4], body_ipos[GM_MAXB][3], body_iquat[GM_MAXB][4], body_mass[GM_MAXB], body_inertia[GM_MAXB][3];
  float body_invw[GM_MAXB][2], jnt_axis[GM_MAXB][3], jnt_range[GM_MAXB][2];
  int jnt_limited[GM_MAXB];
  // dofs
  int dof_body[GM_MAXV], dof_parent[GM_MAXV], dof_flossrow[GM_MAXV];
  float dof_damping[GM_MAXV], dof_floss[GM_MAXV], dof_armature[GM_MAXV], dof_invw[GM_MAXV];
  // actuators (affine: force = gain0 * ctrl + bias0 + bias1 * q + bias2 * qvel)
  int act_dof[GM_MAXU], act_qadr[GM_MAXU], act_ctrllimited[GM_MAXU], act_forcelimited[GM_MAXU];
  float act_gain0[GM_MAXU], act_bias[GM_MAXU][3], act_gear[GM_MAXU], act_ctrlrange[GM_MAXU][2], act_forcerange[GM_MAXU][2];
  // collision geoms against the plane, in MuJoCo's geom order
  int geom_type[GM_MAXG], geom_body[GM_MAXG], geom_dim[GM_MAXG];
  float geom_pos[GM_MAXG][3], geom_mat[GM_MAXG][9], geom_size[GM_MAXG][3], geom_margin[GM_MAXG];
  float geom_friction[GM_MAXG][5], geom_K[GM_MAXG], geom_B[GM_MAXG], geom_solimp[GM_MAXG][5];   // already mixed with the plane's parameters
  float plane_n[3], plane_pos[3], plane_frame[9];
  float lim_K, lim_B, lim_imp0, lim_solimp[5];        // default solref / solimp of friction-loss and limit rows
};

struct GenArgs {
  const GenModel* model;
  float* qpos; float* qvel; float* warm; const float* ctrl; int* info;   // info[N][4] = ncon, nefc, Newton iterations (last substep), overflow flag
  int* ovf;                                                                // work list of the second tier: count, then (env, substeps left) pairs
  int num_envs, nstep;
};

// ---------------------------------------------------------------------------------------------- small algebra (device)
__device__ __forceinline__ void g_quat2mat(const float* q, float* m) {
  const float w = q[0], x = q[1], y = q[2], z = q[3];
  m[0] = w * w + x * x - y * y - z * z; m[1] = 2.f * (x * y - w * z); m[2] = 2.f * (x * z + w * y);
  m[3] = 2.f * (x * y + w * z); m[4] = w * w - x * x + y * y - z * z; m[5] = 2.f * (y * z - w * x);
  m[6] = 2.f * (x * z - w * y); m[7] = 2.f * (y * z + w * x); m[8] = w * w - x * x - y * y + z * z;
}
__device__ __forceinline__ void g_mulquat(float* r, const float* a, const float* b) {
  const float w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  const float x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  const float y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  const float z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  r[0] = w; r[1] = x; r[2] = y; r[3] = z;
}
__device__ __forceinline__ void g_normquat(float* q) {
  const float n = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < 1e-15f) { q[0] = 1.f; q[1] = q[2] = q[3] = 0.f; return; }
  const float inv = 1.f / n;
  q[0] *= inv; q[1] *= inv; q[2] *= inv; q[3] *= inv;
}
__device__ __forceinline__ void g_matvec(float* r, const float* m, const float* v) {
  const float x = m[0] * v[0] + m[1] * v[1] + m[2] * v[2], y = m[3] * v[0] + m[4] * v[1] + m[5] * v[2], z = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
__device__ __forceinline__ void g_cross(float* r, const float* a, const float* b) {
  const float x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
__device__ __forceinline__ float g_dot3(const float* a, const float* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
// spatial inertia (Ixx Iyy Izz Ixy Ixz Iyz, m*r(3), m) times motion vector [w; v]
__device__ __forceinline__ void g_inertvec(float* res, const float* i, const float* v) {
  res[0] = i[0] * v[0] + i[3] * v[1] + i[4] * v[2] - i[8] * v[4] + i[7] * v[5];
  res[1] = i[3] * v[0] + i[1] * v[1] + i[5] * v[2] + i[8] * v[3] - i[6] * v[5];
  res[2] = i[4] * v[0] + i[5] * v[1] + i[2] * v[2] - i[7] * v[3] + i[6] * v[4];
  res[3] = i[8] * v[1] - i[7] * v[2] + i[9] * v[3];
  res[4] = i[6] * v[2] - i[8] * v[0] + i[9] * v[4];
  res[5] = i[7] * v[0] - i[6] * v[1] + i[9] * v[5];
}
__device__ __forceinline__ void g_crossmotion(float* r, const float* vel, const float* v) {
  float a[3], b[3], c[3];
  g_cross(a, vel, v); g_cross(b, vel, v + 3); g_cross(c, vel + 3, v);
  r[0] = a[0]; r[1] = a[1]; r[2] = a[2]; r[3] = b[0] + c[0]; r[4] = b[1] + c[1]; r[5] = b[2] + c[2];
}
__device__ __forceinline__ void g_crossforce(float* r, const float* vel, const float* f) {
  float a[3], b[3], c[3];
  g_cross(a, vel, f); g_cross(b, vel + 3, f + 3); g_cross(c, vel, f + 3);
  r[0] = a[0] + b[0]; r[1] = a[1] + b[1]; r[2] = a[2] + b[2]; r[3] = c[0]; r[4] = c[1]; r[5] = c[2];
}
__device__ __forceinline__ float g_warpsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __noinline__ float g_impedance_pow(float x, float mid, float power) {      // general solimp power: cold, out of line
  return x <= mid ? powf(x, power) / powf(mid, power - 1.f) : 1.f - powf(1.f - x, power) / powf(1.f - mid, power - 1.f);
}
__device__ __forceinline__ float g_impedance(const float* si, float pos) {      // si: dmin dmax width mid power (clamped on the host)
  if (si[0] == si[1] || si[2] <= 1e-15f) return 0.5f * (si[0] + si[1]);
  const float x = fabsf(pos / si[2]);
  if (x >= 1.f) return si[1];
  if (x <= 0.f) return si[0];
  float y;
  if (si[4] == 1.f) y = x;
  else if (si[4] == 2.f) y = x <= si[3] ? x * x / si[3] : 1.f - (1.f - x) * (1.f - x) / (1.f - si[3]);      // MuJoCo's default power
  else y = g_impedance_pow(x, si[3], si[4]);
  return si[0] + y * (si[1] - si[0]);
}

// ---------------------------------------------------------------------------------------------- per-environment shared memory
template <int MC, int MR>
struct alignas(16) EnvMemT {
  // fp64 islands of the Newton solver: the iterate and the constraint residual jar = J qacc - aref.  On a sliding elliptic
  // contact the force is Dm * mu * (mu * jar_n - mu * |friction . jar_t|) with Dm ~ 4e6 for the feet (impratio 100): the
  // difference of two O(10) numbers has to be right to 1e-9 for a force good to 1e-3 N, which fp32 residuals miss by three
  // orders of magnitude (measured: fp32 floor 1e-3 relative in qvel, with these two arrays in fp64 2.6e-6).
  double xd[GM_MAXV], jar[MR];
  float qpos[GM_MAXQ + 3], qvel[GM_MAXV], warm[GM_MAXV], ctrl[GM_MAXU + 2];
  float xpos[GM_MAXB][3], xmat[GM_MAXB][9];
  float com[4];
  float cdof[GM_MAXV][6];
  // Two phases share one block of memory (shared memory per environment decides how many warps an SM holds: 24.1 KB -> 19.1 KB
  // = 8 -> 12 warps): `k` is dead once the smooth dynamics are done (before collision), `s` is written from the constraint rows on
  alignas(16) union {
    struct { float xquat[GM_MAXB][4], xipos[GM_MAXB][3], ximat[GM_MAXB][9], cinert[GM_MAXB][10], crb[GM_MAXB][10], cdofdot[GM_MAXV][6],
                   cvel[GM_MAXB][6], cacc[GM_MAXB][6], cfrc[GM_MAXB][6]; } k;
    struct { float H[GM_MAXV][GM_MAXV], aref[MR], D[MR], R[MR], jv[MR], force[MR], Hd[MR], floss[MR]; } s;
  };
  alignas(16) float M[GM_MAXV][GM_MAXV];
  float smooth[GM_MAXV], qaccs[GM_MAXV], qacc[GM_MAXV], grad[GM_MAXV], dir[GM_MAXV], Ma[GM_MAXV], vec[GM_MAXV];
  // contacts
  int ncon, nefc, nlim, overflow;
  float cpos[MC][3], cdist[MC], cmu[MC];
  unsigned char cgeom[MC], cadr[MC], cdim[MC], czone[MC], crow[MR];   // crow: contact that owns a row (row addresses < 256 in both tiers)
  float cDm[MC], ccoef[MC], cb[MC][GM_MAXV];          // elliptic-cone curvature data (g_cone)
  // rows
  float J[MR][GM_MAXV];
  unsigned char rtype[2 * GM_MAXV], rdof[2 * GM_MAXV];  // simple rows (friction loss, limits): one dof each, J = rsgn * e_dof
  float rsgn[2 * GM_MAXV];
};


// x <- A^-1 x for the symmetric positive definite n x n matrix A (lower triangle read) by the whole warp, entirely in
// registers: lane i holds row i and its right-hand side, elimination of column j broadcasts pivot row j by shuffles and
// clears the column in ALL other rows (Gauss-Jordan: no pivoting needed for an SPD matrix, no back substitution, no
// transposed access).  ~n^2/2 shuffle + FMA pairs and a dependent chain of ~50 cycles per column, against ~4000 instructions
// and two shared-memory round trips per column of the left-looking Cholesky this replaces.  Returns false if a pivot is <= 0.
template <int N>
__device__ __noinline__ bool g_solve_spd_n(const float (*A)[GM_MAXV], float* x, int n, int lane) {
  float h[N];
  const int i = lane;
#pragma unroll
  for (int k = 0; k < N; k++) h[k] = (k == i) ? 1.f : 0.f;           // rows / columns beyond n: identity
  if (i < n) {
#pragma unroll
    for (int k = 0; k < N; k++) if (k < n) h[k] = k <= i ? A[i][k] : A[k][i];
  }
  float b = i < n ? x[i] : 0.f, d = 1.f;
  bool ok = true;
#pragma unroll
  for (int j = 0; j < N; j++) {
    float pj = __shfl_sync(0xffffffffu, h[j], j);
    ok &= pj > 1e-30f;
    pj = fmaxf(pj, 1e-30f);
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(pj));
    r = fmaf(r, fmaf(-pj, r, 1.f), r);                       // one Newton step: correctly rounded to within an ulp, no slow path
    d = i == j ? r : d;                                      // row j is final after its own column: 1 / its pivot
    const float f = i == j ? 0.f : h[j] * r;
#pragma unroll
    for (int k = j + 1; k < N; k++) h[k] = fmaf(-f, __shfl_sync(0xffffffffu, h[k], j), h[k]);
    b = fmaf(-f, __shfl_sync(0xffffffffu, b, j), b);
  }
  if (i < n) x[i] = b * d;
  __syncwarp();
  return ok;
}
// sizes the code is instantiated for: the quadrupeds / hexapods of the reference (nv = 18, 24); anything smaller pads
__device__ __forceinline__ bool g_solve_spd(const float (*A)[GM_MAXV], float* x, int n, int lane) {
  return n <= 18 ? g_solve_spd_n<18>(A, x, n, lane) : g_solve_spd_n<GM_MAXV>(A, x, n, lane);
}

// (row, column) of the p-th entry of a lower triangle, p < 16 * 17 / 2
__constant__ unsigned char g_tri_r[136] = {0, 1,1, 2,2,2, 3,3,3,3, 4,4,4,4,4, 5,5,5,5,5,5, 6,6,6,6,6,6,6, 7,7,7,7,7,7,7,7, 8,8,8,8,8,8,8,8,8, 9,9,9,9,9,9,9,9,9,9,
  10,10,10,10,10,10,10,10,10,10,10, 11,11,11,11,11,11,11,11,11,11,11,11, 12,12,12,12,12,12,12,12,12,12,12,12,12, 13,13,13,13,13,13,13,13,13,13,13,13,13,13,
  14,14,14,14,14,14,14,14,14,14,14,14,14,14,14, 15,15,15,15,15,15,15,15,15,15,15,15,15,15,15,15};
__constant__ unsigned char g_tri_c[136] = {0, 0,1, 0,1,2, 0,1,2,3, 0,1,2,3,4, 0,1,2,3,4,5, 0,1,2,3,4,5,6, 0,1,2,3,4,5,6,7, 0,1,2,3,4,5,6,7,8, 0,1,2,3,4,5,6,7,8,9,
  0,1,2,3,4,5,6,7,8,9,10, 0,1,2,3,4,5,6,7,8,9,10,11, 0,1,2,3,4,5,6,7,8,9,10,11,12, 0,1,2,3,4,5,6,7,8,9,10,11,12,13,
  0,1,2,3,4,5,6,7,8,9,10,11,12,13,14, 0,1,2,3,4,5,6,7,8,9,10,11,12,13,14,15};

// ---------------------------------------------------------------------------------------------- constraint rows
// sqrt of a double from the fp32 rsqrt plus one Newton step in fp64 (relative error ~1e-14): the fp64 sqrt / divide are
// software sequences of a few hundred cycles, this is 5 DFMA-class operations
__device__ __forceinline__ double g_sqrt64(double x, double* rinv) {
  if (!(x > 0.0)) { *rinv = 0.0; return 0.0; }
  const double r0 = (double)rsqrtf((float)x);
  const double r = r0 * (1.5 - 0.5 * x * r0 * r0);
  *rinv = r;
  return x * r;
}

// One elliptic contact at residual j[0..dim): zone (0 top: no force, 1 bottom: quadratic in every row, 2 middle: on the cone),
// forces f, cost; for zone 2 also the curvature data of the Hessian Dm g g' + c F (I - uu') F  (g = mu (e0 - F u)):
// uh[j] = friction_j * u_j (j >= 1) and the scalar c = -Dm mu (N - mu T) / T >= 0.
__device__ __forceinline__ int g_cone(int dim, const double* j, const float* fri, float mu, const float* D, float Dm, float* f, float* cost, float* uh, float* ccoef) {
  if (dim == 1) {
    if (j[0] < 0.0) { const float jj = (float)j[0]; f[0] = -D[0] * jj; *cost = 0.5f * D[0] * jj * jj; return 1; }
    f[0] = 0.f; *cost = 0.f; return 0;
  }
  double U[6], T2 = 0.0, rT;
#pragma unroll
  for (int k = 1; k < 6; k++) if (k < dim) { U[k] = j[k] * (double)fri[k - 1]; T2 += U[k] * U[k]; }
  const double N = j[0] * (double)mu, T = g_sqrt64(T2, &rT);
  if (N >= (double)mu * T || (T <= 0.0 && N >= 0.0)) {
#pragma unroll
    for (int k = 0; k < 6; k++) f[k] = 0.f;
    *cost = 0.f;
    return 0;
  }
  if ((double)mu * N + T <= 0.0 || (T <= 0.0 && N < 0.0)) {
    float cs = 0.f;
#pragma unroll
    for (int k = 0; k < 6; k++) if (k < dim) { const float jj = (float)j[k]; f[k] = -D[k] * jj; cs += 0.5f * D[k] * jj * jj; }
    *cost = cs;
    return 1;
  }
  const float NmT = (float)(N - (double)mu * T);
  const float f0 = -Dm * NmT * mu;
  f[0] = f0;
#pragma unroll
  for (int k = 1; k < 6; k++) if (k < dim) { const float u = (float)(U[k] * rT) * fri[k - 1]; uh[k] = u; f[k] = -f0 * u; }
  *cost = 0.5f * Dm * NmT * NmT;
  *ccoef = -Dm * mu * NmT * (float)rT;
  return 2;
}

// friction-loss / joint-limit row at residual j: force, cost, curvature
__device__ __forceinline__ float g_simple(int type, float D, float R, float fl, float j, float* cost, float* hd) {
  if (type == R_FRICTION) {
    const float bound = R * fl;
    if (j <= -bound) { *cost = -0.5f * R * fl * fl - fl * j; *hd = 0.f; return fl; }
    if (j >= bound) { *cost = -0.5f * R * fl * fl + fl * j; *hd = 0.f; return -fl; }
  } else if (!(j < 0.f)) { *cost = 0.f; *hd = 0.f; return 0.f; }
  *cost = 0.5f * D * j * j; *hd = D;
  return -D * j;
}

// Row evaluation at the residual e.jar + alpha * e.s.jv, one copy of the code for its three uses (the kernel is bound by
// instruction fetch):  mode 0 = line search: returns sum_r force_r jv_r, stores nothing;  mode 1 = forces into e.s.force, returns
// the cost;  mode 2 = mode 1 plus the curvature data (e.s.Hd / e.czone / e.ccoef / e.cb).  alpha is ignored unless mode == 0.
template <class EnvMem>
__device__ __noinline__ float g_rows(const GenModel& m, EnvMem& e, int mode, float alpha, int lane) {
  float acc = 0.f;
  const int nsimple = m.nfloss + e.nlim;
  for (int r = lane; r < nsimple; r += 32) {
    float cs, hd;
    const float jv = mode == 0 ? e.s.jv[r] : 0.f;
    const float f = g_simple(e.rtype[r], e.s.D[r], e.s.R[r], e.s.floss[r], (float)(mode == 0 ? e.jar[r] + (double)alpha * (double)jv : e.jar[r]), &cs, &hd);
    if (mode == 0) acc += f * jv;
    else { e.s.force[r] = f; acc += cs; if (mode == 2) e.s.Hd[r] = hd; }
  }
  for (int c = lane; c < e.ncon; c += 32) {
    const int a = e.cadr[c], dim = e.cdim[c];
    const float* fri = m.geom_friction[e.cgeom[c]];
    double j[6];
    float jv[6];
#pragma unroll
    for (int k = 0; k < 6; k++) if (k < dim) { j[k] = e.jar[a + k]; if (mode == 0) { jv[k] = e.s.jv[a + k]; j[k] += (double)alpha * (double)jv[k]; } }
    float f[6], uh[6], cs, cc = 0.f;
    const int zone = g_cone(dim, j, fri, e.cmu[c], e.s.D + a, e.cDm[c], f, &cs, uh, &cc);
    if (mode == 0) {
#pragma unroll
      for (int k = 0; k < 6; k++) if (k < dim) acc += f[k] * jv[k];
      continue;
    }
#pragma unroll
    for (int k = 0; k < 6; k++) if (k < dim) e.s.force[a + k] = f[k];
    acc += cs;
    if (mode == 2) {
      e.czone[c] = zone;
      if (zone == 2) {
        e.ccoef[c] = cc;
        const int body = m.geom_body[e.cgeom[c]], nd = m.body_ndof[body];
#pragma unroll 4
        for (int q = 0; q < nd; q++) {                     // b = sum_k friction_k u_k J_k, on the dofs of the body's chain
          const int i = m.body_dofs[body][q];
          float t = 0.f;
#pragma unroll
          for (int k = 1; k < 6; k++) if (k < dim) t += uh[k] * e.J[a + k][i];
          e.cb[c][i] = t;
        }
      }
    }
  }
  return g_warpsum(acc);
}

// ================================================================================================ the kernel
// Two capacity tiers.  FIRST: <GM_MAXCON contacts, GM_MAXROW rows>, GM_WARPS environments per CTA, one warp per environment
// of the batch.  An environment whose substep needs more (a robot lying on its boxes and cylinders: ~1 % of tumbling robots)
// is handed over untouched at the START of that substep through the work list A.ovf = {count, (env, substeps left) ...};
// the second launch (!FIRST: <GM_BIGCON, GM_BIGROW>, one environment per CTA at a time) finishes those.
template <int MC, int MR, int WARPS, bool FIRST>
__global__ void __launch_bounds__(WARPS * 32) nm_generic_step_kernel(const GenArgs A) {
  typedef EnvMemT<MC, MR> EnvMem;
  extern __shared__ __align__(16) unsigned char gm_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  EnvMem& e = *reinterpret_cast<EnvMem*>(gm_smem + sizeof(EnvMem) * warp);
  const GenModel& m = *A.model;
  const int nv = m.nv, nq = m.nq, nb = m.nbody, nu = m.nu;
  const float h = m.timestep;
  const int nwork = FIRST ? A.num_envs : min(A.ovf[0], A.num_envs);
  // The warps of a CTA run in LOCKSTEP through the substeps and the Newton iterations (CTA barriers at the top of every
  // Newton trip): the kernel is bound by instruction fetch (ncu: 72 % of the stall samples were no_inst with eight warps per SM
  // each at its own place in ~50 KB of code against a 32 KB instruction cache); warps that walk the code together share the fetches.
  // `live` = this warp has an environment to work on; warps without one (tail of the batch, handed over) still meet the barriers.
  for (int base = blockIdx.x * WARPS; base < nwork; base += gridDim.x * WARPS) {        // (FIRST: the grid covers the batch, one trip)
  const int item = base + warp;
  bool live = item < nwork;
  const int env = !live ? 0 : FIRST ? item : A.ovf[1 + 2 * item];
  const int nstep = FIRST ? A.nstep : (live ? A.ovf[2 + 2 * item] : 0);          // uniform over the CTA (the second tier has one warp per CTA)
  __syncwarp();

  if (live) {
    for (int i = lane; i < nq; i += 32) e.qpos[i] = A.qpos[(size_t)env * nq + i];
    for (int i = lane; i < nv; i += 32) { e.qvel[i] = A.qvel[(size_t)env * nv + i]; e.warm[i] = A.warm[(size_t)env * nv + i]; }
    for (int i = lane; i < nu; i += 32) e.ctrl[i] = A.ctrl[(size_t)env * nu + i];
  }
  int niter_last = 0;
  __syncwarp();

  for (int sub = 0; sub < nstep; sub++) {
    if (live) do {
    // ------------------------------------------------------------------ divergence guard (≙ mj_checkPos / mj_checkVel)
    {
      bool bad = false;
      for (int i = lane; i < nq; i += 32) bad |= !(fabsf(e.qpos[i]) < 1e10f);
      for (int i = lane; i < nv; i += 32) bad |= !(fabsf(e.qvel[i]) < 1e10f);
      if (__any_sync(0xffffffffu, bad)) {
        for (int i = lane; i < nq; i += 32) e.qpos[i] = m.qpos0[i];
        for (int i = lane; i < nv; i += 32) { e.qvel[i] = 0.f; e.warm[i] = 0.f; }
      }
      __syncwarp();
    }
    // ------------------------------------------------------------------ P1 kinematics, level by level
    if (lane == 0) {
      e.xpos[0][0] = e.xpos[0][1] = e.xpos[0][2] = 0.f;
      e.k.xquat[0][0] = 1.f; e.k.xquat[0][1] = e.k.xquat[0][2] = e.k.xquat[0][3] = 0.f;
      g_quat2mat(e.k.xquat[0], e.xmat[0]);
    }
    __syncwarp();
    for (int lev = 1; lev <= m.maxdepth; lev++) {
      for (int b = 1 + lane; b < nb; b += 32) {
        if (m.body_depth[b] != lev) continue;
        float pos[3], quat[4];
        if (b == 1) {                                  // free joint
          pos[0] = e.qpos[0]; pos[1] = e.qpos[1]; pos[2] = e.qpos[2];
          quat[0] = e.qpos[3]; quat[1] = e.qpos[4]; quat[2] = e.qpos[5]; quat[3] = e.qpos[6];
          g_normquat(quat);
        } else {
          const int p = m.body_parent[b];
          float t[3];
          g_matvec(t, e.xmat[p], m.body_pos[b]);
          pos[0] = e.xpos[p][0] + t[0]; pos[1] = e.xpos[p][1] + t[1]; pos[2] = e.xpos[p][2] + t[2];
          g_mulquat(quat, e.k.xquat[p], m.body_quat[b]);
          const float ang = e.qpos[m.body_qadr[b]] - m.qpos0[m.body_qadr[b]];
          float s, c;
          sincosf(0.5f * ang, &s, &c);
          const float ql[4] = {c, m.jnt_axis[b][0] * s, m.jnt_axis[b][1] * s, m.jnt_axis[b][2] * s};
          float qn[4];
          g_mulquat(qn, quat, ql);
          quat[0] = qn[0]; quat[1] = qn[1]; quat[2] = qn[2]; quat[3] = qn[3];      // (joint anchor at the body origin: no offset correction)
        }
        g_normquat(quat);
        for (int k = 0; k < 3; k++) e.xpos[b][k] = pos[k];
        for (int k = 0; k < 4; k++) e.k.xquat[b][k] = quat[k];
        g_quat2mat(quat, e.xmat[b]);
        float t[3], qi[4];
        g_matvec(t, e.xmat[b], m.body_ipos[b]);
        for (int k = 0; k < 3; k++) e.k.xipos[b][k] = pos[k] + t[k];
        g_mulquat(qi, quat, m.body_iquat[b]);
        g_quat2mat(qi, e.k.ximat[b]);
      }
      __syncwarp();
    }
    // ------------------------------------------------------------------ P2 comPos: subtree COM of the root, cinert, cdof
    {
      float mx = 0.f, my = 0.f, mz = 0.f, mm = 0.f;
      for (int b = 1 + lane; b < nb; b += 32) { const float ms = m.body_mass[b]; mx += ms * e.k.xipos[b][0]; my += ms * e.k.xipos[b][1]; mz += ms * e.k.xipos[b][2]; mm += ms; }
      mx = g_warpsum(mx); my = g_warpsum(my); mz = g_warpsum(mz); mm = g_warpsum(mm);
      if (lane == 0) { e.com[0] = mx / mm; e.com[1] = my / mm; e.com[2] = mz / mm; }
      __syncwarp();
    }
    for (int b = lane; b < nb; b += 32) {
      float* ci = e.k.cinert[b];
      if (b == 0) { for (int k = 0; k < 10; k++) ci[k] = 0.f; continue; }
      const float* Rm = e.k.ximat[b];
      const float* I = m.body_inertia[b];
      const float mass = m.body_mass[b];
      const float r[3] = {e.k.xipos[b][0] - e.com[0], e.k.xipos[b][1] - e.com[1], e.k.xipos[b][2] - e.com[2]};
      ci[0] = Rm[0] * Rm[0] * I[0] + Rm[1] * Rm[1] * I[1] + Rm[2] * Rm[2] * I[2];
      ci[1] = Rm[3] * Rm[3] * I[0] + Rm[4] * Rm[4] * I[1] + Rm[5] * Rm[5] * I[2];
      ci[2] = Rm[6] * Rm[6] * I[0] + Rm[7] * Rm[7] * I[1] + Rm[8] * Rm[8] * I[2];
      ci[3] = Rm[0] * Rm[3] * I[0] + Rm[1] * Rm[4] * I[1] + Rm[2] * Rm[5] * I[2];
      ci[4] = Rm[0] * Rm[6] * I[0] + Rm[1] * Rm[7] * I[1] + Rm[2] * Rm[8] * I[2];
      ci[5] = Rm[3] * Rm[6] * I[0] + Rm[4] * Rm[7] * I[1] + Rm[5] * Rm[8] * I[2];
      ci[0] += mass * (r[1] * r[1] + r[2] * r[2]);
      ci[1] += mass * (r[0] * r[0] + r[2] * r[2]);
      ci[2] += mass * (r[0] * r[0] + r[1] * r[1]);
      ci[3] -= mass * r[0] * r[1];
      ci[4] -= mass * r[0] * r[2];
      ci[5] -= mass * r[1] * r[2];
      ci[6] = mass * r[0]; ci[7] = mass * r[1]; ci[8] = mass * r[2];
      ci[9] = mass;
    }
    for (int i = lane; i < nv; i += 32) {
      float* cd = e.cdof[i];
      const int b = m.dof_body[i];
      if (b == 1) {
        if (i < 3) { for (int k = 0; k < 6; k++) cd[k] = 0.f; cd[3 + i] = 1.f; }
        else {
          const int k = i - 3;
          const float ax[3] = {e.xmat[1][k], e.xmat[1][3 + k], e.xmat[1][6 + k]};
          const float off[3] = {e.com[0] - e.xpos[1][0], e.com[1] - e.xpos[1][1], e.com[2] - e.xpos[1][2]};
          cd[0] = ax[0]; cd[1] = ax[1]; cd[2] = ax[2];
          g_cross(cd + 3, ax, off);
        }
      } else {
        float ax[3];
        g_matvec(ax, e.xmat[b], m.jnt_axis[b]);
        const float off[3] = {e.com[0] - e.xpos[b][0], e.com[1] - e.xpos[b][1], e.com[2] - e.xpos[b][2]};
        cd[0] = ax[0]; cd[1] = ax[1]; cd[2] = ax[2];
        g_cross(cd + 3, ax, off);
      }
    }
    __syncwarp();
    // ------------------------------------------------------------------ P3 composite inertias (bottom-up), mass matrix, Cholesky
    for (int b = lane; b < nb; b += 32) for (int k = 0; k < 10; k++) e.k.crb[b][k] = e.k.cinert[b][k];
    __syncwarp();
    for (int lev = m.maxdepth - 1; lev >= 1; lev--) {
      for (int b = 1 + lane; b < nb; b += 32) {
        if (m.body_depth[b] != lev) continue;
        for (int c = b + 1; c < nb; c++)
          if (m.body_parent[c] == b) for (int k = 0; k < 10; k++) e.k.crb[b][k] += e.k.crb[c][k];
      }
      __syncwarp();
    }
    for (int i = lane; i < GM_MAXV * GM_MAXV / 4; i += 32) reinterpret_cast<float4*>(&e.M[0][0])[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncwarp();
    for (int i = lane; i < nv; i += 32) {
      float buf[6];
      g_inertvec(buf, e.k.crb[m.dof_body[i]], e.cdof[i]);
      for (int j = i; j >= 0; j = m.dof_parent[j]) {
        float s = 0.f;
        for (int k = 0; k < 6; k++) s += e.cdof[j][k] * buf[k];
        e.M[i][j] = s; e.M[j][i] = s;
      }
      e.M[i][i] += m.dof_armature[i];
    }
    __syncwarp();
    // ------------------------------------------------------------------ P7 comVel + RNE
    if (lane == 0) for (int k = 0; k < 6; k++) { e.k.cvel[0][k] = 0.f; e.k.cacc[0][k] = k >= 3 ? -m.gravity[k - 3] : 0.f; }
    __syncwarp();
    for (int lev = 1; lev <= m.maxdepth; lev++) {
      for (int b = 1 + lane; b < nb; b += 32) {
        if (m.body_depth[b] != lev) continue;
        float cv[6], ca[6];
        const int p = m.body_parent[b], da = m.body_dofadr[b];
        for (int k = 0; k < 6; k++) { cv[k] = e.k.cvel[p][k]; ca[k] = e.k.cacc[p][k]; }
        if (b == 1) {
          for (int d = 0; d < 3; d++) { for (int k = 0; k < 6; k++) { e.k.cdofdot[da + d][k] = 0.f; cv[k] += e.cdof[da + d][k] * e.qvel[da + d]; } }
          for (int d = 3; d < 6; d++) g_crossmotion(e.k.cdofdot[da + d], cv, e.cdof[da + d]);
          for (int d = 3; d < 6; d++) for (int k = 0; k < 6; k++) cv[k] += e.cdof[da + d][k] * e.qvel[da + d];
          for (int d = 0; d < 6; d++) for (int k = 0; k < 6; k++) ca[k] += e.k.cdofdot[da + d][k] * e.qvel[da + d];
        } else {
          g_crossmotion(e.k.cdofdot[da], cv, e.cdof[da]);
          for (int k = 0; k < 6; k++) { cv[k] += e.cdof[da][k] * e.qvel[da]; ca[k] += e.k.cdofdot[da][k] * e.qvel[da]; }
        }
        float f[6], t1[6], t2[6];
        g_inertvec(f, e.k.cinert[b], ca);
        g_inertvec(t1, e.k.cinert[b], cv);
        g_crossforce(t2, cv, t1);
        for (int k = 0; k < 6; k++) { e.k.cvel[b][k] = cv[k]; e.k.cacc[b][k] = ca[k]; e.k.cfrc[b][k] = f[k] + t2[k]; }
      }
      __syncwarp();
    }
    for (int lev = m.maxdepth - 1; lev >= 1; lev--) {
      for (int b = 1 + lane; b < nb; b += 32) {
        if (m.body_depth[b] != lev) continue;
        for (int c = b + 1; c < nb; c++)
          if (m.body_parent[c] == b) for (int k = 0; k < 6; k++) e.k.cfrc[b][k] += e.k.cfrc[c][k];
      }
      __syncwarp();
    }
    // ------------------------------------------------------------------ P8 passive + actuation + smooth acceleration
    for (int i = lane; i < nv; i += 32) {
      float s = 0.f;
      for (int k = 0; k < 6; k++) s += e.cdof[i][k] * e.k.cfrc[m.dof_body[i]][k];
      e.smooth[i] = -m.dof_damping[i] * e.qvel[i] - s;
    }
    __syncwarp();
    for (int a = lane; a < nu; a += 32) {
      float c = e.ctrl[a];
      if (m.act_ctrllimited[a]) c = fminf(fmaxf(c, m.act_ctrlrange[a][0]), m.act_ctrlrange[a][1]);
      const int dof = m.act_dof[a];
      const float g = m.act_gear[a];
      float f = m.act_gain0[a] * c + m.act_bias[a][0] + m.act_bias[a][1] * (e.qpos[m.act_qadr[a]] * g) + m.act_bias[a][2] * (e.qvel[dof] * g);
      if (m.act_forcelimited[a]) f = fminf(fmaxf(f, m.act_forcerange[a][0]), m.act_forcerange[a][1]);
      e.smooth[dof] += g * f;                          // (one actuator per dof: checked on the host)
    }
    __syncwarp();
    for (int i = lane; i < nv; i += 32) e.qaccs[i] = e.smooth[i];
    __syncwarp();
    g_solve_spd(e.M, e.qaccs, nv, lane);
    // ------------------------------------------------------------------ P4 collision: sphere / box / cylinder against the plane
    {
      int base = 0;
      if (lane == 0) e.overflow = 0;
      for (int g0 = 0; g0 < m.ngeom; g0 += 32) {
        const int g = g0 + lane;
        int cnt = 0;
        float cp[4][3], cd[4];
        if (g < m.ngeom) {
          const int b = m.geom_body[g];
          float gc[3], gm[9];
          g_matvec(gc, e.xmat[b], m.geom_pos[g]);
          gc[0] += e.xpos[b][0]; gc[1] += e.xpos[b][1]; gc[2] += e.xpos[b][2];
          const float* n = m.plane_n;
          const float dif[3] = {gc[0] - m.plane_pos[0], gc[1] - m.plane_pos[1], gc[2] - m.plane_pos[2]};
          const float dist0 = g_dot3(dif, n), margin = m.geom_margin[g];
          const float* size = m.geom_size[g];
          if (m.geom_type[g] == G_SPHERE) {
            const float dist = dist0 - size[0];
            if (dist <= margin) { cd[0] = dist; for (int k = 0; k < 3; k++) cp[0][k] = gc[k] - n[k] * (size[0] + 0.5f * dist); cnt = 1; }
          } else {
            for (int r = 0; r < 3; r++)
              for (int c2 = 0; c2 < 3; c2++) gm[3 * r + c2] = e.xmat[b][3 * r] * m.geom_mat[g][c2] + e.xmat[b][3 * r + 1] * m.geom_mat[g][3 + c2] + e.xmat[b][3 * r + 2] * m.geom_mat[g][6 + c2];
            if (m.geom_type[g] == G_BOX) {
              for (int i = 0; i < 8 && cnt < 4; i++) {
                const float v[3] = {(i & 1) ? size[0] : -size[0], (i & 2) ? size[1] : -size[1], (i & 4) ? size[2] : -size[2]};
                float corner[3];
                g_matvec(corner, gm, v);
                const float ld = g_dot3(n, corner);
                if (dist0 + ld > margin || ld > 0.f) continue;
                cd[cnt] = dist0 + ld;
                for (int k = 0; k < 3; k++) cp[cnt][k] = corner[k] - n[k] * cd[cnt] * 0.5f + gc[k];
                cnt++;
              }
            } else {                                   // cylinder (≙ mjc_PlaneCylinder)
              float axis[3] = {gm[2], gm[5], gm[8]}, vec[3];
              float prjaxis = g_dot3(n, axis);
              if (prjaxis > 0.f) { axis[0] = -axis[0]; axis[1] = -axis[1]; axis[2] = -axis[2]; prjaxis = -prjaxis; }
              for (int k = 0; k < 3; k++) vec[k] = axis[k] * prjaxis - n[k];
              const float len2 = g_dot3(vec, vec);
              if (len2 >= 1e-30f) { const float scl = size[0] / sqrtf(len2); vec[0] *= scl; vec[1] *= scl; vec[2] *= scl; }
              else { vec[0] = gm[0] * size[0]; vec[1] = gm[3] * size[0]; vec[2] = gm[6] * size[0]; }
              const float prjvec = g_dot3(vec, n);
              axis[0] *= size[1]; axis[1] *= size[1]; axis[2] *= size[1];
              prjaxis *= size[1];
              if (dist0 + prjaxis + prjvec <= margin) {
                cd[cnt] = dist0 + prjaxis + prjvec;
                for (int k = 0; k < 3; k++) cp[cnt][k] = gc[k] + vec[k] + axis[k] - n[k] * cd[cnt] * 0.5f;
                cnt++;
                if (dist0 - prjaxis + prjvec <= margin) {
                  cd[cnt] = dist0 - prjaxis + prjvec;
                  for (int k = 0; k < 3; k++) cp[cnt][k] = gc[k] + vec[k] - axis[k] - n[k] * cd[cnt] * 0.5f;
                  cnt++;
                }
                const float prjvec1 = -prjvec * 0.5f;
                if (dist0 + prjaxis + prjvec1 <= margin) {
                  float v1[3];
                  g_cross(v1, vec, axis);
                  const float nn = sqrtf(g_dot3(v1, v1));
                  if (nn < 1e-15f) { v1[0] = 1.f; v1[1] = 0.f; v1[2] = 0.f; } else { v1[0] /= nn; v1[1] /= nn; v1[2] /= nn; }
                  const float sc = size[0] * 0.8660254037844386f;
                  for (int sg = 0; sg < 2; sg++) {
                    cd[cnt] = dist0 + prjaxis + prjvec1;
                    for (int k = 0; k < 3; k++) cp[cnt][k] = gc[k] + (sg ? -v1[k] : v1[k]) * sc + axis[k] - vec[k] * 0.5f - n[k] * cd[cnt] * 0.5f;
                    cnt++;
                  }
                }
              }
            }
          }
        }
        // contacts in geom order: exclusive prefix sum of the counts over the lanes
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        const int first = base + incl - cnt;
        for (int q = 0; q < cnt; q++) {
          const int c = first + q;
          if (c < MC) {
            e.cdist[c] = cd[q]; e.cgeom[c] = g;
            for (int k = 0; k < 3; k++) e.cpos[c][k] = cp[q][k];
          } else e.overflow = 1;
        }
        base += __shfl_sync(0xffffffffu, incl, 31);
      }
      __syncwarp();
      if (lane == 0) e.ncon = min(base, MC);
      __syncwarp();
    }
    // ------------------------------------------------------------------ P5 constraint rows: friction loss, limits, contacts
    {
      // friction-loss rows (host-assigned row per dof)
      for (int i = lane; i < nv; i += 32) {
        const int r = m.dof_flossrow[i];
        if (r < 0) continue;
        const float imp = m.lim_imp0;
        const float R = fmaxf((1.f - imp) * m.dof_invw[i] / imp, 1e-15f);
        e.s.R[r] = R; e.s.D[r] = 1.f / R; e.s.floss[r] = m.dof_floss[i];
        e.s.aref[r] = -m.lim_B * e.qvel[i];
        e.rtype[r] = R_FRICTION; e.rdof[r] = i; e.rsgn[r] = 1.f;
      }
      // joint limits: lower side first, then upper, in joint (= body) order
      int base = m.nfloss;
      for (int b0 = 2; b0 < nb; b0 += 32) {
        const int b = b0 + lane;
        int cnt = 0;
        float dist[2], sgn[2];
        if (b < nb && m.jnt_limited[b]) {
          const float value = e.qpos[m.body_qadr[b]];
          const float dlo = value - m.jnt_range[b][0], dhi = m.jnt_range[b][1] - value;
          if (dlo < 0.f) { dist[cnt] = dlo; sgn[cnt] = 1.f; cnt++; }
          if (dhi < 0.f) { dist[cnt] = dhi; sgn[cnt] = -1.f; cnt++; }
        }
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        const int first = base + incl - cnt;
        for (int q = 0; q < cnt; q++) {
          const int r = first + q, i = m.body_dofadr[b];
          const float imp = g_impedance(m.lim_solimp, dist[q]);
          const float R = fmaxf((1.f - imp) * m.dof_invw[i] / imp, 1e-15f);
          e.s.R[r] = R; e.s.D[r] = 1.f / R; e.s.floss[r] = 0.f;
          e.s.aref[r] = -m.lim_B * (sgn[q] * e.qvel[i]) - m.lim_K * imp * dist[q];
          e.rtype[r] = R_LIMIT; e.rdof[r] = i; e.rsgn[r] = sgn[q];
        }
        base += __shfl_sync(0xffffffffu, incl, 31);
      }
      if (lane == 0) e.nlim = base - m.nfloss;
      __syncwarp();
      // contacts: row addresses (serial prefix over <= 16 contacts), then the Jacobian entries spread over the lanes
      if (lane == 0) {
        int r = m.nfloss + e.nlim, nc = 0;
        for (int c = 0; c < e.ncon; c++) {
          const int dim = m.geom_dim[e.cgeom[c]];
          if (r + dim > MR) { e.overflow = 1; break; }
          e.cadr[c] = r; e.cdim[c] = dim;
          for (int k = 0; k < dim; k++) e.crow[r + k] = c;
          r += dim; nc++;
        }
        e.ncon = nc; e.nefc = r;
      }
      __syncwarp();
      if (FIRST && e.overflow) {                       // hand the environment over as it is at the start of this substep
        for (int i = lane; i < nq; i += 32) A.qpos[(size_t)env * nq + i] = e.qpos[i];
        for (int i = lane; i < nv; i += 32) { A.qvel[(size_t)env * nv + i] = e.qvel[i]; A.warm[(size_t)env * nv + i] = e.warm[i]; }
        if (lane == 0) {
          const int k = atomicAdd(A.ovf, 1);
          A.ovf[1 + 2 * k] = env; A.ovf[2 + 2 * k] = nstep - sub;
        }
        live = false;
        break;
      }
      // contact Jacobians: one (contact, dof) pair per lane trip, all rows of the contact at once
      for (int idx = lane; idx < e.ncon * 32; idx += 32) {   // (contact c, dof lane): no division by nv; nv <= 24 < 32
        const int c = idx >> 5, i = idx & 31;
        if (i >= nv) continue;
        const int a = e.cadr[c], dim = e.cdim[c], b = m.geom_body[e.cgeom[c]];
        float lin[3] = {0.f, 0.f, 0.f}, ang[3] = {0.f, 0.f, 0.f};
        if ((m.body_dofmask[b] >> i) & 1) {
          const float off[3] = {e.cpos[c][0] - e.com[0], e.cpos[c][1] - e.com[1], e.cpos[c][2] - e.com[2]};
          g_cross(lin, e.cdof[i], off);
          lin[0] += e.cdof[i][3]; lin[1] += e.cdof[i][4]; lin[2] += e.cdof[i][5];
          ang[0] = e.cdof[i][0]; ang[1] = e.cdof[i][1]; ang[2] = e.cdof[i][2];
        }
        for (int k = 0; k < dim; k++) {                  // 0 normal, 1-2 tangents, 3 torsion, 4-5 rolling
          const float* fr = m.plane_frame + 3 * (k < 3 ? k : k - 3);
          const float* v = k < 3 ? lin : ang;
          e.J[a + k][i] = fr[0] * v[0] + fr[1] * v[1] + fr[2] * v[2];
        }
      }
      __syncwarp();
      for (int c = lane; c < e.ncon; c += 32) {
        const int a = e.cadr[c], dim = e.cdim[c], g = e.cgeom[c], b = m.geom_body[g];
        const float pos = e.cdist[c] - m.geom_margin[g];
        const float imp = g_impedance(m.geom_solimp[g], pos);
        const float tran = m.body_invw[b][0];
        const float R0 = fmaxf((1.f - imp) * tran / imp, 1e-15f);
        e.s.R[a] = R0;
        if (dim > 1) {
          const float* fri = m.geom_friction[g];
          const float R1 = R0 / fmaxf(m.impratio, 1e-15f);
          e.s.R[a + 1] = R1;
          for (int j = 2; j < dim; j++) e.s.R[a + j] = R1 * fri[0] * fri[0] / (fri[j - 1] * fri[j - 1]);
          e.cmu[c] = fri[0] * sqrtf(R1 / R0);
        } else e.cmu[c] = 0.f;
        for (int j = 0; j < dim; j++) {
          float vel = 0.f;
#pragma unroll
          for (int i = 0; i < GM_MAXV; i++) if (i < nv) vel = fmaf(e.J[a + j][i], e.qvel[i], vel);
          e.s.R[a + j] = fmaxf(e.s.R[a + j], 1e-15f);
          e.s.D[a + j] = 1.f / e.s.R[a + j];
          e.s.aref[a + j] = -m.geom_B[g] * vel - (j == 0 ? m.geom_K[g] * imp * pos : 0.f);
        }
        e.cDm[c] = dim > 1 ? e.s.D[a] / (e.cmu[c] * e.cmu[c] * (1.f + e.cmu[c] * e.cmu[c])) : 0.f;
      }
      __syncwarp();
    }
    } while (0);
    // ------------------------------------------------------------------ P9 Newton solver on the primal problem
    const int ne = live ? e.nefc : 0, nsimple = live ? m.nfloss + e.nlim : 0;
    int niter = 0;
    if (ne == 0) {
      if (live) for (int i = lane; i < nv; i += 32) e.qacc[i] = e.qaccs[i];
      __syncwarp();
    }
    {
      // residual jar = J x - aref in fp64 from the fp64 iterate; simple rows touch one dof, contact rows the dofs of their body's chain
      auto residual = [&](const double* x) {
        for (int r = lane; r < ne; r += 32) {
          double t = -(double)e.s.aref[r];
          if (r < nsimple) t += (double)e.rsgn[r] * x[e.rdof[r]];
          else {
            const int body = m.geom_body[e.cgeom[e.crow[r]]], nd = m.body_ndof[body];
#pragma unroll
            for (int q = 0; q < GM_MAXCHAIN; q++) if (q < nd) { const int i = m.body_dofs[body][q]; t += (double)e.J[r][i] * x[i]; }
          }
          e.jar[r] = t;
        }
      };
      // start from the cheaper of qacc_warmstart and qacc_smooth
      float cost2[2] = {0.f, 0.f};
      if (ne > 0) for (int s = 0; s < 2; s++) {
        const float* q = s == 0 ? e.warm : e.qaccs;
        for (int i = lane; i < nv; i += 32) { e.vec[i] = q[i] - e.qaccs[i]; e.xd[i] = (double)q[i]; }
        __syncwarp();
        float cg = 0.f;
        if (s == 0) for (int i = lane; i < nv; i += 32) {
          float t = 0.f;
#pragma unroll
          for (int k = 0; k < GM_MAXV; k++) if (k < nv) t = fmaf(e.M[i][k], e.vec[k], t);
          cg += 0.5f * e.vec[i] * t;
        }
        residual(e.xd);
        __syncwarp();
        cost2[s] = g_warpsum(cg) + g_rows(m, e, 1, 0.f, lane);
        __syncwarp();
      }
      if (ne > 0 && cost2[0] < cost2[1]) { for (int i = lane; i < nv; i += 32) e.xd[i] = (double)e.warm[i]; }
      __syncwarp();
      float gprev = 1e30f;
      bool active = ne > 0;                                          // (warp-uniform)
      // every trip evaluates forces and gradient at the current iterate FIRST, so whenever a warp stops iterating the forces in
      // e.s.force / e.vec belong to the returned qacc
      for (int it = 0;; it++) {
#ifdef GM_NOLOCKSTEP
        if (!active) break;                                          // (experiment: every warp on its own)
#else
        if (__syncthreads_and(!active)) break;                       // lockstep: all warps of the CTA start a Newton trip together
#endif
        if (active) do {
        niter = it;
        for (int i = lane; i < nv; i += 32) e.vec[i] = (float)(e.xd[i] - (double)e.qaccs[i]);
        residual(e.xd);
        __syncwarp();
        for (int i = lane; i < nv; i += 32) {
          float t = 0.f;
#pragma unroll
          for (int k = 0; k < GM_MAXV; k++) if (k < nv) t = fmaf(e.M[i][k], e.vec[k], t);
          e.Ma[i] = t;
        }
        g_rows(m, e, 2, 0.f, lane);
        __syncwarp();
        float gn = 0.f, gref = 0.f;
        for (int i = lane; i < nv; i += 32) {
          float jf = 0.f;
#pragma unroll 8
          for (int r = nsimple; r < ne; r++) jf = fmaf(e.J[r][i], e.s.force[r], jf);
          const int fr = m.dof_flossrow[i];
          if (fr >= 0) jf += e.s.force[fr];
          e.vec[i] = jf;
        }
        __syncwarp();
        for (int r = m.nfloss + lane; r < nsimple; r += 32) e.vec[e.rdof[r]] += e.rsgn[r] * e.s.force[r];     // limit rows: distinct dofs
        __syncwarp();
        for (int i = lane; i < nv; i += 32) {
          const float jf = e.vec[i], t = e.Ma[i] - jf;
          e.grad[i] = t; gn += t * t; gref += e.Ma[i] * e.Ma[i] + jf * jf;
        }
        gn = g_warpsum(gn); gref = g_warpsum(gref);
        // MuJoCo's test (scaled gradient below opt.tolerance); or the gradient is at the fp32 rounding of its two terms; or it is
        // small and has stopped shrinking (Newton's quadratic phase ended in rounding noise: more steps only wander)
        const bool quad = gn <= 1e-10f * gref;                     // Newton's quadratic phase: full steps, no line search
        if (m.solver_scale * sqrtf(gn) < m.tolerance || gn <= 1e-16f * gref || (quad && gn >= 0.25f * gprev) || it >= m.iterations) { active = false; break; }
        gprev = gn;
        // Hessian (lower triangle) H = M + sum_r Hd_r J_r J_r' + cone terms.  Simple rows add to the diagonal; a contact only
        // touches the dofs of its body's chain (9 of 18 for a leg): a bottom-zone contact adds its rows weighted by D, one on the
        // cone adds Dm mu^2 (J0 - b)(J0 - b)' + c (sum_k fri_k^2 J_k J_k' - b b').  One contact at a time, its chain pairs over the lanes.
        for (int i = lane; i < GM_MAXV * GM_MAXV / 4; i += 32) reinterpret_cast<float4*>(&e.s.H[0][0])[i] = reinterpret_cast<const float4*>(&e.M[0][0])[i];
        __syncwarp();
        for (int i = lane; i < nv; i += 32) { const int fr = m.dof_flossrow[i]; if (fr >= 0) e.s.H[i][i] += e.s.Hd[fr]; }
        __syncwarp();
        for (int r = m.nfloss + lane; r < nsimple; r += 32) e.s.H[e.rdof[r]][e.rdof[r]] += e.s.Hd[r];
        __syncwarp();
        for (int c = 0; c < e.ncon; c++) {
          const int zone = e.czone[c];
          if (zone == 0) continue;
          const int a = e.cadr[c], dim = e.cdim[c], body = m.geom_body[e.cgeom[c]], nd = m.body_ndof[body];
          const float* fri = m.geom_friction[e.cgeom[c]];
          const float w0 = e.cDm[c] * e.cmu[c] * e.cmu[c], cc = e.ccoef[c];
          for (int p = lane; p < nd * (nd + 1) / 2; p += 32) {
            const int i = m.body_dofs[body][g_tri_r[p]], k = m.body_dofs[body][g_tri_c[p]];
            float s = 0.f;
            if (zone == 1) {
#pragma unroll
              for (int j = 0; j < 6; j++) if (j < dim) s += e.s.D[a + j] * e.J[a + j][i] * e.J[a + j][k];
            }
            else {
              const float bi = e.cb[c][i], bk = e.cb[c][k];
              float t = 0.f;
#pragma unroll
              for (int j = 1; j < 6; j++) if (j < dim) t += fri[j - 1] * fri[j - 1] * e.J[a + j][i] * e.J[a + j][k];
              s = w0 * (e.J[a][i] - bi) * (e.J[a][k] - bk) + cc * (t - bi * bk);
            }
            e.s.H[i][k] += s;
          }
          __syncwarp();
        }
        for (int i = lane; i < nv; i += 32) e.dir[i] = -e.grad[i];
        __syncwarp();
        if (!g_solve_spd(e.s.H, e.dir, nv, lane)) { active = false; break; }
        if (quad) {
          // quadratic phase (relative gradient <= 1e-5): full Newton steps without a line search.  At a relative gradient
          // <= 1e-6 the step lands below the fp32 rounding of the gradient (1e-12 relative in exact arithmetic): it is taken and
          // the iteration ends without evaluating the rows again -- the integrator below works from qacc alone
          for (int i = lane; i < nv; i += 32) e.xd[i] += (double)e.dir[i];
          if (gn <= 1e-12f * gref) { niter = it + 1; active = false; }
          __syncwarp();
          break;
        }
        // line search on phi'(alpha) = a1 + alpha a2 - sum_r force_r(jar + alpha jv) jv_r (increasing in alpha), to MuJoCo's
        // relative tolerance class: |phi'| <= 0.1 |phi'(0)| (exactness buys nothing: the outer Newton iteration corrects it)
        for (int r = lane; r < ne; r += 32) {
          float t = 0.f;
          if (r < nsimple) t = e.rsgn[r] * e.dir[e.rdof[r]];
          else {
            const int body = m.geom_body[e.cgeom[e.crow[r]]], nd = m.body_ndof[body];
#pragma unroll
            for (int q = 0; q < GM_MAXCHAIN; q++) if (q < nd) { const int i = m.body_dofs[body][q]; t = fmaf(e.J[r][i], e.dir[i], t); }
          }
          e.s.jv[r] = t;
        }
        float a1 = 0.f, a2 = 0.f;
        for (int i = lane; i < nv; i += 32) {
          float t = 0.f;
#pragma unroll
          for (int k = 0; k < GM_MAXV; k++) if (k < nv) t = fmaf(e.M[i][k], e.dir[k], t);
          a1 += e.dir[i] * e.Ma[i]; a2 += e.dir[i] * t;
        }
        a1 = g_warpsum(a1); a2 = g_warpsum(a2);
        __syncwarp();
        auto dphi = [&](float alpha) -> float { return a1 + alpha * a2 - g_rows(m, e, 0, alpha, lane); };
        float lo = 0.f, hi = 1.f, alpha = 1.f;
        float flo = 0.f;                                             // phi'(0) = dir . grad
        for (int i = lane; i < nv; i += 32) flo += e.dir[i] * e.grad[i];
        flo = g_warpsum(flo);
        if (!(flo < 0.f)) { active = false; break; }
        const float tol = 0.1f * fabsf(flo);
        float fhi = dphi(hi);
        if (fabsf(fhi) > tol) {
          int guard = 0;
          while (fhi < 0.f && guard++ < 30) { lo = hi; flo = fhi; hi *= 2.f; fhi = dphi(hi); }
          alpha = hi;
          if (fhi > tol) {
            int side = 0;
            for (int k = 0; k < 12; k++) {                           // regula falsi with the Illinois modification
              alpha = (lo * fhi - hi * flo) / (fhi - flo);
              if (!(alpha > lo && alpha < hi)) alpha = 0.5f * (lo + hi);
              const float fa = dphi(alpha);
              if (fabsf(fa) <= tol) break;
              if (fa < 0.f) { lo = alpha; flo = fa; if (side == -1) fhi *= 0.5f; side = -1; }
              else { hi = alpha; fhi = fa; if (side == 1) flo *= 0.5f; side = 1; }
            }
          }
        }
        for (int i = lane; i < nv; i += 32) e.xd[i] += (double)alpha * (double)e.dir[i];
        __syncwarp();
        } while (0);
      }
      if (ne > 0) for (int i = lane; i < nv; i += 32) e.qacc[i] = (float)e.xd[i];
      __syncwarp();
    }
    if (live) {
    niter_last = niter;
    for (int i = lane; i < nv; i += 32) e.warm[i] = e.qacc[i];
    // ------------------------------------------------------------------ P11 Euler with implicit joint damping, position integration
    {
      bool any_damp = false;
      for (int i = lane; i < nv; i += 32) any_damp |= m.dof_damping[i] > 0.f;
      any_damp = __any_sync(0xffffffffu, any_damp) && m.eulerdamp;
      if (any_damp) {
        // (M + h D) a = qfrc_smooth + qfrc_constraint = M qacc at the solver's optimum  =>  a = qacc - h (M + h D)^-1 D qacc:
        // the correction is O(h), so its rounding does not matter, and no constraint force has to be formed
        for (int i = lane; i < GM_MAXV * GM_MAXV / 4; i += 32) reinterpret_cast<float4*>(&e.s.H[0][0])[i] = reinterpret_cast<const float4*>(&e.M[0][0])[i];
        __syncwarp();
        for (int i = lane; i < nv; i += 32) e.s.H[i][i] += h * m.dof_damping[i];
        for (int i = lane; i < nv; i += 32) e.vec[i] = m.dof_damping[i] * e.qacc[i];
        __syncwarp();
        g_solve_spd(e.s.H, e.vec, nv, lane);
        for (int i = lane; i < nv; i += 32) e.vec[i] = e.qacc[i] - h * e.vec[i];
        __syncwarp();
      } else {
        for (int i = lane; i < nv; i += 32) e.vec[i] = e.qacc[i];
        __syncwarp();
      }
      for (int i = lane; i < nv; i += 32) e.qvel[i] += h * e.vec[i];
      __syncwarp();
      if (lane == 0) {
        for (int k = 0; k < 3; k++) e.qpos[k] += h * e.qvel[k];
        float w[3] = {e.qvel[3], e.qvel[4], e.qvel[5]};
        const float wn = sqrtf(g_dot3(w, w));
        float quat[4] = {e.qpos[3], e.qpos[4], e.qpos[5], e.qpos[6]};
        g_normquat(quat);
        if (wn >= 1e-15f) {
          float s, c;
          sincosf(0.5f * h * wn, &s, &c);
          const float qr[4] = {c, w[0] / wn * s, w[1] / wn * s, w[2] / wn * s};
          float qn[4];
          g_mulquat(qn, quat, qr);
          for (int k = 0; k < 4; k++) quat[k] = qn[k];
        }
        for (int k = 0; k < 4; k++) e.qpos[3 + k] = quat[k];
      }
      for (int b = 2 + lane; b < nb; b += 32) e.qpos[m.body_qadr[b]] += h * e.qvel[m.body_dofadr[b]];
      __syncwarp();
    }
    }
  }
  if (!live) continue;
  for (int i = lane; i < nq; i += 32) A.qpos[(size_t)env * nq + i] = e.qpos[i];
  for (int i = lane; i < nv; i += 32) { A.qvel[(size_t)env * nv + i] = e.qvel[i]; A.warm[(size_t)env * nv + i] = e.warm[i]; }
  if (A.info != nullptr && lane == 0) {
    A.info[(size_t)env * 4] = e.ncon; A.info[(size_t)env * 4 + 1] = e.nefc; A.info[(size_t)env * 4 + 2] = niter_last; A.info[(size_t)env * 4 + 3] = e.overflow;
  }
  }
}

// ================================================================================================ host side
namespace {
struct Arr { const unsigned char* data = nullptr; int code = -1; long long count = 0; };
bool find(const std::vector<unsigned char>& raw, const char* name, Arr& out) {
  if (raw.size() < 8 || memcmp(raw.data(), "NMB1", 4) != 0) return false;
  unsigned cnt;
  memcpy(&cnt, raw.data() + 4, 4);
  size_t off = 8;
  for (unsigned i = 0; i < cnt; i++) {
    if (off + 80 > raw.size()) return false;
    const char* nm = reinterpret_cast<const char*>(raw.data() + off);
    unsigned code, nd;
    long long dims[4], nbytes;
    memcpy(&code, raw.data() + off + 32, 4);
    memcpy(&nd, raw.data() + off + 36, 4);
    memcpy(dims, raw.data() + off + 40, 32);
    memcpy(&nbytes, raw.data() + off + 72, 8);
    off += 80;
    if (off + (size_t)nbytes > raw.size()) return false;
    if (strncmp(nm, name, 32) == 0) {
      out.data = raw.data() + off; out.code = (int)code; out.count = 1;
      for (unsigned k = 0; k < nd; k++) out.count *= dims[k];
      return true;
    }
    off += (size_t)nbytes + (size_t)((8 - nbytes % 8) % 8);
  }
  return false;
}
void q2m(const double* q, double* mtx) {
  const double w = q[0], x = q[1], y = q[2], z = q[3];
  mtx[0] = w * w + x * x - y * y - z * z; mtx[1] = 2 * (x * y - w * z); mtx[2] = 2 * (x * z + w * y);
  mtx[3] = 2 * (x * y + w * z); mtx[4] = w * w - x * x + y * y - z * z; mtx[5] = 2 * (y * z - w * x);
  mtx[6] = 2 * (x * z - w * y); mtx[7] = 2 * (y * z + w * x); mtx[8] = w * w - x * x - y * y + z * z;
}
double clampd(double x) { return x < 0.0001 ? 0.0001 : (x > 0.9999 ? 0.9999 : x); }
void kb(double timestep, const double* solref, const double* solimp, double& K, double& B) {
  double tc = solref[0], dr = solref[1], dmax = clampd(solimp[1]);
  if (tc > 0) { if (tc < 2 * timestep) tc = 2 * timestep; K = 1.0 / (dmax * dmax * tc * tc * dr * dr); B = 2.0 / (dmax * tc); }
  else { K = -tc / (dmax * dmax); B = -dr / dmax; }
}
}  // namespace

struct nm_gen_model { GenModel host; std::vector<float> qpos0; };
struct nm_gen_batch { const nm_gen_model* model; GenModel* d_model; int* d_ovf; int n, device, sms; GenArgs args; int64_t launches; };
typedef EnvMemT<GM_MAXCON, GM_MAXROW> EnvMemSmall;
typedef EnvMemT<GM_BIGCON, GM_BIGROW> EnvMemBig;

extern "C" int nm_gen_model_from_buffer(const void* data, size_t nbytes, nm_gen_model** out) {
  if (!data || !out || nbytes < 8) return nm_fail(NM_ERR_ARG, "nm_gen_model_from_buffer: bad argument");
  std::vector<unsigned char> raw((const unsigned char*)data, (const unsigned char*)data + nbytes);
#define GET(var, name, type, code_)                                                                       \
  const type* var = nullptr;                                                                             \
  { Arr a_; if (!find(raw, name, a_) || a_.code != code_) return nm_fail(NM_ERR_FORMAT, std::string("nmb: missing array '") + name + "'"); var = (const type*)a_.data; }
  GET(sizes, "sizes", int, 2) GET(oi, "opt_int", int, 2) GET(orl, "opt_real", double, 0) GET(qpos0, "qpos0", double, 0)
  GET(body_parent, "body_parent", int, 2) GET(body_jntadr, "body_jntadr", int, 2) GET(body_jntnum, "body_jntnum", int, 2)
  GET(body_dofadr, "body_dofadr", int, 2) GET(body_dofnum, "body_dofnum", int, 2)
  GET(body_pos, "body_pos", double, 0) GET(body_quat, "body_quat", double, 0) GET(body_ipos, "body_ipos", double, 0) GET(body_iquat, "body_iquat", double, 0)
  GET(body_mass, "body_mass", double, 0) GET(body_inertia, "body_inertia", double, 0) GET(body_invweight0, "body_invweight0", double, 0)
  GET(jnt_type, "jnt_type", int, 2) GET(jnt_qposadr, "jnt_qposadr", int, 2) GET(jnt_dofadr, "jnt_dofadr", int, 2) GET(jnt_pos, "jnt_pos", double, 0)
  GET(jnt_axis, "jnt_axis", double, 0) GET(jnt_limited, "jnt_limited", int, 2) GET(jnt_range, "jnt_range", double, 0)
  GET(dof_body, "dof_body", int, 2) GET(dof_parent, "dof_parent", int, 2) GET(dof_damping, "dof_damping", double, 0)
  GET(dof_frictionloss, "dof_frictionloss", double, 0) GET(dof_armature, "dof_armature", double, 0) GET(dof_invweight0, "dof_invweight0", double, 0)
  GET(act_dof, "act_dof", int, 2) GET(act_gain, "act_gain", double, 0) GET(act_bias, "act_bias", double, 0) GET(act_gear, "act_gear", double, 0)
  GET(act_ctrlrange, "act_ctrlrange", double, 0) GET(act_ctrllimited, "act_ctrllimited", int, 2) GET(act_forcerange, "act_forcerange", double, 0)
  GET(act_forcelimited, "act_forcelimited", int, 2)
  GET(geom_type, "geom_type", int, 2) GET(geom_body, "geom_body", int, 2) GET(geom_condim, "geom_condim", int, 2) GET(geom_priority, "geom_priority", int, 2)
  GET(geom_plane, "geom_plane", int, 2) GET(geom_pos, "geom_pos", double, 0) GET(geom_quat, "geom_quat", double, 0) GET(geom_size, "geom_size", double, 0)
  GET(geom_friction, "geom_friction", double, 0) GET(geom_solref, "geom_solref", double, 0) GET(geom_solimp, "geom_solimp", double, 0)
  GET(geom_margin, "geom_margin", double, 0) GET(geom_gap, "geom_gap", double, 0)
#undef GET
  nm_gen_model* gm = new nm_gen_model();
  GenModel& M = gm->host;
  memset(&M, 0, sizeof(M));
  const int nq = sizes[0], nv = sizes[1], nu = sizes[2], nbody = sizes[3], njnt = sizes[4], ngeom_all = sizes[5];
  auto bail = [&](int code, const char* msg) { delete gm; return nm_fail(code, msg); };
  if (oi[1] != 2) return bail(NM_ERR_UNSUPPORTED, "generic step: this path implements solver=\"Newton\" (PGS models use nm_model_from_buffer / nm_step)");
  if (nbody > GM_MAXB || nv > GM_MAXV || nq > GM_MAXQ || nu > GM_MAXU) return bail(NM_ERR_UNSUPPORTED, "generic step: model too large (<= 15 bodies, 24 dofs, 18 actuators)");
  if (oi[0] != 0) return bail(NM_ERR_UNSUPPORTED, "generic step: only integrator=\"Euler\" is implemented");
  if (oi[2] != 1) return bail(NM_ERR_UNSUPPORTED, "generic step: only cone=\"elliptic\" is implemented with the Newton solver");
  if (nbody < 2 || body_parent[1] != 0 || body_jntnum[1] != 1 || jnt_type[body_jntadr[1]] != 0) return bail(NM_ERR_UNSUPPORTED, "generic step: body 1 must be a free-floating base");
  M.nq = nq; M.nv = nv; M.nu = nu; M.nbody = nbody;
  M.integrator = oi[0]; M.cone = oi[2]; M.iterations = oi[3]; M.eulerdamp = oi[5];
  M.timestep = (float)orl[0]; for (int k = 0; k < 3; k++) M.gravity[k] = (float)orl[1 + k];
  M.tolerance = (float)orl[4]; M.impratio = (float)orl[6];
  M.solver_scale = (float)(1.0 / (orl[7] * (nv > 1 ? nv : 1)));
  for (int i = 0; i < nq; i++) M.qpos0[i] = (float)qpos0[i];
  gm->qpos0.assign(M.qpos0, M.qpos0 + nq);
  M.body_dofadr[0] = -1; M.body_qadr[0] = -1;
  for (int b = 1; b < nbody; b++) {
    if (body_jntnum[b] != 1) return bail(NM_ERR_UNSUPPORTED, "generic step: every body needs exactly one joint");
    const int j = body_jntadr[b];
    if (b > 1 && jnt_type[j] != 3) return bail(NM_ERR_UNSUPPORTED, "generic step: joints other than the base's free joint must be hinges");
    if (b > 1 && body_parent[b] >= b) return bail(NM_ERR_UNSUPPORTED, "generic step: bodies must come after their parents");
    for (int k = 0; k < 3; k++) if (std::fabs(jnt_pos[3 * j + k]) > 1e-12) return bail(NM_ERR_UNSUPPORTED, "generic step: joint anchors must sit at the body origin");
    M.body_parent[b] = body_parent[b];
    M.body_depth[b] = M.body_depth[body_parent[b]] + 1;
    if (M.body_depth[b] > M.maxdepth) M.maxdepth = M.body_depth[b];
    M.body_dofadr[b] = jnt_dofadr[j]; M.body_qadr[b] = jnt_qposadr[j];
    for (int k = 0; k < 3; k++) { M.body_pos[b][k] = (float)body_pos[3 * b + k]; M.body_ipos[b][k] = (float)body_ipos[3 * b + k]; M.body_inertia[b][k] = (float)body_inertia[3 * b + k]; M.jnt_axis[b][k] = (float)jnt_axis[3 * j + k]; }
    for (int k = 0; k < 4; k++) { M.body_quat[b][k] = (float)body_quat[4 * b + k]; M.body_iquat[b][k] = (float)body_iquat[4 * b + k]; }
    M.body_mass[b] = (float)body_mass[b];
    M.body_invw[b][0] = (float)body_invweight0[2 * b]; M.body_invw[b][1] = (float)body_invweight0[2 * b + 1];
    M.jnt_limited[b] = b > 1 ? jnt_limited[j] : 0;
    M.jnt_range[b][0] = (float)jnt_range[2 * j]; M.jnt_range[b][1] = (float)jnt_range[2 * j + 1];
  }
  (void)body_dofnum; (void)njnt;
  int nfl = 0;
  for (int i = 0; i < nv; i++) {
    M.dof_body[i] = dof_body[i]; M.dof_parent[i] = dof_parent[i];
    M.dof_damping[i] = (float)dof_damping[i]; M.dof_floss[i] = (float)dof_frictionloss[i]; M.dof_armature[i] = (float)dof_armature[i];
    M.dof_invw[i] = (float)dof_invweight0[i];
    M.dof_flossrow[i] = dof_frictionloss[i] > 0 ? nfl++ : -1;
  }
  M.nfloss = nfl;
  for (int b = 1; b < nbody; b++) {                    // dofs that move body b: its own chain
    int mask = 0;
    for (int i = M.body_dofadr[b] + (b == 1 ? 5 : 0); i >= 0; i = dof_parent[i]) mask |= 1 << i;
    M.body_dofmask[b] = mask;
    int nd = 0;
    for (int i = 0; i < nv; i++) if ((mask >> i) & 1) { if (nd >= GM_MAXCHAIN) return bail(NM_ERR_UNSUPPORTED, "generic step: kinematic chains of at most 16 dofs"); M.body_dofs[b][nd++] = i; }
    M.body_ndof[b] = nd;
  }
  std::vector<int> used(nv, 0);
  for (int a = 0; a < nu; a++) {
    const int dof = act_dof[a];
    if (dof < 6 || dof >= nv || used[dof]++) return bail(NM_ERR_UNSUPPORTED, "generic step: one actuator per hinge dof");
    M.act_dof[a] = dof; M.act_qadr[a] = M.body_qadr[dof_body[dof]];
    M.act_gain0[a] = (float)act_gain[3 * a];
    for (int k = 0; k < 3; k++) M.act_bias[a][k] = (float)act_bias[3 * a + k];
    M.act_gear[a] = (float)act_gear[a];
    M.act_ctrllimited[a] = act_ctrllimited[a]; M.act_forcelimited[a] = act_forcelimited[a];
    for (int k = 0; k < 2; k++) { M.act_ctrlrange[a][k] = (float)act_ctrlrange[2 * a + k]; M.act_forcerange[a][k] = (float)act_forcerange[2 * a + k]; }
  }
  // plane and the geoms that collide with it
  int plane = -1;
  for (int g = 0; g < ngeom_all; g++) if (geom_type[g] == 0) { if (plane >= 0 || geom_body[g] != 0) return bail(NM_ERR_UNSUPPORTED, "generic step: exactly one world-fixed plane"); plane = g; }
  if (plane < 0) return bail(NM_ERR_UNSUPPORTED, "generic step: the model needs a ground plane");
  {
    double R[9];
    q2m(geom_quat + 4 * plane, R);
    double n[3] = {R[2], R[5], R[8]}, fr[9] = {n[0], n[1], n[2], 0, 0, 0, 0, 0, 0};
    if (n[1] < 0.5 && n[1] > -0.5) fr[4] = 1; else fr[5] = 1;
    const double dd = fr[0] * fr[3] + fr[1] * fr[4] + fr[2] * fr[5];
    for (int c = 0; c < 3; c++) fr[3 + c] -= dd * fr[c];
    const double nn = std::sqrt(fr[3] * fr[3] + fr[4] * fr[4] + fr[5] * fr[5]);
    for (int c = 0; c < 3; c++) fr[3 + c] /= nn;
    fr[6] = fr[1] * fr[5] - fr[2] * fr[4]; fr[7] = fr[2] * fr[3] - fr[0] * fr[5]; fr[8] = fr[0] * fr[4] - fr[1] * fr[3];
    for (int c = 0; c < 9; c++) M.plane_frame[c] = (float)fr[c];
    for (int c = 0; c < 3; c++) { M.plane_n[c] = (float)n[c]; M.plane_pos[c] = (float)geom_pos[3 * plane + c]; }
  }
  int ng = 0;
  for (int g = 0; g < ngeom_all; g++) {
    if (g == plane || geom_plane[g] < 0) continue;
    if (geom_type[g] != G_SPHERE && geom_type[g] != G_BOX && geom_type[g] != G_CYLINDER) return bail(NM_ERR_UNSUPPORTED, "generic step: sphere / box / cylinder geoms only");
    if (ng >= GM_MAXG) return bail(NM_ERR_UNSUPPORTED, "generic step: too many collision geoms");
    const int p = plane;
    double fr3[3], solref[2], solimp[5];
    int dim;
    if (geom_priority[g] == geom_priority[p]) {
      for (int k = 0; k < 3; k++) fr3[k] = std::fmax(geom_friction[3 * g + k], geom_friction[3 * p + k]);
      dim = geom_condim[g] > geom_condim[p] ? geom_condim[g] : geom_condim[p];
      for (int k = 0; k < 2; k++) solref[k] = 0.5 * (geom_solref[2 * g + k] + geom_solref[2 * p + k]);
      for (int k = 0; k < 5; k++) solimp[k] = 0.5 * (geom_solimp[5 * g + k] + geom_solimp[5 * p + k]);
    } else {
      const int w = geom_priority[g] > geom_priority[p] ? g : p;
      for (int k = 0; k < 3; k++) fr3[k] = geom_friction[3 * w + k];
      dim = geom_condim[w];
      for (int k = 0; k < 2; k++) solref[k] = geom_solref[2 * w + k];
      for (int k = 0; k < 5; k++) solimp[k] = geom_solimp[5 * w + k];
    }
    if (dim != 1 && dim != 3 && dim != 4 && dim != 6) return bail(NM_ERR_UNSUPPORTED, "generic step: condim must be 1, 3, 4 or 6");
    M.geom_type[ng] = geom_type[g]; M.geom_body[ng] = geom_body[g]; M.geom_dim[ng] = dim;
    double R[9];
    q2m(geom_quat + 4 * g, R);
    for (int k = 0; k < 9; k++) M.geom_mat[ng][k] = (float)R[k];
    for (int k = 0; k < 3; k++) { M.geom_pos[ng][k] = (float)geom_pos[3 * g + k]; M.geom_size[ng][k] = (float)geom_size[3 * g + k]; }
    M.geom_margin[ng] = (float)(std::fmax(geom_margin[g], geom_margin[p]) - std::fmax(geom_gap[g], geom_gap[p]));
    M.geom_friction[ng][0] = M.geom_friction[ng][1] = (float)fr3[0]; M.geom_friction[ng][2] = (float)fr3[1]; M.geom_friction[ng][3] = M.geom_friction[ng][4] = (float)fr3[2];
    double K, B;
    kb(orl[0], solref, solimp, K, B);
    M.geom_K[ng] = (float)K; M.geom_B[ng] = (float)B;
    M.geom_solimp[ng][0] = (float)clampd(solimp[0]); M.geom_solimp[ng][1] = (float)clampd(solimp[1]); M.geom_solimp[ng][2] = (float)solimp[2];
    M.geom_solimp[ng][3] = (float)clampd(solimp[3]); M.geom_solimp[ng][4] = (float)(solimp[4] < 1 ? 1 : solimp[4]);
    ng++;
  }
  M.ngeom = ng;
  {
    const double dref[2] = {0.02, 1.0}, dimp[5] = {0.9, 0.95, 0.001, 0.5, 2.0};
    double K, B;
    kb(orl[0], dref, dimp, K, B);
    M.lim_K = (float)K; M.lim_B = (float)B; M.lim_imp0 = (float)dimp[0];
    for (int k = 0; k < 5; k++) M.lim_solimp[k] = (float)dimp[k];
  }
  *out = gm;
  return NM_OK;
}

extern "C" void nm_gen_model_destroy(nm_gen_model* m) { delete m; }
extern "C" int nm_gen_model_size(const nm_gen_model* m, const char* what) {
  if (!m || !what) return -1;
  const GenModel& M = m->host;
  if (!strcmp(what, "nq")) return M.nq;
  if (!strcmp(what, "nv")) return M.nv;
  if (!strcmp(what, "nu")) return M.nu;
  if (!strcmp(what, "nbody")) return M.nbody;
  if (!strcmp(what, "ngeom")) return M.ngeom;
  return -1;
}
extern "C" double nm_gen_model_timestep(const nm_gen_model* m) { return m ? m->host.timestep : 0.0; }
extern "C" int nm_gen_model_qpos0(const nm_gen_model* m, float* out, int cap) {
  if (!m || !out) return -1;
  const int n = (int)m->qpos0.size();
  for (int i = 0; i < n && i < cap; i++) out[i] = m->qpos0[i];
  return n;
}

extern "C" int nm_gen_batch_create(const nm_gen_model* m, int num_envs, int device, float* qpos, float* qvel, float* warm, int32_t* info, nm_gen_batch** out) {
  if (!m || !out || num_envs <= 0 || !qpos || !qvel || !warm) return nm_fail(NM_ERR_ARG, "nm_gen_batch_create: bad argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return nm_fail(NM_ERR_CUDA, "no such CUDA device");
  nm_gen_batch* b = new nm_gen_batch();
  memset(b, 0, sizeof(*b));
  b->model = m; b->n = num_envs; b->device = device;
  cudaDeviceProp prop;
  if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess ||
      cudaMalloc(&b->d_model, sizeof(GenModel)) != cudaSuccess ||
      cudaMalloc(&b->d_ovf, sizeof(int) * (1 + 2 * (size_t)num_envs)) != cudaSuccess ||
      cudaMemcpy(b->d_model, &m->host, sizeof(GenModel), cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaFuncSetAttribute(nm_generic_step_kernel<GM_MAXCON, GM_MAXROW, GM_WARPS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(EnvMemSmall) * GM_WARPS)) != cudaSuccess ||
      cudaFuncSetAttribute(nm_generic_step_kernel<GM_BIGCON, GM_BIGROW, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(EnvMemBig)) != cudaSuccess) {
    if (b->d_model) cudaFree(b->d_model);
    if (b->d_ovf) cudaFree(b->d_ovf);
    delete b;
    return nm_fail(NM_ERR_CUDA, "nm_gen_batch_create: CUDA allocation failed");
  }
  b->sms = prop.multiProcessorCount;
  b->args.ovf = b->d_ovf;
  b->args.model = b->d_model; b->args.qpos = qpos; b->args.qvel = qvel; b->args.warm = warm; b->args.info = info; b->args.num_envs = num_envs;
  *out = b;
  return NM_OK;
}
extern "C" void nm_gen_batch_destroy(nm_gen_batch* b) {
  if (!b) return;
  cudaFree(b->d_model);
  cudaFree(b->d_ovf);
  delete b;
}
extern "C" int nm_gen_physics_step(nm_gen_batch* b, const float* ctrl, int nstep, nm_stream stream) {
  if (!b || !ctrl || nstep < 1) return nm_fail(NM_ERR_ARG, "nm_gen_physics_step: bad argument");
  int prev = -1;
  cudaGetDevice(&prev);
  if (prev != b->device) cudaSetDevice(b->device);
  GenArgs a = b->args;
  a.ctrl = ctrl; a.nstep = nstep;
  const int blocks = (b->n + GM_WARPS - 1) / GM_WARPS;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaMemsetAsync(b->d_ovf, 0, sizeof(int), st);
  nm_generic_step_kernel<GM_MAXCON, GM_MAXROW, GM_WARPS, true><<<blocks, GM_WARPS * 32, sizeof(EnvMemSmall) * GM_WARPS, st>>>(a);
  // second tier: the few environments that need more contacts / rows than the first tier holds (usually none: ~3 us)
  const int big = b->n < 4 * b->sms ? b->n : 4 * b->sms;
  nm_generic_step_kernel<GM_BIGCON, GM_BIGROW, 1, false><<<big, 32, sizeof(EnvMemBig), st>>>(a);
  b->launches += 2;
  const cudaError_t err = cudaGetLastError();
  if (prev >= 0 && prev != b->device) cudaSetDevice(prev);
  if (err != cudaSuccess) return nm_fail(NM_ERR_CUDA, std::string("nm_gen_physics_step: ") + cudaGetErrorString(err));
  return NM_OK;
}
extern "C" int64_t nm_gen_batch_launches(const nm_gen_batch* b) { return b ? b->launches : 0; }
