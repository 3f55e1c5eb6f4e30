/* nm_oracle.c — CPU oracle, plain C, fp64.  TEST INFRASTRUCTURE ONLY (see nm_oracle.h).
 *
 * Restates, stage by stage, what the reference executes for one environment step:
 *   env layer : /root/reference/envs/nightmare_v3_env.py:145-371,399-497.  PINNED: the reference's own, unmodified
 *               NightmareV3Env code was run in the build container over a stand-in for its seven MuJoCo calls
 *               (tools/make_refenv_golden.py); this file reproduces its observations and rewards bit for bit and its
 *               flags / counters / commands exactly (tests/test_reference_env_golden.py).
 *   physics   : MuJoCo 3.1.2 mj_step (called at envs/nightmare_v3_env.py:200), published pipeline
 *               (SURVEY.md §3.2 and Appendix A).  MuJoCo is a third-party dependency that is absent
 *               from /root/reference and not installable here -> PARITY UNPINNED against real MuJoCo.
 * Written for clarity, not speed: dense matrices, generic kinematic tree (free/hinge/slide joints).
 */
#include "nm_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <tgmath.h>   /* sqrt/sin/cos/pow/exp/acos/fabs follow the type of `real` */
#undef I              /* complex.h's imaginary unit, pulled in by tgmath.h */

/* Arithmetic type of the restatement.  The oracle proper is fp64 (libnm_oracle.so).  The same source compiled with
 * -DNMO_REAL=float -fsingle-precision-constant (libnm_oracle_f32.so) is the "honest fp32 implementation of the same
 * pipeline": its deviation from the fp64 build on the lockstep suites is the arithmetic floor the fp32 CUDA kernels are
 * measured against (tools/fp32_floor.py, profiles/r02_fp32_floor.md).  The C ABI is double in both builds. */
#ifndef NMO_REAL
#define NMO_REAL double
#endif
typedef NMO_REAL real;

/* ---- floating-point operation counters (libnm_oracle_cnt.so, -DNMO_COUNT_FLOPS; compiled out otherwise).
 * Every add / subtract / multiply / divide / sqrt / sin / cos / exp / pow of the restated pipeline adds 1 to the counter of
 * the stage that is running (an FMA-able multiply-add therefore counts 2).  In the dense linear-algebra loops only
 * multiply-adds whose two operands are both non-zero are counted, so the figure is that of an implementation which, like
 * MuJoCo's tree-sparse L'DL and sparse Jacobians, skips structural zeros; the exhaustive support-vertex scan is counted as the
 * hill climb from vertex 0 that MuJoCo performs (same result).  bench.py reads these as the ALGORITHMIC flops per env-step.
 * Not thread safe: the counting build is driven with nthreads = 1. */
enum { ST_KIN = 0, ST_COMPOS, ST_CRB, ST_FACTOR, ST_COLLISION, ST_MAKECON, ST_PROJECT, ST_COMVEL_RNE, ST_ACTUATION, ST_WARMSTART,
       ST_PGS, ST_NOSLIP, ST_FINISH, ST_SENSOR, ST_INTEGRATE, ST_ENV, ST_COUNT };
#ifdef NMO_COUNT_FLOPS
static double g_flops[ST_COUNT];
static int g_stage = ST_ENV, g_mute = 0;
#define STAGE(s) (g_stage = (s))
#define MUTE(on) (g_mute = (on))
#define FLOP(n) (g_flops[g_stage] += g_mute ? 0.0 : (double)(n))
#define FLOP_NZ(a, b, n) do { if (!g_mute && (a) != 0 && (b) != 0) g_flops[g_stage] += (double)(n); } while (0)
void nmo_flop_counts(double* out, int cap) { for (int i = 0; i < ST_COUNT && i < cap; i++) out[i] = g_flops[i]; }
void nmo_flop_reset(void) { memset(g_flops, 0, sizeof(g_flops)); }
#else
#define STAGE(s) ((void)0)
#define MUTE(on) ((void)0)
#define FLOP(n) ((void)0)
#define FLOP_NZ(a, b, n) ((void)0)
#endif

#define MINVAL 1e-15
#define MAXVAL 1e10
#define TOLPLANEMESH 0.3 /* extra plane-mesh contacts must be this fraction of rbound apart (Appendix A.2) */

enum { JNT_FREE = 0, JNT_BALL = 1, JNT_SLIDE = 2, JNT_HINGE = 3 };
enum { GEOM_PLANE = 0, GEOM_SPHERE = 2, GEOM_CAPSULE = 3, GEOM_CYLINDER = 5, GEOM_BOX = 6, GEOM_MESH = 7 };
enum { SOL_PGS = 0, SOL_CG = 1, SOL_NEWTON = 2 };
enum { CONE_PYRAMIDAL = 0, CONE_ELLIPTIC = 1 };
enum { EFC_FRICTION = 0, EFC_LIMIT = 1, EFC_CONTACT_PYR = 2, EFC_CONTACT_ELL = 3 };
enum { INT_EULER = 0, INT_RK4 = 1, INT_IMPLICIT = 2, INT_IMPLICITFAST = 3 };

/* ------------------------------------------------------------------------------------------ model */
struct nmo_model {
  unsigned char* raw;
  void* owned[64]; int nowned;   /* model arrays converted to `real` */
  int nq, nv, nu, nbody, njnt, ngeom, nsite, nsensor, nhv, nhn;
  int integrator, solver, cone, iterations, noslip_iterations, eulerdamp;
  int planemesh_maxcon; /* contacts per plane-mesh pair (opt_int[7]; 4 when the model file predates the entry) */
  int mpr_iterations;   /* opt_int[8], MuJoCo default 50 */
  /* Details of MuJoCo recalled but not verifiable here, kept as MODEL DATA so that a cross-check against
   * MuJoCo corrects them by re-saving the model file (tools/mujoco_crosscheck.py names the option for each mismatch): */
  int planemesh_allverts;   /* opt_int[9]:  extra plane-mesh contacts drawn from 0 = the support vertex's hull-graph neighbours, 1 = all hull vertices */
  int planemesh_sepvert;    /* opt_int[10]: their minimum separation measured between 0 = contact points, 1 = hull vertices */
  int warm_after_noslip;    /* opt_int[11]: qacc_warmstart saved 0 = before the noslip pass, 1 = after it */
  real planemesh_sep;       /* opt_real[9]:  that separation as a fraction of rbound (0.3) */
  real pyramid_rfac;        /* opt_real[10]: R of a pyramidal contact's edges = this * mu_reg^2 * R[first] (2) */
  real mpr_tolerance;   /* opt_real[8], MuJoCo default 1e-6 */
  real timestep, gravity[3], tolerance, noslip_tolerance, impratio, meaninertia;
  const real *qpos0, *body_pos, *body_quat, *body_ipos, *body_iquat, *body_mass, *body_inertia, *body_invweight0;
  const int *body_parent, *body_rootid, *body_jntadr, *body_jntnum, *body_dofadr, *body_dofnum;
  const int *jnt_type, *jnt_body, *jnt_qposadr, *jnt_dofadr;
  const real *jnt_pos, *jnt_axis;
  const int *dof_body, *dof_jnt, *dof_parent;
  const real *dof_damping, *dof_armature, *dof_frictionloss, *dof_invweight0, *jnt_range;
  const int* jnt_limited;
  const int *act_dof, *act_ctrllimited, *act_forcelimited;
  const real *act_gain, *act_bias, *act_gear, *act_ctrlrange, *act_forcerange;
  const int *geom_type, *geom_body, *geom_condim, *geom_priority, *geom_plane, *geom_hull_adr, *geom_hull_num, *geom_contype, *geom_conaffinity;
  const real *geom_pos, *geom_quat, *geom_size, *geom_friction, *geom_solref, *geom_solimp, *geom_margin, *geom_gap, *geom_rbound, *geom_center;
  const float* hull_vert;
  const int *hull_nbr_adr, *hull_nbr;
  const int *site_body, *sensor_site;
  const real *site_pos, *site_size;
};

static const void* nmb_find(const unsigned char* raw, const char* name, int* code, long long* count) {
  unsigned cnt;
  memcpy(&cnt, raw + 4, 4);
  size_t off = 8;
  for (unsigned i = 0; i < cnt; i++) {
    const char* nm = (const char*)(raw + off);
    unsigned c, nd;
    long long dims[4], nbytes;
    memcpy(&c, raw + off + 32, 4);
    memcpy(&nd, raw + off + 36, 4);
    memcpy(dims, raw + off + 40, 32);
    memcpy(&nbytes, raw + off + 72, 8);
    off += 80;
    if (strncmp(nm, name, 32) == 0) {
      if (code) *code = (int)c;
      long long n = 1;
      for (unsigned k = 0; k < nd; k++) n *= dims[k];
      if (count) *count = n;
      return raw + off;
    }
    off += (size_t)nbytes + (size_t)((8 - nbytes % 8) % 8);
  }
  return NULL;
}

#define GETP(field, type)                                                     \
  do {                                                                        \
    m->field = (const type*)nmb_find(m->raw, #field, NULL, NULL);             \
    if (!m->field) {                                                          \
      snprintf(err, errlen, "nmb: missing array '%s'", #field);               \
      nmo_model_free(m);                                                      \
      return NULL;                                                            \
    }                                                                         \
  } while (0)
/* fp64 array of the model file -> array of `real` owned by the model */
#define GETR(field)                                                           \
  do {                                                                        \
    long long cnt_ = 0;                                                       \
    const double* src_ = (const double*)nmb_find(m->raw, #field, NULL, &cnt_); \
    if (!src_) {                                                              \
      snprintf(err, errlen, "nmb: missing array '%s'", #field);               \
      nmo_model_free(m);                                                      \
      return NULL;                                                            \
    }                                                                         \
    real* dst_ = (real*)malloc(sizeof(real) * (size_t)(cnt_ > 0 ? cnt_ : 1)); \
    for (long long k_ = 0; k_ < cnt_; k_++) dst_[k_] = (real)src_[k_];        \
    m->owned[m->nowned++] = dst_;                                             \
    m->field = dst_;                                                          \
  } while (0)

nmo_model* nmo_model_load(const char* path, char* err, int errlen) {
  char dummy[8];
  if (!err) { err = dummy; errlen = 8; }
  FILE* f = fopen(path, "rb");
  if (!f) { snprintf(err, errlen, "cannot open %s", path); return NULL; }
  fseek(f, 0, SEEK_END);
  long sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  nmo_model* m = (nmo_model*)calloc(1, sizeof(nmo_model));
  m->raw = (unsigned char*)malloc((size_t)sz + 8);
  if (fread(m->raw, 1, (size_t)sz, f) != (size_t)sz || memcmp(m->raw, "NMB1", 4) != 0) {
    fclose(f);
    snprintf(err, errlen, "%s is not an NMB1 file", path);
    nmo_model_free(m);
    return NULL;
  }
  fclose(f);
  const int* sizes = (const int*)nmb_find(m->raw, "sizes", NULL, NULL);
  long long noi = 0;
  const int* oi = (const int*)nmb_find(m->raw, "opt_int", NULL, &noi);
  long long norl = 0;
  const double* orl = (const double*)nmb_find(m->raw, "opt_real", NULL, &norl);
  if (!sizes || !oi || !orl) { snprintf(err, errlen, "nmb: missing header arrays"); nmo_model_free(m); return NULL; }
  m->nq = sizes[0]; m->nv = sizes[1]; m->nu = sizes[2]; m->nbody = sizes[3]; m->njnt = sizes[4];
  m->ngeom = sizes[5]; m->nsite = sizes[6]; m->nsensor = sizes[7]; m->nhv = sizes[8]; m->nhn = sizes[9];
  m->integrator = oi[0]; m->solver = oi[1]; m->cone = oi[2]; m->iterations = oi[3];
  m->noslip_iterations = oi[4]; m->eulerdamp = oi[5];
  m->planemesh_maxcon = (noi > 7 && oi[7] >= 1 && oi[7] <= 4) ? oi[7] : 4;
  m->mpr_iterations = (noi > 8 && oi[8] > 0) ? oi[8] : 50;
  m->planemesh_allverts = noi > 9 ? oi[9] != 0 : 0;
  m->planemesh_sepvert = noi > 10 ? oi[10] != 0 : 0;
  m->warm_after_noslip = noi > 11 ? oi[11] != 0 : 0;
  m->timestep = orl[0]; m->gravity[0] = orl[1]; m->gravity[1] = orl[2]; m->gravity[2] = orl[3];
  m->tolerance = orl[4]; m->noslip_tolerance = orl[5]; m->impratio = orl[6]; m->meaninertia = orl[7];
  m->mpr_tolerance = (norl > 8 && orl[8] > 0) ? (real)orl[8] : (real)1e-6;
  m->planemesh_sep = (norl > 9 && orl[9] > 0) ? (real)orl[9] : (real)TOLPLANEMESH;
  m->pyramid_rfac = (norl > 10 && orl[10] > 0) ? (real)orl[10] : (real)2;
  GETR(qpos0); GETR(body_pos); GETR(body_quat); GETR(body_ipos);
  GETR(body_iquat); GETR(body_mass); GETR(body_inertia); GETR(body_invweight0);
  GETP(body_parent, int); GETP(body_rootid, int); GETP(body_jntadr, int); GETP(body_jntnum, int);
  GETP(body_dofadr, int); GETP(body_dofnum, int);
  GETP(jnt_type, int); GETP(jnt_body, int); GETP(jnt_qposadr, int); GETP(jnt_dofadr, int);
  GETR(jnt_pos); GETR(jnt_axis);
  GETP(dof_body, int); GETP(dof_jnt, int); GETP(dof_parent, int); GETR(dof_damping); GETR(dof_armature);
  GETR(dof_frictionloss); GETR(dof_invweight0); GETR(jnt_range); GETP(jnt_limited, int);
  GETP(act_dof, int); GETP(act_ctrllimited, int); GETP(act_forcelimited, int);
  GETR(act_gain); GETR(act_bias); GETR(act_gear); GETR(act_ctrlrange); GETR(act_forcerange);
  GETP(geom_type, int); GETP(geom_body, int); GETP(geom_condim, int); GETP(geom_priority, int); GETP(geom_plane, int);
  GETP(geom_hull_adr, int); GETP(geom_hull_num, int); GETP(geom_contype, int); GETP(geom_conaffinity, int);
  GETR(geom_pos); GETR(geom_quat); GETR(geom_size); GETR(geom_friction);
  GETR(geom_solref); GETR(geom_solimp); GETR(geom_margin); GETR(geom_gap); GETR(geom_rbound); GETR(geom_center);
  GETP(hull_vert, float); GETP(hull_nbr_adr, int); GETP(hull_nbr, int);
  GETP(site_body, int); GETP(sensor_site, int); GETR(site_pos); GETR(site_size);
  return m;
}

void nmo_model_free(nmo_model* m) {
  if (!m) return;
  for (int i = 0; i < m->nowned; i++) free(m->owned[i]);
  free(m->raw);
  free(m);
}

int nmo_model_size(const nmo_model* m, const char* w) {
  if (!strcmp(w, "nq")) return m->nq;
  if (!strcmp(w, "nv")) return m->nv;
  if (!strcmp(w, "nu")) return m->nu;
  if (!strcmp(w, "nbody")) return m->nbody;
  if (!strcmp(w, "njnt")) return m->njnt;
  if (!strcmp(w, "ngeom")) return m->ngeom;
  if (!strcmp(w, "nsite")) return m->nsite;
  if (!strcmp(w, "nsensor")) return m->nsensor;
  return -1;
}

/* ------------------------------------------------------------------------------------------ small math */
static inline real dot3(const real* a, const real* b) { FLOP(5); return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline void cross3(real* r, const real* a, const real* b) {
  FLOP(9);
  real x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
static inline real normalize3(real* v) {
  FLOP(4);
  real n = sqrt(dot3(v, v));
  if (n < MINVAL) { v[0] = 1; v[1] = 0; v[2] = 0; return 0; }
  v[0] /= n; v[1] /= n; v[2] /= n;
  return n;
}
static inline void normalize4(real* q) {
  FLOP(12);
  real n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < MINVAL) { q[0] = 1; q[1] = q[2] = q[3] = 0; return; }
  q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n;
}
static inline void mul_quat(real* r, const real* a, const real* b) {
  FLOP(28);
  real w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  real x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  real y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  real z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  r[0] = w; r[1] = x; r[2] = y; r[3] = z;
}
static inline void quat2mat(real* m, const real* q) {
  FLOP(31);
  real w = q[0], x = q[1], y = q[2], z = q[3];
  m[0] = w * w + x * x - y * y - z * z; m[1] = 2 * (x * y - w * z); m[2] = 2 * (x * z + w * y);
  m[3] = 2 * (x * y + w * z); m[4] = w * w - x * x + y * y - z * z; m[5] = 2 * (y * z - w * x);
  m[6] = 2 * (x * z - w * y); m[7] = 2 * (y * z + w * x); m[8] = w * w - x * x - y * y + z * z;
}
static inline void mat_vec3(real* r, const real* m, const real* v) {
  FLOP(15);
  real x = m[0] * v[0] + m[1] * v[1] + m[2] * v[2];
  real y = m[3] * v[0] + m[4] * v[1] + m[5] * v[2];
  real z = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
static inline void axisangle2quat(real* q, const real* axis, real angle) {
  FLOP(7);
  real s = sin(0.5 * angle);
  q[0] = cos(0.5 * angle); q[1] = axis[0] * s; q[2] = axis[1] * s; q[3] = axis[2] * s;
}
/* rotate vector by quaternion (≙ mju_rotVecQuat, env.py:217-219) */
static inline void rot_vec_quat(real* r, const real* v, const real* q) {
  real m[9];
  quat2mat(m, q);
  mat_vec3(r, m, v);
}
/* spatial inertia (10 numbers: Ixx Iyy Izz Ixy Ixz Iyz, m*r(3), m) times motion vector [w; v] */
static void mul_inert_vec(real* res, const real* i, const real* v) {
  FLOP(42);
  res[0] = i[0] * v[0] + i[3] * v[1] + i[4] * v[2] - i[8] * v[4] + i[7] * v[5];
  res[1] = i[3] * v[0] + i[1] * v[1] + i[5] * v[2] + i[8] * v[3] - i[6] * v[5];
  res[2] = i[4] * v[0] + i[5] * v[1] + i[2] * v[2] - i[7] * v[3] + i[6] * v[4];
  res[3] = i[8] * v[1] - i[7] * v[2] + i[9] * v[3];
  res[4] = i[6] * v[2] - i[8] * v[0] + i[9] * v[4];
  res[5] = i[7] * v[0] - i[6] * v[1] + i[9] * v[5];
}
/* motion x motion */
static void cross_motion(real* r, const real* vel, const real* v) {
  FLOP(3);
  real a[3], b[3], c[3];
  cross3(a, vel, v);          /* w x v_ang */
  cross3(b, vel, v + 3);      /* w x v_lin */
  cross3(c, vel + 3, v);      /* vlin x v_ang */
  r[0] = a[0]; r[1] = a[1]; r[2] = a[2];
  r[3] = b[0] + c[0]; r[4] = b[1] + c[1]; r[5] = b[2] + c[2];
}
/* motion x* force */
static void cross_force(real* r, const real* vel, const real* f) {
  FLOP(3);
  real a[3], b[3], c[3];
  cross3(a, vel, f);          /* w x f_ang */
  cross3(b, vel + 3, f + 3);  /* v x f_lin */
  cross3(c, vel, f + 3);      /* w x f_lin */
  r[0] = a[0] + b[0]; r[1] = a[1] + b[1]; r[2] = a[2] + b[2];
  r[3] = c[0]; r[4] = c[1]; r[5] = c[2];
}

/* ------------------------------------------------------------------------------------------ per-env data */
typedef struct {
  real dist, pos[3], frame[9], mu, solref[2], solimp[5], margin;
  real friction[5];      /* tangent 1, tangent 2, torsional, rolling 1, rolling 2 (elliptic cones, condim up to 6) */
  int geom1, geom2, body1, body2, vert, efc_address, dim;
} contact_t;

typedef struct {
  /* state */
  real *qpos, *qvel, *qacc_warmstart, *ctrl, time;
  /* position-dependent */
  real *xpos, *xquat, *xmat, *xipos, *ximat, *xanchor, *xaxis, *site_xpos, *subtree_com, *cinert, *crb, *cdof;
  real *M, *L;               /* dense mass matrix and its Cholesky factor (lower) */
  real *MH, *LH;             /* M - h*qDeriv of the implicit velocity update and its factor */
  /* velocity-dependent */
  real *cvel, *cdof_dot, *qfrc_bias, *qfrc_passive, *qfrc_actuator, *qfrc_smooth, *qacc_smooth;
  real *qfrc_constraint, *qacc, *act_force;
  /* contacts / constraints */
  int ncon, nefc;
  contact_t con[NMO_MAXCON];
  real *efc_J, *efc_pos, *efc_margin, *efc_diagApprox, *efc_R, *efc_D, *efc_aref, *efc_vel, *efc_b, *efc_force, *efc_AR;
  real *efc_frictionloss, *efc_jar;
  int *efc_type, *efc_id;    /* row type (EFC_*) and the dof / joint / contact it belongs to */
  int solver_niter, noslip_niter, warm_used, nwarn;
  long nmpr, nmpr_hit, nsupp;   /* narrow-phase (MPR) calls, those that found a contact, support-function evaluations so far */
  real* sensordata;
  real* scratch;   /* >= 8*nv + MAXEFC*nv */
} data_t;

#define MAXEFC (6 * NMO_MAXCON + 64)

/* env layer carry state (envs/nightmare_v3_env.py:56-97) */
typedef struct {
  real actions[18], prev_actions[18], dof_pos[18], dof_vel[18], commands[3];
  real base_lin_vel[3], base_ang_vel[3], projected_gravity[3], base_height;
  real tibia_f[6], feet_f[6], body_f, dof_acc[18];
  real feet_air_time[6];
  int last_contacts[6], last_contacts_filt[6];
  real episode_sums[NMO_NREW], sums_at_reset[NMO_NREW];
  int64_t ep_len;
  int reset_buf, time_out;
} envstate_t;

/* nmo_envcfg (double, C ABI) converted to the arithmetic type */
typedef struct {
  int decimation, num_actions, tibia_contact_mode, body_contact_mode, add_noise, resample_period, strict_reference;
  real action_scale, clip_actions, p_gain, clip_obs;
  real default_pos[18];
  real obs_lin_vel, obs_ang_vel, obs_dof_pos, obs_dof_vel;
  real max_lin_vel_x, max_ang_vel, max_episode_length, max_episode_length_s;
  real termination_contact_force, tibia_max_contact_force, body_max_contact_force;
  real tracking_sigma, base_height_target, max_contact_force, dt;
  real rew_scale[NMO_NREW];
  real noise_vec[66];
} cfg_t;

static void cfg_convert(cfg_t* o, const nmo_envcfg* c) {
  memset(o, 0, sizeof(*o));
  o->decimation = c->decimation; o->num_actions = c->num_actions; o->tibia_contact_mode = c->tibia_contact_mode;
  o->body_contact_mode = c->body_contact_mode; o->add_noise = c->add_noise; o->resample_period = c->resample_period;
  o->strict_reference = c->strict_reference;
  o->action_scale = (real)c->action_scale; o->clip_actions = (real)c->clip_actions; o->p_gain = (real)c->p_gain; o->clip_obs = (real)c->clip_obs;
  for (int k = 0; k < 18; k++) o->default_pos[k] = (real)c->default_pos[k];
  o->obs_lin_vel = (real)c->obs_lin_vel; o->obs_ang_vel = (real)c->obs_ang_vel; o->obs_dof_pos = (real)c->obs_dof_pos; o->obs_dof_vel = (real)c->obs_dof_vel;
  o->max_lin_vel_x = (real)c->max_lin_vel_x; o->max_ang_vel = (real)c->max_ang_vel;
  o->max_episode_length = (real)c->max_episode_length; o->max_episode_length_s = (real)c->max_episode_length_s;
  o->termination_contact_force = (real)c->termination_contact_force; o->tibia_max_contact_force = (real)c->tibia_max_contact_force;
  o->body_max_contact_force = (real)c->body_max_contact_force;
  o->tracking_sigma = (real)c->tracking_sigma; o->base_height_target = (real)c->base_height_target; o->max_contact_force = (real)c->max_contact_force;
  o->dt = (real)c->dt;
  for (int k = 0; k < NMO_NREW; k++) o->rew_scale[k] = (real)c->rew_scale[k];
  for (int k = 0; k < 66; k++) o->noise_vec[k] = (real)c->noise_vec[k];
}

struct nmo_batch {
  const nmo_model* m;
  int n;
  uint64_t seed;
  cfg_t cfg;
  data_t* d;
  envstate_t* e;
  int64_t step_counter;
};

static real* dalloc(size_t n) { return (real*)calloc(n ? n : 1, sizeof(real)); }

static void data_init(const nmo_model* m, data_t* d) {
  int nv = m->nv, nb = m->nbody;
  memset(d, 0, sizeof(*d));
  d->qpos = dalloc(m->nq); d->qvel = dalloc(nv); d->qacc_warmstart = dalloc(nv); d->ctrl = dalloc(m->nu);
  d->xpos = dalloc(3 * nb); d->xquat = dalloc(4 * nb); d->xmat = dalloc(9 * nb); d->xipos = dalloc(3 * nb);
  d->ximat = dalloc(9 * nb); d->xanchor = dalloc(3 * m->njnt); d->xaxis = dalloc(3 * m->njnt);
  d->site_xpos = dalloc(3 * m->nsite); d->subtree_com = dalloc(3 * nb); d->cinert = dalloc(10 * nb);
  d->crb = dalloc(10 * nb); d->cdof = dalloc(6 * nv); d->M = dalloc(nv * nv); d->L = dalloc(nv * nv); d->MH = dalloc(nv * nv); d->LH = dalloc(nv * nv);
  d->cvel = dalloc(6 * nb); d->cdof_dot = dalloc(6 * nv); d->qfrc_bias = dalloc(nv); d->qfrc_passive = dalloc(nv);
  d->qfrc_actuator = dalloc(nv); d->qfrc_smooth = dalloc(nv); d->qacc_smooth = dalloc(nv);
  d->qfrc_constraint = dalloc(nv); d->qacc = dalloc(nv); d->act_force = dalloc(m->nu);
  d->efc_J = dalloc(MAXEFC * nv); d->efc_pos = dalloc(MAXEFC); d->efc_margin = dalloc(MAXEFC);
  d->efc_diagApprox = dalloc(MAXEFC); d->efc_R = dalloc(MAXEFC); d->efc_D = dalloc(MAXEFC);
  d->efc_aref = dalloc(MAXEFC); d->efc_vel = dalloc(MAXEFC); d->efc_b = dalloc(MAXEFC);
  d->efc_force = dalloc(MAXEFC); d->efc_AR = dalloc((size_t)MAXEFC * MAXEFC);
  d->efc_frictionloss = dalloc(MAXEFC); d->efc_jar = dalloc(MAXEFC);
  d->efc_type = (int*)calloc(MAXEFC, sizeof(int)); d->efc_id = (int*)calloc(MAXEFC, sizeof(int));
  d->sensordata = dalloc(m->nsensor);
  d->scratch = dalloc(16 * nv + (size_t)MAXEFC * nv + 16 * nb + 4 * (size_t)nv * nv + 8 * MAXEFC);
  memcpy(d->qpos, m->qpos0, sizeof(real) * m->nq);
}

static void data_free(data_t* d) {
  real** p[] = {&d->qpos, &d->qvel, &d->qacc_warmstart, &d->ctrl, &d->xpos, &d->xquat, &d->xmat, &d->xipos, &d->ximat,
                  &d->xanchor, &d->xaxis, &d->site_xpos, &d->subtree_com, &d->cinert, &d->crb, &d->cdof, &d->M, &d->L, &d->MH, &d->LH,
                  &d->cvel, &d->cdof_dot, &d->qfrc_bias, &d->qfrc_passive, &d->qfrc_actuator, &d->qfrc_smooth,
                  &d->qacc_smooth, &d->qfrc_constraint, &d->qacc, &d->act_force, &d->efc_J, &d->efc_pos, &d->efc_margin,
                  &d->efc_diagApprox, &d->efc_R, &d->efc_D, &d->efc_aref, &d->efc_vel, &d->efc_b, &d->efc_force,
                  &d->efc_AR, &d->sensordata, &d->scratch, &d->efc_frictionloss, &d->efc_jar};
  for (size_t i = 0; i < sizeof(p) / sizeof(p[0]); i++) free(*p[i]);
  free(d->efc_type); free(d->efc_id);
}

/* ------------------------------------------------------------------------------------------ P1 kinematics */
static void kinematics(const nmo_model* m, data_t* d) {
  STAGE(ST_KIN);
  real* xpos = d->xpos; real* xquat = d->xquat; real* xmat = d->xmat;
  xpos[0] = xpos[1] = xpos[2] = 0;
  xquat[0] = 1; xquat[1] = xquat[2] = xquat[3] = 0;
  quat2mat(xmat, xquat);
  memset(d->xipos, 0, 3 * sizeof(real));
  quat2mat(d->ximat, xquat);
  for (int i = 1; i < m->nbody; i++) {
    int pid = m->body_parent[i], ja = m->body_jntadr[i], jn = m->body_jntnum[i];
    real pos[3], quat[4];
    if (jn == 1 && m->jnt_type[ja] == JNT_FREE) {
      int qa = m->jnt_qposadr[ja];
      memcpy(pos, d->qpos + qa, 3 * sizeof(real));
      memcpy(quat, d->qpos + qa + 3, 4 * sizeof(real));
      normalize4(quat);
      memcpy(d->xanchor + 3 * ja, pos, 3 * sizeof(real));
      d->xaxis[3 * ja] = 0; d->xaxis[3 * ja + 1] = 0; d->xaxis[3 * ja + 2] = 1;
    } else {
      real t[3];
      mat_vec3(t, xmat + 9 * pid, m->body_pos + 3 * i);
      for (int k = 0; k < 3; k++) pos[k] = xpos[3 * pid + k] + t[k];
      FLOP(3);
      mul_quat(quat, xquat + 4 * pid, m->body_quat + 4 * i);
      for (int j = ja; j < ja + jn; j++) {
        real mat[9], anchor[3], axis[3];
        quat2mat(mat, quat);
        mat_vec3(anchor, mat, m->jnt_pos + 3 * j);
        for (int k = 0; k < 3; k++) anchor[k] += pos[k];
        FLOP(7);
        mat_vec3(axis, mat, m->jnt_axis + 3 * j);
        memcpy(d->xanchor + 3 * j, anchor, sizeof(anchor));
        memcpy(d->xaxis + 3 * j, axis, sizeof(axis));
        real q = d->qpos[m->jnt_qposadr[j]] - m->qpos0[m->jnt_qposadr[j]];
        if (m->jnt_type[j] == JNT_HINGE) {
          real qloc[4], qn[4], off[3];
          axisangle2quat(qloc, m->jnt_axis + 3 * j, q);
          mul_quat(qn, quat, qloc);
          memcpy(quat, qn, sizeof(qn));
          quat2mat(mat, quat);
          mat_vec3(off, mat, m->jnt_pos + 3 * j);      /* off-centre rotation correction */
          for (int k = 0; k < 3; k++) pos[k] = anchor[k] - off[k];
        } else if (m->jnt_type[j] == JNT_SLIDE) {
          for (int k = 0; k < 3; k++) pos[k] += axis[k] * q;
        }
      }
    }
    normalize4(quat);
    memcpy(xpos + 3 * i, pos, sizeof(pos));
    memcpy(xquat + 4 * i, quat, sizeof(quat));
    quat2mat(xmat + 9 * i, quat);
    real t[3], qi[4];
    mat_vec3(t, xmat + 9 * i, m->body_ipos + 3 * i);
    for (int k = 0; k < 3; k++) d->xipos[3 * i + k] = pos[k] + t[k];
    FLOP(3);
    mul_quat(qi, quat, m->body_iquat + 4 * i);
    quat2mat(d->ximat + 9 * i, qi);
  }
  for (int s = 0; s < m->nsite; s++) {
    int b = m->site_body[s];
    real t[3];
    mat_vec3(t, xmat + 9 * b, m->site_pos + 3 * s);
    for (int k = 0; k < 3; k++) d->site_xpos[3 * s + k] = xpos[3 * b + k] + t[k];
    FLOP(3);
  }
}

/* ------------------------------------------------------------------------------------------ P2 comPos */
static void com_pos(const nmo_model* m, data_t* d) {
  STAGE(ST_COMPOS);
  int nb = m->nbody;
  FLOP(3 * nb + 4 * (nb - 1) + 3 * nb);   /* mass-weighted positions, subtree accumulation, division */
  real* smass = d->scratch;
  for (int i = 0; i < nb; i++) {
    smass[i] = m->body_mass[i];
    for (int k = 0; k < 3; k++) d->subtree_com[3 * i + k] = m->body_mass[i] * d->xipos[3 * i + k];
  }
  for (int i = nb - 1; i > 0; i--) {
    int p = m->body_parent[i];
    smass[p] += smass[i];
    for (int k = 0; k < 3; k++) d->subtree_com[3 * p + k] += d->subtree_com[3 * i + k];
  }
  for (int i = 0; i < nb; i++) {
    if (smass[i] < MINVAL) memcpy(d->subtree_com + 3 * i, d->xipos + 3 * i, 3 * sizeof(real));
    else for (int k = 0; k < 3; k++) d->subtree_com[3 * i + k] /= smass[i];
  }
  for (int i = 1; i < nb; i++) {
    const real* R = d->ximat + 9 * i;
    const real* I = m->body_inertia + 3 * i;
    const real* c = d->subtree_com + 3 * m->body_rootid[i];
    real mass = m->body_mass[i], r[3];
    for (int k = 0; k < 3; k++) r[k] = d->xipos[3 * i + k] - c[k];
    real* ci = d->cinert + 10 * i;
    /* R diag(I) R^T */
    ci[0] = R[0] * R[0] * I[0] + R[1] * R[1] * I[1] + R[2] * R[2] * I[2];
    ci[1] = R[3] * R[3] * I[0] + R[4] * R[4] * I[1] + R[5] * R[5] * I[2];
    ci[2] = R[6] * R[6] * I[0] + R[7] * R[7] * I[1] + R[8] * R[8] * I[2];
    ci[3] = R[0] * R[3] * I[0] + R[1] * R[4] * I[1] + R[2] * R[5] * I[2];
    ci[4] = R[0] * R[6] * I[0] + R[1] * R[7] * I[1] + R[2] * R[8] * I[2];
    ci[5] = R[3] * R[6] * I[0] + R[4] * R[7] * I[1] + R[5] * R[8] * I[2];
    /* parallel axis */
    ci[0] += mass * (r[1] * r[1] + r[2] * r[2]);
    ci[1] += mass * (r[0] * r[0] + r[2] * r[2]);
    ci[2] += mass * (r[0] * r[0] + r[1] * r[1]);
    ci[3] -= mass * r[0] * r[1];
    ci[4] -= mass * r[0] * r[2];
    ci[5] -= mass * r[1] * r[2];
    ci[6] = mass * r[0]; ci[7] = mass * r[1]; ci[8] = mass * r[2];
    ci[9] = mass;
    FLOP(3 + 6 * 8 + 9 + 12 + 3);         /* r, R diag(I) R^T (6 entries), parallel-axis terms, m*r */
  }
  memset(d->cinert, 0, 10 * sizeof(real));
  for (int j = 0; j < m->njnt; j++) {
    int b = m->jnt_body[j], da = m->jnt_dofadr[j];
    const real* c = d->subtree_com + 3 * m->body_rootid[b];
    real off[3];
    for (int k = 0; k < 3; k++) off[k] = c[k] - d->xanchor[3 * j + k];
    FLOP(3);
    if (m->jnt_type[j] == JNT_FREE) {
      for (int k = 0; k < 3; k++) {
        real* cd = d->cdof + 6 * (da + k);
        memset(cd, 0, 6 * sizeof(real));
        cd[3 + k] = 1;
      }
      for (int k = 0; k < 3; k++) {
        real* cd = d->cdof + 6 * (da + 3 + k);
        real ax[3] = {d->xmat[9 * b + k], d->xmat[9 * b + 3 + k], d->xmat[9 * b + 6 + k]};
        memcpy(cd, ax, sizeof(ax));
        cross3(cd + 3, ax, off);
      }
    } else if (m->jnt_type[j] == JNT_HINGE) {
      real* cd = d->cdof + 6 * da;
      memcpy(cd, d->xaxis + 3 * j, 3 * sizeof(real));
      cross3(cd + 3, d->xaxis + 3 * j, off);
    } else { /* slide */
      real* cd = d->cdof + 6 * da;
      cd[0] = cd[1] = cd[2] = 0;
      memcpy(cd + 3, d->xaxis + 3 * j, 3 * sizeof(real));
    }
  }
}

/* ------------------------------------------------------------------------------------------ P3 CRBA + factor */
static int cholesky(real* L, const real* A, int n) {
  memcpy(L, A, sizeof(real) * n * n);
  for (int j = 0; j < n; j++) {
    real s = L[j * n + j];
    for (int k = 0; k < j; k++) { FLOP_NZ(L[j * n + k], L[j * n + k], 2); s -= L[j * n + k] * L[j * n + k]; }
    if (s < MINVAL) return -1;
    s = sqrt(s);
    FLOP(1);
    L[j * n + j] = s;
    for (int i = j + 1; i < n; i++) {
      real t = L[i * n + j];
      for (int k = 0; k < j; k++) { FLOP_NZ(L[i * n + k], L[j * n + k], 2); t -= L[i * n + k] * L[j * n + k]; }
      FLOP_NZ(t, 1, 1);
      L[i * n + j] = t / s;
    }
    for (int k = j + 1; k < n; k++) L[j * n + k] = 0;
  }
  return 0;
}
static void chol_solve(const real* L, int n, real* x) {
  for (int i = 0; i < n; i++) {
    real s = x[i];
    for (int k = 0; k < i; k++) { FLOP_NZ(L[i * n + k], x[k], 2); s -= L[i * n + k] * x[k]; }
    FLOP_NZ(s, 1, 1);
    x[i] = s / L[i * n + i];
  }
  for (int i = n - 1; i >= 0; i--) {
    real s = x[i];
    for (int k = i + 1; k < n; k++) { FLOP_NZ(L[k * n + i], x[k], 2); s -= L[k * n + i] * x[k]; }
    FLOP_NZ(s, 1, 1);
    x[i] = s / L[i * n + i];
  }
}

static void crb(const nmo_model* m, data_t* d) {
  STAGE(ST_CRB);
  int nv = m->nv, nb = m->nbody;
  memcpy(d->crb, d->cinert, sizeof(real) * 10 * nb);
  for (int i = nb - 1; i > 0; i--) {
    int p = m->body_parent[i];
    if (p > 0) { FLOP(10); for (int k = 0; k < 10; k++) d->crb[10 * p + k] += d->crb[10 * i + k]; }
  }
  memset(d->M, 0, sizeof(real) * nv * nv);
  for (int i = 0; i < nv; i++) {
    real buf[6];
    mul_inert_vec(buf, d->crb + 10 * m->dof_body[i], d->cdof + 6 * i);
    for (int j = i; j >= 0; j = m->dof_parent[j]) {
      real s = 0;
      for (int k = 0; k < 6; k++) { FLOP_NZ(d->cdof[6 * j + k], buf[k], 2); s += d->cdof[6 * j + k] * buf[k]; }
      d->M[i * nv + j] = d->M[j * nv + i] = s;
    }
    d->M[i * nv + i] += m->dof_armature[i];
  }
  STAGE(ST_FACTOR);
  if (cholesky(d->L, d->M, nv) != 0) d->nwarn++;
}

/* ------------------------------------------------------------------------------------------ P4 collision */
static void make_frame(real* f) {
  /* f[0:3] = normal; build tangents (≙ mju_makeFrame) */
  normalize3(f);
  real* y = f + 3;
  y[0] = 0; y[1] = 0; y[2] = 0;
  if (f[1] < 0.5 && f[1] > -0.5) y[1] = 1; else y[2] = 1;
  real dd = dot3(f, y);
  for (int k = 0; k < 3; k++) y[k] -= dd * f[k];
  normalize3(y);
  cross3(f + 6, f, y);
}

static void mix_params(const nmo_model* m, int g1, int g2, contact_t* c) {
  int p1 = m->geom_priority[g1], p2 = m->geom_priority[g2];
  const real *f1 = m->geom_friction + 3 * g1, *f2 = m->geom_friction + 3 * g2;
  {
    const real* fw = p1 == p2 ? NULL : (p1 > p2 ? f1 : f2);
    real fr[3];
    for (int k = 0; k < 3; k++) fr[k] = fw ? fw[k] : (f1[k] > f2[k] ? f1[k] : f2[k]);
    c->friction[0] = c->friction[1] = fr[0]; c->friction[2] = fr[1]; c->friction[3] = c->friction[4] = fr[2];
  }
  if (p1 == p2) {
    c->mu = f1[0] > f2[0] ? f1[0] : f2[0];
    c->dim = m->geom_condim[g1] > m->geom_condim[g2] ? m->geom_condim[g1] : m->geom_condim[g2];
    for (int k = 0; k < 2; k++) c->solref[k] = 0.5 * (m->geom_solref[2 * g1 + k] + m->geom_solref[2 * g2 + k]);
    for (int k = 0; k < 5; k++) c->solimp[k] = 0.5 * (m->geom_solimp[5 * g1 + k] + m->geom_solimp[5 * g2 + k]);
  } else {
    int g = p1 > p2 ? g1 : g2;
    c->mu = m->geom_friction[3 * g];
    c->dim = m->geom_condim[g];
    memcpy(c->solref, m->geom_solref + 2 * g, 2 * sizeof(real));
    memcpy(c->solimp, m->geom_solimp + 5 * g, 5 * sizeof(real));
  }
  c->margin = (m->geom_margin[g1] > m->geom_margin[g2] ? m->geom_margin[g1] : m->geom_margin[g2]) -
              (m->geom_gap[g1] > m->geom_gap[g2] ? m->geom_gap[g1] : m->geom_gap[g2]);
}


/* ------------------------------------------------------------------------------------------ P4b convex-convex narrow phase
 * MuJoCo 3.1.2 collides two convex meshes with libccd's Minkowski Portal Refinement (mjc_Convex -> ccdMPRPenetration,
 * libccd is vendored by MuJoCo, absent from /root/reference): centres = the geoms' frame origins (the mesh centre of
 * mass), support = hull vertex maximising the direction, tolerance opt.mpr_tolerance = 1e-6, at most opt.mpr_iterations = 50
 * refinement steps; one contact per pair (multiccd is off by default): dist = -depth, normal = direction from geom1 to geom2,
 * pos = midpoint of the two witness points.  Restated from libccd's published algorithm (src/mpr.c, vec3.c): portal
 * discovery, portal refinement, penetration from the final portal.  Reached from models/nightmare_v3/mjmodel.xml:47
 * (tibia contype=2 / conaffinity=3 => the 15 tibia-tibia pairs). */
#include <float.h>
#define CCD_EPS ((real)(sizeof(real) == 8 ? DBL_EPSILON : FLT_EPSILON))
typedef struct { real v[3], v1[3], v2[3]; } supp_t;                 /* point of the Minkowski difference obj1 - obj2 and its two witnesses */
typedef struct { const nmo_model* m; const real *R, *p; int adr, num; real center[3]; } hull_t;

static inline int ccd_zero(real x) { return fabs(x) < CCD_EPS; }
static inline int ccd_eq(real a_, real b_) {
  real ab = fabs(a_ - b_);
  if (ab < CCD_EPS) return 1;
  real a = fabs(a_), b = fabs(b_);
  return b > a ? ab < CCD_EPS * b : ab < CCD_EPS * a;
}
static inline int ccd_vec_eq(const real* a, const real* b) { return ccd_eq(a[0], b[0]) && ccd_eq(a[1], b[1]) && ccd_eq(a[2], b[2]); }
static inline void vsub(real* r, const real* a, const real* b) { FLOP(3); r[0] = a[0] - b[0]; r[1] = a[1] - b[1]; r[2] = a[2] - b[2]; }
static inline void ccd_normalize(real* v) { FLOP(9); real n = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); v[0] /= n; v[1] /= n; v[2] /= n; }

/* support point of a hull in world direction `dir`: hill climb on the hull's vertex graph from vertex 0, the way MuJoCo's
 * mesh support function walks mesh_graph (a local maximum of a linear function on a convex polytope is the global one;
 * neighbours are visited in list order, a neighbour replaces the current best only if strictly better) */
static long g_nsupp = 0;
static void hull_support(const hull_t* o, const real* dir, real* out) {
  const real* R = o->R;
  const nmo_model* m = o->m;
  g_nsupp++;
  real dl[3] = {R[0] * dir[0] + R[3] * dir[1] + R[6] * dir[2], R[1] * dir[0] + R[4] * dir[1] + R[7] * dir[2], R[2] * dir[0] + R[5] * dir[1] + R[8] * dir[2]};
  FLOP(15 + 5 + 18);
  int best = 0;
  const float* h = m->hull_vert + 3 * o->adr;
  real bv = dl[0] * (real)h[0] + dl[1] * (real)h[1] + dl[2] * (real)h[2];
  for (;;) {
    int nb = best;
    for (int e = m->hull_nbr_adr[o->adr + best]; e < m->hull_nbr_adr[o->adr + best + 1]; e++) {
      int v = m->hull_nbr[e];
      h = m->hull_vert + 3 * (o->adr + v);
      real val = dl[0] * (real)h[0] + dl[1] * (real)h[1] + dl[2] * (real)h[2];
      FLOP(5);
      if (val > bv) { bv = val; nb = v; }
    }
    if (nb == best) break;
    best = nb;
  }
  h = m->hull_vert + 3 * (o->adr + best);
  real lv[3] = {(real)h[0], (real)h[1], (real)h[2]};
  mat_vec3(out, R, lv);
  for (int k = 0; k < 3; k++) out[k] += o->p[k];
}
static void mpr_support(const hull_t* a, const hull_t* b, const real* dir, supp_t* s) {
  real nd[3] = {-dir[0], -dir[1], -dir[2]};
  hull_support(a, dir, s->v1);
  hull_support(b, nd, s->v2);
  vsub(s->v, s->v1, s->v2);
}
static void portal_dir(const supp_t* P, real* dir) {
  real a[3], b[3];
  vsub(a, P[2].v, P[1].v);
  vsub(b, P[3].v, P[1].v);
  cross3(dir, a, b);
  ccd_normalize(dir);
}
static int portal_reach_tolerance(const supp_t* P, const supp_t* v4, const real* dir, real tol) {
  real dv1 = dot3(P[1].v, dir), dv2 = dot3(P[2].v, dir), dv3 = dot3(P[3].v, dir), dv4 = dot3(v4->v, dir);
  real d1 = dv4 - dv1, d2 = dv4 - dv2, d3 = dv4 - dv3;
  if (d2 < d1) d1 = d2;
  if (d3 < d1) d1 = d3;
  return ccd_eq(d1, tol) || d1 < tol;
}
static void expand_portal(supp_t* P, const supp_t* v4) {
  real v4v0[3];
  cross3(v4v0, v4->v, P[0].v);
  if (dot3(P[1].v, v4v0) > 0) {
    if (dot3(P[2].v, v4v0) > 0) P[1] = *v4; else P[3] = *v4;
  } else {
    if (dot3(P[3].v, v4v0) > 0) P[2] = *v4; else P[1] = *v4;
  }
}
static real point_seg_dist2(const real* x0, const real* b, real* wit) {      /* P = origin */
  real d[3];
  vsub(d, b, x0);
  real t = -dot3(x0, d) / dot3(d, d);
  if (t < 0 || ccd_zero(t)) { memcpy(wit, x0, 3 * sizeof(real)); return dot3(x0, x0); }
  if (t > 1 || ccd_eq(t, 1)) { memcpy(wit, b, 3 * sizeof(real)); return dot3(b, b); }
  for (int k = 0; k < 3; k++) wit[k] = d[k] * t + x0[k];
  return dot3(wit, wit);
}
/* squared distance of the origin from the triangle (x0, B, C) and the closest point (≙ ccdVec3PointTriDist2) */
static real origin_tri_dist2(const real* x0, const real* B, const real* C, real* wit) {
  real d1[3], d2[3];
  vsub(d1, B, x0);
  vsub(d2, C, x0);
  real v = dot3(d1, d1), w = dot3(d2, d2), p = dot3(x0, d1), q = dot3(x0, d2), r = dot3(d1, d2);
  real den = w * v - r * r, s, t;
  if (ccd_zero(den)) { s = t = -1; }
  else { s = (q * r - w * p) / den; t = (-s * r - q) / w; }
  if ((ccd_zero(s) || s > 0) && (ccd_eq(s, 1) || s < 1) && (ccd_zero(t) || t > 0) && (ccd_eq(t, 1) || t < 1) && (ccd_eq(t + s, 1) || t + s < 1)) {
    for (int k = 0; k < 3; k++) wit[k] = x0[k] + s * d1[k] + t * d2[k];
    return dot3(wit, wit);
  }
  real w2[3];
  real dist = point_seg_dist2(x0, B, wit);
  real dd = point_seg_dist2(x0, C, w2);
  if (dd < dist) { dist = dd; memcpy(wit, w2, sizeof(w2)); }
  dd = point_seg_dist2(B, C, w2);
  if (dd < dist) { dist = dd; memcpy(wit, w2, sizeof(w2)); }
  return dist;
}
static void mpr_find_pos(const supp_t* P, real* pos) {
  real dir[3], vec[3], b[4];
  portal_dir(P, dir);
  cross3(vec, P[1].v, P[2].v); b[0] = dot3(vec, P[3].v);
  cross3(vec, P[3].v, P[2].v); b[1] = dot3(vec, P[0].v);
  cross3(vec, P[0].v, P[1].v); b[2] = dot3(vec, P[3].v);
  cross3(vec, P[2].v, P[1].v); b[3] = dot3(vec, P[0].v);
  real sum = b[0] + b[1] + b[2] + b[3];
  if (ccd_zero(sum) || sum < 0) {
    b[0] = 0;
    cross3(vec, P[2].v, P[3].v); b[1] = dot3(vec, dir);
    cross3(vec, P[3].v, P[1].v); b[2] = dot3(vec, dir);
    cross3(vec, P[1].v, P[2].v); b[3] = dot3(vec, dir);
    sum = b[1] + b[2] + b[3];
  }
  real inv = 1 / sum;
  for (int k = 0; k < 3; k++) {
    real p1 = 0, p2 = 0;
    for (int i = 0; i < 4; i++) { p1 += P[i].v1[k] * b[i]; p2 += P[i].v2[k] * b[i]; }
    pos[k] = (p1 * inv + p2 * inv) * (real)0.5;
  }
  FLOP(60);
}
/* returns 1 with depth / dir / pos when the hulls intersect, 0 otherwise (≙ ccdMPRPenetration == 0 and dir != 0) */
static int mpr_penetration(const nmo_model* m, const hull_t* A, const hull_t* B, real* depth, real* dir_out, real* pos) {
  supp_t P[4], v4;
  real dir[3], va[3], vb[3], dot;
  /* ---- portal discovery */
  memcpy(P[0].v1, A->center, 3 * sizeof(real));
  memcpy(P[0].v2, B->center, 3 * sizeof(real));
  vsub(P[0].v, P[0].v1, P[0].v2);
  real zero3[3] = {0, 0, 0};
  if (ccd_vec_eq(P[0].v, zero3)) P[0].v[0] += CCD_EPS * 10;
  for (int k = 0; k < 3; k++) dir[k] = -P[0].v[k];
  ccd_normalize(dir);
  mpr_support(A, B, dir, &P[1]);
  dot = dot3(P[1].v, dir);
  if (ccd_zero(dot) || dot < 0) return 0;
  cross3(dir, P[0].v, P[1].v);
  if (ccd_zero(dot3(dir, dir))) {
    if (ccd_vec_eq(P[1].v, zero3)) return 0;                       /* touching contact on v1: depth 0, direction undefined -> MuJoCo drops it */
    /* origin on the v0-v1 segment */
    for (int k = 0; k < 3; k++) pos[k] = (P[1].v1[k] + P[1].v2[k]) * (real)0.5;
    memcpy(dir_out, P[1].v, 3 * sizeof(real));
    *depth = sqrt(dot3(dir_out, dir_out));
    ccd_normalize(dir_out);
    return 1;
  }
  ccd_normalize(dir);
  mpr_support(A, B, dir, &P[2]);
  dot = dot3(P[2].v, dir);
  if (ccd_zero(dot) || dot < 0) return 0;
  vsub(va, P[1].v, P[0].v);
  vsub(vb, P[2].v, P[0].v);
  cross3(dir, va, vb);
  ccd_normalize(dir);
  if (dot3(dir, P[0].v) > 0) {                                     /* portal faces oriented away from the origin */
    supp_t t = P[1]; P[1] = P[2]; P[2] = t;
    for (int k = 0; k < 3; k++) dir[k] = -dir[k];
  }
  for (;;) {
    mpr_support(A, B, dir, &P[3]);
    dot = dot3(P[3].v, dir);
    if (ccd_zero(dot) || dot < 0) return 0;
    int cont = 0;
    cross3(va, P[1].v, P[3].v);
    dot = dot3(va, P[0].v);
    if (dot < 0 && !ccd_zero(dot)) { P[2] = P[3]; cont = 1; }       /* origin outside (v1, v0, v3) */
    if (!cont) {
      cross3(va, P[3].v, P[2].v);
      dot = dot3(va, P[0].v);
      if (dot < 0 && !ccd_zero(dot)) { P[1] = P[3]; cont = 1; }     /* origin outside (v3, v0, v2) */
    }
    if (!cont) break;
    vsub(va, P[1].v, P[0].v);
    vsub(vb, P[2].v, P[0].v);
    cross3(dir, va, vb);
    ccd_normalize(dir);
  }
  /* ---- portal refinement: does the portal enclose the origin? */
  for (;;) {
    portal_dir(P, dir);
    dot = dot3(dir, P[1].v);
    if (ccd_zero(dot) || dot > 0) break;                            /* origin inside the portal */
    mpr_support(A, B, dir, &v4);
    dot = dot3(v4.v, dir);
    if (!(ccd_zero(dot) || dot > 0) || portal_reach_tolerance(P, &v4, dir, m->mpr_tolerance)) return 0;
    expand_portal(P, &v4);
  }
  /* ---- penetration depth / direction / position from the refined portal */
  for (unsigned long it = 0;; it++) {
    portal_dir(P, dir);
    mpr_support(A, B, dir, &v4);
    if (portal_reach_tolerance(P, &v4, dir, m->mpr_tolerance) || it > (unsigned long)m->mpr_iterations) {
      real wit[3];
      *depth = sqrt(origin_tri_dist2(P[1].v, P[2].v, P[3].v, wit));
      if (ccd_zero(*depth)) return 0;                               /* touching: direction undefined -> no contact in MuJoCo */
      memcpy(dir_out, wit, sizeof(wit));
      ccd_normalize(dir_out);
      mpr_find_pos(P, pos);
      return 1;
    }
    expand_portal(P, &v4);
  }
}

#ifdef NMO_COUNT_FLOPS
/* operations of the hill climb from hull vertex 0 to the support vertex (what mjc_PlaneConvex does): one 5-flop dot
 * product per neighbour examined; rotating the direction into the geom frame costs 15 */
static void count_hill_climb(const nmo_model* m, int adr, int num, const real* R, const real* n) {
  real dl[3] = {R[0] * n[0] + R[3] * n[1] + R[6] * n[2], R[1] * n[0] + R[4] * n[1] + R[7] * n[2], R[2] * n[0] + R[5] * n[1] + R[8] * n[2]};
  FLOP(15 + 5);
  int best = 0;
  real bv = dl[0] * m->hull_vert[3 * adr] + dl[1] * m->hull_vert[3 * adr + 1] + dl[2] * m->hull_vert[3 * adr + 2];
  for (;;) {
    int nb = best;
    for (int e = m->hull_nbr_adr[adr + best]; e < m->hull_nbr_adr[adr + best + 1]; e++) {
      int v = m->hull_nbr[e];
      real val = dl[0] * m->hull_vert[3 * (adr + v)] + dl[1] * m->hull_vert[3 * (adr + v) + 1] + dl[2] * m->hull_vert[3 * (adr + v) + 2];
      FLOP(5);
      if (val < bv) { bv = val; nb = v; }
    }
    if (nb == best) break;
    best = nb;
  }
  (void)num;
}
#endif

static void collision(const nmo_model* m, data_t* d) {
  STAGE(ST_COLLISION);
  d->ncon = 0;
  for (int g = 0; g < m->ngeom; g++) {
    int pg = m->geom_plane[g];
    if (pg < 0) continue;
    int pb = m->geom_body[pg], b = m->geom_body[g];
    /* plane frame in world */
    real pq[4], pm[9], ppos[3], t[3];
    mul_quat(pq, d->xquat + 4 * pb, m->geom_quat + 4 * pg);
    quat2mat(pm, pq);
    mat_vec3(t, d->xmat + 9 * pb, m->geom_pos + 3 * pg);
    for (int k = 0; k < 3; k++) ppos[k] = d->xpos[3 * pb + k] + t[k];
    real n[3] = {pm[2], pm[5], pm[8]};
    real margin = m->geom_margin[g] > m->geom_margin[pg] ? m->geom_margin[g] : m->geom_margin[pg];
    const real* R = d->xmat + 9 * b;
    const real* p = d->xpos + 3 * b;
    if (m->geom_type[g] == GEOM_MESH) {
      int adr = m->geom_hull_adr[g], num = m->geom_hull_num[g];
      /* support vertex along -n: exhaustive argmin of n.(v - ppos), lowest index wins ties
         (MuJoCo hill-climbs the hull graph; identical on a convex hull except for exact ties) */
      int best = -1;
      real bestd = 0, bestw[3] = {0, 0, 0};
#ifdef NMO_COUNT_FLOPS
      count_hill_climb(m, adr, num, R, n);
#endif
      MUTE(1);                                   /* the exhaustive scan is counted as the hill climb above */
      for (int v = 0; v < num; v++) {
        real lv[3] = {m->hull_vert[3 * (adr + v)], m->hull_vert[3 * (adr + v) + 1], m->hull_vert[3 * (adr + v) + 2]};
        real w[3];
        mat_vec3(w, R, lv);
        for (int k = 0; k < 3; k++) w[k] += p[k];
        real dist = (w[0] - ppos[0]) * n[0] + (w[1] - ppos[1]) * n[1] + (w[2] - ppos[2]) * n[2];
        if (best < 0 || dist < bestd) { best = v; bestd = dist; memcpy(bestw, w, sizeof(w)); }
      }
      MUTE(0);
      FLOP(15 + 3 + 8);                          /* support vertex to world, distance to the plane */
      if (best < 0 || bestd > margin) continue;
      int first = d->ncon, cnt = 0;
      for (int pass = 0; pass < 2; pass++) {
        /* pass 0: the support vertex.  pass 1: its hull-graph neighbours, or every other hull vertex in index order (model option) */
        const int allv = pass == 1 && m->planemesh_allverts;
        int lo = pass == 0 || allv ? 0 : m->hull_nbr_adr[adr + best];
        int hi = pass == 0 ? 1 : allv ? num : m->hull_nbr_adr[adr + best + 1];
        for (int e = lo; e < hi && cnt < m->planemesh_maxcon && d->ncon < NMO_MAXCON; e++) {
          int v = pass == 0 ? best : allv ? e : m->hull_nbr[e];
          if (allv && v == best) continue;
          real w[3], dist;
          if (pass == 0) { memcpy(w, bestw, sizeof(w)); dist = bestd; }
          else {
            real lv[3] = {m->hull_vert[3 * (adr + v)], m->hull_vert[3 * (adr + v) + 1], m->hull_vert[3 * (adr + v) + 2]};
            mat_vec3(w, R, lv);
            for (int k = 0; k < 3; k++) w[k] += p[k];
            dist = (w[0] - ppos[0]) * n[0] + (w[1] - ppos[1]) * n[1] + (w[2] - ppos[2]) * n[2];
            FLOP(3 + 8);
            if (dist > margin) continue;
          }
          real cp[3];
          for (int k = 0; k < 3; k++) cp[k] = w[k] - 0.5 * dist * n[k];
          FLOP(7 + 9 * cnt);
          int tooclose = 0;
          for (int c = first; c < first + cnt; c++) {
            /* separation between the contact points, or between the hull vertices they came from (vertex = point + dist/2 n) */
            const real hv = m->planemesh_sepvert ? (real)0.5 : 0;
            real dx = (d->con[c].pos[0] + hv * d->con[c].dist * n[0]) - (cp[0] + hv * dist * n[0]);
            real dy = (d->con[c].pos[1] + hv * d->con[c].dist * n[1]) - (cp[1] + hv * dist * n[1]);
            real dz = (d->con[c].pos[2] + hv * d->con[c].dist * n[2]) - (cp[2] + hv * dist * n[2]);
            if (sqrt(dx * dx + dy * dy + dz * dz) < m->planemesh_sep * m->geom_rbound[g]) tooclose = 1;
          }
          if (tooclose) continue;
          contact_t* c = d->con + d->ncon++;
          cnt++;
          c->dist = dist;
          memcpy(c->pos, cp, sizeof(cp));
          memcpy(c->frame, n, sizeof(n));
          make_frame(c->frame);
          c->geom1 = pg; c->geom2 = g; c->body1 = pb; c->body2 = b; c->vert = v;
          mix_params(m, pg, g, c);
        }
      }
    } else if (m->geom_type[g] == GEOM_SPHERE) {
      real gc[3];
      mat_vec3(gc, R, m->geom_pos + 3 * g);
      for (int k = 0; k < 3; k++) gc[k] += p[k];
      real dist = (gc[0] - ppos[0]) * n[0] + (gc[1] - ppos[1]) * n[1] + (gc[2] - ppos[2]) * n[2] - m->geom_size[3 * g];
      if (dist > margin || d->ncon >= NMO_MAXCON) continue;
      contact_t* c = d->con + d->ncon++;
      c->dist = dist;
      for (int k = 0; k < 3; k++) c->pos[k] = gc[k] - n[k] * (m->geom_size[3 * g] + 0.5 * dist);
      memcpy(c->frame, n, sizeof(n));
      make_frame(c->frame);
      c->geom1 = pg; c->geom2 = g; c->body1 = pb; c->body2 = b; c->vert = -1;
      mix_params(m, pg, g, c);
    } else if (m->geom_type[g] == GEOM_BOX || m->geom_type[g] == GEOM_CYLINDER) {
      /* geom frame in world */
      real gq[4], gm[9], gc[3], cpos[4][3], cdist[4];
      int cnt = 0;
      mul_quat(gq, d->xquat + 4 * b, m->geom_quat + 4 * g);
      quat2mat(gm, gq);
      mat_vec3(gc, R, m->geom_pos + 3 * g);
      for (int k = 0; k < 3; k++) gc[k] += p[k];
      const real* size = m->geom_size + 3 * g;
      real dif[3];
      vsub(dif, gc, ppos);
      const real dist0 = dot3(dif, n);
      if (m->geom_type[g] == GEOM_BOX) {
        /* ≙ mjc_PlaneBox: the corners below the margin that point towards the plane, in corner order, at most 4 */
        for (int i = 0; i < 8 && cnt < 4; i++) {
          real vec[3] = {(i & 1) ? size[0] : -size[0], (i & 2) ? size[1] : -size[1], (i & 4) ? size[2] : -size[2]}, corner[3];
          mat_vec3(corner, gm, vec);
          const real ldist = dot3(n, corner);
          if (dist0 + ldist > margin || ldist > 0) continue;
          cdist[cnt] = dist0 + ldist;
          for (int k = 0; k < 3; k++) cpos[cnt][k] = corner[k] - n[k] * cdist[cnt] * (real)0.5 + gc[k];
          FLOP(8);
          cnt++;
        }
      } else {
        /* ≙ mjc_PlaneCylinder: two points on the rim line nearest the plane, then two more on the near disk */
        real axis[3] = {gm[2], gm[5], gm[8]}, vec[3];
        real prjaxis = dot3(n, axis);
        if (prjaxis > 0) { for (int k = 0; k < 3; k++) axis[k] = -axis[k]; prjaxis = -prjaxis; }
        for (int k = 0; k < 3; k++) vec[k] = axis[k] * prjaxis - n[k];
        const real len2 = dot3(vec, vec);
        if (len2 >= MINVAL * MINVAL) { const real scl = size[0] / sqrt(len2); for (int k = 0; k < 3; k++) vec[k] *= scl; }
        else { vec[0] = gm[0] * size[0]; vec[1] = gm[3] * size[0]; vec[2] = gm[6] * size[0]; }
        const real prjvec = dot3(vec, n);
        for (int k = 0; k < 3; k++) axis[k] *= size[1];
        prjaxis *= size[1];
        FLOP(30);
        if (dist0 + prjaxis + prjvec <= margin) {
          cdist[cnt] = dist0 + prjaxis + prjvec;
          for (int k = 0; k < 3; k++) cpos[cnt][k] = gc[k] + vec[k] + axis[k] - n[k] * cdist[cnt] * (real)0.5;
          cnt++;
          if (dist0 - prjaxis + prjvec <= margin) {
            cdist[cnt] = dist0 - prjaxis + prjvec;
            for (int k = 0; k < 3; k++) cpos[cnt][k] = gc[k] + vec[k] - axis[k] - n[k] * cdist[cnt] * (real)0.5;
            cnt++;
          }
          const real prjvec1 = -prjvec * (real)0.5;
          if (dist0 + prjaxis + prjvec1 <= margin) {
            real vec1[3];
            cross3(vec1, vec, axis);
            normalize3(vec1);
            const real sc = size[0] * sqrt((real)3.0) * (real)0.5;
            for (int k = 0; k < 3; k++) vec1[k] *= sc;
            for (int sgn = 0; sgn < 2; sgn++) {
              cdist[cnt] = dist0 + prjaxis + prjvec1;
              for (int k = 0; k < 3; k++)
                cpos[cnt][k] = gc[k] + (sgn ? -vec1[k] : vec1[k]) + axis[k] - vec[k] * (real)0.5 - n[k] * cdist[cnt] * (real)0.5;
              cnt++;
            }
            FLOP(40);
          }
        }
      }
      for (int q = 0; q < cnt && d->ncon < NMO_MAXCON; q++) {
        contact_t* c = d->con + d->ncon++;
        c->dist = cdist[q];
        memcpy(c->pos, cpos[q], 3 * sizeof(real));
        memcpy(c->frame, n, sizeof(n));
        make_frame(c->frame);
        c->geom1 = pg; c->geom2 = g; c->body1 = pb; c->body2 = b; c->vert = q;
        mix_params(m, pg, g, c);
      }
    }
  }
  /* convex-convex pairs, after all pairs with the world body: MuJoCo orders contacts by the pair's (body1, body2) signature */
  for (int g1 = 0; g1 < m->ngeom; g1++) {
    if (m->geom_type[g1] != GEOM_MESH || m->geom_hull_num[g1] <= 0) continue;
    for (int g2 = g1 + 1; g2 < m->ngeom; g2++) {
      if (m->geom_type[g2] != GEOM_MESH || m->geom_hull_num[g2] <= 0) continue;
      int b1 = m->geom_body[g1], b2 = m->geom_body[g2];
      if (b1 == b2 || b1 == 0 || b2 == 0) continue;
      if (!((m->geom_contype[g1] & m->geom_conaffinity[g2]) || (m->geom_contype[g2] & m->geom_conaffinity[g1]))) continue;
      if (m->body_parent[b1] == b2 || m->body_parent[b2] == b1) continue;          /* parent-child filter */
      hull_t A = {m, d->xmat + 9 * b1, d->xpos + 3 * b1, m->geom_hull_adr[g1], m->geom_hull_num[g1], {0, 0, 0}};
      hull_t B = {m, d->xmat + 9 * b2, d->xpos + 3 * b2, m->geom_hull_adr[g2], m->geom_hull_num[g2], {0, 0, 0}};
      mat_vec3(A.center, A.R, m->geom_center + 3 * g1);
      mat_vec3(B.center, B.R, m->geom_center + 3 * g2);
      for (int k = 0; k < 3; k++) { A.center[k] += A.p[k]; B.center[k] += B.p[k]; }
      real margin = m->geom_margin[g1] > m->geom_margin[g2] ? m->geom_margin[g1] : m->geom_margin[g2];
      real cc[3];
      vsub(cc, A.center, B.center);
      FLOP(6 + 8);
      if (sqrt(dot3(cc, cc)) > margin + m->geom_rbound[g1] + m->geom_rbound[g2]) continue;   /* bounding spheres */
      real depth, dir[3], pos[3];
      d->nmpr++;
      const long s0 = g_nsupp;
      const int hit = mpr_penetration(m, &A, &B, &depth, dir, pos);
      d->nsupp += g_nsupp - s0;
      if (!hit || d->ncon >= NMO_MAXCON) continue;
      d->nmpr_hit++;
      contact_t* c = d->con + d->ncon++;
      c->dist = margin - depth;
      memcpy(c->pos, pos, sizeof(pos));
      memcpy(c->frame, dir, sizeof(dir));
      make_frame(c->frame);
      c->geom1 = g1; c->geom2 = g2; c->body1 = b1; c->body2 = b2; c->vert = -1;
      mix_params(m, g1, g2, c);
    }
  }
}

/* ------------------------------------------------------------------------------------------ P5 constraints */
static void jac_point(const nmo_model* m, const data_t* d, int body, const real* point, real* jacp /*3 x nv*/) {
  int nv = m->nv;
  memset(jacp, 0, sizeof(real) * 3 * nv);
  if (body <= 0) return;
  real off[3];
  const real* c = d->subtree_com + 3 * m->body_rootid[body];
  for (int k = 0; k < 3; k++) off[k] = point[k] - c[k];
  int b = body;
  while (b > 0 && m->body_dofnum[b] == 0) b = m->body_parent[b];
  if (b <= 0) return;
  for (int i = m->body_dofadr[b] + m->body_dofnum[b] - 1; i >= 0; i = m->dof_parent[i]) {
    real t[3];
    cross3(t, d->cdof + 6 * i, off);
    FLOP(3);
    for (int k = 0; k < 3; k++) jacp[k * nv + i] = d->cdof[6 * i + 3 + k] + t[k];
  }
}

static real impedance(const real* solimp, real pos, real margin) {
  FLOP(8);
  real dmin = solimp[0], dmax = solimp[1], width = solimp[2], mid = solimp[3], power = solimp[4];
  if (dmin < 0.0001) dmin = 0.0001; if (dmin > 0.9999) dmin = 0.9999;
  if (dmax < 0.0001) dmax = 0.0001; if (dmax > 0.9999) dmax = 0.9999;
  if (mid < 0.0001) mid = 0.0001; if (mid > 0.9999) mid = 0.9999;
  if (power < 1) power = 1;
  if (dmin == dmax || width <= MINVAL) return 0.5 * (dmin + dmax);
  real x = fabs((pos - margin) / width);
  if (x >= 1) return dmax;
  if (x <= 0) return dmin;
  real y;
  if (power == 1) y = x;
  else if (x <= mid) y = pow(x, power) / pow(mid, power - 1);
  else y = 1 - pow(1 - x, power) / pow(1 - mid, power - 1);
  return dmin + y * (dmax - dmin);
}

/* rotational Jacobian of a body (3 x nv): column i = angular part of cdof i for the dofs of the body's chain */
static void jac_rot(const nmo_model* m, const data_t* d, int body, real* jacr) {
  int nv = m->nv;
  memset(jacr, 0, sizeof(real) * 3 * nv);
  int b = body;
  while (b > 0 && m->body_dofnum[b] == 0) b = m->body_parent[b];
  if (b <= 0) return;
  for (int i = m->body_dofadr[b] + m->body_dofnum[b] - 1; i >= 0; i = m->dof_parent[i])
    for (int k = 0; k < 3; k++) jacr[k * nv + i] = d->cdof[6 * i + k];
}

/* stiffness / damping of the reference acceleration from solref (≙ MuJoCo's getsolparam, with the refsafe clamp) */
static void sol_kb(const nmo_model* m, const real* solref, const real* solimp, real* K, real* B) {
  real tc = solref[0], dr = solref[1], dmax = solimp[1];
  if (dmax < 0.0001) dmax = 0.0001; if (dmax > 0.9999) dmax = 0.9999;
  if (tc > 0) {
    if (tc < 2 * m->timestep) tc = 2 * m->timestep;
    *K = 1.0 / (dmax * dmax * tc * tc * dr * dr);
    *B = 2.0 / (dmax * tc);
  } else {
    *K = -tc / (dmax * dmax);
    *B = -dr / dmax;
  }
}

static const real DEF_SOLREF[2] = {0.02, 1.0}, DEF_SOLIMP[5] = {0.9, 0.95, 0.001, 0.5, 2.0};   /* MuJoCo defaults (dof / joint solref, solimp) */

static void make_constraint(const nmo_model* m, data_t* d) {
  STAGE(ST_MAKECON);
  int nv = m->nv;
  d->nefc = 0;
  real* jac1 = d->scratch;            /* 3 x nv */
  real* jac2 = d->scratch + 3 * nv;   /* 3 x nv */
  /* ---- rows come in MuJoCo's order: dof friction loss, joint limits, contacts (mj_makeConstraint) */
  for (int i = 0; i < nv; i++) {
    if (!(m->dof_frictionloss[i] > 0) || d->nefc >= MAXEFC) continue;
    int e = d->nefc++;
    memset(d->efc_J + e * nv, 0, sizeof(real) * nv);
    d->efc_J[e * nv + i] = 1;
    d->efc_pos[e] = 0; d->efc_margin[e] = 0; d->efc_diagApprox[e] = m->dof_invweight0[i];
    d->efc_frictionloss[e] = m->dof_frictionloss[i];
    d->efc_type[e] = EFC_FRICTION; d->efc_id[e] = i;
  }
  for (int j = 0; j < m->njnt; j++) {
    if (!m->jnt_limited[j] || (m->jnt_type[j] != JNT_HINGE && m->jnt_type[j] != JNT_SLIDE)) continue;
    const real value = d->qpos[m->jnt_qposadr[j]];
    for (int side = -1; side <= 1; side += 2) {            /* lower limit first, then upper */
      const real dist = side * (m->jnt_range[2 * j + (side + 1) / 2] - value);
      if (!(dist < 0) || d->nefc >= MAXEFC) continue;      /* joint margin 0 */
      int e = d->nefc++;
      memset(d->efc_J + e * nv, 0, sizeof(real) * nv);
      d->efc_J[e * nv + m->jnt_dofadr[j]] = -side;
      d->efc_pos[e] = dist; d->efc_margin[e] = 0; d->efc_diagApprox[e] = m->dof_invweight0[m->jnt_dofadr[j]];
      d->efc_frictionloss[e] = 0;
      d->efc_type[e] = EFC_LIMIT; d->efc_id[e] = j;
    }
  }
  for (int e = 0; e < d->nefc; e++) {                      /* impedance, R, reference acceleration of the rows above */
    real K, B;
    sol_kb(m, DEF_SOLREF, DEF_SOLIMP, &K, &B);
    const real imp = impedance(DEF_SOLIMP, d->efc_pos[e], d->efc_margin[e]);
    real R = (1 - imp) * d->efc_diagApprox[e] / imp;
    if (R < MINVAL) R = MINVAL;
    d->efc_R[e] = R; d->efc_D[e] = 1 / R;
    real vel = 0;
    for (int i = 0; i < nv; i++) vel += d->efc_J[e * nv + i] * d->qvel[i];
    d->efc_vel[e] = vel;
    d->efc_aref[e] = -B * vel - K * imp * (d->efc_pos[e] - d->efc_margin[e]);
    FLOP(12);
  }
  if (m->cone == CONE_ELLIPTIC) {
    /* ---- elliptic cones: one row per contact dimension (normal, 2 tangents, torsion, 2 rolling), condim 1 / 3 / 4 / 6 */
    real* jr1 = d->scratch + 6 * nv;
    real* jr2 = d->scratch + 9 * nv;
    for (int ci = 0; ci < d->ncon; ci++) {
      contact_t* c = d->con + ci;
      c->efc_address = -1;
      const int dim = c->dim;
      if ((dim != 1 && dim != 3 && dim != 4 && dim != 6) || d->nefc + dim > MAXEFC) { d->nwarn++; continue; }
      const int e0 = c->efc_address = d->nefc;
      jac_point(m, d, c->body1, c->pos, jac1);
      jac_point(m, d, c->body2, c->pos, jac2);
      if (dim > 3) { jac_rot(m, d, c->body1, jr1); jac_rot(m, d, c->body2, jr2); }
      const real tran = m->body_invweight0[2 * c->body1] + m->body_invweight0[2 * c->body2];
      const real rot = m->body_invweight0[2 * c->body1 + 1] + m->body_invweight0[2 * c->body2 + 1];
      for (int r = 0; r < dim; r++) {
        int e = d->nefc++;
        const real* fr = c->frame + 3 * (r < 3 ? r : r - 3);
        const real *ja = r < 3 ? jac1 : jr1, *jb = r < 3 ? jac2 : jr2;
        for (int i = 0; i < nv; i++) {
          real s2 = 0;
          for (int k = 0; k < 3; k++) s2 += fr[k] * (jb[k * nv + i] - ja[k * nv + i]);
          d->efc_J[e * nv + i] = s2;
        }
        FLOP(6 * nv);
        d->efc_pos[e] = r == 0 ? c->dist : 0;
        d->efc_margin[e] = r == 0 ? c->margin : 0;
        d->efc_diagApprox[e] = r < 3 ? tran : rot;
        d->efc_frictionloss[e] = 0;
        d->efc_type[e] = EFC_CONTACT_ELL; d->efc_id[e] = ci;
      }
      real K, B;
      sol_kb(m, c->solref, c->solimp, &K, &B);
      const real imp = impedance(c->solimp, c->dist, c->margin);
      for (int r = 0; r < dim; r++) {
        int e = e0 + r;
        real vel = 0;
        for (int i = 0; i < nv; i++) vel += d->efc_J[e * nv + i] * d->qvel[i];
        d->efc_vel[e] = vel;
        d->efc_aref[e] = -B * vel - (r == 0 ? K * imp * (c->dist - c->margin) : 0);   /* friction dimensions have no position term */
        FLOP(2 * nv + 4);
      }
      real R0 = (1 - imp) * tran / imp;
      if (R0 < MINVAL) R0 = MINVAL;
      d->efc_R[e0] = R0;
      if (dim > 1) {
        /* friction rows: R = R_normal / impratio, scaled so that R_j mu_j^2 is the same for every friction dimension */
        const real R1 = R0 / (m->impratio > MINVAL ? m->impratio : MINVAL);
        d->efc_R[e0 + 1] = R1;
        for (int r = 2; r < dim; r++) d->efc_R[e0 + r] = R1 * c->friction[0] * c->friction[0] / (c->friction[r - 1] * c->friction[r - 1]);
        c->mu = c->friction[0] * sqrt(R1 / R0);             /* regularised friction coefficient of the cone */
      }
      for (int r = 0; r < dim; r++) { if (d->efc_R[e0 + r] < MINVAL) d->efc_R[e0 + r] = MINVAL; d->efc_D[e0 + r] = 1 / d->efc_R[e0 + r]; }
    }
    return;
  }
  for (int ci = 0; ci < d->ncon; ci++) {
    contact_t* c = d->con + ci;
    c->efc_address = -1;
    if (c->dim != 3 || d->nefc + 4 > MAXEFC) { d->nwarn++; continue; }   /* oracle scope: condim 3, pyramidal */
    c->efc_address = d->nefc;
    jac_point(m, d, c->body1, c->pos, jac1);
    jac_point(m, d, c->body2, c->pos, jac2);
    real Jc[3][64];   /* rows in contact frame; nv <= 64 */
    for (int r = 0; r < 3; r++)
      for (int i = 0; i < nv; i++) {
        real s = 0;
        for (int k = 0; k < 3; k++) { FLOP_NZ(jac2[k * nv + i] - jac1[k * nv + i], 1, 3); s += c->frame[3 * r + k] * (jac2[k * nv + i] - jac1[k * nv + i]); }
        Jc[r][i] = s;
      }
    real tran = m->body_invweight0[2 * c->body1] + m->body_invweight0[2 * c->body2];
    for (int r = 0; r < 4; r++) {
      int e = d->nefc++;
      int t = 1 + r / 2;
      real sgn = (r % 2 == 0) ? 1.0 : -1.0;
      for (int i = 0; i < nv; i++) { FLOP_NZ(Jc[0][i], 1, 2); d->efc_J[e * nv + i] = Jc[0][i] + sgn * c->mu * Jc[t][i]; }
      FLOP(3);
      d->efc_pos[e] = c->dist;
      d->efc_margin[e] = c->margin;
      d->efc_diagApprox[e] = tran + c->mu * c->mu * tran;
      d->efc_frictionloss[e] = 0;
      d->efc_type[e] = EFC_CONTACT_PYR; d->efc_id[e] = ci;
    }
  }
  /* impedance, R, D, reference acceleration */
  for (int ci = 0; ci < d->ncon; ci++) {
    contact_t* c = d->con + ci;
    if (c->efc_address < 0) continue;
    int e0 = c->efc_address;
    real tc = c->solref[0], dr = c->solref[1], dmax = c->solimp[1];
    if (dmax < 0.0001) dmax = 0.0001; if (dmax > 0.9999) dmax = 0.9999;
    real K, B;
    if (tc > 0) {
      if (tc < 2 * m->timestep) tc = 2 * m->timestep;   /* refsafe */
      K = 1.0 / (dmax * dmax * tc * tc * dr * dr);
      B = 2.0 / (dmax * tc);
    } else {
      K = -tc / (dmax * dmax);
      B = -dr / dmax;
    }
    for (int r = 0; r < 4; r++) {
      int e = e0 + r;
      real imp = impedance(c->solimp, d->efc_pos[e], d->efc_margin[e]);
      real R = (1 - imp) * d->efc_diagApprox[e] / imp;
      if (R < MINVAL) R = MINVAL;
      d->efc_R[e] = R;
      real vel = 0;
      FLOP(4 + 6);
      for (int i = 0; i < nv; i++) { FLOP_NZ(d->efc_J[e * nv + i], d->qvel[i], 2); vel += d->efc_J[e * nv + i] * d->qvel[i]; }
      d->efc_vel[e] = vel;
      d->efc_aref[e] = -B * vel - K * imp * (d->efc_pos[e] - d->efc_margin[e]);
    }
    /* pyramidal cone: all edges share R = 2 mu^2 R[first]  (mu regularised by impratio) */
    real mureg = c->mu / sqrt(m->impratio > MINVAL ? m->impratio : 1.0);
    real Rpy = m->pyramid_rfac * mureg * mureg * d->efc_R[e0];
    if (Rpy < MINVAL) Rpy = MINVAL;
    for (int r = 0; r < 4; r++) { d->efc_R[e0 + r] = Rpy; d->efc_D[e0 + r] = 1.0 / Rpy; }
  }
}

/* ------------------------------------------------------------------------------------------ P6 projectConstraint */
static void project_constraint(const nmo_model* m, data_t* d) {
  STAGE(ST_PROJECT);
  int nv = m->nv, ne = d->nefc;
  real* X = d->scratch + 16 * nv;  /* ne x nv : rows = M^-1 J_e^T */
  for (int e = 0; e < ne; e++) {
    memcpy(X + e * nv, d->efc_J + e * nv, sizeof(real) * nv);
    chol_solve(d->L, nv, X + e * nv);
  }
  for (int a = 0; a < ne; a++)
    for (int b = 0; b < ne; b++) {
      real s = 0;
      /* A is symmetric: an implementation computes the upper triangle once (count b >= a only) */
      for (int i = 0; i < nv; i++) { if (b >= a) FLOP_NZ(d->efc_J[a * nv + i], X[b * nv + i], 2); s += d->efc_J[a * nv + i] * X[b * nv + i]; }
      d->efc_AR[a * ne + b] = s + (a == b ? d->efc_R[a] : 0);
    }
}

/* ------------------------------------------------------------------------------------------ P7 comVel + rne */
static void com_vel(const nmo_model* m, data_t* d) {
  STAGE(ST_COMVEL_RNE);
  memset(d->cvel, 0, 6 * sizeof(real));
  for (int i = 1; i < m->nbody; i++) {
    real cvel[6];
    memcpy(cvel, d->cvel + 6 * m->body_parent[i], sizeof(cvel));
    int bda = m->body_dofadr[i];
    for (int j = m->body_jntadr[i]; j >= 0 && j < m->body_jntadr[i] + m->body_jntnum[i]; j++) {
      if (m->jnt_type[j] == JNT_FREE) {
        for (int k = 0; k < 3; k++) {
          memset(d->cdof_dot + 6 * (bda + k), 0, 6 * sizeof(real));
          for (int c = 0; c < 6; c++) { FLOP_NZ(d->cdof[6 * (bda + k) + c], d->qvel[bda + k], 2); cvel[c] += d->cdof[6 * (bda + k) + c] * d->qvel[bda + k]; }
        }
        bda += 3;
        for (int k = 0; k < 3; k++) cross_motion(d->cdof_dot + 6 * (bda + k), cvel, d->cdof + 6 * (bda + k));
        for (int k = 0; k < 3; k++)
          for (int c = 0; c < 6; c++) { FLOP_NZ(d->cdof[6 * (bda + k) + c], d->qvel[bda + k], 2); cvel[c] += d->cdof[6 * (bda + k) + c] * d->qvel[bda + k]; }
        bda += 3;
      } else {
        cross_motion(d->cdof_dot + 6 * bda, cvel, d->cdof + 6 * bda);
        FLOP(12);
        for (int c = 0; c < 6; c++) cvel[c] += d->cdof[6 * bda + c] * d->qvel[bda];
        bda++;
      }
    }
    memcpy(d->cvel + 6 * i, cvel, sizeof(cvel));
  }
}

static void rne(const nmo_model* m, data_t* d, real* result) {
  int nb = m->nbody, nv = m->nv;
  real* cacc = d->scratch;            /* 6 x nb */
  real* cfrc = d->scratch + 6 * nb;   /* 6 x nb */
  memset(cacc, 0, 6 * sizeof(real));
  for (int k = 0; k < 3; k++) cacc[3 + k] = -m->gravity[k];
  memset(cfrc, 0, 6 * sizeof(real));
  for (int i = 1; i < nb; i++) {
    real tmp[6], tmp1[6];
    memcpy(cacc + 6 * i, cacc + 6 * m->body_parent[i], 6 * sizeof(real));
    for (int j = 0; j < m->body_dofnum[i]; j++) {
      int dd = m->body_dofadr[i] + j;
      for (int c = 0; c < 6; c++) { FLOP_NZ(d->cdof_dot[6 * dd + c], d->qvel[dd], 2); cacc[6 * i + c] += d->cdof_dot[6 * dd + c] * d->qvel[dd]; }
    }
    mul_inert_vec(cfrc + 6 * i, d->cinert + 10 * i, cacc + 6 * i);
    mul_inert_vec(tmp, d->cinert + 10 * i, d->cvel + 6 * i);
    cross_force(tmp1, d->cvel + 6 * i, tmp);
    for (int c = 0; c < 6; c++) cfrc[6 * i + c] += tmp1[c];
    FLOP(6);
  }
  for (int i = nb - 1; i > 0; i--) {
    int p = m->body_parent[i];
    if (p > 0) { FLOP(6); for (int c = 0; c < 6; c++) cfrc[6 * p + c] += cfrc[6 * i + c]; }
  }
  for (int i = 0; i < nv; i++) {
    real s = 0;
    for (int c = 0; c < 6; c++) { FLOP_NZ(d->cdof[6 * i + c], 1, 2); s += d->cdof[6 * i + c] * cfrc[6 * m->dof_body[i] + c]; }
    result[i] = s;
  }
}

/* ------------------------------------------------------------------------------------------ P8 actuation / acceleration */
static void fwd_velocity_actuation_acceleration(const nmo_model* m, data_t* d) {
  int nv = m->nv;
  com_vel(m, d);
  for (int i = 0; i < nv; i++) d->qfrc_passive[i] = -m->dof_damping[i] * d->qvel[i];
  rne(m, d, d->qfrc_bias);
  STAGE(ST_ACTUATION);
  FLOP(nv);
  memset(d->qfrc_actuator, 0, sizeof(real) * nv);
  for (int a = 0; a < m->nu; a++) {
    real ctrl = d->ctrl[a];
    if (m->act_ctrllimited[a]) {
      if (ctrl < m->act_ctrlrange[2 * a]) ctrl = m->act_ctrlrange[2 * a];
      if (ctrl > m->act_ctrlrange[2 * a + 1]) ctrl = m->act_ctrlrange[2 * a + 1];
    }
    int dof = m->act_dof[a], jid = m->dof_jnt[dof];
    real gear = m->act_gear[a];
    real length = d->qpos[m->jnt_qposadr[jid]] * gear, velocity = d->qvel[dof] * gear;
    real f = m->act_gain[3 * a] * ctrl + m->act_bias[3 * a] + m->act_bias[3 * a + 1] * length + m->act_bias[3 * a + 2] * velocity;
    if (m->act_forcelimited[a]) {
      if (f < m->act_forcerange[2 * a]) f = m->act_forcerange[2 * a];
      if (f > m->act_forcerange[2 * a + 1]) f = m->act_forcerange[2 * a + 1];
    }
    d->act_force[a] = f;
    FLOP(10);
    d->qfrc_actuator[dof] += gear * f;
  }
  for (int i = 0; i < nv; i++) {
    d->qfrc_smooth[i] = d->qfrc_passive[i] - d->qfrc_bias[i] + d->qfrc_actuator[i];
    d->qacc_smooth[i] = d->qfrc_smooth[i];
  }
  FLOP(2 * nv);
  chol_solve(d->L, nv, d->qacc_smooth);
}

/* ------------------------------------------------------------------------------------------ P9 constraint solve */
static void dual_finish(const nmo_model* m, data_t* d) {
  STAGE(ST_FINISH);
  int nv = m->nv, ne = d->nefc;
  for (int i = 0; i < nv; i++) {
    real s = 0;
    for (int e = 0; e < ne; e++) { FLOP_NZ(d->efc_J[e * nv + i], d->efc_force[e], 2); s += d->efc_J[e * nv + i] * d->efc_force[e]; }
    d->qfrc_constraint[i] = s;
    d->qacc[i] = s;
  }
  chol_solve(d->L, nv, d->qacc);
  for (int i = 0; i < nv; i++) d->qacc[i] += d->qacc_smooth[i];
  FLOP(nv);
}

/* ------------------------------------------------------------------------------------------ P9b Newton solver (primal)
 * MuJoCo's default solver (models/anymal_c/anymal_c.xml leaves `solver` at Newton, :4 selects elliptic cones): minimise
 *     cost(qacc) = 1/2 (qacc - qacc_smooth)' M (qacc - qacc_smooth) + sum_i s_i(J qacc - aref)
 * a convex function with a unique minimiser.  s_i by row type (MuJoCo's constraint update):
 *   friction loss : quadratic 1/2 D jar^2 inside |jar| < R*floss, linear (force = -+floss) outside;
 *   limit / frictionless row : 1/2 D jar^2 for jar < 0, else 0;
 *   elliptic contact (dim rows): with N = mu*jar_0, T = |friction_j * jar_j|: 0 in the top zone (N >= mu T), 1/2 sum D_j jar_j^2
 *     in the bottom zone (mu N + T <= 0), 1/2 Dm (N - mu T)^2 with Dm = D_0 / (mu^2 (1 + mu^2)) in between.
 * MuJoCo iterates Newton steps with an approximate line search until the scaled improvement or gradient drops below
 * opt.tolerance (1e-8); this restatement uses exact Newton steps with an exact line search and iterates to the minimiser itself,
 * i.e. MuJoCo's result is reproduced up to ITS solver tolerance (parity is stated to that tolerance). */
static real newton_rows(const nmo_model* m, data_t* d, const real* jar, real* force, real* Hd, real* Hcone /* ncon x 36 */, int* cone_on) {
  real cost = 0;
  const int ne = d->nefc;
  if (Hd) for (int e = 0; e < ne; e++) Hd[e] = 0;
  for (int e = 0; e < ne; e++) {
    const real D = d->efc_D[e], R = d->efc_R[e];
    switch (d->efc_type[e]) {
      case EFC_FRICTION: {
        const real fl = d->efc_frictionloss[e], bound = R * fl;
        if (jar[e] <= -bound) { force[e] = fl; cost += -(real)0.5 * R * fl * fl - fl * jar[e]; }
        else if (jar[e] >= bound) { force[e] = -fl; cost += -(real)0.5 * R * fl * fl + fl * jar[e]; }
        else { force[e] = -D * jar[e]; cost += (real)0.5 * D * jar[e] * jar[e]; if (Hd) Hd[e] = D; }
        FLOP(8);
        break;
      }
      case EFC_LIMIT:
      case EFC_CONTACT_PYR:
        if (jar[e] < 0) { force[e] = -D * jar[e]; cost += (real)0.5 * D * jar[e] * jar[e]; if (Hd) Hd[e] = D; }
        else force[e] = 0;
        FLOP(4);
        break;
      case EFC_CONTACT_ELL: {
        const int ci = d->efc_id[e];
        const contact_t* c = d->con + ci;
        if (c->efc_address != e) break;                       /* handled at the contact's first row */
        const int dim = c->dim;
        const real mu = c->mu;
        real U[6];
        U[0] = jar[e] * mu;
        real T2 = 0;
        for (int j = 1; j < dim; j++) { U[j] = jar[e + j] * c->friction[j - 1]; T2 += U[j] * U[j]; }
        const real N = U[0], T = sqrt(T2);
        if (cone_on) cone_on[ci] = 0;
        FLOP(6 * dim);
        if (dim == 1) {                                       /* frictionless contact */
          if (jar[e] < 0) { force[e] = -D * jar[e]; cost += (real)0.5 * D * jar[e] * jar[e]; if (Hd) Hd[e] = D; } else force[e] = 0;
        } else if (N >= mu * T || (T <= 0 && N >= 0)) {       /* top zone: satisfied */
          for (int j = 0; j < dim; j++) force[e + j] = 0;
        } else if (mu * N + T <= 0 || (T <= 0 && N < 0)) {    /* bottom zone: quadratic in every dimension */
          for (int j = 0; j < dim; j++) {
            force[e + j] = -d->efc_D[e + j] * jar[e + j];
            cost += (real)0.5 * d->efc_D[e + j] * jar[e + j] * jar[e + j];
            if (Hd) Hd[e + j] = d->efc_D[e + j];
          }
        } else {                                              /* middle zone: on the cone */
          const real Dm = D / (mu * mu * (1 + mu * mu)), NmT = N - mu * T;
          cost += (real)0.5 * Dm * NmT * NmT;
          force[e] = -Dm * NmT * mu;
          for (int j = 1; j < dim; j++) force[e + j] = -force[e] / T * U[j] * c->friction[j - 1];
          if (Hcone) {
            /* Hessian of 1/2 Dm (N - mu T)^2 in jar: Dm (g g' - mu (N - mu T) d2T),  g = (mu, -mu f_j U_j / T),
             * d2T_jk = f_j f_k (delta_jk - U_j U_k / T^2) / T */
            real g[6];
            g[0] = mu;
            for (int j = 1; j < dim; j++) g[j] = -mu * c->friction[j - 1] * U[j] / T;
            real* H = Hcone + 36 * ci;
            for (int a = 0; a < dim; a++)
              for (int b = 0; b < dim; b++) {
                real h = g[a] * g[b];
                if (a > 0 && b > 0)
                  h += -mu * NmT * c->friction[a - 1] * c->friction[b - 1] * ((a == b ? 1 : 0) - U[a] * U[b] / T2) / T;
                H[6 * a + b] = Dm * h;
              }
            cone_on[ci] = 1;
            FLOP(12 * dim * dim);
          }
        }
        break;
      }
    }
  }
  return cost;
}

static void solve_newton(const nmo_model* m, data_t* d) {
  const int nv = m->nv, ne = d->nefc;
  real* work = d->scratch + 16 * nv + (size_t)MAXEFC * nv + 16 * m->nbody;      /* 4 nv^2 + 8 MAXEFC reals */
  real *H = work, *LH = work + nv * nv, *tmpM = work + 2 * nv * nv;
  real *jar = work + 4 * nv * nv, *jv = jar + MAXEFC, *Hd = jv + MAXEFC, *force = d->efc_force;
  real Hcone[NMO_MAXCON * 36];
  int cone_on[NMO_MAXCON];
  real qacc[64], grad[64], dir[64], dq[64], Ma[64], Md[64];
  (void)tmpM;
  const real scale = 1.0 / (m->meaninertia * (nv > 1 ? nv : 1));
  /* cost of a candidate start */
  #define TOTAL_COST(q_, out_)                                                                            \
    do {                                                                                                  \
      real c_ = 0;                                                                                        \
      for (int i_ = 0; i_ < nv; i_++) dq[i_] = (q_)[i_] - d->qacc_smooth[i_];                             \
      for (int i_ = 0; i_ < nv; i_++) { real s_ = 0; for (int k_ = 0; k_ < nv; k_++) s_ += d->M[i_ * nv + k_] * dq[k_]; c_ += (real)0.5 * dq[i_] * s_; } \
      for (int e_ = 0; e_ < ne; e_++) { real s_ = -d->efc_aref[e_]; for (int i_ = 0; i_ < nv; i_++) s_ += d->efc_J[e_ * nv + i_] * (q_)[i_]; jar[e_] = s_; } \
      c_ += newton_rows(m, d, jar, force, NULL, NULL, NULL);                                              \
      (out_) = c_;                                                                                        \
    } while (0)
  STAGE(ST_WARMSTART);
  real cw, cs;
  TOTAL_COST(d->qacc_warmstart, cw);
  TOTAL_COST(d->qacc_smooth, cs);
  FLOP(2 * (2 * nv * nv + 2 * ne * nv));
  if (cw < cs) { memcpy(qacc, d->qacc_warmstart, sizeof(real) * nv); d->warm_used = 1; }
  else memcpy(qacc, d->qacc_smooth, sizeof(real) * nv);
  STAGE(ST_PGS);
  for (int it = 0; it < m->iterations; it++) {
    /* gradient and Hessian at qacc */
    for (int i = 0; i < nv; i++) dq[i] = qacc[i] - d->qacc_smooth[i];
    for (int i = 0; i < nv; i++) { real s2 = 0; for (int k = 0; k < nv; k++) s2 += d->M[i * nv + k] * dq[k]; Ma[i] = s2; }
    for (int e = 0; e < ne; e++) { real s2 = -d->efc_aref[e]; for (int i = 0; i < nv; i++) s2 += d->efc_J[e * nv + i] * qacc[i]; jar[e] = s2; }
    newton_rows(m, d, jar, force, Hd, Hcone, cone_on);
    real gn = 0;
    for (int i = 0; i < nv; i++) {
      real s2 = Ma[i];
      for (int e = 0; e < ne; e++) s2 -= d->efc_J[e * nv + i] * force[e];
      grad[i] = s2; gn += s2 * s2;
    }
    FLOP(2 * nv * nv + 4 * ne * nv);
    d->solver_niter = it;
    if (getenv("NMO_NEWTON_TRACE")) fprintf(stderr, "it %d scaled grad %.3e\n", it, (double)(scale * sqrt(gn)));
    if (scale * sqrt(gn) < m->tolerance * (real)1e-3) break;            /* converged far below MuJoCo's own tolerance */
    memcpy(H, d->M, sizeof(real) * nv * nv);
    for (int e = 0; e < ne; e++) {
      if (Hd[e] == 0) continue;
      const real* J = d->efc_J + e * nv;
      for (int i = 0; i < nv; i++) { if (J[i] == 0) continue; const real t = Hd[e] * J[i]; for (int k = 0; k < nv; k++) H[i * nv + k] += t * J[k]; }
      FLOP(2 * nv * nv);
    }
    for (int ci = 0; ci < d->ncon; ci++) {
      const contact_t* c = d->con + ci;
      if (c->efc_address < 0 || d->efc_type[c->efc_address] != EFC_CONTACT_ELL || !cone_on[ci]) continue;
      const real* Jc = d->efc_J + c->efc_address * nv;
      for (int a = 0; a < c->dim; a++)
        for (int b = 0; b < c->dim; b++) {
          const real h = Hcone[36 * ci + 6 * a + b];
          if (h == 0) continue;
          for (int i = 0; i < nv; i++) { const real t = h * Jc[a * nv + i]; if (t == 0) continue; for (int k = 0; k < nv; k++) H[i * nv + k] += t * Jc[b * nv + k]; }
        }
      FLOP(2 * c->dim * c->dim * nv * nv);
    }
    if (cholesky(LH, H, nv) != 0) { d->nwarn++; break; }
    for (int i = 0; i < nv; i++) dir[i] = -grad[i];
    chol_solve(LH, nv, dir);
    /* exact line search: phi'(alpha) = dir' M (dq + alpha dir) - sum force_e(jar + alpha jv) jv_e is increasing in alpha */
    for (int e = 0; e < ne; e++) { real s2 = 0; for (int i = 0; i < nv; i++) s2 += d->efc_J[e * nv + i] * dir[i]; jv[e] = s2; }
    real a1 = 0, a2 = 0;
    for (int i = 0; i < nv; i++) { real s2 = 0; for (int k = 0; k < nv; k++) s2 += d->M[i * nv + k] * dir[k]; Md[i] = s2; a1 += dir[i] * Ma[i]; a2 += dir[i] * s2; }
    FLOP(2 * ne * nv + 2 * nv * nv);
    real* jt = Hd;                                                        /* reuse: trial jar */
    #define DPHI(alpha_, out_)                                                                            \
      do {                                                                                                \
        for (int e_ = 0; e_ < ne; e_++) jt[e_] = jar[e_] + (alpha_) * jv[e_];                            \
        newton_rows(m, d, jt, force, NULL, NULL, NULL);                                                   \
        real s_ = a1 + (alpha_) * a2;                                                                     \
        for (int e_ = 0; e_ < ne; e_++) s_ -= force[e_] * jv[e_];                                         \
        (out_) = s_;                                                                                      \
      } while (0)
    real lo = 0, hi = 1, flo, fhi, alpha = 1;
    DPHI(0, flo);
    if (!(flo < 0)) break;                                                /* no descent left: at the minimum to rounding */
    DPHI(hi, fhi);
    int guard = 0;
    while (fhi < 0 && guard++ < 40) { lo = hi; flo = fhi; hi *= 2; DPHI(hi, fhi); }
    if (fhi < 0) alpha = hi;
    else {
      /* regula falsi with the Illinois modification on the bracket [lo, hi] */
      int side = 0;
      alpha = hi;
      for (int k = 0; k < 60; k++) {
        alpha = (lo * fhi - hi * flo) / (fhi - flo);
        if (!(alpha > lo && alpha < hi)) alpha = (real)0.5 * (lo + hi);
        real fa;
        DPHI(alpha, fa);
        if (fabs(fa) <= (real)1e-13 * fabs(a1) || hi - lo <= (real)1e-15 * hi) break;
        if (fa < 0) { lo = alpha; flo = fa; if (side == -1) fhi *= (real)0.5; side = -1; }
        else { hi = alpha; fhi = fa; if (side == 1) flo *= (real)0.5; side = 1; }
      }
    }
    if (getenv("NMO_NEWTON_TRACE")) {
      fprintf(stderr, "   alpha %.4e lo %.3e hi %.3e a1 %.3e zones:", (double)alpha, (double)lo, (double)hi, (double)a1);
      for (int ci = 0; ci < d->ncon; ci++) fprintf(stderr, " %d", cone_on[ci]);
      int nq = 0; for (int e = 0; e < ne; e++) if (d->efc_type[e] == EFC_FRICTION && fabs(jar[e]) < d->efc_R[e] * d->efc_frictionloss[e]) nq++;
      fprintf(stderr, " fl-quadratic %d\n", nq);
    }
    for (int i = 0; i < nv; i++) qacc[i] += alpha * dir[i];
    #undef DPHI
  }
  #undef TOTAL_COST
  /* forces and accelerations at the solution */
  STAGE(ST_FINISH);
  for (int e = 0; e < ne; e++) { real s2 = -d->efc_aref[e]; for (int i = 0; i < nv; i++) s2 += d->efc_J[e * nv + i] * qacc[i]; jar[e] = s2; d->efc_jar[e] = s2; }
  newton_rows(m, d, jar, force, NULL, NULL, NULL);
  for (int i = 0; i < nv; i++) {
    real s2 = 0;
    for (int e = 0; e < ne; e++) s2 += d->efc_J[e * nv + i] * force[e];
    d->qfrc_constraint[i] = s2;
    d->qacc[i] = qacc[i];
  }
  FLOP(4 * ne * nv);
  memcpy(d->qacc_warmstart, d->qacc, sizeof(real) * nv);
}

static void fwd_constraint(const nmo_model* m, data_t* d) {
  int nv = m->nv, ne = d->nefc;
  d->solver_niter = d->noslip_niter = 0;
  d->warm_used = 0;
  if (ne == 0) {
    memcpy(d->qacc, d->qacc_smooth, sizeof(real) * nv);
    memcpy(d->qacc_warmstart, d->qacc_smooth, sizeof(real) * nv);
    memset(d->qfrc_constraint, 0, sizeof(real) * nv);
    return;
  }
  if (m->solver == SOL_NEWTON) { solve_newton(m, d); return; }
  const real* AR = d->efc_AR;
  real* f = d->efc_force;
  STAGE(ST_WARMSTART);
  /* b = J qacc_smooth - aref */
  for (int e = 0; e < ne; e++) {
    real s = 0;
    FLOP(1);
    for (int i = 0; i < nv; i++) { FLOP_NZ(d->efc_J[e * nv + i], 1, 2); s += d->efc_J[e * nv + i] * d->qacc_smooth[i]; }
    d->efc_b[e] = s - d->efc_aref[e];
  }
  /* warm start: forces implied by qacc_warmstart, kept only if their dual cost beats f = 0 */
  for (int e = 0; e < ne; e++) {
    real jar = -d->efc_aref[e];
    FLOP(2);
    for (int i = 0; i < nv; i++) { FLOP_NZ(d->efc_J[e * nv + i], d->qacc_warmstart[i], 2); jar += d->efc_J[e * nv + i] * d->qacc_warmstart[i]; }
    f[e] = jar < 0 ? -d->efc_D[e] * jar : 0;
  }
  real cost = 0;
  for (int a = 0; a < ne; a++) {
    real s = 0;
    for (int b = 0; b < ne; b++) { FLOP_NZ(AR[a * ne + b], f[b], 2); s += AR[a * ne + b] * f[b]; }
    FLOP(5);
    cost += 0.5 * f[a] * s + f[a] * d->efc_b[a];
  }
  if (cost > 0) memset(f, 0, sizeof(real) * ne); else d->warm_used = 1;

  real scale = 1.0 / (m->meaninertia * (nv > 1 ? nv : 1));
  /* PGS (pyramidal rows are scalar inequality constraints) */
  STAGE(ST_PGS);
  for (int it = 0; it < m->iterations; it++) {
    real improvement = 0;
    for (int e = 0; e < ne; e++) {
      real res = d->efc_b[e];
      FLOP(2 * ne + 10);                        /* residual of a dense row of A, update, cost change */
      for (int b = 0; b < ne; b++) res += AR[e * ne + b] * f[b];
      real old = f[e];
      f[e] -= res / AR[e * ne + e];
      if (f[e] < 0) f[e] = 0;
      real delta = f[e] - old;
      real change = 0.5 * delta * delta * AR[e * ne + e] + delta * res;
      if (change > 1e-10) { f[e] = old; change = 0; }
      improvement -= change;
    }
    d->solver_niter++;
    if (improvement * scale < m->tolerance) break;
  }
  dual_finish(m, d);
  memcpy(d->qacc_warmstart, d->qacc, sizeof(real) * nv);   /* saved BEFORE noslip */

  /* noslip post-processing: friction dimensions re-solved without regularisation */
  if (m->noslip_iterations > 0) {
    STAGE(ST_NOSLIP);
    for (int it = 0; it < m->noslip_iterations; it++) {
      real improvement = 0;
      for (int ci = 0; ci < d->ncon; ci++) {
        int e0 = d->con[ci].efc_address;
        if (e0 < 0) continue;
        for (int j = e0; j < e0 + 4; j += 2) {
          real res[2], old[2] = {f[j], f[j + 1]};
          FLOP(2 * (2 * ne + 2) + 40);            /* two dense-row residuals, the 2x2 pair problem, cost change */
          for (int k = 0; k < 2; k++) {
            real s = d->efc_b[j + k];
            for (int b = 0; b < ne; b++) s += AR[(j + k) * ne + b] * f[b];
            res[k] = s - d->efc_R[j + k] * f[j + k];
          }
          real Ac[4] = {AR[j * ne + j] - d->efc_R[j], AR[j * ne + j + 1], AR[(j + 1) * ne + j], AR[(j + 1) * ne + j + 1] - d->efc_R[j + 1]};
          real bc[2] = {res[0] - Ac[0] * old[0] - Ac[1] * old[1], res[1] - Ac[2] * old[0] - Ac[3] * old[1]};
          real mid = 0.5 * (old[0] + old[1]);
          real K1 = Ac[0] + Ac[3] - Ac[1] - Ac[2];
          real K0 = mid * (Ac[0] - Ac[3]) + bc[0] - bc[1];
          if (K1 < MINVAL) { f[j] = f[j + 1] = mid; }
          else {
            real x = -K0 / K1;
            if (x < -mid) { f[j] = 0; f[j + 1] = 2 * mid; }
            else if (x > mid) { f[j] = 2 * mid; f[j + 1] = 0; }
            else { f[j] = mid + x; f[j + 1] = mid - x; }
          }
          real dl[2] = {f[j] - old[0], f[j + 1] - old[1]};
          real change = 0.5 * (dl[0] * (Ac[0] * dl[0] + Ac[1] * dl[1]) + dl[1] * (Ac[2] * dl[0] + Ac[3] * dl[1])) + dl[0] * res[0] + dl[1] * res[1];
          if (change > 1e-10) { f[j] = old[0]; f[j + 1] = old[1]; change = 0; }
          improvement -= change;
        }
      }
      d->noslip_niter++;
      if (improvement * scale < m->noslip_tolerance) break;
    }
    dual_finish(m, d);
    if (m->warm_after_noslip) memcpy(d->qacc_warmstart, d->qacc, sizeof(real) * nv);   /* model option: saved AFTER noslip */
  }
}

/* ------------------------------------------------------------------------------------------ P10 touch sensors */
static real ray_sphere(const real* center, real radius, const real* pnt, const real* vec) {
  real dif[3] = {pnt[0] - center[0], pnt[1] - center[1], pnt[2] - center[2]};
  real a = dot3(vec, vec), b = dot3(vec, dif), c = dot3(dif, dif) - radius * radius;
  real det = b * b - a * c;
  if (det < MINVAL || a < MINVAL) return -1;
  det = sqrt(det);
  real x0 = (-b - det) / a, x1 = (-b + det) / a;
  if (x0 >= 0) return x0;
  if (x1 >= 0) return x1;
  return -1;
}

static void sensor_touch(const nmo_model* m, data_t* d) {
  STAGE(ST_SENSOR);
  for (int s = 0; s < m->nsensor; s++) {
    int site = m->sensor_site[s], body = m->site_body[site];
    real sum = 0;
    for (int ci = 0; ci < d->ncon; ci++) {
      const contact_t* c = d->con + ci;
      if (c->efc_address < 0 || (c->body1 != body && c->body2 != body)) continue;
      real fn = 0;
      if (d->efc_type[c->efc_address] == EFC_CONTACT_ELL) fn = d->efc_force[c->efc_address];     /* elliptic: the normal row */
      else for (int r = 0; r < 4; r++) fn += d->efc_force[c->efc_address + r];
      if (fn <= 0) continue;
      FLOP(3 + 3 + 25);
      real ray[3] = {c->frame[0] * fn, c->frame[1] * fn, c->frame[2] * fn};
      normalize3(ray);
      if (c->body2 == body) { ray[0] = -ray[0]; ray[1] = -ray[1]; ray[2] = -ray[2]; }
      if (ray_sphere(d->site_xpos + 3 * site, m->site_size[site], c->pos, ray) >= 0) sum += fn;
    }
    d->sensordata[s] = sum;
  }
}

/* ------------------------------------------------------------------------------------------ forward + integrate */
static void forward(const nmo_model* m, data_t* d) {
  kinematics(m, d);
  com_pos(m, d);
  crb(m, d);
  collision(m, d);
  make_constraint(m, d);
  if (m->solver != SOL_NEWTON) project_constraint(m, d);     /* the dual solvers need A = J M^-1 J' + R; Newton works on J itself */
  fwd_velocity_actuation_acceleration(m, d);
  fwd_constraint(m, d);
  sensor_touch(m, d);
}

static void integrate_pos(const nmo_model* m, real* qpos, const real* qvel, real h) {
  FLOP(2 * m->nv);
  for (int j = 0; j < m->njnt; j++) {
    int qa = m->jnt_qposadr[j], da = m->jnt_dofadr[j];
    if (m->jnt_type[j] == JNT_FREE) {
      for (int k = 0; k < 3; k++) qpos[qa + k] += h * qvel[da + k];
      real w[3] = {qvel[da + 3], qvel[da + 4], qvel[da + 5]}, qr[4], qn[4];
      real angle = h * normalize3(w);
      axisangle2quat(qr, w, angle);
      normalize4(qpos + qa + 3);
      mul_quat(qn, qpos + qa + 3, qr);
      memcpy(qpos + qa + 3, qn, sizeof(qn));
    } else {
      qpos[qa] += h * qvel[da];
    }
  }
}

static int bad_state(const nmo_model* m, const data_t* d) {
  for (int i = 0; i < m->nq; i++) if (!(fabs(d->qpos[i]) < MAXVAL)) return 1;
  for (int i = 0; i < m->nv; i++) if (!(fabs(d->qvel[i]) < MAXVAL)) return 1;
  return 0;
}

static void reset_data(const nmo_model* m, data_t* d) {
  memcpy(d->qpos, m->qpos0, sizeof(real) * m->nq);
  memset(d->qvel, 0, sizeof(real) * m->nv);
  memset(d->qacc_warmstart, 0, sizeof(real) * m->nv);
  d->time = 0;
  d->nwarn++;
}

static void step1(const nmo_model* m, data_t* d) {
  int nv = m->nv;
  real h = m->timestep;
  if (bad_state(m, d)) reset_data(m, d);      /* ≙ mj_checkPos / mj_checkVel */
  forward(m, d);
  for (int i = 0; i < nv; i++)
    if (!(fabs(d->qacc[i]) < MAXVAL)) { reset_data(m, d); forward(m, d); break; }   /* ≙ mj_checkAcc */
  real* qacc = d->scratch;
  STAGE(ST_INTEGRATE);
  FLOP(2 * nv + 2 * nv + 3 * m->nu);
  if (m->integrator == INT_IMPLICITFAST || m->integrator == INT_IMPLICIT) {
    /* implicitfast: qDeriv = d(qfrc_smooth)/d(qvel) restricted to actuator + passive terms (diagonal here) */
    real* A = d->MH;
    real* L = d->LH;
    memcpy(A, d->M, sizeof(real) * nv * nv);
    for (int i = 0; i < nv; i++) A[i * nv + i] += h * m->dof_damping[i];
    for (int a = 0; a < m->nu; a++) {
      int dof = m->act_dof[a];
      real g = m->act_gear[a];
      A[dof * nv + dof] -= h * m->act_bias[3 * a + 2] * g * g;
    }
    cholesky(L, A, nv);
    for (int i = 0; i < nv; i++) qacc[i] = d->qfrc_smooth[i] + d->qfrc_constraint[i];
    chol_solve(L, nv, qacc);
  } else { /* Euler (with implicit joint damping when eulerdamp is enabled) */
    int any = 0;
    for (int i = 0; i < nv; i++) if (m->dof_damping[i] > 0) any = 1;
    if (any && m->eulerdamp) {
      real* A = d->MH;
      real* L = d->LH;
      memcpy(A, d->M, sizeof(real) * nv * nv);
      for (int i = 0; i < nv; i++) A[i * nv + i] += h * m->dof_damping[i];
      cholesky(L, A, nv);
      for (int i = 0; i < nv; i++) qacc[i] = d->qfrc_smooth[i] + d->qfrc_constraint[i];
      chol_solve(L, nv, qacc);
    } else {
      memcpy(qacc, d->qacc, sizeof(real) * nv);
    }
  }
  for (int i = 0; i < nv; i++) d->qvel[i] += h * qacc[i];
  integrate_pos(m, d->qpos, d->qvel, h);
  d->time += h;
}

/* ------------------------------------------------------------------------------------------ Philox4x32-10 */
void nmo_philox4x32(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t out[4]) {
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
static inline real u01(uint32_t x) { return (real)(x >> 8) * (1.0 / 16777216.0); }

/* ------------------------------------------------------------------------------------------ batch */
nmo_batch* nmo_batch_create(const nmo_model* m, int n, uint64_t seed, const nmo_envcfg* cfg) {
  if (m->nv > 64) return NULL;
  nmo_batch* b = (nmo_batch*)calloc(1, sizeof(nmo_batch));
  b->m = m; b->n = n; b->seed = seed;
  if (cfg) cfg_convert(&b->cfg, cfg);
  b->d = (data_t*)calloc(n, sizeof(data_t));
  b->e = (envstate_t*)calloc(n, sizeof(envstate_t));
  for (int i = 0; i < n; i++) { data_init(m, b->d + i); b->e[i].reset_buf = 1; }
  return b;
}
void nmo_batch_free(nmo_batch* b) {
  if (!b) return;
  for (int i = 0; i < b->n; i++) data_free(b->d + i);
  free(b->d); free(b->e); free(b);
}
void nmo_set_state(nmo_batch* b, const double* qpos, const double* qvel, const double* warm) {
  const int nq = b->m->nq, nv = b->m->nv;
  for (int i = 0; i < b->n; i++) {
    if (qpos) for (int k = 0; k < nq; k++) b->d[i].qpos[k] = (real)qpos[(size_t)i * nq + k];
    if (qvel) for (int k = 0; k < nv; k++) b->d[i].qvel[k] = (real)qvel[(size_t)i * nv + k];
    if (warm) for (int k = 0; k < nv; k++) b->d[i].qacc_warmstart[k] = (real)warm[(size_t)i * nv + k];
  }
}
void nmo_get_state(const nmo_batch* b, double* qpos, double* qvel, double* warm) {
  const int nq = b->m->nq, nv = b->m->nv;
  for (int i = 0; i < b->n; i++) {
    if (qpos) for (int k = 0; k < nq; k++) qpos[(size_t)i * nq + k] = (double)b->d[i].qpos[k];
    if (qvel) for (int k = 0; k < nv; k++) qvel[(size_t)i * nv + k] = (double)b->d[i].qvel[k];
    if (warm) for (int k = 0; k < nv; k++) warm[(size_t)i * nv + k] = (double)b->d[i].qacc_warmstart[k];
  }
}

/* contiguous env slices per thread, last thread takes the remainder (env.py:195-204) */
typedef struct { nmo_batch* b; int lo, hi, nstep, mode; const float* actions; int act_stride; float* obs; float* rew; int64_t* done; } job_t;

static void env_step_one(nmo_batch* b, int i, const float* act, float* obs, float* rew, int64_t* done);

static void* worker(void* arg) {
  job_t* j = (job_t*)arg;
  for (int i = j->lo; i < j->hi; i++) {
    if (j->mode == 0) for (int s = 0; s < j->nstep; s++) step1(j->b->m, j->b->d + i);
    else if (j->mode == 1) forward(j->b->m, j->b->d + i);
    else env_step_one(j->b, i, j->actions + (size_t)i * j->act_stride, j->obs + (size_t)i * 66, j->rew + i, j->done + i);
  }
  return NULL;
}

static void run_parallel(job_t proto, int nthreads) {
  int n = proto.b->n;
  if (nthreads < 1) nthreads = 1;
  if (nthreads > n) nthreads = n;
  if (nthreads == 1) { proto.lo = 0; proto.hi = n; worker(&proto); return; }
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nthreads);
  job_t* jobs = (job_t*)malloc(sizeof(job_t) * nthreads);
  int chunk = n / nthreads;
  for (int t = 0; t < nthreads; t++) {
    jobs[t] = proto;
    jobs[t].lo = t * chunk;
    jobs[t].hi = (t == nthreads - 1) ? n : (t + 1) * chunk;
    pthread_create(th + t, NULL, worker, jobs + t);
  }
  for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
  free(th); free(jobs);
}

static void set_ctrl(nmo_batch* b, const double* ctrl) {
  const int nu = b->m->nu;
  if (ctrl) for (int i = 0; i < b->n; i++) for (int k = 0; k < nu; k++) b->d[i].ctrl[k] = (real)ctrl[(size_t)i * nu + k];
}
void nmo_physics_step(nmo_batch* b, const double* ctrl, int nstep, int nthreads) {
  set_ctrl(b, ctrl);
  job_t j; memset(&j, 0, sizeof(j));
  j.b = b; j.nstep = nstep; j.mode = 0;
  run_parallel(j, nthreads);
}
void nmo_forward(nmo_batch* b, const double* ctrl, int nthreads) {
  set_ctrl(b, ctrl);
  job_t j; memset(&j, 0, sizeof(j));
  j.b = b; j.mode = 1;
  run_parallel(j, nthreads);
}

#define RET(name_, ptr_, cnt_)                                                         \
  if (!strcmp(name, name_)) {                                                          \
    int c_ = (cnt_);                                                                   \
    if (c_ > cap) c_ = cap;                                                            \
    for (int k_ = 0; k_ < c_; k_++) out[k_] = (double)(ptr_)[k_];                       \
    return (cnt_);                                                                     \
  }

int nmo_get_array(const nmo_batch* b, int env, const char* name, double* out, int cap) {
  const nmo_model* m = b->m;
  const data_t* d = b->d + env;
  int nv = m->nv, nb = m->nbody;
  RET("qpos", d->qpos, m->nq) RET("qvel", d->qvel, nv) RET("qacc_warmstart", d->qacc_warmstart, nv) RET("ctrl", d->ctrl, m->nu)
  RET("xpos", d->xpos, 3 * nb) RET("xquat", d->xquat, 4 * nb) RET("xmat", d->xmat, 9 * nb) RET("xipos", d->xipos, 3 * nb)
  RET("ximat", d->ximat, 9 * nb) RET("xanchor", d->xanchor, 3 * m->njnt) RET("xaxis", d->xaxis, 3 * m->njnt)
  RET("site_xpos", d->site_xpos, 3 * m->nsite) RET("subtree_com", d->subtree_com, 3 * nb) RET("cinert", d->cinert, 10 * nb)
  RET("cdof", d->cdof, 6 * nv) RET("M", d->M, nv * nv) RET("cvel", d->cvel, 6 * nb) RET("cdof_dot", d->cdof_dot, 6 * nv)
  RET("qfrc_bias", d->qfrc_bias, nv) RET("qfrc_actuator", d->qfrc_actuator, nv) RET("qfrc_smooth", d->qfrc_smooth, nv)
  RET("qacc_smooth", d->qacc_smooth, nv) RET("qfrc_constraint", d->qfrc_constraint, nv) RET("qacc", d->qacc, nv)
  RET("sensordata", d->sensordata, m->nsensor)
  RET("efc_J", d->efc_J, d->nefc * nv) RET("efc_pos", d->efc_pos, d->nefc) RET("efc_R", d->efc_R, d->nefc)
  RET("efc_aref", d->efc_aref, d->nefc) RET("efc_b", d->efc_b, d->nefc) RET("efc_force", d->efc_force, d->nefc)
  RET("efc_AR", d->efc_AR, d->nefc * d->nefc) RET("efc_vel", d->efc_vel, d->nefc)
  if (!strcmp(name, "ncon")) { if (cap > 0) out[0] = d->ncon; return 1; }
  if (!strcmp(name, "nefc")) { if (cap > 0) out[0] = d->nefc; return 1; }
  if (!strcmp(name, "time")) { if (cap > 0) out[0] = d->time; return 1; }
  if (!strcmp(name, "solver_niter")) { if (cap > 0) out[0] = d->solver_niter; if (cap > 1) out[1] = d->noslip_niter; if (cap > 2) out[2] = d->warm_used; return 3; }
  if (!strcmp(name, "nwarn")) { if (cap > 0) out[0] = d->nwarn; return 1; }
  if (!strcmp(name, "nmpr")) { if (cap > 0) out[0] = (double)d->nmpr; if (cap > 1) out[1] = (double)d->nmpr_hit; if (cap > 2) out[2] = (double)d->nsupp; return 3; }
  if (!strcmp(name, "contact_frame")) {   /* per contact: normal(3) */
    for (int c = 0; c < d->ncon && 3 * c + 2 < cap; c++) { out[3 * c] = d->con[c].frame[0]; out[3 * c + 1] = d->con[c].frame[1]; out[3 * c + 2] = d->con[c].frame[2]; }
    return d->ncon * 3;
  }
  if (!strcmp(name, "contact")) {   /* per contact: geom1 geom2 vert dist pos(3) */
    int c_ = d->ncon * 7;
    for (int c = 0; c < d->ncon && 7 * c + 6 < cap; c++) {
      out[7 * c] = d->con[c].geom1; out[7 * c + 1] = d->con[c].geom2; out[7 * c + 2] = d->con[c].vert;
      out[7 * c + 3] = d->con[c].dist; out[7 * c + 4] = d->con[c].pos[0]; out[7 * c + 5] = d->con[c].pos[1]; out[7 * c + 6] = d->con[c].pos[2];
    }
    return c_;
  }
  return -1;
}

/* ------------------------------------------------------------------------------------------ env layer */
enum { RW_ACTION_RATE = 0, RW_ANG_VEL_XY, RW_BASE_HEIGHT, RW_BODY_CONTACT_FORCES, RW_COLLISION, RW_DEFAULT_POSITION,
       RW_DOF_ACC, RW_DOF_VEL, RW_FEET_AIR_TIME, RW_FEET_CONTACT_FORCES, RW_FEET_STUMBLE, RW_LIN_VEL_Z, RW_ORIENTATION,
       RW_STAND_STILL, RW_TERMINATION, RW_TORQUES, RW_TRACKING_ANG_VEL, RW_TRACKING_LIN_VEL };

/* ≙ _resample_commands (env.py:321-333) with the counter-based RNG this project defines
   (the reference uses the unseeded global numpy MT19937, which cannot be reproduced) */
static void resample_commands(nmo_batch* b, int i, int phase) {
  const cfg_t* c = &b->cfg;
  envstate_t* e = b->e + i;
  uint32_t r[4];
  nmo_philox4x32((uint32_t)b->seed, (uint32_t)i, (uint32_t)b->step_counter, (uint32_t)((uint64_t)b->step_counter >> 32), (uint32_t)phase,
                 (uint32_t)(b->seed >> 32), r);
  e->commands[0] = u01(r[0]) * 2 * c->max_lin_vel_x - c->max_lin_vel_x;
  e->commands[1] = 0;
  e->commands[2] = u01(r[1]) * 2 * c->max_ang_vel - c->max_ang_vel;
  real nrm = sqrt(e->commands[0] * e->commands[0] + e->commands[1] * e->commands[1]);
  real keep = nrm > 0.02 ? 1.0 : 0.0;
  e->commands[0] *= keep; e->commands[1] *= keep;
}

static void reset_one(nmo_batch* b, int i) {
  /* env.py:348-361: only qpos/qvel are restored; warm start, time and ctrl survive (quirk Q3) */
  const nmo_model* m = b->m;
  data_t* d = b->d + i;
  envstate_t* e = b->e + i;
  memcpy(d->qpos, m->qpos0, sizeof(real) * m->nq);
  memset(d->qvel, 0, sizeof(real) * m->nv);
  resample_commands(b, i, 1);
  memset(e->feet_air_time, 0, sizeof(e->feet_air_time));
  e->ep_len = 0;
  e->reset_buf = 1;
}

static real reward_term(nmo_batch* b, envstate_t* e, int k) {
  const cfg_t* c = &b->cfg;
  real s = 0;
  switch (k) {
    case RW_ACTION_RATE: for (int j = 0; j < 18; j++) { real x = e->prev_actions[j] - e->actions[j]; s += x * x; } return s;
    case RW_ANG_VEL_XY: return e->base_ang_vel[0] * e->base_ang_vel[0] + e->base_ang_vel[1] * e->base_ang_vel[1];
    case RW_BASE_HEIGHT: { real x = e->base_height - c->base_height_target; return x * x; }
    case RW_BODY_CONTACT_FORCES:
      if (c->tibia_contact_mode == 1) for (int j = 0; j < 6; j++) s += e->tibia_f[j];
      if (c->body_contact_mode == 1) s += e->body_f;
      return s;
    case RW_DEFAULT_POSITION: for (int j = 0; j < 18; j++) { real x = e->dof_pos[j] - c->default_pos[j]; s += x * x; } return s;
    case RW_DOF_ACC: for (int j = 0; j < 18; j++) s += e->dof_acc[j] * e->dof_acc[j]; return s;
    case RW_DOF_VEL: for (int j = 0; j < 18; j++) s += e->dof_vel[j] * e->dof_vel[j]; return s;
    case RW_FEET_AIR_TIME: {   /* stateful, env.py:447-477 */
      for (int j = 0; j < 6; j++) {
        int contact = e->feet_f[j] > 1.0;
        int filt = contact || e->last_contacts[j];
        e->feet_air_time[j] += c->dt;
        e->feet_air_time[j] *= (filt == e->last_contacts_filt[j]) ? 1.0 : 0.0;
        e->last_contacts[j] = contact;
        e->last_contacts_filt[j] = filt;
        real t = e->feet_air_time[j];
        real r = (t > 1.0 ? (t - 1.0) : 0.0) + (t < 0.5 ? (0.5 - t) : 0.0);
        s += r * r;
      }
      return s;
    }
    case RW_FEET_CONTACT_FORCES:
      for (int j = 0; j < 6; j++) { real x = (e->feet_f[j] - c->max_contact_force) * (e->feet_f[j] > c->max_contact_force ? 1.0 : 0.0); s += x * x; }
      return s;
    case RW_LIN_VEL_Z: return e->base_lin_vel[2] * e->base_lin_vel[2];
    case RW_ORIENTATION: return e->projected_gravity[0] * e->projected_gravity[0] + e->projected_gravity[1] * e->projected_gravity[1];
    case RW_STAND_STILL: {
      for (int j = 0; j < 18; j++) s += fabs(e->dof_pos[j] - c->default_pos[j]);
      real nrm = sqrt(e->commands[0] * e->commands[0] + e->commands[1] * e->commands[1]);
      return s * (nrm < 0.01 ? 1.0 : 0.0);
    }
    case RW_TORQUES: return 0.0;   /* qfrc_applied is never written (quirk Q7) */
    case RW_TRACKING_ANG_VEL: { real x = e->commands[2] - e->base_ang_vel[2]; return exp(-x * x / c->tracking_sigma); }
    case RW_TRACKING_LIN_VEL: {
      real x = e->commands[0] - e->base_lin_vel[0], y = e->commands[1] - e->base_lin_vel[1];
      return exp(-(x * x + y * y) / c->tracking_sigma);
    }
    default: return 0.0;   /* collision / feet_stumble have no function in the reference */
  }
}

static void env_step_one(nmo_batch* b, int i, const float* act, float* obs, float* rew, int64_t* done) {
  const nmo_model* m = b->m;
  const cfg_t* c = &b->cfg;
  data_t* d = b->d + i;
  envstate_t* e = b->e + i;
  STAGE(ST_ENV);
  /* E1/E3 36+54, E7 3 rotations 3*(31+15), E9 36+6, E11 ~20, E14 active terms ~200, E15 ~70 + clip */
  FLOP(90 + 138 + 42 + 20 + 200 + 70);
  /* E1 */
  real prev_dof_vel[18];
  for (int j = 0; j < 18; j++) {
    e->prev_actions[j] = e->actions[j];
    /* the reference scales and clips the policy's float32 tensor in float32 (numpy: float32 array * python float stays
     * float32, envs/nightmare_v3_env.py:155-156) and only then mixes it with float64 buffers */
    const float a32 = act[j] * (float)c->action_scale, lim = (float)c->clip_actions;
    e->actions[j] = (real)(a32 < -lim ? -lim : (a32 > lim ? lim : a32));
    prev_dof_vel[j] = e->dof_vel[j];
  }
  /* E3/E4: PD law from the carried (possibly stale) dof_pos */
  for (int j = 0; j < 18; j++) d->ctrl[j] = ((e->actions[j] - c->default_pos[j]) - e->dof_pos[j]) * c->p_gain;
  /* E5 */
  for (int s = 0; s < c->decimation; s++) step1(m, d);
  STAGE(ST_ENV);
  /* E6 */
  e->ep_len += 1;
  /* E7: base frame quantities; cvel/xipos/sensordata are one substep stale (quirk Q4) */
  int fj = -1;
  for (int j = 0; j < m->njnt; j++) if (m->jnt_type[j] == JNT_FREE) { fj = j; break; }
  int bb = m->jnt_body[fj], qa = m->jnt_qposadr[fj];
  real bq[4] = {d->qpos[qa + 3], -d->qpos[qa + 4], -d->qpos[qa + 5], -d->qpos[qa + 6]};
  real grav[3] = {0, 0, -9.81};
  rot_vec_quat(e->base_lin_vel, d->cvel + 6 * bb + 3, bq);
  rot_vec_quat(e->base_ang_vel, d->cvel + 6 * bb, bq);
  rot_vec_quat(e->projected_gravity, grav, bq);
  /* E8 */
  for (int j = 0; j < 18; j++) { e->dof_pos[j] = d->qpos[m->nq - 18 + j]; e->dof_vel[j] = d->qvel[m->nv - 18 + j]; }
  e->base_height = d->xipos[3 * bb + 2];
  for (int j = 0; j < 6; j++) { e->tibia_f[j] = d->sensordata[j]; e->feet_f[j] = d->sensordata[6 + j]; }
  e->body_f = d->sensordata[12];
  /* E9 */
  for (int j = 0; j < 18; j++) e->dof_acc[j] = (e->dof_vel[j] - prev_dof_vel[j]) / c->dt;
  for (int j = 0; j < 6; j++) e->tibia_f[j] *= (e->feet_f[j] == 0) ? 1.0 : 0.0;
  /* E10 */
  if (c->resample_period > 0 && e->ep_len % c->resample_period == 0) resample_commands(b, i, 0);
  /* E11 */
  e->time_out = (real)e->ep_len > c->max_episode_length;
  int reset = e->time_out;
  real fmax = e->feet_f[0], tmax = e->tibia_f[0];
  for (int j = 1; j < 6; j++) { if (e->feet_f[j] > fmax) fmax = e->feet_f[j]; if (e->tibia_f[j] > tmax) tmax = e->tibia_f[j]; }
  reset |= fmax > c->termination_contact_force;
  if (c->tibia_contact_mode == 2) reset |= tmax > c->tibia_max_contact_force;
  if (c->body_contact_mode == 2) reset |= e->body_f > c->body_max_contact_force;
  {
    const real* pg = e->projected_gravity;
    real nrm = sqrt(pg[0] * pg[0] + pg[1] * pg[1] + pg[2] * pg[2]);
    reset |= acos(-pg[2] / nrm) > 60.0 * M_PI / 180.0;
  }
  e->reset_buf = reset;
  /* E13: reset_idx runs BEFORE this step's rewards are accumulated (env.py:274 precedes :277-288), so the
     logged episode sums exclude the terminal step and that step's reward opens the next episode's sum */
  if (reset) {
    reset_one(b, i);
    memcpy(e->sums_at_reset, e->episode_sums, sizeof(e->episode_sums));
    memset(e->episode_sums, 0, sizeof(e->episode_sums));
  }
  /* E14: rewards from pre-reset buffers, post-reset commands (quirk Q1) */
  real total = 0;
  for (int k = 0; k < NMO_NREW; k++) {
    if (k == RW_TERMINATION || c->rew_scale[k] == 0) continue;
    real r = reward_term(b, e, k) * c->rew_scale[k];
    total += r;
    e->episode_sums[k] += r;
  }
  if (c->rew_scale[RW_TERMINATION] != 0) {
    real r = (real)(e->reset_buf * (e->time_out ? 0 : 1)) * c->rew_scale[RW_TERMINATION];
    total += r;
    e->episode_sums[RW_TERMINATION] += r;
  }
  /* E15 */
  real o[66];
  for (int k = 0; k < 3; k++) {
    o[k] = e->base_lin_vel[k] * c->obs_lin_vel;
    o[3 + k] = e->base_ang_vel[k] * c->obs_ang_vel;
    o[6 + k] = e->projected_gravity[k];
  }
  o[9] = e->commands[0] * c->obs_lin_vel; o[10] = e->commands[1] * c->obs_lin_vel; o[11] = e->commands[2] * c->obs_ang_vel;
  for (int j = 0; j < 18; j++) {
    o[12 + j] = (e->dof_pos[j] - c->default_pos[j]) * c->obs_dof_pos;
    o[30 + j] = e->dof_vel[j] * c->obs_dof_vel;
    o[48 + j] = e->actions[j];
  }
  if (c->add_noise) {
    for (int k = 0; k < 66; k += 4) {
      uint32_t r[4];
      nmo_philox4x32((uint32_t)b->seed, (uint32_t)i, (uint32_t)b->step_counter, (uint32_t)((uint64_t)b->step_counter >> 32),
                     (uint32_t)(2 + k / 4), (uint32_t)(b->seed >> 32), r);
      for (int q = 0; q < 4 && k + q < 66; q++) o[k + q] += (2 * u01(r[q]) - 1) * c->noise_vec[k + q];
    }
  }
  for (int k = 0; k < 66; k++) {
    real v = o[k] < -c->clip_obs ? -c->clip_obs : (o[k] > c->clip_obs ? c->clip_obs : o[k]);
    obs[k] = (float)v;
  }
  *rew = (float)total;
  *done = e->reset_buf;
}

void nmo_env_step(nmo_batch* b, const float* actions, int act_stride, float* obs, float* rew, int64_t* done,
                  float* time_outs, double* ep_sum_means, int* num_reset, int nthreads) {
  b->step_counter += 1;      /* E6: common_step_counter (used as the RNG counter of this step) */
  /* episode sums of envs that reset this step must be captured before they are zeroed: do the
     per-env work first, then the cross-env reduction (env.py:363-367) */
  job_t j; memset(&j, 0, sizeof(j));
  j.b = b; j.mode = 2; j.actions = actions; j.act_stride = act_stride; j.obs = obs; j.rew = rew; j.done = done;
  run_parallel(j, nthreads);
  int nres = 0;
  double acc[NMO_NREW];
  memset(acc, 0, sizeof(acc));
  for (int i = 0; i < b->n; i++) {
    if (time_outs) time_outs[i] = b->e[i].time_out ? 1.0f : 0.0f;
    if (b->e[i].reset_buf) {
      nres++;
      for (int k = 0; k < NMO_NREW; k++) acc[k] += (double)b->e[i].sums_at_reset[k];
    }
  }
  /* extras["episode"]["rew_k"] = mean over reset envs / max_episode_length_s (env.py:366) */
  if (ep_sum_means)
    for (int k = 0; k < NMO_NREW; k++) ep_sum_means[k] = nres ? acc[k] / nres / (double)b->cfg.max_episode_length_s : 0.0;
  if (num_reset) *num_reset = nres;
}

void nmo_env_reset_idx(nmo_batch* b, const int64_t* ids, int n) {
  for (int k = 0; k < n; k++) {
    int i = (int)ids[k];
    reset_one(b, i);
    memset(b->e[i].episode_sums, 0, sizeof(b->e[i].episode_sums));
  }
}

int nmo_env_get(const nmo_batch* b, const char* name, double* out, int cap) {
  int n = b->n;
#define EGET(nm, expr, per)                                                     \
  if (!strcmp(name, nm)) {                                                      \
    for (int i = 0; i < n; i++)                                                 \
      for (int k = 0; k < (per); k++)                                           \
        if (i * (per) + k < cap) out[i * (per) + k] = (double)(expr);           \
    return n * (per);                                                           \
  }
  EGET("ep_len", b->e[i].ep_len, 1) EGET("commands", b->e[i].commands[k], 3) EGET("actions", b->e[i].actions[k], 18)
  EGET("dof_pos", b->e[i].dof_pos[k], 18) EGET("dof_vel", b->e[i].dof_vel[k], 18)
  EGET("episode_sums", b->e[i].episode_sums[k], NMO_NREW) EGET("feet_air_time", b->e[i].feet_air_time[k], 6)
  EGET("last_contacts", b->e[i].last_contacts[k], 6) EGET("last_contacts_filt", b->e[i].last_contacts_filt[k], 6)
  EGET("reset_buf", b->e[i].reset_buf, 1) EGET("time_out", b->e[i].time_out, 1)
  EGET("tibia_f", b->e[i].tibia_f[k], 6) EGET("feet_f", b->e[i].feet_f[k], 6) EGET("body_f", b->e[i].body_f, 1)
  EGET("base_lin_vel", b->e[i].base_lin_vel[k], 3) EGET("base_ang_vel", b->e[i].base_ang_vel[k], 3)
  EGET("projected_gravity", b->e[i].projected_gravity[k], 3)
  if (!strcmp(name, "step_counter")) { if (cap > 0) out[0] = (double)b->step_counter; return 1; }
  return -1;
}

int nmo_env_set(nmo_batch* b, const char* name, const double* in, int count) {
  int n = b->n;
#define ESET(nm, lhs, type, per)                                                \
  if (!strcmp(name, nm)) {                                                      \
    if (count != n * (per)) return -2;                                          \
    for (int i = 0; i < n; i++)                                                 \
      for (int k = 0; k < (per); k++) lhs = (type)in[i * (per) + k];            \
    return 0;                                                                   \
  }
  ESET("ep_len", b->e[i].ep_len, int64_t, 1) ESET("commands", b->e[i].commands[k], real, 3)
  ESET("actions", b->e[i].actions[k], real, 18) ESET("dof_pos", b->e[i].dof_pos[k], real, 18)
  ESET("dof_vel", b->e[i].dof_vel[k], real, 18) ESET("episode_sums", b->e[i].episode_sums[k], real, NMO_NREW)
  ESET("feet_air_time", b->e[i].feet_air_time[k], real, 6) ESET("last_contacts", b->e[i].last_contacts[k], int, 6)
  ESET("last_contacts_filt", b->e[i].last_contacts_filt[k], int, 6)
  if (!strcmp(name, "step_counter")) { b->step_counter = (int64_t)in[0]; return 0; }
  return -1;
}
