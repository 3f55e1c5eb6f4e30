// nm_abi.cu — host side of the C ABI declared in include/nightmare_b200.h.
//
// Parses a compiled model (.nmb, produced by nightmare_rl_b200/mjcf.py), checks that it has the
// topology the kernels are written for, digests it into the NmDevModel constant table and launches
// the kernels.  Nothing here computes physics on the CPU: there is no fallback path — if CUDA is
// unavailable every entry point fails with NM_ERR_CUDA.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/nightmare_b200.h"
#include "nm_device.hpp"

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
int nm_fail(int code, const std::string& msg) { return fail(code, msg); }   // shared with nm_policy.cu
extern "C" const char* nm_last_error(void) { return g_err.c_str(); }

#define CUDA_OK(expr)                                                                                   \
  do {                                                                                                  \
    cudaError_t e_ = (expr);                                                                            \
    if (e_ != cudaSuccess) return fail(NM_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); \
  } while (0)

// ------------------------------------------------------------------------------------------------ NMB container
struct NmbArray { const unsigned char* data; int code; std::vector<long long> dims; long long count; };

struct nm_model {
  std::vector<unsigned char> raw;
  int nq, nv, nu, nbody, njnt, ngeom, nsite, nsensor, nhv, nhn, nleg;
  double timestep;
  NmDevModel dev;                       // host copy of the constant table
  std::vector<float4> hull4;
  std::vector<float> smap;       // support maps of the legs' hulls (nm_device.hpp, NM_SMAP_N)
  std::vector<int> nbr_adr, nbr;
  std::vector<std::string> names[6];    // body joint geom site actuator sensor
  std::vector<float> qpos0;
};

static bool nmb_find(const std::vector<unsigned char>& raw, const char* name, NmbArray& out) {
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
      out.data = raw.data() + off;
      out.code = (int)code;
      out.dims.assign(dims, dims + nd);
      out.count = 1;
      for (unsigned k = 0; k < nd; k++) out.count *= dims[k];
      return true;
    }
    off += (size_t)nbytes + (size_t)((8 - nbytes % 8) % 8);
  }
  return false;
}

template <typename T>
static bool get(const nm_model* m, const char* name, int code, const T*& ptr, long long* count = nullptr) {
  NmbArray a;
  if (!nmb_find(m->raw, name, a) || a.code != code) return false;
  ptr = reinterpret_cast<const T*>(a.data);
  if (count) *count = a.count;
  return true;
}

static void quat2mat(const double* q, double* m) {
  double w = q[0], x = q[1], y = q[2], z = q[3];
  m[0] = w * w + x * x - y * y - z * z; m[1] = 2 * (x * y - w * z); m[2] = 2 * (x * z + w * y);
  m[3] = 2 * (x * y + w * z); m[4] = w * w - x * x + y * y - z * z; m[5] = 2 * (y * z - w * x);
  m[6] = 2 * (x * z - w * y); m[7] = 2 * (y * z + w * x); m[8] = w * w - x * x - y * y + z * z;
}

// inertia about the COM expressed in the link frame: R diag(I) R^T -> xx yy zz xy xz yz
static void local_inertia(const double* iquat, const double* diag, float* out) {
  double R[9];
  quat2mat(iquat, R);
  double M[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double s = 0;
      for (int k = 0; k < 3; k++) s += R[3 * i + k] * diag[k] * R[3 * j + k];
      M[3 * i + j] = s;
    }
  out[0] = (float)M[0]; out[1] = (float)M[4]; out[2] = (float)M[8]; out[3] = (float)M[1]; out[4] = (float)M[2]; out[5] = (float)M[5];
}

static int build_device_model(nm_model* m) {
  const int *sizes, *oi;
  const double* orl;
  long long n_oi = 0, n_orl = 0;
  if (!get(m, "sizes", 2, sizes) || !get(m, "opt_int", 2, oi, &n_oi) || !get(m, "opt_real", 0, orl, &n_orl))
    return fail(NM_ERR_FORMAT, "nmb: missing header arrays (sizes/opt_int/opt_real)");
  m->nq = sizes[0]; m->nv = sizes[1]; m->nu = sizes[2]; m->nbody = sizes[3]; m->njnt = sizes[4];
  m->ngeom = sizes[5]; m->nsite = sizes[6]; m->nsensor = sizes[7]; m->nhv = sizes[8]; m->nhn = sizes[9];
  m->timestep = orl[0];
#define NEED(var, name, code, type)        \
  const type* var;                         \
  if (!get(m, name, code, var)) return fail(NM_ERR_FORMAT, std::string("nmb: missing array '") + name + "'")
  NEED(qpos0, "qpos0", 0, double); NEED(body_parent, "body_parent", 2, int); NEED(body_jntadr, "body_jntadr", 2, int);
  NEED(body_jntnum, "body_jntnum", 2, int); NEED(body_pos, "body_pos", 0, double); NEED(body_quat, "body_quat", 0, double);
  NEED(body_ipos, "body_ipos", 0, double); NEED(body_iquat, "body_iquat", 0, double); NEED(body_mass, "body_mass", 0, double);
  NEED(body_inertia, "body_inertia", 0, double); NEED(body_invweight0, "body_invweight0", 0, double);
  NEED(jnt_type, "jnt_type", 2, int); NEED(jnt_qposadr, "jnt_qposadr", 2, int); NEED(jnt_dofadr, "jnt_dofadr", 2, int);
  NEED(jnt_pos, "jnt_pos", 0, double); NEED(jnt_axis, "jnt_axis", 0, double);
  NEED(dof_damping, "dof_damping", 0, double); NEED(dof_armature, "dof_armature", 0, double);
  NEED(act_dof, "act_dof", 2, int); NEED(act_gain, "act_gain", 0, double); NEED(act_bias, "act_bias", 0, double);
  NEED(act_gear, "act_gear", 0, double); NEED(act_ctrlrange, "act_ctrlrange", 0, double); NEED(act_ctrllimited, "act_ctrllimited", 2, int);
  NEED(act_forcerange, "act_forcerange", 0, double); NEED(act_forcelimited, "act_forcelimited", 2, int);
  NEED(geom_type, "geom_type", 2, int); NEED(geom_body, "geom_body", 2, int); NEED(geom_condim, "geom_condim", 2, int);
  NEED(geom_priority, "geom_priority", 2, int); NEED(geom_plane, "geom_plane", 2, int); NEED(geom_pos, "geom_pos", 0, double);
  NEED(geom_quat, "geom_quat", 0, double); NEED(geom_friction, "geom_friction", 0, double); NEED(geom_solref, "geom_solref", 0, double);
  NEED(geom_solimp, "geom_solimp", 0, double); NEED(geom_margin, "geom_margin", 0, double); NEED(geom_gap, "geom_gap", 0, double);
  NEED(geom_rbound, "geom_rbound", 0, double); NEED(geom_hull_adr, "geom_hull_adr", 2, int); NEED(geom_hull_num, "geom_hull_num", 2, int);
  NEED(hull_vert, "hull_vert", 1, float); NEED(hull_nbr_adr, "hull_nbr_adr", 2, int); NEED(hull_nbr, "hull_nbr", 2, int);
  NEED(site_body, "site_body", 2, int); NEED(site_pos, "site_pos", 0, double); NEED(site_size, "site_size", 0, double);
  NEED(sensor_site, "sensor_site", 2, int);
#undef NEED
  const int integrator = oi[0], solver = oi[1], cone = oi[2];
  if (solver != 0) return fail(NM_ERR_UNSUPPORTED, "only solver=\"PGS\" is implemented by nm_step / nm_physics_step; Newton models load through nm_gen_model_from_buffer");
  if (cone != 0) return fail(NM_ERR_UNSUPPORTED, "only cone=\"pyramidal\" is implemented by the step kernels");
  if (integrator != 0 && integrator != 3) return fail(NM_ERR_UNSUPPORTED, "only integrator Euler/implicitfast is implemented");

  // ---- topology: world -> base(free) -> nleg x (hinge, hinge, hinge)
  if (m->nbody < 2 || body_parent[1] != 0 || body_jntnum[1] != 1 || jnt_type[body_jntadr[1]] != 0)
    return fail(NM_ERR_UNSUPPORTED, "body 1 must be a free-floating base attached to the world");
  if ((m->nbody - 2) % 3 != 0) return fail(NM_ERR_UNSUPPORTED, "expected 3 links per leg");
  const int nleg = (m->nbody - 2) / 3;
  if (nleg < 1 || nleg > 6) return fail(NM_ERR_UNSUPPORTED, "1..6 legs supported");
  if (m->nq != 7 + 3 * nleg || m->nv != 6 + 3 * nleg) return fail(NM_ERR_UNSUPPORTED, "nq/nv do not match base + 3 hinges per leg");
  m->nleg = nleg;
  NmDevModel& D = m->dev;
  memset(&D, 0, sizeof(D));
  D.nleg = nleg;
  for (int l = 0; l < NM_OCT; l++) {
    NmLeg& L = D.leg[l];
    for (int j = 0; j < 3; j++) {
      L.rc[j][0] = L.rc[j][4] = L.rc[j][8] = 1.f;
      L.rc_ident[j] = 1;
      L.axis[j][2] = 1.f;
      L.clo[j] = L.flo[j] = -INFINITY;
      L.chi[j] = L.fhi[j] = INFINITY;
    }
    L.site_r[0] = L.site_r[1] = -1.f;
    L.isleg = l < nleg ? 1.f : 0.f;
  }
  std::vector<int> last_link(nleg);
  for (int k = 0; k < nleg; k++) {
    NmLeg& L = D.leg[k];
    for (int j = 0; j < 3; j++) {
      const int b = 2 + 3 * k + j;
      const int want_parent = j == 0 ? 1 : b - 1;
      if (body_parent[b] != want_parent || body_jntnum[b] != 1) return fail(NM_ERR_UNSUPPORTED, "legs must be serial chains of single-joint links");
      const int jid = body_jntadr[b];
      if (jnt_type[jid] != 3) return fail(NM_ERR_UNSUPPORTED, "leg joints must be hinges");
      if (jnt_dofadr[jid] != 6 + 3 * k + j || jnt_qposadr[jid] != 7 + 3 * k + j) return fail(NM_ERR_UNSUPPORTED, "unexpected dof order");
      for (int c = 0; c < 3; c++)
        if (std::fabs(jnt_pos[3 * jid + c]) > 1e-12) return fail(NM_ERR_UNSUPPORTED, "joint anchors must sit at the link origin");
      for (int c = 0; c < 3; c++) { L.pos[j][c] = (float)body_pos[3 * b + c]; L.axis[j][c] = (float)jnt_axis[3 * jid + c]; L.ipos[j][c] = (float)body_ipos[3 * b + c]; }
      double R[9];
      quat2mat(body_quat + 4 * b, R);
      bool ident = true;
      for (int c = 0; c < 9; c++) { L.rc[j][c] = (float)R[c]; ident &= std::fabs(R[c] - (c % 4 == 0 ? 1.0 : 0.0)) < 1e-14; }
      L.rc_ident[j] = ident ? 1 : 0;
      local_inertia(body_iquat + 4 * b, body_inertia + 3 * b, L.iloc[j]);
      L.mass[j] = (float)body_mass[b];
      L.qref[j] = (float)qpos0[7 + 3 * k + j];
      L.damping[j] = (float)dof_damping[6 + 3 * k + j];
      L.armature[j] = (float)dof_armature[6 + 3 * k + j];
    }
    last_link[k] = 2 + 3 * k + 2;
  }
  for (int a = 0; a < m->nu; a++) {
    const int dof = act_dof[a];
    if (dof < 6 || dof >= m->nv) return fail(NM_ERR_UNSUPPORTED, "actuator on a non-hinge dof");
    if (dof != 6 + a || m->nu != 3 * nleg) return fail(NM_ERR_UNSUPPORTED, "actuator i must drive hinge i (ctrl layout of env.py:192)");
    NmLeg& L = D.leg[(dof - 6) / 3];
    const int j = (dof - 6) % 3;
    L.gain0[j] = (float)act_gain[3 * a]; L.bias0[j] = (float)act_bias[3 * a]; L.bias1[j] = (float)act_bias[3 * a + 1];
    L.bias2[j] = (float)act_bias[3 * a + 2]; L.gear[j] = (float)act_gear[a];
    if (act_ctrllimited[a]) { L.clo[j] = (float)act_ctrlrange[2 * a]; L.chi[j] = (float)act_ctrlrange[2 * a + 1]; }
    if (act_forcelimited[a]) { L.flo[j] = (float)act_forcerange[2 * a]; L.fhi[j] = (float)act_forcerange[2 * a + 1]; }
  }
  // base
  double total_mass = 0;
  for (int b = 1; b < m->nbody; b++) total_mass += body_mass[b];
  for (int c = 0; c < 3; c++) D.b_ipos[c] = (float)body_ipos[3 + c];
  local_inertia(body_iquat + 4, body_inertia + 3, D.b_iloc);
  D.b_mass = (float)body_mass[1];
  D.total_mass = (float)total_mass;
  for (int i = 0; i < m->nq; i++) D.qpos0[i] = (float)qpos0[i];
  m->qpos0.assign(D.qpos0, D.qpos0 + m->nq);

  // ---- collision geoms: exactly one static plane; convex hulls on the base and on the last links
  int plane = -1;
  for (int g = 0; g < m->ngeom; g++)
    if (geom_type[g] == 0) {
      if (plane >= 0 || geom_body[g] != 0) return fail(NM_ERR_UNSUPPORTED, "exactly one world-fixed plane geom is supported");
      plane = g;
    }
  const double opt_timestep = orl[0], impratio = orl[6], meaninertia = orl[7];
  const double pyramid_rfac = (n_orl > 10 && orl[10] > 0) ? orl[10] : 2.0;       // model option (opt_real[10]): R of pyramid edges = this * mu_reg^2 * R[first]
  if (plane >= 0) {
    double R[9];
    quat2mat(geom_quat + 4 * plane, R);
    double n[3] = {R[2], R[5], R[8]};
    double fr[9] = {n[0], n[1], n[2], 0, 0, 0, 0, 0, 0};
    if (n[1] < 0.5 && n[1] > -0.5) fr[4] = 1; else fr[5] = 1;       // ≙ mju_makeFrame
    double dd = fr[0] * fr[3] + fr[1] * fr[4] + fr[2] * fr[5];
    for (int c = 0; c < 3; c++) fr[3 + c] -= dd * fr[c];
    double nn = std::sqrt(fr[3] * fr[3] + fr[4] * fr[4] + fr[5] * fr[5]);
    for (int c = 0; c < 3; c++) fr[3 + c] /= nn;
    fr[6] = fr[1] * fr[5] - fr[2] * fr[4]; fr[7] = fr[2] * fr[3] - fr[0] * fr[5]; fr[8] = fr[0] * fr[4] - fr[1] * fr[3];
    for (int c = 0; c < 9; c++) D.frame[c] = (float)fr[c];
    for (int c = 0; c < 3; c++) D.plane_n[c] = (float)n[c];
    D.plane_d = (float)(n[0] * geom_pos[3 * plane] + n[1] * geom_pos[3 * plane + 1] + n[2] * geom_pos[3 * plane + 2]);
  } else {
    D.plane_n[2] = 1.f; D.frame[2] = 1.f; D.frame[4] = 1.f; D.frame[6] = -1.f;
  }
  for (int g = 0; g < m->ngeom; g++) {
    if (geom_type[g] == 0 || geom_plane[g] < 0) continue;
    if (geom_type[g] != 7) return fail(NM_ERR_UNSUPPORTED, "only convex-mesh geoms collide with the plane in this kernel");
    int lane = -1;
    if (geom_body[g] == 1) lane = 6;
    for (int k = 0; k < nleg; k++) if (geom_body[g] == last_link[k]) lane = k;
    if (lane < 0) return fail(NM_ERR_UNSUPPORTED, "plane-colliding geoms must sit on the base or on a leg's last link");
    NmGeom& G = D.leg[lane].geom;
    if (G.has) return fail(NM_ERR_UNSUPPORTED, "at most one colliding geom per body");
    const int p = plane;
    double mu, solref[2], solimp[5];
    int dim;
    if (geom_priority[g] == geom_priority[p]) {
      mu = std::fmax(geom_friction[3 * g], geom_friction[3 * p]);
      dim = geom_condim[g] > geom_condim[p] ? geom_condim[g] : geom_condim[p];
      for (int c = 0; c < 2; c++) solref[c] = 0.5 * (geom_solref[2 * g + c] + geom_solref[2 * p + c]);
      for (int c = 0; c < 5; c++) solimp[c] = 0.5 * (geom_solimp[5 * g + c] + geom_solimp[5 * p + c]);
    } else {
      const int w = geom_priority[g] > geom_priority[p] ? g : p;
      mu = geom_friction[3 * w]; dim = geom_condim[w];
      for (int c = 0; c < 2; c++) solref[c] = geom_solref[2 * w + c];
      for (int c = 0; c < 5; c++) solimp[c] = geom_solimp[5 * w + c];
    }
    if (dim != 3) return fail(NM_ERR_UNSUPPORTED, "only condim=3 contacts are implemented");
    auto clampd = [](double x) { return x < 0.0001 ? 0.0001 : (x > 0.9999 ? 0.9999 : x); };
    G.has = 1;
    G.hull_adr = geom_hull_adr[g]; G.hull_num = geom_hull_num[g]; G.start = 0;
    G.rbound = (float)geom_rbound[g];
    G.margin = (float)(std::fmax(geom_margin[g], geom_margin[p]) - std::fmax(geom_gap[g], geom_gap[p]));
    G.mu = (float)mu;
    const double tran = body_invweight0[2 * geom_body[g]] + body_invweight0[2 * geom_body[p]];
    const double mureg = mu / std::sqrt(impratio > 1e-15 ? impratio : 1.0);
    G.rfac = (float)(pyramid_rfac * mureg * mureg * (1.0 + mu * mu) * tran);
    double tc = solref[0], dr = solref[1], dmax = clampd(solimp[1]);
    if (tc > 0) {
      if (tc < 2 * opt_timestep) tc = 2 * opt_timestep;
      G.K = (float)(1.0 / (dmax * dmax * tc * tc * dr * dr));
      G.B = (float)(2.0 / (dmax * tc));
    } else {
      G.K = (float)(-tc / (dmax * dmax));
      G.B = (float)(-dr / dmax);
    }
    G.dmin = (float)clampd(solimp[0]); G.dmax = (float)dmax; G.width = (float)solimp[2];
    G.mid = (float)clampd(solimp[3]); G.power = (float)(solimp[4] < 1 ? 1 : solimp[4]);
  }
  // ---- convex-convex pairs between the legs' hulls (reference mjmodel.xml:47: tibia contype=2 / conaffinity=3).  MuJoCo's
  // filter: (contype1 & conaffinity2) || (contype2 & conaffinity1), different bodies, not parent and child.
  {
    const int *contype = nullptr, *conaff = nullptr;
    const double* center = nullptr;
    const bool have = get(m, "geom_contype", 2, contype) && get(m, "geom_conaffinity", 2, conaff) && get(m, "geom_center", 0, center);
    std::vector<int> geom_of(NM_OCT, -1);
    for (int g = 0; g < m->ngeom; g++) {
      if (geom_type[g] != 7) continue;
      for (int k = 0; k < nleg; k++) if (geom_body[g] == last_link[k] && D.leg[k].geom.has) geom_of[k] = g;
      if (geom_body[g] == 1) geom_of[6] = g;
    }
    D.pair_mask = 0;
    if (have) {
      int idx = 0;
      for (int i = 0; i < 6; i++)
        for (int j = i + 1; j < 6; j++, idx++) {
          if (i >= nleg || j >= nleg || geom_of[i] < 0 || geom_of[j] < 0) continue;
          const int gi = geom_of[i], gj = geom_of[j];
          if (!((contype[gi] & conaff[gj]) || (contype[gj] & conaff[gi]))) continue;
          const NmGeom &A = D.leg[i].geom, &B = D.leg[j].geom;
          if (A.mu != B.mu || A.K != B.K || A.B != B.B || A.dmin != B.dmin || A.dmax != B.dmax || A.width != B.width || A.mid != B.mid ||
              A.power != B.power || A.margin != B.margin || geom_priority[gi] != geom_priority[gj])
            return fail(NM_ERR_UNSUPPORTED, "leg hulls that collide with each other must share their contact parameters");
          D.pair_mask |= 1 << idx;
        }
      // base hull against leg hulls, or any other convex-convex pair, is not handled by the kernel: refuse rather than ignore
      for (int g1 = 0; g1 < m->ngeom; g1++)
        for (int g2 = g1 + 1; g2 < m->ngeom; g2++) {
          if (geom_type[g1] != 7 || geom_type[g2] != 7 || geom_body[g1] == geom_body[g2]) continue;
          if (!((contype[g1] & conaff[g2]) || (contype[g2] & conaff[g1]))) continue;
          if (body_parent[geom_body[g1]] == geom_body[g2] || body_parent[geom_body[g2]] == geom_body[g1]) continue;
          bool l1 = false, l2 = false;
          for (int k = 0; k < nleg; k++) { l1 |= geom_of[k] == g1; l2 |= geom_of[k] == g2; }
          if (!(l1 && l2)) return fail(NM_ERR_UNSUPPORTED, "convex-convex pairs are implemented between the legs' last-link hulls only");
        }
    }
    for (int k = 0; k < nleg; k++) {
      NmGeom& G = D.leg[k].geom;
      if (!G.has || geom_of[k] < 0) continue;
      const int g = geom_of[k];
      const float* hv = hull_vert + 3 * (size_t)G.hull_adr;
      double c[3] = {0, 0, 0};
      if (have) for (int a = 0; a < 3; a++) c[a] = center[3 * g + a];
      else { for (int v = 0; v < G.hull_num; v++) for (int a = 0; a < 3; a++) c[a] += hv[3 * v + a] / G.hull_num; }
      for (int a = 0; a < 3; a++) G.center[a] = (float)c[a];
      // bounding capsule: principal axis of the vertex cloud (power iteration on its covariance), extent along it, largest
      // distance from the axis
      double mean[3] = {0, 0, 0}, C[9] = {0};
      for (int v = 0; v < G.hull_num; v++) for (int a = 0; a < 3; a++) mean[a] += hv[3 * v + a] / (double)G.hull_num;
      for (int v = 0; v < G.hull_num; v++)
        for (int a = 0; a < 3; a++) for (int b2 = 0; b2 < 3; b2++) C[3 * a + b2] += (hv[3 * v + a] - mean[a]) * (hv[3 * v + b2] - mean[b2]);
      double u[3] = {1, 1, 1};
      for (int it = 0; it < 200; it++) {
        double w[3] = {C[0] * u[0] + C[1] * u[1] + C[2] * u[2], C[3] * u[0] + C[4] * u[1] + C[5] * u[2], C[6] * u[0] + C[7] * u[1] + C[8] * u[2]};
        const double nn = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
        if (nn < 1e-30) break;
        for (int a = 0; a < 3; a++) u[a] = w[a] / nn;
      }
      // ... then tightened: the axis (direction and offset) is moved by a deterministic local search to minimise the largest
      // distance of a vertex from it (the principal axis of the hexapod's tibia leaves 43 mm because of the bracket at its top;
      // the best axis 30 mm), and the segment is shortened at both ends by what the spherical caps cover.  A tight capsule
      // matters: walking robots keep adjacent tibias 20-40 mm apart, and every pair of overlapping capsules costs a
      // support-function test over both hulls in the step kernel.
      auto radius_of = [&](const double* ax, const double* c0) {
        double r = 0;
        for (int v = 0; v < G.hull_num; v++) {
          const double d[3] = {hv[3 * v] - c0[0], hv[3 * v + 1] - c0[1], hv[3 * v + 2] - c0[2]};
          const double t = d[0] * ax[0] + d[1] * ax[1] + d[2] * ax[2];
          const double pr[3] = {d[0] - t * ax[0], d[1] - t * ax[1], d[2] - t * ax[2]};
          r = std::fmax(r, pr[0] * pr[0] + pr[1] * pr[1] + pr[2] * pr[2]);
        }
        return std::sqrt(r);
      };
      {
        double best = radius_of(u, mean);
        unsigned long long rs = 0x9E3779B97F4A7C15ull + (unsigned long long)k;
        auto rnd = [&]() { rs ^= rs << 13; rs ^= rs >> 7; rs ^= rs << 17; return (double)(rs >> 11) * (1.0 / 9007199254740992.0) * 2.0 - 1.0; };
        double step_u = 0.05, step_c = 0.25 * best;
        for (int it = 0; it < 6000; it++) {
          double nu[3] = {u[0] + step_u * rnd(), u[1] + step_u * rnd(), u[2] + step_u * rnd()};
          const double nn = std::sqrt(nu[0] * nu[0] + nu[1] * nu[1] + nu[2] * nu[2]);
          for (int a = 0; a < 3; a++) nu[a] /= nn;
          const double nc[3] = {mean[0] + step_c * rnd(), mean[1] + step_c * rnd(), mean[2] + step_c * rnd()};
          const double r = radius_of(nu, nc);
          if (r < best) { best = r; for (int a = 0; a < 3; a++) { u[a] = nu[a]; mean[a] = nc[a]; } }
          if (it % 1000 == 999) { step_u *= 0.5; step_c *= 0.5; }
        }
      }
      double rmax = radius_of(u, mean) * 1.0001 + 1e-6;
      double t0 = -1e30, t1 = 1e30, tlo = 1e30, thi = -1e30;           // segment [t0, t1]: every vertex within rmax of it
      for (int v = 0; v < G.hull_num; v++) {
        const double d[3] = {hv[3 * v] - mean[0], hv[3 * v + 1] - mean[1], hv[3 * v + 2] - mean[2]};
        const double t = d[0] * u[0] + d[1] * u[1] + d[2] * u[2];
        const double pr[3] = {d[0] - t * u[0], d[1] - t * u[1], d[2] - t * u[2]};
        const double reach = std::sqrt(std::fmax(rmax * rmax - (pr[0] * pr[0] + pr[1] * pr[1] + pr[2] * pr[2]), 0.0));
        t1 = std::fmin(t1, t + reach);        // the lower end may sit as high as this and still cover vertex v with its cap ...
        t0 = std::fmax(t0, t - reach);        // ... and the upper end as low as this
        tlo = std::fmin(tlo, t); thi = std::fmax(thi, t);
      }
      double tmin = std::fmin(t1, 0.5 * (tlo + thi)), tmax = std::fmax(t0, 0.5 * (tlo + thi));
      for (int a = 0; a < 3; a++) { G.cap_a[a] = (float)(mean[a] + tmin * u[a]); G.cap_b[a] = (float)(mean[a] + tmax * u[a]); }
      G.cap_r = (float)rmax;
      G.cap_il2 = (float)(1.0 / std::fmax((tmax - tmin) * (tmax - tmin), 1e-12));
      G.cap_len = (float)((tmax - tmin) * 1.0001);
      {
        // support map (nm_device.hpp, NM_SMAP_N): node values in double, rounded up by 2 um so that the fp32 interpolation in
        // the kernel stays an upper bound
        G.smap_adr = (int)m->smap.size();
        m->smap.resize(m->smap.size() + NM_SMAP_FLOATS);
        float* T = m->smap.data() + G.smap_adr;
        for (int ax = 0; ax < 3; ax++)
          for (int sg = 0; sg < 2; sg++)
            for (int iv = 0; iv <= NM_SMAP_N; iv++)
              for (int iu = 0; iu <= NM_SMAP_N; iu++) {
                double cdir[3];
                cdir[ax] = sg ? -1.0 : 1.0;
                cdir[(ax + 1) % 3] = -1.0 + 2.0 * iu / NM_SMAP_N;
                cdir[(ax + 2) % 3] = -1.0 + 2.0 * iv / NM_SMAP_N;
                double h = -1e300;
                for (int v = 0; v < G.hull_num; v++) h = std::fmax(h, cdir[0] * hv[3 * v] + cdir[1] * hv[3 * v + 1] + cdir[2] * hv[3 * v + 2]);
                T[((2 * ax + sg) * (NM_SMAP_N + 1) + iv) * (NM_SMAP_N + 1) + iu] = (float)(h + 2e-6 + 1e-6 * std::fabs(h));
              }
      }
      const double mureg = G.mu / std::sqrt(impratio > 1e-15 ? impratio : 1.0);
      G.rfac_self = (float)(pyramid_rfac * mureg * mureg * (1.0 + (double)G.mu * G.mu) * body_invweight0[2 * geom_body[g]]);
    }
    long long n_or = 0;
    const double* orl2 = nullptr;
    get(m, "opt_real", 0, orl2, &n_or);
    D.mpr_iterations = (n_oi > 8 && oi[8] > 0) ? oi[8] : 50;
    D.mpr_tolerance = (n_or > 8 && orl2[8] > 0) ? (float)orl2[8] : 1e-6f;
  }
  // ---- touch sensors: [nleg x slot0 | nleg x slot1 | base]  (mjmodel.xml:157-169, env.py:224-226)
  if (m->nsensor != 0) {
    if (m->nsensor != 2 * nleg + 1) return fail(NM_ERR_UNSUPPORTED, "touch sensors must be laid out as [legs..., feet..., base]");
    for (int s = 0; s < m->nsensor; s++) {
      const int site = sensor_site[s];
      const int lane = s < 2 * nleg ? s % nleg : 6, slot = s < nleg ? 0 : (s < 2 * nleg ? 1 : 0);
      const int want_body = lane == 6 ? 1 : last_link[lane];
      if (site_body[site] != want_body) return fail(NM_ERR_UNSUPPORTED, "touch sensor site is not on the expected body");
      NmLeg& L = D.leg[lane];
      for (int c = 0; c < 3; c++) L.site_pos[slot][c] = (float)site_pos[3 * site + c];
      L.site_r[slot] = (float)site_size[site];
    }
  }
  for (int c = 0; c < 3; c++) D.gravity[c] = (float)orl[1 + c];
  D.timestep = (float)opt_timestep;
  D.tolerance = (float)orl[4];
  D.noslip_tolerance = (float)orl[5];
  D.solver_scale = (float)(1.0 / (meaninertia * (m->nv > 1 ? m->nv : 1)));
  D.iterations = oi[3];
  D.noslip_iterations = oi[4];
  D.planemesh_maxcon = (n_oi > 7 && oi[7] >= 1 && oi[7] <= NM_MAXC) ? oi[7] : NM_MAXC;
  D.planemesh_allverts = n_oi > 9 ? oi[9] != 0 : 0;
  D.planemesh_sepvert = n_oi > 10 ? oi[10] != 0 : 0;
  D.warm_after_noslip = n_oi > 11 ? oi[11] != 0 : 0;
  D.planemesh_sep = (n_orl > 9 && orl[9] > 0) ? (float)orl[9] : 0.3f;
  D.integrator = integrator;
  D.imp_act = integrator == 3 ? 1.f : 0.f;
  D.imp_damp = (integrator == 3 || oi[5]) ? 1.f : 0.f;
  m->hull4.resize(m->nhv);
  for (int i = 0; i < m->nhv; i++) m->hull4[i] = make_float4(hull_vert[3 * i], hull_vert[3 * i + 1], hull_vert[3 * i + 2], 0.f);
  m->nbr_adr.assign(hull_nbr_adr, hull_nbr_adr + m->nhv + 1);
  // adjacency offsets are stored relative to the geom's first vertex so that the kernel indexes [hull_adr + v]
  m->nbr.assign(hull_nbr, hull_nbr + m->nhn);
  // names
  static const char* keys[6] = {"names_body", "names_joint", "names_geom", "names_site", "names_actuator", "names_sensor"};
  for (int k = 0; k < 6; k++) {
    NmbArray a;
    m->names[k].clear();
    if (!nmb_find(m->raw, keys[k], a) || a.count == 0) continue;
    std::string s(reinterpret_cast<const char*>(a.data), (size_t)a.count), cur;
    for (char ch : s) { if (ch == '\n') { m->names[k].push_back(cur); cur.clear(); } else cur.push_back(ch); }
    m->names[k].push_back(cur);
  }
  return NM_OK;
}

extern "C" int nm_model_from_buffer(const void* data, size_t nbytes, nm_model** out) {
  if (!data || !out) return fail(NM_ERR_ARG, "null argument");
  if (nbytes < 8 || memcmp(data, "NMB1", 4) != 0) return fail(NM_ERR_FORMAT, "buffer is not an NMB1 compiled model");
  nm_model* m = new nm_model();
  m->raw.assign(static_cast<const unsigned char*>(data), static_cast<const unsigned char*>(data) + nbytes);
  int rc = build_device_model(m);
  if (rc != NM_OK) { delete m; return rc; }
  *out = m;
  return NM_OK;
}

extern "C" int nm_model_load(const char* path, nm_model** out) {
  if (!path || !out) return fail(NM_ERR_ARG, "null argument");
  FILE* f = fopen(path, "rb");
  if (!f) return fail(NM_ERR_IO, std::string("cannot open ") + path);
  std::vector<unsigned char> buf;
  unsigned char tmp[65536];
  size_t n;
  while ((n = fread(tmp, 1, sizeof(tmp), f)) > 0) buf.insert(buf.end(), tmp, tmp + n);
  fclose(f);
  return nm_model_from_buffer(buf.data(), buf.size(), out);
}

extern "C" void nm_model_destroy(nm_model* m) { delete m; }

extern "C" int nm_model_size(const nm_model* m, const char* w) {
  if (!m || !w) return -1;
  std::string s(w);
  if (s == "nq") return m->nq; if (s == "nv") return m->nv; if (s == "nu") return m->nu; if (s == "nbody") return m->nbody;
  if (s == "njnt") return m->njnt; if (s == "ngeom") return m->ngeom; if (s == "nsite") return m->nsite;
  if (s == "nsensor") return m->nsensor; if (s == "nleg") return m->nleg;
  return -1;
}
extern "C" double nm_model_timestep(const nm_model* m) { return m ? m->timestep : 0.0; }

extern "C" int nm_name2id(const nm_model* m, int objtype, const char* name) {
  if (!m || !name) return -1;
  int k;
  switch (objtype) { case 1: k = 0; break; case 3: k = 1; break; case 5: k = 2; break; case 6: k = 3; break; case 19: k = 4; break; case 20: k = 5; break; default: return -1; }
  for (size_t i = 0; i < m->names[k].size(); i++) if (m->names[k][i] == name) return (int)i;
  return -1;
}

extern "C" int nm_model_qpos0(const nm_model* m, float* out, int cap) {
  if (!m || !out) return NM_ERR_ARG;
  for (int i = 0; i < m->nq && i < cap; i++) out[i] = m->qpos0[i];
  return m->nq;
}

extern "C" int nm_model_support_map(const nm_model* m, int leg, float* table, int cap, float* verts, int vcap, int* nvert) {
  if (!m || leg < 0 || leg >= m->nleg) return 0;
  const NmGeom& G = m->dev.leg[leg].geom;
  if (nvert) *nvert = 0;
  if (!G.has || G.hull_num <= 0 || (size_t)G.smap_adr + NM_SMAP_FLOATS > m->smap.size()) return 0;
  if (table) for (int i = 0; i < NM_SMAP_FLOATS && i < cap; i++) table[i] = m->smap[(size_t)G.smap_adr + i];
  if (verts) for (int v = 0; v < G.hull_num && v < vcap; v++) { verts[3 * v] = m->hull4[G.hull_adr + v].x; verts[3 * v + 1] = m->hull4[G.hull_adr + v].y; verts[3 * v + 2] = m->hull4[G.hull_adr + v].z; }
  if (nvert) *nvert = G.hull_num;
  return NM_SMAP_N;
}

// ------------------------------------------------------------------------------------------------ batch
struct nm_batch {
  const nm_model* model;
  int n, device;
  NmKernelArgs args;
  NmDevModel* d_model;
  NmDevCfg* d_cfg;
  float4* d_hull;
  float* d_smap;
  unsigned short *d_nbr16, *d_nadr16;
  cudaEvent_t host_ev;   // nm_step_host waits for the step kernel's outputs, not for the extras latch launched behind it
  int* d_hint;
  float* d_acc;          // [2][19] double-buffered episode accumulators
  int parity;
  float* d_stage_actions;
  size_t stage_cap;
  int64_t launches;
  float* rec_ring;       // env-0 recorder ring (caller-owned) and its write cursor
  int rec_cap;
  int64_t rec_count;
};

// every entry point runs with the batch's device current and restores the caller's (a batch on cuda:1 used from a thread whose
// current device is 0 would otherwise launch on a foreign stream)
struct DevGuard {
  int prev = -1, dev;
  explicit DevGuard(int d) : dev(d) { if (cudaGetDevice(&prev) != cudaSuccess) prev = -1; if (prev != dev) cudaSetDevice(dev); }
  ~DevGuard() { if (prev >= 0 && prev != dev) cudaSetDevice(prev); }
};

static void fill_cfg(const nm_envcfg& c, NmDevCfg& d) {
  memset(&d, 0, sizeof(d));
  d.decimation = c.decimation; d.tibia_mode = c.tibia_contact_mode; d.body_mode = c.body_contact_mode;
  d.add_noise = c.add_noise; d.resample_period = c.resample_period; d.strict = c.strict_reference ? 1 : 0;
  d.action_scale = (float)c.action_scale; d.clip_actions = (float)c.clip_actions; d.p_gain = (float)c.p_gain; d.clip_obs = (float)c.clip_obs;
  for (int i = 0; i < 18; i++) d.default_pos[i] = (float)c.default_pos[i];
  d.obs_lin_vel = (float)c.obs_lin_vel; d.obs_ang_vel = (float)c.obs_ang_vel; d.obs_dof_pos = (float)c.obs_dof_pos; d.obs_dof_vel = (float)c.obs_dof_vel;
  d.max_lin_vel_x = (float)c.max_lin_vel_x; d.max_ang_vel = (float)c.max_ang_vel;
  d.max_episode_length = (float)c.max_episode_length;
  d.inv_episode_length_s = (float)(1.0 / c.max_episode_length_s);
  d.term_force = (float)c.termination_contact_force; d.tibia_max_force = (float)c.tibia_max_contact_force; d.body_max_force = (float)c.body_max_contact_force;
  d.inv_tracking_sigma = (float)(1.0 / c.tracking_sigma); d.base_height_target = (float)c.base_height_target;
  d.max_contact_force = (float)c.max_contact_force; d.dt = (float)c.dt; d.inv_dt = (float)(1.0 / c.dt);
  for (int i = 0; i < 18; i++) d.rew_scale[i] = (float)c.rew_scale[i];
  for (int i = 0; i < 66; i++) d.noise_vec[i] = (float)c.noise_vec[i];
}

static int batch_create_impl(const nm_model* m, int num_envs, int device, uint64_t seed, const nm_envcfg* cfg, const nm_buffers* bufs, nm_batch* b);

extern "C" int nm_batch_create(const nm_model* m, int num_envs, int device, uint64_t seed, const nm_envcfg* cfg,
                               const nm_buffers* bufs, nm_batch** out) {
  if (!m || !bufs || !out || num_envs <= 0) return fail(NM_ERR_ARG, "nm_batch_create: bad argument");
  nm_batch* b = new nm_batch();
  memset(b, 0, sizeof(*b));
  const int rc = batch_create_impl(m, num_envs, device, seed, cfg, bufs, b);
  if (rc != NM_OK) { nm_batch_destroy(b); return rc; }        // frees whatever was allocated before the failure
  *out = b;
  return NM_OK;
}

static int batch_create_impl(const nm_model* m, int num_envs, int device, uint64_t seed, const nm_envcfg* cfg, const nm_buffers* bufs, nm_batch* b) {
  if (m->nleg != 6 || m->nq != NM_NQ || m->nv != NM_NV)
    return fail(NM_ERR_UNSUPPORTED, "the env-step kernel is laid out for the 6-leg / 18-dof Nightmare topology");
  if (!bufs->qpos || !bufs->qvel || !bufs->warm || !bufs->sensordata) return fail(NM_ERR_ARG, "physics buffers must be non-null");
  if (cfg && (!bufs->actions || !bufs->dof_pos || !bufs->dof_vel || !bufs->commands || !bufs->episode_length || !bufs->episode_sums ||
              !bufs->feet_air_time || !bufs->contact_bits || !bufs->obs || !bufs->rew || !bufs->done || !bufs->time_outs || !bufs->episode_acc || !bufs->ep_means || !bufs->time_outs_latched))
    return fail(NM_ERR_ARG, "env buffers must be non-null when an env config is given");
  if (cfg && (cfg->decimation < 1 || cfg->num_actions < NM_NDOF)) return fail(NM_ERR_ARG, "env config: decimation>=1 and num_actions>=18 required");
  int ndev = 0;
  CUDA_OK(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(NM_ERR_CUDA, "no such CUDA device");
  CUDA_OK(cudaSetDevice(device));
  b->model = m; b->n = num_envs; b->device = device;
  CUDA_OK(cudaMalloc(&b->d_model, sizeof(NmDevModel)));
  CUDA_OK(cudaMemcpy(b->d_model, &m->dev, sizeof(NmDevModel), cudaMemcpyHostToDevice));
  CUDA_OK(cudaMalloc(&b->d_cfg, sizeof(NmDevCfg)));
  NmDevCfg hc;
  if (cfg) fill_cfg(*cfg, hc); else memset(&hc, 0, sizeof(hc));
  CUDA_OK(cudaMemcpy(b->d_cfg, &hc, sizeof(NmDevCfg), cudaMemcpyHostToDevice));
  CUDA_OK(cudaMalloc(&b->d_hull, sizeof(float4) * (m->hull4.size() + 1)));
  CUDA_OK(cudaMemcpy(b->d_hull, m->hull4.data(), sizeof(float4) * m->hull4.size(), cudaMemcpyHostToDevice));
  CUDA_OK(cudaMalloc(&b->d_smap, sizeof(float) * (m->smap.size() + 4)));
  if (!m->smap.empty()) CUDA_OK(cudaMemcpy(b->d_smap, m->smap.data(), sizeof(float) * m->smap.size(), cudaMemcpyHostToDevice));
  // limits the packed support-vertex hint (9 bits vertex, 6 bits degree) and the 16-bit adjacency rely on
  for (int g = 0; g < NM_OCT; g++) {
    const NmGeom& G = m->dev.leg[g].geom;
    if (!G.has) continue;
    if (G.hull_num > 0x1ff) return fail(NM_ERR_UNSUPPORTED, "convex hulls with more than 511 vertices are not supported");
    for (int v = 0; v < G.hull_num; v++)
      if (m->nbr_adr[G.hull_adr + v + 1] - m->nbr_adr[G.hull_adr + v] > 63) return fail(NM_ERR_UNSUPPORTED, "hull vertex with more than 63 neighbours");
  }
  CUDA_OK(cudaMalloc(&b->d_hint, sizeof(int) * (size_t)num_envs * NM_OCT));
  CUDA_OK(cudaMemset(b->d_hint, 0, sizeof(int) * (size_t)num_envs * NM_OCT));
  CUDA_OK(cudaMalloc(&b->d_acc, sizeof(float) * 2 * (NM_NREW + 1)));
  CUDA_OK(cudaMemset(b->d_acc, 0, sizeof(float) * 2 * (NM_NREW + 1)));
  NmKernelArgs& a = b->args;
  a.hull_hint = b->d_hint;
  a.ep_means = bufs->ep_means; a.time_outs_latched = bufs->time_outs_latched;
  a.model = b->d_model; a.cfg = b->d_cfg; a.hull_vert = b->d_hull; a.hull_smap = b->d_smap;
  {
    // compact adjacency (16-bit ids and offsets) for the support-vertex walk.  NM_HULL_SMEM=1 makes every CTA stage the tables
    // (52 KB for the hexapod) in shared memory: measured (round 2, gpurun_out/r02_qb17.log, r02_phase8*.log) the walk itself
    // drops from 4630 to 3210 cycles per substep but the step does not get faster (89.1 -> 91.0 us at 4096 envs, equal at
    // 16 384 / 131 072): with the CTA's warps in lockstep the phase is bound by instruction issue, not by load latency.  Off by default.
    const size_t ne = m->nbr.size(), nv1 = m->nbr_adr.size();
    if (ne >= 65536) return fail(NM_ERR_UNSUPPORTED, "hull adjacency with more than 65535 directed edges is not supported");
    const size_t ne_pad = (ne + 8 + 7) & ~(size_t)7, na_pad = (nv1 + 7) & ~(size_t)7;
    std::vector<unsigned short> n16(ne_pad, 0), a16(na_pad, 0);
    for (size_t i = 0; i < ne; i++) n16[i] = (unsigned short)m->nbr[i];
    for (size_t i = 0; i < nv1; i++) a16[i] = (unsigned short)m->nbr_adr[i];
    CUDA_OK(cudaMalloc(&b->d_nbr16, 2 * ne_pad));
    CUDA_OK(cudaMemcpy(b->d_nbr16, n16.data(), 2 * ne_pad, cudaMemcpyHostToDevice));
    CUDA_OK(cudaMalloc(&b->d_nadr16, 2 * na_pad));
    CUDA_OK(cudaMemcpy(b->d_nadr16, a16.data(), 2 * na_pad, cudaMemcpyHostToDevice));
    a.hull_nbr16 = b->d_nbr16; a.hull_nadr16 = b->d_nadr16;
    a.hull_nv = (int)m->hull4.size(); a.hull_ne_pad = (int)ne_pad; a.hull_na_pad = (int)na_pad;
    { const char* pf = getenv("NM_PAIR_FILTER_OFF"); a.pair_filter_off = (pf && pf[0] == '1') ? 1 : 0; }
    const char* sw = getenv("NM_HULL_SMEM");
    const size_t bytes = 16 * m->hull4.size() + 2 * ne_pad + 2 * na_pad;
    a.hull_smem = (bytes <= 96 * 1024 && sw && sw[0] == '1') ? 1 : 0;
  }
  a.num_envs = num_envs; a.nstep = cfg ? cfg->decimation : 1; a.step_counter = 0; a.env_offset = 0; a.seed = seed;
  a.qpos = bufs->qpos; a.qvel = bufs->qvel; a.warm = bufs->warm; a.actions = bufs->actions; a.dof_pos = bufs->dof_pos;
  a.dof_vel = bufs->dof_vel; a.commands = bufs->commands; a.episode_length = reinterpret_cast<long long*>(bufs->episode_length);
  a.episode_sums = bufs->episode_sums; a.feet_air_time = bufs->feet_air_time; a.contact_bits = bufs->contact_bits;
  a.obs = bufs->obs; a.rew = bufs->rew; a.done = reinterpret_cast<long long*>(bufs->done); a.time_outs = bufs->time_outs;
  a.sensordata = bufs->sensordata; a.episode_acc = bufs->episode_acc; a.debug = bufs->debug;
  a.in_actions = nullptr; a.act_stride = 0; a.in_ctrl = nullptr;
  a.host_obs = nullptr; a.host_rew = nullptr; a.host_done = nullptr;
  a.dr = nullptr; a.dr_on_reset = 0; a.rec_row = nullptr;
  return NM_OK;
}

extern "C" void nm_batch_destroy(nm_batch* b) {
  if (!b) return;
  cudaFree(b->d_model); cudaFree(b->d_cfg); cudaFree(b->d_hull); cudaFree(b->d_smap); cudaFree(b->d_nbr16); cudaFree(b->d_nadr16); if (b->host_ev) cudaEventDestroy(b->host_ev); cudaFree(b->d_hint); cudaFree(b->d_acc);
  if (b->d_stage_actions) cudaFree(b->d_stage_actions);
  delete b;
}

extern "C" int nm_batch_set_domain_randomization(nm_batch* b, float* dr, const float* ranges, int resample_on_reset) {
  if (!b) return fail(NM_ERR_ARG, "null batch");
  b->args.dr = dr;
  b->args.dr_on_reset = (dr && ranges && resample_on_reset) ? 1 : 0;
  for (int i = 0; i < 6; i++) b->args.dr_range[i] = ranges ? ranges[i] : 1.f;
  return NM_OK;
}

extern "C" int nm_batch_set_recorder(nm_batch* b, float* ring, int capacity) {
  if (!b || (ring && capacity <= 0)) return fail(NM_ERR_ARG, "nm_batch_set_recorder: bad argument");
  b->rec_ring = ring; b->rec_cap = ring ? capacity : 0; b->rec_count = 0;
  return NM_OK;
}

static inline float* next_rec_row(nm_batch* b) {
  if (!b->rec_ring) return nullptr;
  return b->rec_ring + (size_t)(b->rec_count++ % b->rec_cap) * NM_REC_STRIDE;
}

extern "C" int nm_batch_set_env_offset(nm_batch* b, int64_t first) {
  if (!b) return fail(NM_ERR_ARG, "null batch");
  b->args.env_offset = first;
  return NM_OK;
}

extern "C" int nm_step(nm_batch* b, const float* actions, int act_stride, int64_t step_counter, nm_stream stream) {
  if (!b || !actions) return fail(NM_ERR_ARG, "nm_step: null argument");
  if (!b->args.obs) return fail(NM_ERR_ARG, "nm_step: batch was created without env buffers");
  if (act_stride < NM_NDOF) return fail(NM_ERR_ARG, "nm_step: actions need at least 18 columns");
  DevGuard guard(b->device);
  NmKernelArgs a = b->args;
  a.in_actions = actions; a.act_stride = act_stride; a.step_counter = step_counter;
  a.rec_row = next_rec_row(b);
  a.acc_cur = b->d_acc + (NM_NREW + 1) * b->parity;
  a.acc_next = b->d_acc + (NM_NREW + 1) * (b->parity ^ 1);
  b->parity ^= 1;
  nm_launch_step(a, true, stream);
  nm_launch_finalize(a, stream);
  b->launches += 2;
  CUDA_OK(cudaGetLastError());
  return NM_OK;
}

extern "C" int nm_physics_step(nm_batch* b, const float* ctrl, int nstep, nm_stream stream) {
  if (!b || !ctrl || nstep < 1) return fail(NM_ERR_ARG, "nm_physics_step: bad argument");
  DevGuard guard(b->device);
  NmKernelArgs a = b->args;
  a.in_ctrl = ctrl; a.nstep = nstep;
  nm_launch_step(a, false, stream);
  b->launches++;
  CUDA_OK(cudaGetLastError());
  return NM_OK;
}

extern "C" int nm_reset_idx(nm_batch* b, const int64_t* env_ids, int n, int64_t step_counter, nm_stream stream) {
  if (!b || (n > 0 && !env_ids)) return fail(NM_ERR_ARG, "nm_reset_idx: bad argument");
  if (!b->args.obs) return fail(NM_ERR_ARG, "nm_reset_idx: batch was created without env buffers");
  if (n <= 0) return NM_OK;
  DevGuard guard(b->device);
  NmKernelArgs a = b->args;
  a.step_counter = step_counter;
  nm_launch_reset(a, reinterpret_cast<const long long*>(env_ids), n, stream);
  b->launches++;
  CUDA_OK(cudaGetLastError());
  return NM_OK;
}

// pinned (page-locked) host memory is mapped into the device address space under UVA: returns its device alias or null
static void* mapped_alias(const void* host_ptr) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, host_ptr) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  if (at.type != cudaMemoryTypeHost || at.devicePointer == nullptr) return nullptr;
  return at.devicePointer;
}

extern "C" int nm_step_host(nm_batch* b, const float* h_actions, int act_stride, int64_t step_counter, float* h_obs, float* h_rew,
                            int64_t* h_done, nm_stream stream) {
  if (!b || !h_actions || !h_obs || !h_rew || !h_done) return fail(NM_ERR_ARG, "nm_step_host: null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // Zero-copy path: when all four host buffers are pinned the kernel reads the actions and writes obs / rew / dones
  // directly through their device aliases (coalesced rows over PCIe), so the step needs no copy-engine transfers at all.
  void *ma = mapped_alias(h_actions), *mo = mapped_alias(h_obs), *mr = mapped_alias(h_rew), *md = mapped_alias(h_done);
  if (ma && mo && mr && md) {
    if (!b->args.obs) return fail(NM_ERR_ARG, "nm_step_host: batch was created without env buffers");
    if (act_stride < NM_NDOF) return fail(NM_ERR_ARG, "nm_step_host: actions need at least 18 columns");
    DevGuard guard(b->device);
    NmKernelArgs a = b->args;
    a.in_actions = static_cast<const float*>(ma); a.act_stride = act_stride; a.step_counter = step_counter;
    a.rec_row = next_rec_row(b);
    a.host_obs = static_cast<float*>(mo); a.host_rew = static_cast<float*>(mr); a.host_done = static_cast<long long*>(md);
    a.acc_cur = b->d_acc + (NM_NREW + 1) * b->parity;
    a.acc_next = b->d_acc + (NM_NREW + 1) * (b->parity ^ 1);
    b->parity ^= 1;
    nm_launch_step(a, true, stream);
    // the host buffers are complete when the step kernel is: wait for an event recorded right behind it and let the extras
    // latch (device-side state, stream-ordered before anything that reads it) run while the caller already has its outputs
    if (!b->host_ev) CUDA_OK(cudaEventCreateWithFlags(&b->host_ev, cudaEventDisableTiming));
    CUDA_OK(cudaEventRecord(b->host_ev, st));
    nm_launch_finalize(a, stream);
    b->launches += 2;
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaEventSynchronize(b->host_ev));
    return NM_OK;
  }
  // Pageable host memory: staged copies on the same stream
  DevGuard guard(b->device);
  const size_t need = (size_t)b->n * act_stride * sizeof(float);
  if (need > b->stage_cap) {
    if (b->d_stage_actions) cudaFree(b->d_stage_actions);
    CUDA_OK(cudaMalloc(&b->d_stage_actions, need));
    b->stage_cap = need;
  }
  CUDA_OK(cudaMemcpyAsync(b->d_stage_actions, h_actions, need, cudaMemcpyHostToDevice, st));
  int rc = nm_step(b, b->d_stage_actions, act_stride, step_counter, stream);
  if (rc != NM_OK) return rc;
  CUDA_OK(cudaMemcpyAsync(h_obs, b->args.obs, (size_t)b->n * NM_NOBS * sizeof(float), cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaMemcpyAsync(h_rew, b->args.rew, (size_t)b->n * sizeof(float), cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaMemcpyAsync(h_done, b->args.done, (size_t)b->n * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaStreamSynchronize(st));
  return NM_OK;
}

extern "C" int64_t nm_batch_launches(const nm_batch* b) { return b ? b->launches : 0; }

extern "C" double nm_measure_fp32_peak(nm_stream stream) { return nm_run_ffma_peak(stream); }
