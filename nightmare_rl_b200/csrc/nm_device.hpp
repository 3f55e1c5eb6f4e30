// nm_device.hpp — constant tables shared by the host ABI (nm_abi.cu) and the step kernels (nm_kernels.cu).
//
// Topology the kernels are specialised for ("legged-star"): one free-floating base body plus up to six
// serial chains of three single-hinge links (coxa, femur, tibia), joint anchors at the link origins;
// one convex collision hull on every last link and one on the base, colliding with a static plane;
// touch sites on the last links and on the base.  That is exactly models/nightmare_v3/mjmodel.xml:32-170
// of the reference.  nm_batch_create() verifies the compiled model against it and fails loudly otherwise.
#pragma once
#include <cstdint>

#define NM_OCT 8            // lanes per environment: 6 leg lanes + base-geom lane + spare
#define NM_MAXC 4           // contacts per collision geom
#define NM_NOBS_DEV 66      // observation entries per env (envs/nightmare_v3_config.py:11)
#define NM_MAXPAIR 4        // simultaneous convex-convex (tibia-tibia) contacts per environment; further ones are dropped
// Support map of a hull: h(c) = max over hull vertices of c.v, tabulated at the nodes of an N x N grid on each face of the cube
// |c|_inf = 1 (hull frame).  h is convex and positively homogeneous, so on a face (a plane) the bilinear interpolant of the node
// values is an UPPER bound of h, and h(d) = |d|_inf h(d / |d|_inf): four loads instead of a scan over the hull's ~260 vertices
// (mean slack 0.1 mm, worst 3-4 mm at N = 16 on the hexapod's tibia; the exact answer comes from MPR for the pairs it keeps).
#define NM_SMAP_N 16
#define NM_SMAP_FLOATS (6 * (NM_SMAP_N + 1) * (NM_SMAP_N + 1))
#define NM_DBG 320          // floats per env of the optional debug record (== NM_DBG_STRIDE of the C ABI)
#ifndef NM_BLOCK
#define NM_BLOCK 64         // threads per CTA = 8 environments
#endif

struct NmGeom {
  int has;                  // lane owns a collision geom
  int hull_adr, hull_num, start;
  float rbound, margin, mu;
  float rfac;               // R = rfac * (1-imp)/imp   (pyramidal: 2 mu_reg^2 * (1+mu^2) * invweight)
  float K, B;               // reference-acceleration spring/damper from solref
  float dmin, dmax, width, mid, power;
  // convex-convex pairs (mjmodel.xml:47: the tibia geoms collide with each other)
  float center[3];          // MPR centre: the mesh's centre of mass, body frame (MuJoCo: the geom frame origin)
  float cap_a[3], cap_b[3]; // bounding capsule of the hull, body frame: segment end points ...
  float cap_r;              // ... and radius (conservative broad phase: hulls whose capsules do not overlap cannot intersect)
  float cap_il2;            // 1 / |cap_b - cap_a|^2
  float cap_len;            // |cap_b - cap_a|
  int smap_adr;             // first entry of this hull's support map in NmKernelArgs::hull_smap (see NM_SMAP_N)
  float rfac_self;          // this body's share of a pair contact's R: 2 mu_reg^2 (1+mu^2) * body_invweight0
};

struct NmLeg {
  float pos[3][3];          // link origin in parent frame
  float rc[3][9];           // constant rotation of the link frame (body quat)
  int   rc_ident[3];
  float axis[3][3];         // hinge axis, link frame
  float ipos[3][3];         // link COM, link frame
  float iloc[3][6];         // link inertia about its COM in the link frame: xx yy zz xy xz yz
  float mass[3];
  float qref[3];
  float gain0[3], bias0[3], bias1[3], bias2[3], gear[3];
  float clo[3], chi[3], flo[3], fhi[3];
  float damping[3], armature[3];
  float site_pos[2][3];     // touch sites on the last link: slot 0 ("tibia"/"base"), slot 1 ("foot")
  float site_r[2];          // < 0: absent
  float isleg;              // 1 for real legs, 0 for the base-geom / spare lanes
  NmGeom geom;
};

struct alignas(16) NmDevModel {
  NmLeg leg[NM_OCT];
  float b_ipos[3], b_iloc[6], b_mass, total_mass;
  float plane_n[3], plane_d, frame[9];
  float gravity[3];
  float timestep, tolerance, noslip_tolerance, solver_scale;
  int iterations, noslip_iterations, nleg, integrator;
  int planemesh_maxcon;     // contacts per plane-mesh pair (opt_int[7] of the model file, 1..NM_MAXC)
  int pair_mask;            // bit idx(i,j), i<j lexicographic over the leg lanes: the hulls of legs i and j collide
  int mpr_iterations;       // opt.mpr_iterations (50)
  // recalled-but-unverified MuJoCo details as model data (nightmare_rl_b200/mjcf.py lists them with their defaults):
  int planemesh_allverts;   // opt_int[9]:  extra plane-mesh contacts from 0 = hull-graph neighbours of the support vertex, 1 = all hull vertices
  int planemesh_sepvert;    // opt_int[10]: minimum separation measured between 0 = contact points, 1 = hull vertices
  int warm_after_noslip;    // opt_int[11]: qacc_warmstart saved 0 = before noslip, 1 = after
  float planemesh_sep;      // opt_real[9]:  separation as a fraction of rbound (0.3)
  float mpr_tolerance;      // opt.mpr_tolerance (1e-6)
  float imp_damp, imp_act;  // which velocity derivatives enter the implicit velocity update (implicitfast: both; Euler+eulerdamp: damping)
  float qpos0[32];
};

// env-layer scalars in fp32 (from nm_envcfg)
struct alignas(16) NmDevCfg {
  int decimation, tibia_mode, body_mode, add_noise, resample_period;
  int strict;               // 1 (default): extras latched only on steps where >= 1 env reset (quirk Q10); 0: time_outs refreshed every step
  int pad[2];
  float action_scale, clip_actions, p_gain, clip_obs;
  float default_pos[18];
  float obs_lin_vel, obs_ang_vel, obs_dof_pos, obs_dof_vel;
  float max_lin_vel_x, max_ang_vel, max_episode_length, inv_episode_length_s;
  float term_force, tibia_max_force, body_max_force;
  float inv_tracking_sigma, base_height_target, max_contact_force, dt, inv_dt;
  float rew_scale[18];
  float noise_vec[66];
};

struct NmKernelArgs {
  const NmDevModel* model;
  const NmDevCfg* cfg;
  const float4* hull_vert;
  const float* hull_smap;    // support maps of the legs' hulls (NM_SMAP_FLOATS each, at NmGeom::smap_adr)
  int pair_filter_off;       // test switch (NM_PAIR_FILTER_OFF=1 at batch creation): every pair of overlapping capsules goes to MPR
  // compact adjacency for the support-vertex walk: 16-bit neighbour ids and list offsets (+ hull_vert); small enough
  // (52 KB for the hexapod) to be staged in shared memory by every CTA when hull_smem != 0
  const unsigned short* hull_nbr16;   // [hull_ne_pad]
  const unsigned short* hull_nadr16;  // [hull_nv + 1, padded]
  int hull_nv, hull_ne_pad, hull_na_pad, hull_smem;
  int* hull_hint;            // [N, NM_OCT] last support vertex per collision hull (library-owned scratch)
  int num_envs;
  int nstep;
  long long step_counter;
  long long env_offset;
  unsigned long long seed;
  // state
  float* qpos; float* qvel; float* warm;
  float* actions; float* dof_pos; float* dof_vel; float* commands;
  long long* episode_length; float* episode_sums; float* feet_air_time; int* contact_bits;
  float* obs; float* rew; long long* done; float* time_outs; float* sensordata; float* episode_acc; float* debug;
  float* ep_means; float* time_outs_latched;   // extras, refreshed only on steps where >= 1 env reset (env.py:363-371)
  float* rec_row;                              // env-0 recorder (env.py:261-272): [done, qpos(25), qvel(24)] BEFORE reset_idx, or null
  float* acc_cur; float* acc_next;             // library-owned double-buffered accumulators behind episode_acc
  // host-resident callers (nm_step_host, zero-copy): pinned host memory mapped into the device address space, or null
  float* host_obs; float* host_rew; long long* host_done;
  // domain randomisation (opt-in): [N,4] = (friction scale, kv scale, base-mass scale, unused), or null
  float* dr; int dr_on_reset; float dr_range[6];
  // inputs
  const float* in_actions; int act_stride;   // env mode
  const float* in_ctrl;                       // physics-only mode
};

void nm_launch_step(const NmKernelArgs& a, bool env_mode, void* stream);
void nm_launch_finalize(const NmKernelArgs& a, void* stream);
void nm_launch_reset(const NmKernelArgs& a, const long long* env_ids, int n, void* stream);
double nm_run_ffma_peak(void* stream);
