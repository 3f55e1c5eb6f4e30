/* nm_oracle.h — CPU oracle (TEST INFRASTRUCTURE ONLY, never on the product path).
 *
 * Plain-C fp64 restatement of the reference hot path:
 *   - env layer  : /root/reference/envs/nightmare_v3_env.py:145-371,399-497  (rows E1-E18 of SURVEY.md §8a)
 *   - physics    : MuJoCo 3.1.2 mj_step as configured by models/nightmare_v3/mjmodel.xml:2-3
 *                  (third-party, pinned `mujoco<=3.1.2` in requirements.txt:2, NOT vendored in the
 *                  reference tree; restated from its published pipeline).
 *
 * ENV LAYER PINNED: checked against the outputs of the reference's own, unmodified NightmareV3Env code run in the
 * build container over a stand-in for its MuJoCo calls (tools/make_refenv_golden.py -> tests/golden/
 * reference_env_on_oracle_physics.npz; tests/test_reference_env_golden.py: observations and rewards bit-identical).
 * PHYSICS PARITY UNPINNED: the reference holds no golden vectors / tests for this path and MuJoCo cannot be
 * installed in this environment, so the mj_step restatement is validated by physical invariants
 * (tests/test_oracle_physics.py) and by the behaviour of the reference's scripted gait on it (tests/test_gait.py)
 * only — see DESIGN.md "Oracle".
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may use it.
 */
#ifndef NM_ORACLE_H
#define NM_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NMO_NREW 18      /* reward terms, alphabetical (envs/helpers.py:7 iterates dir()) */
#define NMO_MAXCON 64

typedef struct nmo_model nmo_model;
typedef struct nmo_batch nmo_batch;

/* Scalars the env layer reads from NightmareV3Config (envs/nightmare_v3_config.py:4-100). */
typedef struct {
  int32_t decimation;              /* control.decimation                      :45 */
  int32_t num_actions;             /* env.num_actions (>= 18 columns accepted)  :13 */
  int32_t tibia_contact_mode;      /* env.tibia_contact_mode                  :18 */
  int32_t body_contact_mode;       /* env.body_contact_mode                   :20 */
  int32_t add_noise;               /* noise.add_noise                         :49 */
  int32_t resample_period;         /* int(commands.resampling_time / dt)      env.py:235 */
  int32_t strict_reference;        /* reserved */
  int32_t pad0;
  double action_scale, clip_actions, p_gain, clip_obs;
  double default_pos[18];
  double obs_lin_vel, obs_ang_vel, obs_dof_pos, obs_dof_vel;
  double max_lin_vel_x, max_ang_vel;
  double max_episode_length;       /* ceil(episode_length_s / dt)             env.py:101 */
  double max_episode_length_s;
  double termination_contact_force, tibia_max_contact_force, body_max_contact_force;
  double tracking_sigma, base_height_target, max_contact_force;
  double dt;                       /* timestep * decimation                   env.py:99 */
  double rew_scale[NMO_NREW];      /* already multiplied by dt; 0 = inactive  env.py:123-128 */
  double noise_vec[66];            /* env.py:109-119 */
} nmo_envcfg;

nmo_model* nmo_model_load(const char* nmb_path, char* err, int errlen);
void       nmo_model_free(nmo_model*);
int        nmo_model_size(const nmo_model*, const char* what);   /* "nq","nv","nu","nbody","nsensor",... */

nmo_batch* nmo_batch_create(const nmo_model*, int num_envs, uint64_t seed, const nmo_envcfg* cfg);
void       nmo_batch_free(nmo_batch*);

/* raw state access, row-major [n, nq] / [n, nv] */
void nmo_set_state(nmo_batch*, const double* qpos, const double* qvel, const double* warm);
void nmo_get_state(const nmo_batch*, double* qpos, double* qvel, double* warm);

/* ≙ mj.mj_step(model, data[i], nstep) for every env (env.py:200); ctrl is [n, nu]. */
void nmo_physics_step(nmo_batch*, const double* ctrl, int nstep, int nthreads);
/* ≙ mj.mj_forward: fills all intermediates without integrating. */
void nmo_forward(nmo_batch*, const double* ctrl, int nthreads);

/* copy a named intermediate of one env into out (returns element count, <0 if unknown name) */
int nmo_get_array(const nmo_batch*, int env, const char* name, double* out, int cap);

/* ≙ NightmareV3Env.step (env.py:145-311).  actions: float32 [n, act_stride]. */
void nmo_env_step(nmo_batch*, const float* actions, int act_stride, float* obs, float* rew,
                  int64_t* done, float* time_outs, double* ep_sum_means /*[NMO_NREW] or NULL*/,
                  int* num_reset, int nthreads);
/* ≙ reset_idx (env.py:335-371) for explicit ids (used by reset()). */
void nmo_env_reset_idx(nmo_batch*, const int64_t* ids, int n);
/* env carry state: name in {"ep_len","commands","actions","dof_pos","dof_vel","episode_sums","step_counter",
   "feet_air_time","last_contacts","last_contacts_filt"}; values as double */
int nmo_env_get(const nmo_batch*, const char* name, double* out, int cap);
int nmo_env_set(nmo_batch*, const char* name, const double* in, int count);

/* Philox4x32-10 (shared definition with the CUDA kernels; exposed for known-answer tests) */
void nmo_philox4x32(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t out[4]);

#ifdef __cplusplus
}
#endif
#endif
