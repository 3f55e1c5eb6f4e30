/* nightmare_b200.h — C ABI of the B200-native batched Nightmare-v3 environment step.
 *
 * The reference has no native boundary of its own: its env layer (Python) drives MuJoCo 3.1.2
 * through pybind11.  Each entry point below names the reference call site(s) it replaces
 * (paths relative to /root/reference).  All pointers in nm_buffers are DEVICE pointers owned by the
 * caller (torch tensors in the Python host layer); the library never synchronises the stream.
 *
 * Error convention: functions return 0 on success or a negative nm_status; nm_last_error() gives
 * the message of the last failure on the calling thread.  No exceptions cross this boundary.
 */
#ifndef NIGHTMARE_B200_H
#define NIGHTMARE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NM_NDOF 18        /* actuated hinge dofs  (envs/nightmare_v3_env.py:40  num_dof = nv - 6) */
#define NM_NQ 25          /* models/nightmare_v3/mjmodel.xml: free joint (7) + 18 hinges */
#define NM_NV 24
#define NM_NOBS 66        /* envs/nightmare_v3_config.py:11 */
#define NM_NSENSOR 13     /* mjmodel.xml:156-170 */
#define NM_NREW 18        /* reward terms in alphabetical order (envs/helpers.py:7) */
#define NM_MAXCON_GEOM 4  /* plane-mesh: support vertex + up to 3 more */
#define NM_DBG_STRIDE 320 /* floats per env in the optional debug buffer */
#define NM_REC_STRIDE 52  /* floats per row of the env-0 recorder ring: done flag, qpos[25], qvel[24], 2 pad */

typedef enum {
  NM_OK = 0,
  NM_ERR_IO = -1,          /* file missing / unreadable                         */
  NM_ERR_FORMAT = -2,      /* not an NMB1 compiled model / missing arrays       */
  NM_ERR_UNSUPPORTED = -3, /* model topology/options outside the kernel's scope */
  NM_ERR_ARG = -4,
  NM_ERR_CUDA = -5,
  NM_ERR_NAME = -6
} nm_status;

typedef struct nm_model nm_model;
typedef struct nm_batch nm_batch;
typedef void* nm_stream;   /* cudaStream_t */

/* Scalars of NightmareV3Config the step reads (envs/nightmare_v3_config.py:4-100), pre-digested
 * the way NightmareV3Env.__init__ does it (envs/nightmare_v3_env.py:99-137).  Same layout as the
 * oracle's nmo_envcfg. */
typedef struct {
  int32_t decimation, num_actions, tibia_contact_mode, body_contact_mode, add_noise, resample_period;
  int32_t strict_reference;   /* 1: extras["time_outs"] / ["episode"] latched only on steps where an env reset (quirk Q10, :363-371);
                                 0: time_outs refreshed every step (training mode, identical whenever >= 1 env resets) */
  int32_t pad0;
  double action_scale, clip_actions, p_gain, clip_obs;
  double default_pos[NM_NDOF];
  double obs_lin_vel, obs_ang_vel, obs_dof_pos, obs_dof_vel;
  double max_lin_vel_x, max_ang_vel;
  double max_episode_length, max_episode_length_s;
  double termination_contact_force, tibia_max_contact_force, body_max_contact_force;
  double tracking_sigma, base_height_target, max_contact_force;
  double dt;
  double rew_scale[NM_NREW];
  double noise_vec[NM_NOBS];
} nm_envcfg;

/* Caller-owned device buffers, row-major [num_envs, K]. */
typedef struct {
  /* physics state  (≙ MjData.qpos / qvel / qacc_warmstart, envs/nightmare_v3_env.py:38) */
  float* qpos;           /* [N,25] */
  float* qvel;           /* [N,24] */
  float* warm;           /* [N,24] */
  /* env carry state (≙ buffers of envs/nightmare_v3_env.py:56-97) */
  float* actions;        /* [N,18] clipped actions of the previous step */
  float* dof_pos;        /* [N,18] */
  float* dof_vel;        /* [N,18] */
  float* commands;       /* [N,3]  */
  int64_t* episode_length; /* [N] (episode_length_buf, int64 like the reference :88) */
  float* episode_sums;   /* [N,18] indexed by reward-term id */
  float* feet_air_time;  /* [N,6]  */
  int32_t* contact_bits; /* [N] bits 0-5 last_contacts, 8-13 last_contacts_filt */
  /* step outputs (≙ return tuple of step(), envs/nightmare_v3_env.py:311) */
  float* obs;            /* [N,66] */
  float* rew;            /* [N]    */
  int64_t* done;         /* [N]    reset_buf */
  float* time_outs;      /* [N]    */
  float* sensordata;     /* [N,13] touch sensors of the last substep */
  float* episode_acc;    /* [NM_NREW+1] sum over the envs reset THIS step of their episode sums, then their count */
  float* debug;          /* [N,NM_DBG_STRIDE] or NULL */
  /* extras (≙ self.extras, envs/nightmare_v3_env.py:363-371): refreshed only on steps where >= 1 env reset */
  float* ep_means;       /* [NM_NREW] mean episode sum of the envs that reset / max_episode_length_s   (:366) */
  float* time_outs_latched; /* [N] time_out_buf as of the last step that reset an env                  (:371) */
} nm_buffers;

const char* nm_last_error(void);

/* ≙ mj.MjModel.from_xml_path(cfg.env.model_path)           envs/nightmare_v3_env.py:37
 * The MJCF is compiled by the host layer (nightmare_rl_b200/mjcf.py) into an .nmb file / buffer. */
int  nm_model_load(const char* nmb_path, nm_model** out);
int  nm_model_from_buffer(const void* data, size_t nbytes, nm_model** out);
void nm_model_destroy(nm_model*);
/* ≙ model.nq / nv / nu / nbody ...  ("nq","nv","nu","nbody","ngeom","nsensor","nleg")   env.py:40-41 */
int  nm_model_size(const nm_model*, const char* what);
/* ≙ model.opt.timestep                                     envs/nightmare_v3_env.py:99 */
double nm_model_timestep(const nm_model*);
/* ≙ mj.mj_name2id(model, objtype, name) (mjtObj numbering)  envs/nightmare_v3_env.py:48 */
int  nm_name2id(const nm_model*, int objtype, const char* name);
/* ≙ model.qpos0                                            envs/nightmare_v3_env.py:349 */
int  nm_model_qpos0(const nm_model*, float* out, int cap);
/* Test access to the support map of leg `leg`'s collision hull (the conservative first stage of the tibia-tibia narrow phase that
 * replaces mj_collision's convex-convex broad phase for models/nightmare_v3/mjmodel.xml:47): copies up to `cap` floats of the
 * table the step kernel reads -- 6 cube faces x (n+1) x (n+1) nodes, node value >= max over hull vertices of c.v at the cube
 * point c -- and up to `vcap` xyz triples of the hull's vertices (body frame).  Returns the grid size n (0: the leg has no
 * hull), *nvert = number of hull vertices.  Host only, no GPU needed. */
int  nm_model_support_map(const nm_model*, int leg, float* table, int cap, float* verts, int vcap, int* nvert);

/* ≙ [mj.MjData(model) for _ in range(num_envs)]             envs/nightmare_v3_env.py:38 */
int  nm_batch_create(const nm_model*, int num_envs, int device, uint64_t seed, const nm_envcfg* cfg,
                     const nm_buffers* bufs, nm_batch** out);
void nm_batch_destroy(nm_batch*);
/* global env id of local env 0 (multi-GPU sharding keeps RNG streams keyed by GLOBAL env id) */
int  nm_batch_set_env_offset(nm_batch*, int64_t first_global_env);

/* Domain randomisation — NOT in the reference (BASELINE config 3 asks for it).  dr: DEVICE float32 [N,4] =
 * (contact-friction scale, actuator-kv scale, base-mass scale, unused), caller-owned; NULL switches it off (default).
 * ranges = {mu_lo, mu_hi, kv_lo, kv_hi, mass_lo, mass_hi}: when resample_on_reset != 0 an env that resets inside nm_step
 * draws new scales uniformly from them (Philox phase 3).  With all scales 1 the step is bit-identical to DR off. */
int  nm_batch_set_domain_randomization(nm_batch*, float* dr, const float* ranges, int resample_on_reset);

/* ≙ the env-0 state recorder                                envs/nightmare_v3_env.py:261-272
 * ring: DEVICE float32 [capacity, NM_REC_STRIDE], caller-owned (NULL switches it off).  Every nm_step / nm_step_host then
 * writes row (number of steps since this call) % capacity = {done flag of env 0, its qpos[25], its qvel[24]} as they are
 * BEFORE reset_idx runs (the reference appends the row at :272, ahead of reset_idx at :274). */
int  nm_batch_set_recorder(nm_batch*, float* ring, int capacity);

/* ≙ NightmareV3Env.step(actions)                            envs/nightmare_v3_env.py:145-311
 * actions: device float32 [N, act_stride], first 18 columns used (:156).  step_counter is the
 * value of common_step_counter AFTER this step's increment (:213); it is the RNG counter. */
int  nm_step(nm_batch*, const float* actions, int act_stride, int64_t step_counter, nm_stream stream);
/* ≙ for i: data[i].ctrl = ctrl[i]; mj.mj_step(model, data[i], nstep)   envs/nightmare_v3_env.py:191-200
 * Raw physics (parity harness): ctrl device float32 [N,18]. */
int  nm_physics_step(nm_batch*, const float* ctrl, int nstep, nm_stream stream);
/* ≙ reset_idx(env_ids) state part: data[i].qpos = qpos0; data[i].qvel = 0      envs/nightmare_v3_env.py:348-361
 * env_ids: device int64 [n]. Also resamples commands (phase 1 of step_counter), zeroes feet_air_time,
 * episode_length and episode_sums, sets done=1. */
int  nm_reset_idx(nm_batch*, const int64_t* env_ids, int n, int64_t step_counter, nm_stream stream);

/* Same as nm_step but with HOST buffers (the way the reference's step is called: CPU tensor in, CPU tensors out,
 * envs/nightmare_v3_env.py:155,:311), then synchronises the stream.  Pinned (page-locked) buffers are accessed zero-copy:
 * the kernel reads the actions and writes obs/rew/done through their device aliases, no copy-engine transfer is issued
 * (the device-resident obs/rew/done buffers are updated as well).  Pageable buffers are staged with async copies. */
int  nm_step_host(nm_batch*, const float* h_actions, int act_stride, int64_t step_counter,
                  float* h_obs, float* h_rew, int64_t* h_done, nm_stream stream);

/* ---- PPO policy forward (≙ rsl_rl v1.0.2 PPO.act: actor_critic.act / evaluate / get_actions_log_prob, reached from
 * train.py:54 runner.learn and play.py:122 `nn.act(obs)`; network shape from envs/nightmare_v3_config.py:105-109).
 * One launch computes, for n observations: mean = actor(obs), value = critic(obs), actions = mean + std*eps with
 * eps ~ N(0,1) from Philox keyed by (seed, env_offset + row, step), and log_prob = sum log N(actions; mean, std).
 * TF32 tensor-core MMAs with fp32 accumulation; ELU hidden activations. */
typedef struct nm_policy nm_policy;
typedef struct {
  int32_t num_layers;    /* linear layers, 1..6 */
  int32_t dims[7];       /* dims[0] = inputs, dims[i] = width after layer i (<= 128 each) */
} nm_mlp_shape;
int  nm_policy_create(const nm_mlp_shape* actor, const nm_mlp_shape* critic, int device, nm_policy** out);
void nm_policy_destroy(nm_policy*);
/* which: 0 actor, 1 critic (floats of the flat parameter vector), 2 number of actions */
int  nm_policy_param_count(const nm_policy*, int which);
/* actor_params / critic_params: DEVICE float32, the module's parameters flattened in PyTorch order
 * (layer0.weight [out][in], layer0.bias, layer1.weight, ...); std: DEVICE float32 [num_actions]. */
int  nm_policy_load_weights(nm_policy*, const float* actor_params, const float* critic_params, const float* std, nm_stream stream);
/* obs: DEVICE float32 [n, obs_stride]; outputs DEVICE float32: actions [n,A], mean [n,A], value [n], logp [n].
 * deterministic != 0: actions = mean (≙ act_inference, play scripts). */
int  nm_policy_act(nm_policy*, const float* obs, int obs_stride, int n, uint64_t seed, int64_t step, int64_t env_offset,
                   int deterministic, float* actions, float* mean, float* value, float* logp, nm_stream stream);
/* nm_policy_act that also writes the rollout buffer's copy of the observations (obs_copy [n, obs_dim]) and the
 * per-env std row (sigma_out [n,A]); either may be NULL. */
int  nm_policy_act_store(nm_policy*, const float* obs, int obs_stride, int n, uint64_t seed, int64_t step, int64_t env_offset,
                         int deterministic, float* actions, float* mean, float* value, float* logp, float* obs_copy,
                         float* sigma_out, nm_stream stream);
int64_t nm_policy_launches(const nm_policy*);

/* Same contract on Blackwell's tcgen05 tensor cores (csrc/nm_policy_tc5.cu): 128-env tiles, A/B operands in shared memory
 * (UMMA canonical K-major layout), fp32 accumulators in tensor memory, tcgen05.mma.kind::tf32 issued 3x on hi/lo splits.
 * Limits: <= 72 inputs, hidden widths <= 64, <= 32 outputs (the reference's 66-54-42-30-18 / 1 networks fit). */
typedef struct nm_policy_tc5 nm_policy_tc5;
int  nm_policy_tc5_create(const nm_mlp_shape* actor, const nm_mlp_shape* critic, int device, nm_policy_tc5** out);
void nm_policy_tc5_destroy(nm_policy_tc5*);
int  nm_policy_tc5_load_weights(nm_policy_tc5*, const float* actor_params, const float* critic_params, const float* std, nm_stream stream);
int  nm_policy_tc5_act(nm_policy_tc5*, const float* obs, int obs_stride, int n, uint64_t seed, int64_t step, int64_t env_offset,
                       int deterministic, float* actions, float* mean, float* value, float* logp, float* obs_copy, float* sigma_out,
                       nm_stream stream);

/* ---- rollout bookkeeping (≙ rsl_rl v1.0.2 PPO.process_env_step + RolloutStorage.add_transitions, and the episode
 * statistics OnPolicyRunner.learn keeps; reached from train.py:54).  All pointers DEVICE; source pointers that are NULL
 * mean "already written in place by nm_policy_act[_store]" (obs / std / actions / mean / value / logp). */
typedef struct {
  int32_t n, obs_dim, act_dim, ring_cap;
  float gamma, pad0;
  const float* obs; const float* actions; const float* mean; const float* std; const float* value; const float* logp;
  const float* rew; const int64_t* done; const float* time_outs;          /* time_outs may be NULL */
  float* s_obs; float* s_actions; float* s_mu; float* s_sigma; float* s_values; float* s_logp; float* s_rewards; uint8_t* s_dones;
  float* cur_rew; float* cur_len; float* ring_rew; float* ring_len; int64_t* ring_count;   /* all NULL: no statistics */
  const float* ep_means; float* ep_acc; int32_t n_ep, pad1;  /* ep_acc[0..n_ep) += ep_means, ep_acc[n_ep] += 1; NULL: off */
} nm_rollout_slot;
int  nm_rollout_store(const nm_rollout_slot* slot, nm_stream stream);

/* ---- PPO loss head (≙ the distribution / loss part of rsl_rl v1.0.2 PPO.update, train.py:54): forward sums and the
 * gradients w.r.t. mu, value and std in one launch.  All pointers DEVICE float32; n samples, act_dim actions.
 * out[0..2] = sum surrogate, sum value loss, sum KL; g_mu [n,A], g_value [n] and g_std [A] are d(loss)/d(.) for
 * loss = mean(surrogate) + value_coef*mean(value loss) - entropy_coef*mean(entropy), EXCEPT the entropy term of g_std
 * (-entropy_coef/std, sample independent), which the caller adds. */
typedef struct {
  int32_t n, act_dim, use_clipped_value_loss, pad0;
  float clip, value_coef, pad1, pad2;
  const float* mu; const float* value; const float* std; const float* actions; const float* old_logp; const float* old_mu;
  const float* old_sigma; const float* adv; const float* ret; const float* tgt_val;
  float* out; float* g_mu; float* g_value; float* g_std;
} nm_ppo_head_args;
int  nm_ppo_head(const nm_ppo_head_args* head, nm_stream stream);

/* ---- one PPO mini-batch in one launch (csrc/nm_ppo_grad.cu; ≙ rsl_rl v1.0.2 PPO.update's loop body up to loss.backward(),
 * train.py:54): rows idx[0..n) of the flat rollout buffers are gathered, actor and critic evaluated (3xTF32 tensor-core
 * forward), the loss head of nm_ppo_head applied, and d(loss)/d(parameter) accumulated for every parameter (TF32 backward).
 * All pointers DEVICE.  idx may be NULL (rows 0..n-1).  *_params / g_*: the network's parameters / gradients flattened in
 * PyTorch order (layer0.weight [out][in], layer0.bias, ...); the g_* buffers and out are ZEROED by the call.
 * out[0..2] = sum surrogate, sum value loss, sum KL over the n samples; g_std includes the entropy term. */
typedef struct {
  int32_t n, obs_dim, act_dim, use_clipped_value_loss;
  float clip, value_coef, entropy_coef, pad0;
  const int64_t* idx;
  const float* obs; const float* critic_obs; const float* actions; const float* old_logp; const float* old_mu; const float* old_sigma;
  const float* adv; const float* ret; const float* tgt_val;
  const float* actor_params; const float* critic_params; const float* std;
  float* g_actor; float* g_critic; float* g_std; float* out;
} nm_ppo_grad_args;
int  nm_ppo_grad(const nm_mlp_shape* actor, const nm_mlp_shape* critic, const nm_ppo_grad_args* args, nm_stream stream);

/* ---- optimiser tail of one PPO mini-batch (≙ rsl_rl v1.0.2 PPO.update after loss.backward(), train.py:54): KL-adaptive
 * learning rate (x / 1.5 on KL outside [desired/2, 2*desired], bounds [1e-5, 1e-2]), clip_grad_norm_ and torch.optim.Adam's
 * update, over flat DEVICE float32 vectors, in one single-CTA launch.  sums = out[] of nm_ppo_grad (sum surrogate, sum value
 * loss, sum KL), n_samples its n.  step and lr are DEVICE scalars (in/out); loss_acc[0..1] += mean value loss, mean surrogate
 * (may be NULL).  grads are left clipped. */
typedef struct {
  int32_t n_params, n_samples, adaptive, pad0;
  float desired_kl, max_grad_norm, beta1, beta2, eps, pad1, pad2, pad3;
  float* params; float* grads; float* exp_avg; float* exp_avg_sq; float* step; float* lr; const float* sums; float* loss_acc;
} nm_ppo_adam_args;
int  nm_ppo_adam(const nm_ppo_adam_args* args, nm_stream stream);

/* ---- GAE(lambda) over a stored rollout (≙ rsl_rl v1.0.2 RolloutStorage.compute_returns, train.py:54).  All pointers
 * DEVICE; rewards / values / returns / advantages float32 [T, n], dones uint8 [T, n], last_values float32 [n].
 * Writes returns and the RAW advantages (returns - values) and moments[0..1] = their sum and sum of squares (fp64),
 * from which the caller normalises them. */
int  nm_gae(int T, int n, const float* rewards, const uint8_t* dones, const float* values, const float* last_values, float gamma,
            float lam, float* returns, float* advantages, double* moments, nm_stream stream);

/* ---------------------------------------------------------------------------------------------------------------------
 * Second model family of the reference: models/anymal_c (BASELINE configs[3]) -- MuJoCo's Newton solver with elliptic
 * cones (anymal_c.xml:4), joint damping + friction loss (:9), condim-6 feet with priority (:20-21), position actuators with
 * a force range (:26), joint limits, box / cylinder / sphere geoms against the plane, Euler with implicit joint damping.
 * Same call sites as above (mj.MjModel.from_xml_path, mj.mj_step: envs/nightmare_v3_env.py:37,200; simple_test.py:39).
 * Scope: a free-floating base plus hinge joints, collisions with ONE static plane (the model's geom-geom self collisions
 * are not generated).  Models with solver="PGS" go through nm_model_* / nm_step above; each loader names the other one
 * in its error message. */
typedef struct nm_gen_model nm_gen_model;
typedef struct nm_gen_batch nm_gen_batch;
/* ≙ mj.MjModel.from_xml_path (compiled .nmb buffer, see nm_model_from_buffer) */
int  nm_gen_model_from_buffer(const void* data, size_t nbytes, nm_gen_model** out);
void nm_gen_model_destroy(nm_gen_model*);
/* ≙ model.nq / nv / nu / nbody; "ngeom" = collision geoms kept (plane excluded) */
int  nm_gen_model_size(const nm_gen_model*, const char* what);
/* ≙ model.opt.timestep */
double nm_gen_model_timestep(const nm_gen_model*);
/* ≙ model.qpos0 */
int  nm_gen_model_qpos0(const nm_gen_model*, float* out, int cap);
/* ≙ [mj.MjData(model) for _ in range(num_envs)]: state in caller-owned DEVICE float32 buffers qpos [N,nq], qvel [N,nv],
 * warm [N,nv] (qacc_warmstart); info: DEVICE int32 [N,4] = {ncon, nefc, Newton iterations, overflow flag} of the last
 * substep, or NULL. */
int  nm_gen_batch_create(const nm_gen_model*, int num_envs, int device, float* qpos, float* qvel, float* warm, int32_t* info,
                         nm_gen_batch** out);
void nm_gen_batch_destroy(nm_gen_batch*);
/* ≙ for i: data[i].ctrl = ctrl[i]; mj.mj_step(model, data[i], nstep)      ctrl: DEVICE float32 [N,nu] */
int  nm_gen_physics_step(nm_gen_batch*, const float* ctrl, int nstep, nm_stream stream);
int64_t nm_gen_batch_launches(const nm_gen_batch*);

/* number of kernel launches issued by this batch so far (bench.py "gpu_launches") */
int64_t nm_batch_launches(const nm_batch*);

/* Measurement utility (no reference counterpart): FFMA micro-benchmark on the current device, returns
 * TFLOP/s (2 flops per FMA).  bench.py uses it as the FP32-pipe roofline denominator. */
double nm_measure_fp32_peak(nm_stream stream);

#ifdef __cplusplus
}
#endif
#endif
