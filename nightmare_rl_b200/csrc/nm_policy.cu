// nm_policy.cu — batched forward pass of the PPO actor-critic MLPs on tensor cores + Gaussian action sampling.
//
// Replaces, per rollout step, what rsl_rl v1.0.2's PPO.act does with three PyTorch module calls
// (reference call sites: train.py:40,54 -> OnPolicyRunner.learn -> alg.act; play.py:122 `nn.act(obs)`):
//   mean   = actor(obs)              66 -> 54 -> 42 -> 30 -> 18, ELU     (envs/nightmare_v3_config.py:107-109)
//   value  = critic(obs)             66 -> 54 -> 42 -> 30 -> 1,  ELU
//   action = mean + std * eps,  log_prob = sum_j log N(action_j; mean_j, std_j)
// in ONE launch: each warp owns a tile of 16 environments, both networks' weights live in shared memory (staged
// once per CTA), activations ping-pong between two per-warp shared-memory tiles, and every layer is a chain of
// mma.sync.m16n8k8 TF32 tensor-core instructions with fp32 accumulation, issued as 3xTF32 (hi/lo split of both
// operands) so that the result matches the fp32 autograd path PPO.update compares it with.  The GEMMs are tiny
// (15 kFLOP per env and network): the kernel is latency/bandwidth bound (264 B in, ~160 B out per env) and the
// tensor cores are used because the contraction is dense, not because they are the bottleneck.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/nightmare_b200.h"

#define NMP_MAXL 6           // layers per network
#define NMP_MAXW 128         // widest layer (padded)
#define NMP_WARPS 4          // warps per CTA = 64 environments
#define NMP_LDA (NMP_MAXW + 4)   // activation tile leading dimension: == 4 (mod 32) -> conflict-free A fragments

int nm_fail(int code, const std::string& msg);   // nm_abi.cu

struct NmpNet {
  int nl;
  int kin[NMP_MAXL], kout[NMP_MAXL], kpad[NMP_MAXL], npad[NMP_MAXL], ldw[NMP_MAXL];
  int woff[NMP_MAXL], boff[NMP_MAXL];   // offsets (floats) into the packed buffer
  int src_w[NMP_MAXL], src_b[NMP_MAXL]; // offsets into the PyTorch-layout flat parameter vector
  int total, src_total;
};

struct NmpArgs {
  NmpNet actor, critic;
  const float* packed;       // [actor.total + critic.total + act_dim] packed weights, biases, then std
  int packed_floats;
  const float* obs; int obs_stride; int n;
  unsigned long long seed; long long step, env_offset;
  int deterministic, act_dim;
  float* actions; float* mean; float* value; float* logp;
  float* obs_copy; float* sigma_out;     // optional: row of the rollout buffer that receives the observations / std
};

__device__ __forceinline__ unsigned f2tf32(float x) {
  unsigned r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void mma_tf32(float* c, const unsigned* a, unsigned b0, unsigned b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Philox4x32-10, same definition as the step kernel (nm_kernels.cu)
__device__ __noinline__ void nmp_philox(unsigned k0, unsigned k1, unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned* out) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    unsigned hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    unsigned hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    unsigned n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// One MLP over a 16-row tile held in `xin` (smem, [16][NMP_LDA], zero padded to kpad[0]); result left in the
// returned tile (first kout[last] columns valid).
__device__ __forceinline__ float* mlp_tile(const NmpNet& net, const float* wbase, float* xin, float* xout, int lane) {
  const int g = lane >> 2, t = lane & 3;
  for (int l = 0; l < net.nl; l++) {
    const float* W = wbase + net.woff[l];
    const float* B = wbase + net.boff[l];
    const int ldw = net.ldw[l], nk = net.kpad[l] >> 3, nn = net.npad[l] >> 3;
    const bool last = l == net.nl - 1;
    for (int nt = 0; nt < nn; nt++) {
      const int col = nt * 8 + 2 * t;
      float c[4] = {B[col], B[col + 1], B[col], B[col + 1]};
      for (int kt = 0; kt < nk; kt++) {
        // 3xTF32: x = hi + lo with both halves representable in TF32; hi*hi + hi*lo + lo*hi recovers fp32-level
        // accuracy (the dropped lo*lo term is ~2^-22 relative), so rollout log-probs agree with the fp32 autograd path
        const int kc = kt * 8 + t;
        const float af[4] = {xin[g * NMP_LDA + kc], xin[(g + 8) * NMP_LDA + kc], xin[g * NMP_LDA + kc + 4], xin[(g + 8) * NMP_LDA + kc + 4]};
        const float bf0 = W[kc * ldw + nt * 8 + g], bf1 = W[(kc + 4) * ldw + nt * 8 + g];
        unsigned ah[4], al[4];
#pragma unroll
        for (int i = 0; i < 4; i++) { ah[i] = f2tf32(af[i]); al[i] = f2tf32(af[i] - __uint_as_float(ah[i])); }
        const unsigned bh0 = f2tf32(bf0), bh1 = f2tf32(bf1);
        const unsigned bl0 = f2tf32(bf0 - __uint_as_float(bh0)), bl1 = f2tf32(bf1 - __uint_as_float(bh1));
        mma_tf32(c, al, bh0, bh1);
        mma_tf32(c, ah, bl0, bl1);
        mma_tf32(c, ah, bh0, bh1);
      }
      if (!last) {
#pragma unroll
        for (int i = 0; i < 4; i++) c[i] = c[i] > 0.f ? c[i] : expm1f(c[i]);      // ELU(alpha = 1)
      }
      xout[g * NMP_LDA + col] = c[0]; xout[g * NMP_LDA + col + 1] = c[1];
      xout[(g + 8) * NMP_LDA + col] = c[2]; xout[(g + 8) * NMP_LDA + col + 1] = c[3];
    }
    __syncwarp();
    float* tmp = xin; xin = xout; xout = tmp;
  }
  return xin;
}

__global__ void __launch_bounds__(NMP_WARPS * 32) nm_policy_kernel(const NmpArgs A) {
  extern __shared__ __align__(16) float smem[];
  float* wsm = smem;                                            // packed weights of both nets + std
  float* tiles = smem + ((A.packed_floats + 3) & ~3);           // per warp: 2 activation tiles + 16 log-prob slots
  for (int i = threadIdx.x; i < A.packed_floats; i += blockDim.x) wsm[i] = __ldg(A.packed + i);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* t0 = tiles + warp * (2 * 16 * NMP_LDA + 16);
  float* t1 = t0 + 16 * NMP_LDA;
  float* lps = t1 + 16 * NMP_LDA;
  const int row0 = (blockIdx.x * NMP_WARPS + warp) * 16;
  if (row0 >= A.n) return;
  const int kin = A.actor.kin[0], kp = A.actor.kpad[0];
  // stage the observations of the 16 envs (zero padded); kept in t0 for the critic pass as well
  for (int idx = lane; idx < 16 * kp; idx += 32) {
    const int r = idx / kp, c = idx - r * kp;
    const int e = row0 + r;
    t0[r * NMP_LDA + c] = (c < kin && e < A.n) ? A.obs[(size_t)e * A.obs_stride + c] : 0.f;
  }
  if (lane < 16) lps[lane] = 0.f;
  __syncwarp();
  if (A.obs_copy != nullptr) {                               // the rollout buffer's copy of what the policy saw
    for (int idx = lane; idx < 16 * kin; idx += 32) {
      const int r = idx / kin, c = idx - r * kin;
      if (row0 + r < A.n) A.obs_copy[(size_t)(row0 + r) * kin + c] = t0[r * NMP_LDA + c];
    }
  }
  // ---- critic first, then the actor.  The ping-pong overwrites the observation tile, and the networks are tiny,
  // so the observations are simply staged a second time for the actor pass.
  float* vout = mlp_tile(A.critic, wsm + A.actor.total, t0, t1, lane);
  if (lane < 16 && row0 + lane < A.n) A.value[row0 + lane] = vout[lane * NMP_LDA];
  __syncwarp();
  for (int idx = lane; idx < 16 * kp; idx += 32) {
    const int r = idx / kp, c = idx - r * kp;
    const int e = row0 + r;
    t0[r * NMP_LDA + c] = (c < kin && e < A.n) ? A.obs[(size_t)e * A.obs_stride + c] : 0.f;
  }
  __syncwarp();
  float* mout = mlp_tile(A.actor, wsm, t0, t1, lane);
  // ---- Gaussian sampling: one Philox block (4 normals) per (env, group of 4 action columns)
  const float* stdv = wsm + A.actor.total + A.critic.total;
  const int ngrp = (A.act_dim + 3) >> 2;
  for (int task = lane; task < 16 * ngrp; task += 32) {
    const int r = task / ngrp, q = task - r * ngrp;
    const int e = row0 + r;
    if (e >= A.n) continue;
    float z[4] = {0.f, 0.f, 0.f, 0.f};
    if (!A.deterministic) {
      unsigned rn[4];
      const long long genv = A.env_offset + e;
      nmp_philox((unsigned)A.seed, (unsigned)genv, (unsigned)A.step, (unsigned)((unsigned long long)A.step >> 32), 0x40000000u + (unsigned)q,
                 (unsigned)(A.seed >> 32) ^ (unsigned)((unsigned long long)genv >> 32), rn);
      // Box-Muller on (0,1] uniforms
      const float u0 = ((float)(rn[0] >> 8) + 1.f) * (1.f / 16777216.f), u1 = (float)(rn[1] >> 8) * (1.f / 16777216.f);
      const float u2 = ((float)(rn[2] >> 8) + 1.f) * (1.f / 16777216.f), u3 = (float)(rn[3] >> 8) * (1.f / 16777216.f);
      const float ra = sqrtf(-2.f * logf(u0)), rb = sqrtf(-2.f * logf(u2));
      float s0, c0, s1, c1;
      sincospif(2.f * u1, &s0, &c0);
      sincospif(2.f * u3, &s1, &c1);
      z[0] = ra * c0; z[1] = ra * s0; z[2] = rb * c1; z[3] = rb * s1;
    }
    float lp = 0.f;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int j = q * 4 + k;
      if (j >= A.act_dim) break;
      const float m = mout[r * NMP_LDA + j], s = stdv[j];
      A.mean[(size_t)e * A.act_dim + j] = m;
      A.actions[(size_t)e * A.act_dim + j] = fmaf(s, z[k], m);
      if (A.sigma_out != nullptr) A.sigma_out[(size_t)e * A.act_dim + j] = s;
      lp += -0.5f * z[k] * z[k] - logf(s) - 0.91893853320467274f;
    }
    atomicAdd(lps + r, lp);
  }
  __syncwarp();
  if (lane < 16 && row0 + lane < A.n) A.logp[row0 + lane] = lps[lane];
}

// pack PyTorch-layout parameters ([out][in] weights, [out] biases) into the transposed, zero-padded layout
__global__ void nm_policy_pack_kernel(NmpNet net, const float* src, float* dst) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < net.total; i += gridDim.x * blockDim.x) {
    float v = 0.f;
    for (int l = 0; l < net.nl; l++) {
      if (i >= net.woff[l] && i < net.woff[l] + net.kpad[l] * net.ldw[l]) {
        const int k = (i - net.woff[l]) / net.ldw[l], n = (i - net.woff[l]) % net.ldw[l];
        if (k < net.kin[l] && n < net.kout[l]) v = src[net.src_w[l] + n * net.kin[l] + k];
      } else if (i >= net.boff[l] && i < net.boff[l] + net.npad[l]) {
        const int n = i - net.boff[l];
        if (n < net.kout[l]) v = src[net.src_b[l] + n];
      }
    }
    dst[i] = v;
  }
}

// ------------------------------------------------------------------------------------------------ host side / C ABI
struct nm_policy {
  NmpNet actor, critic;
  int act_dim, obs_dim, device;
  float* d_packed;
  int packed_floats;
  size_t smem_bytes;
  int64_t launches;
};

static int layout_net(const nm_mlp_shape* s, NmpNet& n) {
  memset(&n, 0, sizeof(n));
  if (!s || s->num_layers < 1 || s->num_layers > NMP_MAXL) return -1;
  n.nl = s->num_layers;
  int off = 0, src = 0;
  for (int l = 0; l < n.nl; l++) {
    n.kin[l] = s->dims[l]; n.kout[l] = s->dims[l + 1];
    if (n.kin[l] < 1 || n.kout[l] < 1 || n.kin[l] > NMP_MAXW || n.kout[l] > NMP_MAXW) return -1;
    n.kpad[l] = (n.kin[l] + 7) & ~7;
    n.npad[l] = (n.kout[l] + 7) & ~7;
    if (l > 0 && n.kpad[l] != n.npad[l - 1]) return -1;
    n.ldw[l] = ((n.npad[l] + 23) / 32) * 32 + 8;          // == 8 (mod 32): conflict-free B fragments
    n.woff[l] = off; off += n.kpad[l] * n.ldw[l];
    n.boff[l] = off; off += n.npad[l];
    n.src_w[l] = src; src += n.kin[l] * n.kout[l];
    n.src_b[l] = src; src += n.kout[l];
  }
  n.total = (off + 3) & ~3;
  n.src_total = src;
  return 0;
}

extern "C" int nm_policy_create(const nm_mlp_shape* actor, const nm_mlp_shape* critic, int device, nm_policy** out) {
  if (!actor || !critic || !out) return nm_fail(NM_ERR_ARG, "nm_policy_create: null argument");
  nm_policy* p = new nm_policy();
  memset(p, 0, sizeof(*p));
  if (layout_net(actor, p->actor) != 0 || layout_net(critic, p->critic) != 0) {
    delete p;
    return nm_fail(NM_ERR_UNSUPPORTED, "nm_policy_create: 1..6 layers of width 1..128 supported");
  }
  if (p->actor.kin[0] != p->critic.kin[0]) { delete p; return nm_fail(NM_ERR_UNSUPPORTED, "actor and critic must read the same observation"); }
  p->obs_dim = p->actor.kin[0];
  p->act_dim = p->actor.kout[p->actor.nl - 1];
  if (p->act_dim > 64) { delete p; return nm_fail(NM_ERR_UNSUPPORTED, "at most 64 actions"); }
  p->device = device;
  p->packed_floats = p->actor.total + p->critic.total + ((p->act_dim + 3) & ~3);
  p->smem_bytes = sizeof(float) * (size_t)(((p->packed_floats + 3) & ~3) + NMP_WARPS * (2 * 16 * NMP_LDA + 16));
  if (p->smem_bytes > 227 * 1024) { delete p; return nm_fail(NM_ERR_UNSUPPORTED, "policy weights do not fit in shared memory"); }
  if (cudaSetDevice(device) != cudaSuccess || cudaMalloc(&p->d_packed, sizeof(float) * p->packed_floats) != cudaSuccess ||
      cudaMemset(p->d_packed, 0, sizeof(float) * p->packed_floats) != cudaSuccess ||
      cudaFuncSetAttribute(nm_policy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem_bytes) != cudaSuccess) {
    delete p;
    return nm_fail(NM_ERR_CUDA, "nm_policy_create: CUDA allocation failed");
  }
  *out = p;
  return NM_OK;
}

extern "C" void nm_policy_destroy(nm_policy* p) {
  if (!p) return;
  cudaFree(p->d_packed);
  delete p;
}

extern "C" int nm_policy_param_count(const nm_policy* p, int which) {
  if (!p) return -1;
  return which == 0 ? p->actor.src_total : (which == 1 ? p->critic.src_total : p->act_dim);
}

extern "C" int nm_policy_load_weights(nm_policy* p, const float* actor_params, const float* critic_params, const float* std, nm_stream stream) {
  if (!p || !actor_params || !critic_params || !std) return nm_fail(NM_ERR_ARG, "nm_policy_load_weights: null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  nm_policy_pack_kernel<<<32, 256, 0, st>>>(p->actor, actor_params, p->d_packed);
  nm_policy_pack_kernel<<<32, 256, 0, st>>>(p->critic, critic_params, p->d_packed + p->actor.total);
  if (cudaMemcpyAsync(p->d_packed + p->actor.total + p->critic.total, std, sizeof(float) * p->act_dim, cudaMemcpyDeviceToDevice, st) != cudaSuccess ||
      cudaGetLastError() != cudaSuccess)
    return nm_fail(NM_ERR_CUDA, "nm_policy_load_weights: launch failed");
  p->launches += 2;
  return NM_OK;
}

extern "C" int nm_policy_act_store(nm_policy* p, const float* obs, int obs_stride, int n, uint64_t seed, int64_t step, int64_t env_offset,
                                   int deterministic, float* actions, float* mean, float* value, float* logp, float* obs_copy,
                                   float* sigma_out, nm_stream stream);

extern "C" int nm_policy_act(nm_policy* p, const float* obs, int obs_stride, int n, uint64_t seed, int64_t step, int64_t env_offset,
                             int deterministic, float* actions, float* mean, float* value, float* logp, nm_stream stream) {
  return nm_policy_act_store(p, obs, obs_stride, n, seed, step, env_offset, deterministic, actions, mean, value, logp, nullptr, nullptr, stream);
}

extern "C" int nm_policy_act_store(nm_policy* p, const float* obs, int obs_stride, int n, uint64_t seed, int64_t step, int64_t env_offset,
                                   int deterministic, float* actions, float* mean, float* value, float* logp, float* obs_copy,
                                   float* sigma_out, nm_stream stream) {
  if (!p || !obs || !actions || !mean || !value || !logp || n <= 0) return nm_fail(NM_ERR_ARG, "nm_policy_act: bad argument");
  if (obs_stride < p->obs_dim) return nm_fail(NM_ERR_ARG, "nm_policy_act: obs_stride smaller than the observation size");
  NmpArgs a;
  a.actor = p->actor; a.critic = p->critic; a.packed = p->d_packed; a.packed_floats = p->packed_floats;
  a.obs = obs; a.obs_stride = obs_stride; a.n = n; a.seed = seed; a.step = step; a.env_offset = env_offset;
  a.deterministic = deterministic; a.act_dim = p->act_dim;
  a.actions = actions; a.mean = mean; a.value = value; a.logp = logp; a.obs_copy = obs_copy; a.sigma_out = sigma_out;
  const int per_cta = NMP_WARPS * 16;
  nm_policy_kernel<<<(n + per_cta - 1) / per_cta, NMP_WARPS * 32, p->smem_bytes, static_cast<cudaStream_t>(stream)>>>(a);
  p->launches++;
  if (cudaGetLastError() != cudaSuccess) return nm_fail(NM_ERR_CUDA, "nm_policy_act: launch failed");
  return NM_OK;
}

extern "C" int64_t nm_policy_launches(const nm_policy* p) { return p ? p->launches : 0; }

// ================================================================================================ rollout slot store
// One launch per env step instead of ~30 tiny PyTorch kernels: writes the transition into row t of the rollout buffer
// (≙ rsl_rl v1.0.2 PPO.process_env_step + RolloutStorage.add_transitions: reward bootstrapped by gamma * V * time_out,
// dones narrowed to uint8) and keeps the runner's episode statistics (running reward / length per env, ring buffer of
// the last `ring_cap` finished episodes) on the device.
__global__ void nm_rollout_store_kernel(const nm_rollout_slot S) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  const int n_obs = S.n * S.obs_dim, n_act = S.n * S.act_dim;
  if (S.obs != nullptr)
    for (int i = tid; i < n_obs; i += nth) S.s_obs[i] = S.obs[i];
  for (int i = tid; i < n_act; i += nth) {
    if (S.actions) S.s_actions[i] = S.actions[i];
    if (S.mean) S.s_mu[i] = S.mean[i];
    if (S.std) S.s_sigma[i] = S.std[i % S.act_dim];
  }
  if (S.ep_acc != nullptr && tid <= S.n_ep) {              // running sum of the env's per-step episode means (runner log)
    if (tid < S.n_ep) S.ep_acc[tid] += S.ep_means[tid];
    else S.ep_acc[S.n_ep] += 1.f;
  }
  for (int e = tid; e < S.n; e += nth) {
    const float v = S.value ? S.value[e] : S.s_values[e];
    if (S.value) S.s_values[e] = v;
    if (S.logp) S.s_logp[e] = S.logp[e];
    const float r = S.rew[e];
    const bool d = S.done[e] != 0;
    S.s_rewards[e] = S.time_outs ? fmaf(S.gamma * v, S.time_outs[e], r) : r;
    S.s_dones[e] = d ? 1 : 0;
    if (S.cur_rew) {
      const float cr = S.cur_rew[e] + r, cl = S.cur_len[e] + 1.f;
      if (d) {
        const unsigned long long k = atomicAdd(reinterpret_cast<unsigned long long*>(S.ring_count), 1ull);
        const int pos = (int)(k % (unsigned long long)S.ring_cap);
        S.ring_rew[pos] = cr; S.ring_len[pos] = cl;
        S.cur_rew[e] = 0.f; S.cur_len[e] = 0.f;
      } else { S.cur_rew[e] = cr; S.cur_len[e] = cl; }
    }
  }
}

extern "C" int nm_rollout_store(const nm_rollout_slot* slot, nm_stream stream) {
  if (!slot || slot->n <= 0 || !slot->rew || !slot->done || !slot->s_obs || !slot->s_sigma || !slot->s_values || !slot->s_rewards || !slot->s_dones)
    return nm_fail(NM_ERR_ARG, "nm_rollout_store: missing buffer");
  if (slot->cur_rew && (!slot->cur_len || !slot->ring_rew || !slot->ring_len || !slot->ring_count || slot->ring_cap <= 0))
    return nm_fail(NM_ERR_ARG, "nm_rollout_store: incomplete episode-statistics buffers");
  const int work = slot->n * slot->obs_dim;
  int blocks = (work + 255) / 256;
  if (blocks > 1184) blocks = 1184;
  nm_rollout_store_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(*slot);
  if (cudaGetLastError() != cudaSuccess) return nm_fail(NM_ERR_CUDA, "nm_rollout_store: launch failed");
  return NM_OK;
}

// ================================================================================================ PPO loss head
// ≙ the distribution / loss part of rsl_rl v1.0.2 PPO.update (reached from train.py:54): Normal log-prob of the stored
// actions, probability ratio, clipped surrogate, clipped value loss, entropy bonus and the KL estimate of the adaptive
// learning-rate rule — forward value AND its gradients w.r.t. the network outputs (mu, value) and the std parameter, in
// one launch, one thread per sample.  It replaces ~120 element-wise / reduction kernels of the autograd graph per
// mini-batch; the MLPs themselves stay with autograd + cuBLAS.
//   out[0] = sum surrogate, out[1] = sum value loss, out[2] = sum KL   (caller divides by n)
__global__ void nm_ppo_head_kernel(const nm_ppo_head_args H) {
  __shared__ float s_red[3];
  __shared__ float s_gstd[64];
  if (threadIdx.x < 3) s_red[threadIdx.x] = 0.f;
  if (threadIdx.x < 64) s_gstd[threadIdx.x] = 0.f;
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int A = H.act_dim;
  float surr = 0.f, vloss = 0.f, kl = 0.f;
  if (i < H.n) {
    const float inv_n = 1.f / (float)H.n;
    float logp = 0.f;
    for (int j = 0; j < A; j++) {
      const float s = H.std[j], m = H.mu[(size_t)i * A + j];
      const float z = (H.actions[(size_t)i * A + j] - m) / s;
      logp += -0.5f * z * z - logf(s) - 0.91893853320467274f;
      const float os = H.old_sigma[(size_t)i * A + j], dm = H.old_mu[(size_t)i * A + j] - m;
      kl += logf(s / os + 1.0e-5f) + (os * os + dm * dm) / (2.f * s * s) - 0.5f;
    }
    const float adv = H.adv[i];
    const float ratio = expf(logp - H.old_logp[i]);
    const float lo = 1.f - H.clip, hi = 1.f + H.clip;
    const float rc = fminf(fmaxf(ratio, lo), hi);
    const float s1 = -adv * ratio, s2 = -adv * rc;
    surr = fmaxf(s1, s2);
    // d surr / d logp: torch.max routes ties to its first argument (the unclipped term); clamp passes the gradient inside [lo, hi]
    float g = 0.f;
    if (s1 >= s2) g = s1;                                   // d(-adv*ratio)/dlogp = -adv*ratio
    else if (ratio >= lo && ratio <= hi) g = s2;
    g *= inv_n;
    for (int j = 0; j < A; j++) {
      const float s = H.std[j];
      const float z = (H.actions[(size_t)i * A + j] - H.mu[(size_t)i * A + j]) / s;
      H.g_mu[(size_t)i * A + j] = g * z / s;
      atomicAdd(s_gstd + j, g * (z * z - 1.f) / s);
    }
    const float v = H.value[i], R = H.ret[i];
    float dv;
    if (H.use_clipped_value_loss) {
      const float tv = H.tgt_val[i];
      const float dcl = fminf(fmaxf(v - tv, -H.clip), H.clip);
      const float vc = tv + dcl;
      const float l1 = (v - R) * (v - R), l2 = (vc - R) * (vc - R);
      vloss = fmaxf(l1, l2);
      if (l1 >= l2) dv = 2.f * (v - R);
      else dv = (v - tv >= -H.clip && v - tv <= H.clip) ? 2.f * (vc - R) : 0.f;
    } else {
      vloss = (R - v) * (R - v);
      dv = 2.f * (v - R);
    }
    H.g_value[i] = H.value_coef * dv * inv_n;
  }
  // block reduction of the three sums, then one atomic per block
  for (int o = 16; o > 0; o >>= 1) {
    surr += __shfl_xor_sync(0xffffffffu, surr, o);
    vloss += __shfl_xor_sync(0xffffffffu, vloss, o);
    kl += __shfl_xor_sync(0xffffffffu, kl, o);
  }
  if ((threadIdx.x & 31) == 0) { atomicAdd(s_red, surr); atomicAdd(s_red + 1, vloss); atomicAdd(s_red + 2, kl); }
  __syncthreads();
  if (threadIdx.x < 3) atomicAdd(H.out + threadIdx.x, s_red[threadIdx.x]);
  if (threadIdx.x < A) atomicAdd(H.g_std + threadIdx.x, s_gstd[threadIdx.x]);
}

extern "C" int nm_ppo_head(const nm_ppo_head_args* h, nm_stream stream) {
  if (!h || h->n <= 0 || h->act_dim < 1 || h->act_dim > 64 || !h->mu || !h->value || !h->std || !h->actions || !h->old_logp || !h->old_mu ||
      !h->old_sigma || !h->adv || !h->ret || !h->tgt_val || !h->out || !h->g_mu || !h->g_value || !h->g_std)
    return nm_fail(NM_ERR_ARG, "nm_ppo_head: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (cudaMemsetAsync(h->out, 0, 3 * sizeof(float), st) != cudaSuccess || cudaMemsetAsync(h->g_std, 0, h->act_dim * sizeof(float), st) != cudaSuccess)
    return nm_fail(NM_ERR_CUDA, "nm_ppo_head: memset failed");
  nm_ppo_head_kernel<<<(h->n + 127) / 128, 128, 0, st>>>(*h);
  if (cudaGetLastError() != cudaSuccess) return nm_fail(NM_ERR_CUDA, "nm_ppo_head: launch failed");
  return NM_OK;
}
