// nm_policy_tc5.cu — the PPO actor-critic forward pass on Blackwell's 5th-generation tensor cores (tcgen05 + TMEM).
//
// Same contract as nm_policy.cu (≙ rsl_rl v1.0.2 PPO.act, reference call sites train.py:54 / play.py:122; network
// envs/nightmare_v3_config.py:105-109), different engine: one CTA of 128 threads owns a tile of 128 environments;
// every layer is a handful of `tcgen05.mma.cta_group::1.kind::tf32` instructions issued by ONE thread, with
//   A = activations  [128 x K]  in shared memory (UMMA canonical K-major layout, no swizzle),
//   B = weights      [N x K]    in shared memory (same layout, staged once per network),
//   D = accumulator  [128 x N]  fp32 in TENSOR MEMORY (64 TMEM columns),
// completion is signalled through `tcgen05.commit` -> mbarrier, and the epilogue (bias, ELU, hi/lo split for the next
// layer, or Gaussian sampling after the last one) reads the accumulator back with `tcgen05.ld.32x32b` — thread r of
// the CTA gets row r, i.e. environment r of the tile.  Each MMA is issued three times on TF32 hi/lo splits of both
// operands (3xTF32), which restores fp32-level accuracy so rollout log-probs match the fp32 autograd path of PPO.update.
//
// The contraction is tiny (15 kFLOP per env and network); the kernel is latency bound.  It exists because the one dense
// GEMM of the pipeline belongs on the tensor cores of the machine it runs on (north_star), not for throughput.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstring>
#include <string>

#include "../../include/nightmare_b200.h"

#define T5_ROWS 128
#define T5_THREADS 256
#define T5_MAXL 6
#define T5_TMEM_COLS 64

int nm_fail(int code, const std::string& msg);   // nm_abi.cu

struct T5Layer { int kin, kout, kpad, npad, w_off, b_off, src_w, src_b; };   // w_off: floats into one weight plane
struct T5Net { int nl, plane_floats, bias_floats, src_total; T5Layer L[T5_MAXL]; };

struct T5Args {
  T5Net actor, critic;
  const float* packed;     // [actor: W_hi | W_lo | bias][critic: W_hi | W_lo | bias][std]
  const float* obs; int obs_stride; int n;
  unsigned long long seed; long long step, env_offset;
  int deterministic, act_dim;
  int split;               // 1: grid.y = 2, CTA (x, 0) runs the critic and CTA (x, 1) the actor of tile x (small batches: twice the CTAs, half the layer chain)
  float* actions; float* mean; float* value; float* logp; float* obs_copy; float* sigma_out;
};

// ---------------------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned t5_tf32(float x) { unsigned r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return r; }
// hi/lo split for the per-element hot loops: cvt.rna.tf32 is a ~4-instruction sequence on sm_100a, a mask is one.  hi = x with the
// 13 low mantissa bits cleared (exact TF32), lo = x - hi (exact in fp32); the tensor core ignores lo's own low bits (< 2^-20 |x|).
__device__ __forceinline__ unsigned t5_hi(float x) { return __float_as_uint(x) & 0xffffe000u; }

// UMMA shared-memory descriptor, K-major, SWIZZLE_NONE: 8-row x 16-byte core matrices; LBO = byte distance between the
// two 16-byte K chunks of one MMA, SBO = byte distance between 8-row groups (cute/arch/mma_sm100_desc.hpp SmemDescriptor)
__device__ __forceinline__ unsigned long long t5_desc(unsigned addr, unsigned lbo_bytes, unsigned sbo_bytes) {
  unsigned long long d = 0;
  d |= (unsigned long long)((addr >> 4) & 0x3fffu);
  d |= (unsigned long long)((lbo_bytes >> 4) & 0x3fffu) << 16;
  d |= (unsigned long long)((sbo_bytes >> 4) & 0x3fffu) << 32;
  d |= 1ull << 46;                                           // descriptor version 1 (Blackwell)
  return d;                                                  // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}
// instruction descriptor for kind::tf32: D = F32, A = B = TF32, both K-major, M = 128, N = n
__device__ __forceinline__ unsigned t5_idesc(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(n >> 3) << 17) | ((unsigned)(T5_ROWS >> 4) << 24);
}
__device__ __forceinline__ void t5_mma(unsigned tmem_d, unsigned long long a, unsigned long long b, unsigned idesc, unsigned accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(a), "l"(b), "r"(idesc), "r"(accumulate) : "memory");
}
// issue one 8-column accumulator load; the registers are valid only after t5_ld_wait (which names them, so that neither the
// compiler nor ptxas can schedule a consumer ahead of the wait)
__device__ __forceinline__ void t5_ld8_issue(unsigned taddr, unsigned* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ void t5_ld_wait(unsigned* r /*[32]*/) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), "+r"(r[10]),
                 "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]),
                 "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :: "memory");
}

__device__ __noinline__ void t5_philox(unsigned k0, unsigned k1, unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned* out) {
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

// float index of element (row, col) of a [rows x kpad] operand in the canonical K-major no-swizzle layout
__device__ __forceinline__ int t5_idx(int row, int col, int kpad) { return (row >> 3) * (kpad >> 2) * 32 + (col >> 2) * 32 + (row & 7) * 4 + (col & 3); }

// One network over the CTA's 128-row tile.  On entry A_hi/A_lo hold the observations (layout for kpad of layer 0).
// Hidden layers rewrite A_hi/A_lo; the last layer's accumulator row is returned in out[] (first kout entries).
// one thread: TMA bulk copy (cp.async.bulk, SASS UBLKCP) of a network's packed weights [W_hi | W_lo | bias] into smem
__device__ __forceinline__ void t5_issue_weights(const T5Net& net, const float* gsrc, float* W, unsigned long long* wbar) {
  const unsigned bytes = (unsigned)(2 * net.plane_floats + net.bias_floats) * 4u;
  const unsigned b32 = smem_u32(wbar);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // earlier generic-proxy reads of W precede the async write
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b32), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(W)), "l"(gsrc), "r"(bytes),
               "r"(b32)
               : "memory");
}
__device__ __forceinline__ void t5_wait(unsigned long long* bar, unsigned parity) {
  unsigned done = 0, spins = 0;
  const unsigned b32 = smem_u32(bar);
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(b32), "r"(parity) : "memory");
    if (++spins > (1u << 24)) { asm volatile("trap;"); }                  // never hang the GPU on a bad descriptor
  }
}

// One network over the CTA's 128-row tile.  Layer 0 reads the observation operand (X), hidden layers read and rewrite the
// hidden operand (H; in place is safe: the MMAs of a layer have completed before its epilogue runs).  256 threads: warp w
// owns accumulator lanes 32*(w&3).. (rows of the tile) and the column half (w>>2) of every layer's output; the last
// layer's half row is returned in out[0..16).
__device__ __forceinline__ void t5_net(const T5Net& net, const float* X_hi, const float* X_lo, float* H_hi, float* H_lo, float* W, unsigned tmem,
                                       unsigned long long* bar, unsigned& phase, unsigned long long* wbar, unsigned wparity, float* out) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row = (warp & 3) * 32 + lane, half = warp >> 2;
  t5_wait(wbar, wparity);                                    // this network's weights have landed (bulk copy issued earlier)
  const float* bias = W + 2 * net.plane_floats;
  for (int l = 0; l < net.nl; l++) {
    const T5Layer& L = net.L[l];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic-proxy smem writes -> visible to the tensor core
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const unsigned idesc = t5_idesc(L.npad);
      const unsigned sbo = (unsigned)(L.kpad >> 2) * 128u;
      const unsigned a_hi = smem_u32(l == 0 ? X_hi : H_hi), a_lo = smem_u32(l == 0 ? X_lo : H_lo);
      const unsigned w_hi = smem_u32(W + L.w_off), w_lo = smem_u32(W + net.plane_floats + L.w_off);
      const int nks = L.kpad >> 3;
      for (int ks = 0; ks < nks; ks++) {
        const unsigned off = (unsigned)ks * 256u;                          // 8 TF32 = two 16-byte chunks = 2 x LBO
        const unsigned long long dah = t5_desc(a_hi + off, 128u, sbo), dal = t5_desc(a_lo + off, 128u, sbo);
        const unsigned long long dbh = t5_desc(w_hi + off, 128u, sbo), dbl = t5_desc(w_lo + off, 128u, sbo);
        t5_mma(tmem, dal, dbh, idesc, ks > 0 ? 1u : 0u);
        t5_mma(tmem, dah, dbl, idesc, 1u);
        t5_mma(tmem, dah, dbh, idesc, 1u);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    }
    t5_wait(bar, phase);                                     // everyone waits for the accumulator of this layer
    phase ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const bool last = l == net.nl - 1;
    const int kp_next = L.npad, ncol = L.npad >> 1, cbeg = half * ncol;      // this warp's column range [cbeg, cbeg + ncol)
    const unsigned taddr = tmem + ((unsigned)((warp & 3) * 32) << 16) + (unsigned)cbeg;
    const int rbase = (row >> 3) * (kp_next >> 2) * 32 + (row & 7) * 4;
    unsigned acc[32];                                          // this thread's half row of the accumulator: all loads in flight, one wait
#pragma unroll
    for (int i = 0; i < 32; i++) acc[i] = 0u;
#pragma unroll
    for (int cc = 0; cc < T5_TMEM_COLS / 16; cc++)
      if (cc * 8 < ncol) t5_ld8_issue(taddr + (unsigned)(cc * 8), acc + cc * 8);
    t5_ld_wait(acc);
#pragma unroll
    for (int cc = 0; cc < T5_TMEM_COLS / 16; cc++) {           // compile-time column blocks keep out[] in registers
      if (cc * 8 < ncol) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
          const int col = cbeg + cc * 8 + i;
          float x = __uint_as_float(acc[cc * 8 + i]) + bias[L.b_off + col];
          if (last) { if (cc < 2) out[(cc < 2 ? cc : 0) * 8 + i] = x; }
          else {
            x = (col < L.kout) ? (x > 0.f ? x : __expf(x) - 1.f) : 0.f;      // ELU; padded columns stay exactly zero
            const unsigned h = t5_hi(x);
            const int idx = rbase + (col >> 2) * 32 + (col & 3);
            H_hi[idx] = __uint_as_float(h);
            H_lo[idx] = x - __uint_as_float(h);
          }
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");        // TMEM reads done before the next MMA overwrites D
  }
  __syncthreads();
}

__global__ void __launch_bounds__(T5_THREADS, 1) nm_policy_tc5_kernel(const T5Args A) {
  extern __shared__ __align__(1024) float t5_smem[];
  __shared__ __align__(8) unsigned long long bar, wbar;
  __shared__ unsigned tmem_base_s;
  __shared__ float lp_s[T5_ROWS];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int kp0 = A.actor.L[0].kpad, kin = A.actor.L[0].kin;
  float* X_hi = t5_smem;                                       // observations  [128 x 72], hi / lo TF32 planes
  float* X_lo = X_hi + T5_ROWS * 72;
  float* H_hi = X_lo + T5_ROWS * 72;                           // hidden activations [128 x <=64]
  float* H_lo = H_hi + T5_ROWS * 64;
  float* W = H_lo + T5_ROWS * 64;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"((unsigned)T5_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&wbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < T5_ROWS) lp_s[tid] = 0.f;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tmem = tmem_base_s;
  const int row0 = blockIdx.x * T5_ROWS;
  unsigned phase = 0;
  const int a_floats = 2 * A.actor.plane_floats + A.actor.bias_floats;
  const bool do_critic = !A.split || blockIdx.y == 0, do_actor = !A.split || blockIdx.y == 1;
  if (tid == 0) t5_issue_weights(do_critic ? A.critic : A.actor, do_critic ? A.packed + a_floats : A.packed, W, &wbar);     // lands while the observations are staged
  // observations: split into TF32 hi/lo planes in the UMMA operand layout.  A warp covers 8 rows x 4 columns per trip
  // (lane = 4 * (row & 7) + (col & 3)): that is exactly one shared-memory bank per lane in the canonical layout, the global
  // reads are eight 16-byte row segments, and the 18 trips of a row group are independent loads in flight together (the first
  // version walked a flat index: a division per element, 8-way bank conflicts, one exposed load latency per element --
  // half of the kernel's time at 4096 observations, profiles/r02_policy_tc5_metrics.txt)
  {
    const int nq = kp0 >> 2;                                   // column quads (18 for 66 -> 72 inputs)
    for (int g = warp; g < T5_ROWS / 8; g += T5_THREADS / 32) {
      const int r = g * 8 + (lane >> 2), e = row0 + r, cl = lane & 3;
      const float* src = A.obs + (size_t)(e < A.n ? e : 0) * A.obs_stride;
      float xv[18];
#pragma unroll
      for (int q = 0; q < 18; q++) {
        const int c = q * 4 + cl;
        xv[q] = (q < nq && c < kin && e < A.n) ? __ldg(src + c) : 0.f;
      }
      const int jb = (r >> 3) * nq * 32 + (r & 7) * 4 + cl;
#pragma unroll
      for (int q = 0; q < 18; q++) {
        if (q < nq) {
          const float x = xv[q];
          const unsigned h = t5_tf32(x);
          X_hi[jb + q * 32] = __uint_as_float(h);
          X_lo[jb + q * 32] = x - __uint_as_float(h);          // exact in fp32; the tensor core drops lo's own low bits (< 2^-21 |x|)
          const int c = q * 4 + cl;
          if (A.obs_copy != nullptr && do_critic && c < kin && e < A.n) A.obs_copy[(size_t)e * kin + c] = x;
        }
      }
    }
  }
  float vout[16], mout[16];
  unsigned wpar = 0u;
  if (do_critic) {
    t5_net(A.critic, X_hi, X_lo, H_hi, H_lo, W, tmem, &bar, phase, &wbar, wpar, vout);
    wpar ^= 1u;
    if (do_actor && tid == 0) t5_issue_weights(A.actor, A.packed, W, &wbar);   // critic MMAs are complete: reuse the buffer
  }
  if (do_actor) t5_net(A.actor, X_hi, X_lo, H_hi, H_lo, W, tmem, &bar, phase, &wbar, wpar, mout);

  // ---- epilogue: thread (row, half) = environment row of the tile, action columns [16*half, 16*half + 16)
  const int row = (warp & 3) * 32 + lane, half = warp >> 2;
  const int e = row0 + row;
  if (e < A.n && do_critic && half == 0) A.value[e] = vout[0];
  if (e < A.n && do_actor) {
    const float* stdv = A.packed + a_floats + 2 * A.critic.plane_floats + A.critic.bias_floats;
    float lp = 0.f;
#pragma unroll
    for (int qq = 0; qq < 4; qq++) {
      const int q = half * 4 + qq;                              // Philox block of action columns 4q .. 4q+3
      if (q * 4 < A.act_dim) {
        float z[4] = {0.f, 0.f, 0.f, 0.f};
        if (!A.deterministic) {
          unsigned rn[4];
          const long long genv = A.env_offset + e;
          t5_philox((unsigned)A.seed, (unsigned)genv, (unsigned)A.step, (unsigned)((unsigned long long)A.step >> 32), 0x40000000u + (unsigned)q,
                    (unsigned)(A.seed >> 32) ^ (unsigned)((unsigned long long)genv >> 32), rn);
          const float u0 = ((float)(rn[0] >> 8) + 1.f) * (1.f / 16777216.f), u1 = (float)(rn[1] >> 8) * (1.f / 16777216.f);
          const float u2 = ((float)(rn[2] >> 8) + 1.f) * (1.f / 16777216.f), u3 = (float)(rn[3] >> 8) * (1.f / 16777216.f);
          const float ra = sqrtf(-2.f * logf(u0)), rb = sqrtf(-2.f * logf(u2));
          float s0, c0, s1, c1;
          sincospif(2.f * u1, &s0, &c0);
          sincospif(2.f * u3, &s1, &c1);
          z[0] = ra * c0; z[1] = ra * s0; z[2] = rb * c1; z[3] = rb * s1;
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const int j = q * 4 + k;
          if (j < A.act_dim) {
            const float m = mout[qq * 4 + k];
            const float s = __ldg(stdv + j);
            A.mean[(size_t)e * A.act_dim + j] = m;
            A.actions[(size_t)e * A.act_dim + j] = fmaf(s, z[k], m);
            if (A.sigma_out != nullptr) A.sigma_out[(size_t)e * A.act_dim + j] = s;
            lp += -0.5f * z[k] * z[k] - logf(s) - 0.91893853320467274f;
          }
        }
      }
    }
    atomicAdd(lp_s + row, lp);
  }
  __syncthreads();
  if (do_actor && tid < T5_ROWS && row0 + tid < A.n) A.logp[row0 + tid] = lp_s[tid];
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((unsigned)T5_TMEM_COLS) : "memory");
}

// pack PyTorch-layout parameters into [W_hi | W_lo | bias] with the canonical UMMA layout per layer
__global__ void nm_policy_tc5_pack_kernel(T5Net net, const float* src, float* dst) {
  const int total = 2 * net.plane_floats + net.bias_floats;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    float v = 0.f;
    if (i < 2 * net.plane_floats) {
      const int plane = i / net.plane_floats, j = i - plane * net.plane_floats;
      for (int l = 0; l < net.nl; l++) {
        const T5Layer& L = net.L[l];
        const int sz = L.npad * L.kpad;
        if (j >= L.w_off && j < L.w_off + sz) {
          const int q = j - L.w_off;                             // invert t5_idx
          const int grp = q / ((L.kpad >> 2) * 32), rem = q - grp * (L.kpad >> 2) * 32;
          const int chunk = rem / 32, r2 = rem - chunk * 32;
          const int n = grp * 8 + (r2 >> 2), k = chunk * 4 + (r2 & 3);
          if (n < L.kout && k < L.kin) {
            const float w = src[L.src_w + n * L.kin + k];
            const float h = __uint_as_float(t5_tf32(w));
            v = plane == 0 ? h : __uint_as_float(t5_tf32(w - h));
          }
        }
      }
    } else {
      const int j = i - 2 * net.plane_floats;
      for (int l = 0; l < net.nl; l++) {
        const T5Layer& L = net.L[l];
        if (j >= L.b_off && j < L.b_off + L.npad && j - L.b_off < L.kout) v = src[L.src_b + j - L.b_off];
      }
    }
    dst[i] = v;
  }
}

// ------------------------------------------------------------------------------------------------ host side
struct nm_policy_tc5 {
  T5Net actor, critic;
  int act_dim, obs_dim, sms;
  float* d_packed;
  int packed_floats;
  size_t smem_bytes;
};

static int t5_layout(const nm_mlp_shape* s, T5Net& n, bool is_actor) {
  memset(&n, 0, sizeof(n));
  if (!s || s->num_layers < 1 || s->num_layers > T5_MAXL) return -1;
  n.nl = s->num_layers;
  int woff = 0, boff = 0, src = 0;
  for (int l = 0; l < n.nl; l++) {
    T5Layer& L = n.L[l];
    L.kin = s->dims[l]; L.kout = s->dims[l + 1];
    L.kpad = l == 0 ? ((L.kin + 7) & ~7) : n.L[l - 1].npad;
    L.npad = (L.kout + 15) & ~15;                              // UMMA M=128 needs N % 16 == 0, 16 <= N <= 256
    if (L.kpad > 72 || L.npad > T5_TMEM_COLS || L.kin < 1 || L.kout < 1) return -1;
    if (l == n.nl - 1 && L.npad > 32) return -1;
    if ((L.npad >> 1) % 8 != 0) return -1;                      // the epilogue splits the columns over two warp groups, 8 at a time
    L.w_off = woff; woff += L.npad * L.kpad;
    L.b_off = boff; boff += L.npad;
    L.src_w = src; src += L.kin * L.kout;
    L.src_b = src; src += L.kout;
  }
  (void)is_actor;
  n.plane_floats = (woff + 255) & ~255;                        // keeps every plane 1024-byte aligned
  n.bias_floats = (boff + 255) & ~255;
  n.src_total = src;
  return 0;
}

extern "C" int nm_policy_tc5_create(const nm_mlp_shape* actor, const nm_mlp_shape* critic, int device, nm_policy_tc5** out) {
  if (!actor || !critic || !out) return nm_fail(NM_ERR_ARG, "nm_policy_tc5_create: null argument");
  nm_policy_tc5* p = new nm_policy_tc5();
  memset(p, 0, sizeof(*p));
  if (t5_layout(actor, p->actor, true) != 0 || t5_layout(critic, p->critic, false) != 0 || actor->dims[0] != critic->dims[0]) {
    delete p;
    return nm_fail(NM_ERR_UNSUPPORTED, "nm_policy_tc5_create: needs <= 72 inputs, hidden widths <= 64, <= 32 outputs");
  }
  p->obs_dim = actor->dims[0];
  p->act_dim = actor->dims[actor->num_layers];
  const int af = 2 * p->actor.plane_floats + p->actor.bias_floats, cf = 2 * p->critic.plane_floats + p->critic.bias_floats;
  p->packed_floats = af + cf + 64;
  const int wmax = af > cf ? af : cf;
  p->smem_bytes = sizeof(float) * (size_t)(2 * T5_ROWS * 72 + 2 * T5_ROWS * 64 + wmax) + 1024;
  if (p->smem_bytes > 227 * 1024) { delete p; return nm_fail(NM_ERR_UNSUPPORTED, "nm_policy_tc5_create: weights do not fit in shared memory"); }
  if (cudaSetDevice(device) != cudaSuccess || cudaMalloc(&p->d_packed, sizeof(float) * p->packed_floats) != cudaSuccess ||
      cudaMemset(p->d_packed, 0, sizeof(float) * p->packed_floats) != cudaSuccess ||
      cudaFuncSetAttribute(nm_policy_tc5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem_bytes) != cudaSuccess) {
    delete p;
    return nm_fail(NM_ERR_CUDA, "nm_policy_tc5_create: CUDA allocation failed");
  }
  cudaDeviceProp prop;
  p->sms = cudaGetDeviceProperties(&prop, device) == cudaSuccess ? prop.multiProcessorCount : 148;
  *out = p;
  return NM_OK;
}

extern "C" void nm_policy_tc5_destroy(nm_policy_tc5* p) {
  if (!p) return;
  cudaFree(p->d_packed);
  delete p;
}

extern "C" int nm_policy_tc5_load_weights(nm_policy_tc5* p, const float* actor_params, const float* critic_params, const float* std, nm_stream stream) {
  if (!p || !actor_params || !critic_params || !std) return nm_fail(NM_ERR_ARG, "nm_policy_tc5_load_weights: null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int af = 2 * p->actor.plane_floats + p->actor.bias_floats, cf = 2 * p->critic.plane_floats + p->critic.bias_floats;
  nm_policy_tc5_pack_kernel<<<64, 256, 0, st>>>(p->actor, actor_params, p->d_packed);
  nm_policy_tc5_pack_kernel<<<64, 256, 0, st>>>(p->critic, critic_params, p->d_packed + af);
  if (cudaMemcpyAsync(p->d_packed + af + cf, std, sizeof(float) * p->act_dim, cudaMemcpyDeviceToDevice, st) != cudaSuccess ||
      cudaGetLastError() != cudaSuccess)
    return nm_fail(NM_ERR_CUDA, "nm_policy_tc5_load_weights: launch failed");
  return NM_OK;
}

extern "C" int nm_policy_tc5_act(nm_policy_tc5* p, const float* obs, int obs_stride, int n, uint64_t seed, int64_t step, int64_t env_offset,
                                 int deterministic, float* actions, float* mean, float* value, float* logp, float* obs_copy, float* sigma_out,
                                 nm_stream stream) {
  if (!p || !obs || !actions || !mean || !value || !logp || n <= 0) return nm_fail(NM_ERR_ARG, "nm_policy_tc5_act: bad argument");
  if (obs_stride < p->obs_dim) return nm_fail(NM_ERR_ARG, "nm_policy_tc5_act: obs_stride smaller than the observation size");
  T5Args a;
  a.actor = p->actor; a.critic = p->critic; a.packed = p->d_packed;
  a.obs = obs; a.obs_stride = obs_stride; a.n = n; a.seed = seed; a.step = step; a.env_offset = env_offset;
  a.deterministic = deterministic; a.act_dim = p->act_dim;
  a.actions = actions; a.mean = mean; a.value = value; a.logp = logp; a.obs_copy = obs_copy; a.sigma_out = sigma_out;
  // a batch whose tiles do not fill the SMs twice over runs the two networks of a tile on two CTAs (4096 observations: 64 CTAs
  // with a 4-layer chain each instead of 32 with 8); larger batches keep one CTA per tile (observations staged once)
  const int tiles = (n + T5_ROWS - 1) / T5_ROWS;
  a.split = 2 * tiles <= p->sms ? 1 : 0;
  nm_policy_tc5_kernel<<<dim3(tiles, a.split ? 2 : 1), T5_THREADS, p->smem_bytes, static_cast<cudaStream_t>(stream)>>>(a);
  if (cudaGetLastError() != cudaSuccess) return nm_fail(NM_ERR_CUDA, "nm_policy_tc5_act: launch failed");
  return NM_OK;
}
