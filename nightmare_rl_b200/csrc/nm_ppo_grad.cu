// nm_ppo_grad.cu — one PPO mini-batch: gather, both MLPs forward, loss head, both MLPs backward, in ONE launch.
//
// ≙ the body of rsl_rl v1.0.2 PPO.update's mini-batch loop up to `loss.backward()` (reached from the reference's
// train.py:54 `ppo_runner.learn`; hyper-parameters envs/nightmare_v3_config.py:117-146): index the rollout buffers with
// the mini-batch's random rows, evaluate actor and critic, Normal log-prob of the stored actions, probability ratio,
// clipped surrogate, clipped value loss, entropy bonus, KL estimate, and the gradient of
//   loss = mean(surrogate) + value_coef * mean(value loss) - entropy_coef * mean(entropy)
// w.r.t. every network parameter and the std vector.  With autograd that is ~200 kernels per mini-batch (gathers, 16
// small GEMMs, 8 split-K weight-gradient GEMMs, element-wise and reduction kernels) and 1.35 ms at 81 920 samples; the
// networks are so small (66-54-42-30-18 / 1, 15 k parameters) that everything fits in one SM's shared memory.
//
// Grid: one persistent CTA (16 warps) per SM; the first `actor_ctas` CTAs work on the actor, the others on the critic, each walking
// 128-sample batches.  Per CTA
// the net's weights, transposed and zero padded, and a gradient accumulator of the same shape live in shared memory for
// the whole launch; every PAIR of warps owns 16 samples of the batch, keeps all of their layer activations in shared
// memory and splits the output tiles of each layer between its two warps (shared memory caps the CTA at 8 sample tiles;
// the second warp per tile doubles the warps the schedulers can choose from).
//   forward   per warp pair, mma.sync.m16n8k8 TF32 issued 3x on hi/lo splits (fp32-level accuracy: the probability ratio
//             must be 1 when the policy has not changed), ELU, activations of every layer kept
//   head      four lanes per sample (8 samples per warp); writes d(loss)/d(output) over the output tile
//   backward  layer by layer: bias gradient = column sums (shared-memory atomics); weight gradient = H^T dZ over the
//             CTA's 128 samples (K = 128) with each 16x8 output tile owned by one warp, accumulated in shared memory
//             without atomics; input gradient per warp pair, ELU derivative applied from the stored activation, written in
//             place over that activation (single-pass TF32, like the TF32 autograd backward it replaces)
//   end       one pass of global atomic adds per CTA into the PyTorch-layout gradient vectors
#include <cuda_runtime.h>

#include <cstdlib>
#include <cstring>
#include <string>

#include "../../include/nightmare_b200.h"

#define PG_MAXL 6
#define PG_MAXW 128
#define PG_WARPS 8            // 16-sample tiles per CTA batch
#define PG_ROWS (PG_WARPS * 16)
#define PG_THREADS (PG_WARPS * 64)   // TWO warps per tile: they split the output tiles of every layer between them

int nm_fail(int code, const std::string& msg);   // nm_abi.cu

struct PgNet {
  int nl;
  int kin[PG_MAXL], kout[PG_MAXL], kpad[PG_MAXL], npad[PG_MAXL], ldw[PG_MAXL], lda[PG_MAXL];
  int woff[PG_MAXL], boff[PG_MAXL];     // offsets into the packed shared-memory image (weights [kpad][ldw], bias [npad])
  int src_w[PG_MAXL], src_b[PG_MAXL];   // offsets into the PyTorch-layout flat parameter vector
  int aoff[PG_MAXL];                    // offset of layer l's INPUT tile inside a warp's activation region
  int ooff, ldo;                        // output tile of the last layer
  int region;                           // floats per warp region
  int total;                            // floats of the packed image
};

struct PgArgs {
  PgNet net[2];
  nm_ppo_grad_args a;
  int actor_ctas;                       // CTAs [0, actor_ctas) work on the actor, the rest on the critic
};

// TF32 operands.  `cvt.rna.tf32.f32` costs ~4 instructions on sm_100a (first ncu capture: 21 % of all executed
// instructions), and the tensor core ignores the 13 low mantissa bits of an fp32 register anyway, so:
//   * single-pass (backward) operands get half a TF32 ulp added to their bit pattern (one IADD), which the hardware's
//     truncation turns into round-to-nearest -- plain truncation biases the gradients (4x the error in the parity test);
//   * the 3xTF32 forward splits x = hi + lo with hi = x & 0xffffe000 (exactly representable in TF32), lo = x - hi
//     (exact in fp32, |lo| < 2^-10 |x|), and lets the hardware truncate lo: the dropped part is < 2^-20 |x|.
__device__ __forceinline__ unsigned pg_raw(float x) { return __float_as_uint(x); }
__device__ __forceinline__ unsigned pg_rn(float x) { return __float_as_uint(x) + 0x1000u; }    // round to nearest (ties away) once truncated
__device__ __forceinline__ unsigned pg_hi(float x) { return __float_as_uint(x) & 0xffffe000u; }
__device__ __forceinline__ void pg_mma(float* c, const unsigned* a, unsigned b0, unsigned b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void pg_prefetch(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// barrier of the two warps that share a tile (ids 1..PG_WARPS; 0 is __syncthreads)
__device__ __forceinline__ void pg_pair_sync(int tile) { asm volatile("bar.sync %0, 64;" ::"r"(tile + 1) : "memory"); }

__global__ void __launch_bounds__(PG_THREADS, 1) nm_ppo_grad_kernel(const PgArgs P) {
  extern __shared__ __align__(16) float smem[];
  const int which = (int)blockIdx.x >= P.actor_ctas ? 1 : 0;
  const int cta = which ? (int)blockIdx.x - P.actor_ctas : (int)blockIdx.x;         // index and count among this network's CTAs
  const int nctas = which ? (int)gridDim.x - P.actor_ctas : P.actor_ctas;
  const PgNet& N = P.net[which];
  const nm_ppo_grad_args& A = P.a;
  const float* __restrict__ src = which ? A.critic_params : A.actor_params;
  const float* __restrict__ xin = which ? A.critic_obs : A.obs;
  float* W = smem;
  float* G = W + N.total;
  float* act = G + N.total;
  float* s_std = act + PG_WARPS * N.region;    // [64]
  float* s_gstd = s_std + 64;                  // [64]
  float* s_red = s_gstd + 64;                  // [4]
  const int tid = threadIdx.x, wid = tid >> 5, warp = wid >> 1, hw = wid & 1, lane = tid & 31, g = lane >> 2, t = lane & 3;
  // `warp` = tile index (16 samples), `hw` = which of the tile's two warps this is, `wid` = warp index in the CTA
  const int nl = N.nl;

  for (int i = tid; i < 2 * N.total; i += blockDim.x) W[i] = 0.f;          // W and G are adjacent
  if (tid < 64) { s_std[tid] = tid < A.act_dim ? A.std[tid] : 1.f; s_gstd[tid] = 0.f; }
  if (tid < 4) s_red[tid] = 0.f;
  __syncthreads();
  for (int l = 0; l < nl; l++) {
    const int kin = N.kin[l], kout = N.kout[l];
    for (int j = tid; j < kin * kout; j += blockDim.x) {
      const int n = j / kin, k = j - n * kin;
      W[N.woff[l] + k * N.ldw[l] + n] = __ldg(src + N.src_w[l] + j);
    }
    for (int n = tid; n < kout; n += blockDim.x) W[N.boff[l] + n] = __ldg(src + N.src_b[l] + n);
  }
  __syncthreads();

  float* R = act + warp * N.region;
  float* O = R + N.ooff;
  const int ldo = N.ldo;
  const float inv_n = 1.f / (float)A.n;
  float acc_s = 0.f, acc_v = 0.f, acc_k = 0.f;

  for (int batch = cta; batch * PG_ROWS < A.n; batch += nctas) {
    const int row0 = batch * PG_ROWS + warp * 16;
    // ---------------------------------------------------------------- gather the 16 observation rows of this warp
    long long ri = -1;
    if (lane < 16 && row0 + lane < A.n) ri = A.idx ? A.idx[row0 + lane] : (long long)(row0 + lane);
    {                                                          // pull the NEXT batch's rows towards L2 while this one computes
      const int nrow = row0 + nctas * PG_ROWS + lane;
      if (hw == 0 && lane < 16 && nrow < A.n) {
        const long long rn = A.idx ? A.idx[nrow] : (long long)nrow;
        const char* po = reinterpret_cast<const char*>(xin + rn * A.obs_dim);
        const int ob = A.obs_dim * 4;
        for (int q = 0; q < ob; q += 128) pg_prefetch(po + q);
        pg_prefetch(po + ob - 1);                              // rows are not line aligned: the tail may sit in one more line
        if (which == 0) {
          const size_t ab = (size_t)rn * A.act_dim * 4;
          const int bb = A.act_dim * 4;
          for (int q = 0; q < bb + 127; q += 128) {
            const int o = q < bb ? q : bb - 1;
            pg_prefetch(reinterpret_cast<const char*>(A.actions) + ab + o);
            pg_prefetch(reinterpret_cast<const char*>(A.old_mu) + ab + o);
            pg_prefetch(reinterpret_cast<const char*>(A.old_sigma) + ab + o);
          }
          pg_prefetch(A.old_logp + rn); pg_prefetch(A.adv + rn);
        } else {
          pg_prefetch(A.ret + rn); pg_prefetch(A.tgt_val + rn);
        }
      }
    }
    {
      const int kin = N.kin[0], kp = N.kpad[0], lda = N.lda[0];
      {                                                        // each warp of the pair stages 8 rows; 32 independent loads per lane
        const int r0 = hw * 8;
        float v[8][4];
#pragma unroll
        for (int rr = 0; rr < 8; rr++) {
          const long long row = __shfl_sync(0xffffffffu, ri, r0 + rr);
#pragma unroll
          for (int q = 0; q < 4; q++) {
            const int c = lane + 32 * q;
            v[rr][q] = (c < kin && row >= 0) ? __ldg(xin + row * A.obs_dim + c) : 0.f;
          }
        }
#pragma unroll
        for (int rr = 0; rr < 8; rr++)
#pragma unroll
          for (int q = 0; q < 4; q++) {
            const int c = lane + 32 * q;
            if (c < kp) R[(r0 + rr) * lda + c] = v[rr][q];
          }
      }
    }
    pg_pair_sync(warp);
    // ---------------------------------------------------------------- forward, all activations kept
    for (int l = 0; l < nl; l++) {
      const float* Wl = W + N.woff[l];
      const float* Bl = W + N.boff[l];
      const float* in = R + N.aoff[l];
      const int ldi = N.lda[l], ldw = N.ldw[l], nk = N.kpad[l] >> 3, nn = N.npad[l] >> 3;
      const bool last = l == nl - 1;
      float* out = last ? O : R + N.aoff[l + 1];
      const int ldout = last ? ldo : N.lda[l + 1];
      const int nhalf = (nn + 1) >> 1, nbeg = hw ? nhalf : 0, nend = hw ? nn : nhalf;      // this warp's share of the output tiles
      for (int nt0 = nbeg; nt0 < nend; nt0 += 4) {             // 4 output tiles per pass share the A fragments
        const int cnt = nend - nt0 < 4 ? nend - nt0 : 4;
        float ch[4][4], cl[4][4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const int col = (nt0 + q) * 8 + 2 * t;
          const float b0 = q < cnt ? Bl[col] : 0.f, b1 = q < cnt ? Bl[col + 1] : 0.f;
          ch[q][0] = b0; ch[q][1] = b1; ch[q][2] = b0; ch[q][3] = b1;
          cl[q][0] = cl[q][1] = cl[q][2] = cl[q][3] = 0.f;
        }
#pragma unroll 2
        for (int kt = 0; kt < nk; kt++) {
          const int kc = kt * 8 + t;
          const float af[4] = {in[g * ldi + kc], in[(g + 8) * ldi + kc], in[g * ldi + kc + 4], in[(g + 8) * ldi + kc + 4]};
          unsigned ah[4], al[4];
#pragma unroll
          for (int i = 0; i < 4; i++) { ah[i] = pg_hi(af[i]); al[i] = pg_raw(af[i] - __uint_as_float(ah[i])); }
#pragma unroll
          for (int q = 0; q < 4; q++) {
            if (q < cnt) {
              const float bf0 = Wl[kc * ldw + (nt0 + q) * 8 + g], bf1 = Wl[(kc + 4) * ldw + (nt0 + q) * 8 + g];
              const unsigned bh0 = pg_hi(bf0), bh1 = pg_hi(bf1);
              const unsigned bl0 = pg_raw(bf0 - __uint_as_float(bh0)), bl1 = pg_raw(bf1 - __uint_as_float(bh1));
              pg_mma(cl[q], al, bh0, bh1);                    // 3xTF32: small terms in their own accumulator chain
              pg_mma(cl[q], ah, bl0, bl1);
              pg_mma(ch[q], ah, bh0, bh1);
            }
          }
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
          if (q < cnt) {
            const int col = (nt0 + q) * 8 + 2 * t;
            float c[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
              c[i] = ch[q][i] + cl[q][i];
              if (!last) c[i] = c[i] > 0.f ? c[i] : __expf(c[i]) - 1.f;      // ELU; |abs error| ~1e-7, expm1f costs 10 % of the kernel
            }
            out[g * ldout + col] = c[0]; out[g * ldout + col + 1] = c[1];
            out[(g + 8) * ldout + col] = c[2]; out[(g + 8) * ldout + col + 1] = c[3];
          }
        }
      }
      pg_pair_sync(warp);
    }
    // ---------------------------------------------------------------- loss head: O <- d(loss)/d(output)
    if (which == 0) {
      const int r = hw * 8 + (lane >> 2), h = lane & 3;        // 4 lanes per sample, 8 samples per warp
      const long long rr = __shfl_sync(0xffffffffu, ri, r);
      const bool valid = rr >= 0;
      const int Ad = A.act_dim;
      float logp = 0.f, kl = 0.f;
      const int quarter_n = (Ad + 3) >> 2;
      if (valid) {
#pragma unroll 3
        for (int j = h; j < Ad; j += 4) {
          const float s = s_std[j], m = O[r * ldo + j];
          const float z = (__ldg(A.actions + rr * Ad + j) - m) / s;
          logp += -0.5f * z * z - logf(s) - 0.91893853320467274f;
          const float os = __ldg(A.old_sigma + rr * Ad + j), dm = __ldg(A.old_mu + rr * Ad + j) - m;
          kl += logf(s / os + 1.0e-5f) + (os * os + dm * dm) / (2.f * s * s) - 0.5f;
          O[r * ldo + j] = z;                                  // the mean is not needed again; keep z for the gradient
        }
      }
      logp += __shfl_xor_sync(0xffffffffu, logp, 1); logp += __shfl_xor_sync(0xffffffffu, logp, 2);
      kl += __shfl_xor_sync(0xffffffffu, kl, 1); kl += __shfl_xor_sync(0xffffffffu, kl, 2);
      float gl = 0.f;
      if (valid) {
        const float adv = __ldg(A.adv + rr);
        const float ratio = expf(logp - __ldg(A.old_logp + rr));
        const float lo = 1.f - A.clip, hi = 1.f + A.clip;
        const float rc = fminf(fmaxf(ratio, lo), hi);
        const float s1 = -adv * ratio, s2 = -adv * rc;
        // torch.max routes ties to the unclipped term; clamp passes the gradient on [lo, hi]
        if (s1 >= s2) gl = s1;
        else if (ratio >= lo && ratio <= hi) gl = s2;
        gl *= inv_n;
        if (h == 0) { acc_s += fmaxf(s1, s2); acc_k += kl; }
      }
      for (int q = 0; q < quarter_n; q++) {                    // same trip count in every lane: the shuffles below are warp-wide
        const int j = 4 * q + h;
        const bool ok = j < Ad;
        float gm = 0.f, gs = 0.f;
        if (ok && valid) {
          const float s = s_std[j], z = O[r * ldo + j];
          gm = gl * z / s;
          gs = gl * (z * z - 1.f) / s;
        }
        if (ok) O[r * ldo + j] = gm;
        for (int o = 4; o < 32; o <<= 1) gs += __shfl_xor_sync(0xffffffffu, gs, o);      // over this warp's 8 samples, per action column
        if ((lane >> 2) == 0 && ok) atomicAdd(s_gstd + j, gs);
      }
    } else {
      const int r = hw * 8 + lane;
      const long long rc_ = __shfl_sync(0xffffffffu, ri, r & 15);
      if (lane < 8) {
        const long long ri = rc_;
        float gv = 0.f;
        if (ri >= 0) {
          const float v = O[r * ldo], Rt = __ldg(A.ret + ri);
          float dv;
          if (A.use_clipped_value_loss) {
            const float tv = __ldg(A.tgt_val + ri);
            const float dcl = fminf(fmaxf(v - tv, -A.clip), A.clip);
            const float vc = tv + dcl;
            const float l1 = (v - Rt) * (v - Rt), l2 = (vc - Rt) * (vc - Rt);
            acc_v += fmaxf(l1, l2);
            if (l1 >= l2) dv = 2.f * (v - Rt);
            else dv = (v - tv >= -A.clip && v - tv <= A.clip) ? 2.f * (vc - Rt) : 0.f;
          } else {
            acc_v += (Rt - v) * (Rt - v);
            dv = 2.f * (v - Rt);
          }
          gv = A.value_coef * dv * inv_n;
        }
        O[r * ldo] = gv;
      }
    }
    pg_pair_sync(warp);
    // ---------------------------------------------------------------- backward
    for (int l = nl - 1; l >= 0; l--) {
      const bool last = l == nl - 1;
      const int doff = last ? N.ooff : N.aoff[l + 1];         // dZ_l sits where layer l's output was
      const int ldd = last ? ldo : N.lda[l + 1];
      const int np = N.npad[l], kp = N.kpad[l], lda = N.lda[l], ldw = N.ldw[l];
      {                                                        // bias gradient: column sums of this warp's 16 rows
        const float* D = R + doff;
        for (int n = lane + 32 * hw; n < np; n += 64) {
          float s = 0.f;
#pragma unroll
          for (int r = 0; r < 16; r++) s += D[r * ldd + n];
          atomicAdd(G + N.boff[l] + n, s);
        }
      }
      __syncthreads();                                         // dZ_l of all 8 warps is in place
      {                                                        // weight gradient over the CTA's 128 samples
        const int mt = (kp + 15) >> 4, nt = np >> 3;
        float* Gl = G + N.woff[l];
        const int npair = (nt + 1) >> 1;                       // units of two neighbouring 16x8 tiles share the A fragments
        for (int unit = wid; unit < mt * npair; unit += 2 * PG_WARPS) {
          const int mi = unit / npair, pj = unit - mi * npair;
          const int m0 = mi * 16, n0 = pj * 16;
          const bool hi_ok = m0 + 8 < kp;                      // rows m0+8.. exist (kp is a multiple of 8)
          const bool two = n0 + 8 < np;
          float c[2][2][4];                                    // [tile][k half][fragment]: four independent accumulator chains
#pragma unroll
          for (int i = 0; i < 16; i++) (&c[0][0][0])[i] = 0.f;
#pragma unroll 2
          for (int reg = 0; reg < PG_WARPS; reg++) {
            const float* Hr = act + reg * N.region + N.aoff[l];
            const float* Dr = act + reg * N.region + doff;
#pragma unroll
            for (int half = 0; half < 2; half++) {
              // the contraction runs over samples, in any order: k slot t <-> row 2t, slot t+4 <-> row 2t+1 makes the
              // four rows a warp touches per load fall into different banks (lda, ldd = 4 * odd)
              const int r0 = half * 8 + 2 * t;
              unsigned a[4];
              a[0] = pg_rn(Hr[r0 * lda + m0 + g]);
              a[1] = hi_ok ? pg_rn(Hr[r0 * lda + m0 + g + 8]) : 0u;
              a[2] = pg_rn(Hr[(r0 + 1) * lda + m0 + g]);
              a[3] = hi_ok ? pg_rn(Hr[(r0 + 1) * lda + m0 + g + 8]) : 0u;
              const unsigned b0 = pg_rn(Dr[r0 * ldd + n0 + g]), b1 = pg_rn(Dr[(r0 + 1) * ldd + n0 + g]);
              pg_mma(c[0][half], a, b0, b1);
              if (two) {
                const unsigned b2 = pg_rn(Dr[r0 * ldd + n0 + 8 + g]), b3 = pg_rn(Dr[(r0 + 1) * ldd + n0 + 8 + g]);
                pg_mma(c[1][half], a, b2, b3);
              }
            }
          }
#pragma unroll
          for (int q = 0; q < 2; q++) {
            if (q == 0 || two) {
              float* p0 = Gl + (m0 + g) * ldw + n0 + q * 8 + 2 * t;
              p0[0] += c[q][0][0] + c[q][1][0]; p0[1] += c[q][0][1] + c[q][1][1];
              if (hi_ok) { float* p1 = p0 + 8 * ldw; p1[0] += c[q][0][2] + c[q][1][2]; p1[1] += c[q][0][3] + c[q][1][3]; }
            }
          }
        }
      }
      __syncthreads();                                         // every warp is done reading H_l and dZ_l
      if (l > 0) {                                             // input gradient, in place over H_l (post-ELU values)
        const float* D = R + doff;
        const float* Wl = W + N.woff[l];
        float* H = R + N.aoff[l];
        const int nk = np >> 3, nc = kp >> 3;
        const int chalf = (nc + 1) >> 1, cbeg = hw ? chalf : 0, cend = hw ? nc : chalf;
        for (int ct0 = cbeg; ct0 < cend; ct0 += 4) {           // 4 output tiles per pass share the dZ fragments
          const int cnt = cend - ct0 < 4 ? cend - ct0 : 4;
          float c[4][4];
#pragma unroll
          for (int i = 0; i < 16; i++) (&c[0][0])[i] = 0.f;
#pragma unroll 2
          for (int kt = 0; kt < nk; kt++) {
            const int kc = kt * 8 + t;
            unsigned a[4];
            a[0] = pg_rn(D[g * ldd + kc]); a[1] = pg_rn(D[(g + 8) * ldd + kc]);
            a[2] = pg_rn(D[g * ldd + kc + 4]); a[3] = pg_rn(D[(g + 8) * ldd + kc + 4]);
#pragma unroll
            for (int q = 0; q < 4; q++) {
              if (q < cnt) {
                const float* wr = Wl + ((ct0 + q) * 8 + g) * ldw + kc;
                pg_mma(c[q], a, pg_rn(wr[0]), pg_rn(wr[4]));
              }
            }
          }
#pragma unroll
          for (int q = 0; q < 4; q++) {
            if (q < cnt) {
              const int col = (ct0 + q) * 8 + 2 * t;
              float* h0 = H + g * lda + col;
              float* h1 = H + (g + 8) * lda + col;
              h0[0] = c[q][0] * (h0[0] > 0.f ? 1.f : h0[0] + 1.f);
              h0[1] = c[q][1] * (h0[1] > 0.f ? 1.f : h0[1] + 1.f);
              h1[0] = c[q][2] * (h1[0] > 0.f ? 1.f : h1[0] + 1.f);
              h1[1] = c[q][3] * (h1[1] > 0.f ? 1.f : h1[1] + 1.f);
            }
          }
        }
        pg_pair_sync(warp);
      }
    }
  }
  // -------------------------------------------------------------------- flush
  for (int o = 16; o > 0; o >>= 1) {
    acc_s += __shfl_xor_sync(0xffffffffu, acc_s, o);
    acc_v += __shfl_xor_sync(0xffffffffu, acc_v, o);
    acc_k += __shfl_xor_sync(0xffffffffu, acc_k, o);
  }
  if (lane == 0) { atomicAdd(s_red, acc_s); atomicAdd(s_red + 1, acc_v); atomicAdd(s_red + 2, acc_k); }
  __syncthreads();
  float* gdst = which ? A.g_critic : A.g_actor;
  for (int l = 0; l < nl; l++) {
    const int kin = N.kin[l], kout = N.kout[l];
    for (int j = tid; j < kin * kout; j += blockDim.x) {
      const int n = j / kin, k = j - n * kin;
      atomicAdd(gdst + N.src_w[l] + j, G[N.woff[l] + k * N.ldw[l] + n]);
    }
    for (int n = tid; n < kout; n += blockDim.x) atomicAdd(gdst + N.src_b[l] + n, G[N.boff[l] + n]);
  }
  if (which == 0) {
    if (tid < A.act_dim) {
      float v = s_gstd[tid];
      if (cta == 0) v += -A.entropy_coef / s_std[tid];          // d(-entropy_coef * mean entropy)/d std, sample independent
      atomicAdd(A.g_std + tid, v);
    }
    if (tid == 0) { atomicAdd(A.out, s_red[0]); atomicAdd(A.out + 2, s_red[2]); }
  } else if (tid == 0) {
    atomicAdd(A.out + 1, s_red[1]);
  }
}

static int pg_layout(const nm_mlp_shape* s, PgNet& n) {
  memset(&n, 0, sizeof(n));
  if (!s || s->num_layers < 1 || s->num_layers > PG_MAXL) return -1;
  n.nl = s->num_layers;
  int off = 0, src = 0, a = 0;
  for (int l = 0; l < n.nl; l++) {
    n.kin[l] = s->dims[l]; n.kout[l] = s->dims[l + 1];
    if (n.kin[l] < 1 || n.kout[l] < 1 || n.kin[l] > PG_MAXW || n.kout[l] > PG_MAXW) return -1;
    n.kpad[l] = (n.kin[l] + 7) & ~7;
    n.npad[l] = (n.kout[l] + 7) & ~7;
    if (l > 0 && n.kpad[l] != n.npad[l - 1]) return -1;
    n.ldw[l] = n.npad[l] + ((n.npad[l] & 15) == 8 ? 0 : 8);      // == 8 (mod 16): B fragments of the forward pass do not collide
    n.lda[l] = n.kpad[l] + 4;                                    // == 4 (mod 8): conflict-free A fragments
    n.woff[l] = off; off += n.kpad[l] * n.ldw[l];
    n.boff[l] = off; off += n.npad[l];
    n.src_w[l] = src; src += n.kin[l] * n.kout[l];
    n.src_b[l] = src; src += n.kout[l];
    n.aoff[l] = a; a += 16 * n.lda[l];
  }
  n.ooff = a;
  n.ldo = n.npad[n.nl - 1] + 4;
  n.region = (a + 16 * n.ldo + 3) & ~3;
  n.total = (off + 3) & ~3;
  return 0;
}

static size_t pg_smem(const PgNet& n) { return sizeof(float) * (size_t)(2 * n.total + PG_WARPS * n.region + 64 + 64 + 4); }

extern "C" int nm_ppo_grad(const nm_mlp_shape* actor, const nm_mlp_shape* critic, const nm_ppo_grad_args* a, nm_stream stream) {
  if (!actor || !critic || !a) return nm_fail(NM_ERR_ARG, "nm_ppo_grad: null argument");
  if (a->n <= 0 || !a->obs || !a->critic_obs || !a->actions || !a->old_logp || !a->old_mu || !a->old_sigma || !a->adv || !a->ret || !a->tgt_val ||
      !a->actor_params || !a->critic_params || !a->std || !a->g_actor || !a->g_critic || !a->g_std || !a->out)
    return nm_fail(NM_ERR_ARG, "nm_ppo_grad: bad argument");
  PgArgs P;
  if (pg_layout(actor, P.net[0]) != 0 || pg_layout(critic, P.net[1]) != 0)
    return nm_fail(NM_ERR_UNSUPPORTED, "nm_ppo_grad: 1..6 layers of width 1..128 supported");
  const PgNet &na = P.net[0], &nc = P.net[1];
  if (na.kin[0] != a->obs_dim || nc.kin[0] != a->obs_dim || na.kout[na.nl - 1] != a->act_dim || nc.kout[nc.nl - 1] != 1 || a->act_dim > 64)
    return nm_fail(NM_ERR_ARG, "nm_ppo_grad: network shapes do not match obs_dim / act_dim (<= 64) / scalar value");
  const size_t smem = pg_smem(na) > pg_smem(nc) ? pg_smem(na) : pg_smem(nc);
  if (smem > 227 * 1024) return nm_fail(NM_ERR_UNSUPPORTED, "nm_ppo_grad: networks do not fit in shared memory");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int dev = 0, sms = 0;
  static int attr_dev_mask = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
    return nm_fail(NM_ERR_CUDA, "nm_ppo_grad: no device");
  if (!(attr_dev_mask & (1 << (dev & 31)))) {
    if (cudaFuncSetAttribute(nm_ppo_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
      return nm_fail(NM_ERR_CUDA, "nm_ppo_grad: cannot raise the shared-memory limit");
    attr_dev_mask |= 1 << (dev & 31);
  }
  const int pa = na.src_w[na.nl - 1] + na.kin[na.nl - 1] * na.kout[na.nl - 1] + na.kout[na.nl - 1];
  const int pc = nc.src_w[nc.nl - 1] + nc.kin[nc.nl - 1] * nc.kout[nc.nl - 1] + nc.kout[nc.nl - 1];
  if (a->g_std + a->act_dim == a->g_actor && a->g_actor + pa == a->g_critic && a->g_critic + pc == a->out) {
    // std | actor | critic | sums in one buffer (how the Python side lays them out): one memset node instead of four
    if (cudaMemsetAsync(a->g_std, 0, sizeof(float) * (a->act_dim + pa + pc + 4), st) != cudaSuccess) return nm_fail(NM_ERR_CUDA, "nm_ppo_grad: memset failed");
  } else if (cudaMemsetAsync(a->g_actor, 0, sizeof(float) * pa, st) != cudaSuccess || cudaMemsetAsync(a->g_critic, 0, sizeof(float) * pc, st) != cudaSuccess ||
             cudaMemsetAsync(a->g_std, 0, sizeof(float) * a->act_dim, st) != cudaSuccess || cudaMemsetAsync(a->out, 0, sizeof(float) * 4, st) != cudaSuccess) {
    return nm_fail(NM_ERR_CUDA, "nm_ppo_grad: memset failed");
  }
  P.a = *a;
  const int batches = (a->n + PG_ROWS - 1) / PG_ROWS;
  // one CTA per SM; the actor (wider output layer, Gaussian head) gets the larger share of the SMs
  int n_act = (sms * 9 + 8) / 16, n_cri = sms - n_act;
  if (const char* e = getenv("NM_PG_ACTOR_CTAS")) { const int v = atoi(e); if (v > 0 && v < sms) { n_act = v; n_cri = sms - v; } }
  if (n_cri < 1) { n_act = 1; n_cri = 1; }
  if (n_act > batches) n_act = batches;
  if (n_cri > batches) n_cri = batches;
  P.actor_ctas = n_act;
  nm_ppo_grad_kernel<<<n_act + n_cri, PG_THREADS, smem, st>>>(P);
  if (cudaGetLastError() != cudaSuccess) return nm_fail(NM_ERR_CUDA, "nm_ppo_grad: launch failed");
  return NM_OK;
}

// ================================================================================================ GAE(lambda)
// ≙ rsl_rl v1.0.2 RolloutStorage.compute_returns (reached from train.py:54; gamma / lam from
// envs/nightmare_v3_config.py:122-123): the backward recursion over the T stored steps, one thread per environment
// (coalesced across environments), plus the sum and sum of squares of the raw advantages in fp64 so that the caller
// can normalise them (locally or, after an all-reduce, globally).  The PyTorch loop is ~10 kernels per step: 9.6 ms for
// 80 steps, against ~0.1 ms here.
__global__ void nm_gae_kernel(int T, int n, const float* __restrict__ rewards, const unsigned char* __restrict__ dones,
                              const float* __restrict__ values, const float* __restrict__ last_values, float gamma, float lam,
                              float* __restrict__ returns, float* __restrict__ advantages, double* moments) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  double s = 0.0, ss = 0.0;
  if (e < n) {
    float adv = 0.f, next_v = last_values[e];
#pragma unroll 8
    for (int t = T - 1; t >= 0; t--) {
      const size_t i = (size_t)t * n + e;
      const float v = values[i];
      const float nt = 1.f - (float)dones[i];
      const float delta = rewards[i] + nt * gamma * next_v - v;
      adv = delta + nt * gamma * lam * adv;
      const float ret = adv + v;
      returns[i] = ret;
      const float a = ret - v;                               // what rsl_rl normalises: returns - values, as stored in fp32
      advantages[i] = a;
      s += (double)a; ss += (double)a * (double)a;
      next_v = v;
    }
  }
  for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); ss += __shfl_xor_sync(0xffffffffu, ss, o); }
  if ((threadIdx.x & 31) == 0) { atomicAdd(moments, s); atomicAdd(moments + 1, ss); }
}

extern "C" int nm_gae(int T, int n, const float* rewards, const uint8_t* dones, const float* values, const float* last_values, float gamma,
                      float lam, float* returns, float* advantages, double* moments, nm_stream stream) {
  if (T <= 0 || n <= 0 || !rewards || !dones || !values || !last_values || !returns || !advantages || !moments)
    return nm_fail(NM_ERR_ARG, "nm_gae: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (cudaMemsetAsync(moments, 0, 2 * sizeof(double), st) != cudaSuccess) return nm_fail(NM_ERR_CUDA, "nm_gae: memset failed");
  nm_gae_kernel<<<(n + 63) / 64, 64, 0, st>>>(T, n, rewards, dones, values, last_values, gamma, lam, returns, advantages, moments);
  if (cudaGetLastError() != cudaSuccess) return nm_fail(NM_ERR_CUDA, "nm_gae: launch failed");
  return NM_OK;
}

// ================================================================================================ optimiser tail
// Everything rsl_rl v1.0.2 PPO.update does per mini-batch AFTER loss.backward() (train.py:54 -> PPO.update): the KL-adaptive
// learning rate (desired_kl rule, bounds [1e-5, 1e-2], envs/nightmare_v3_config.py:125-129), clip_grad_norm_(max_grad_norm)
// and torch.optim.Adam's step -- for the 15 k parameters of this policy ~20 tiny PyTorch kernels, here ONE single-CTA launch
// over the flat parameter / gradient / moment vectors.  Same arithmetic as torch: bias-corrected first and second
// moments, denom = sqrt(v) / sqrt(1 - beta2^t) + eps, p -= lr / (1 - beta1^t) * m / denom.
#define PA_PER 16            // vector elements per thread held in registers on the fast path (n_params <= 16 * 1024)

__global__ void __launch_bounds__(1024, 1) nm_ppo_adam_kernel(const nm_ppo_adam_args A) {
  __shared__ double s_part[32];
  __shared__ float s_coef, s_lr, s_bc1, s_bc2s;
  const int tid = threadIdx.x;
  const bool fast = A.n_params <= PA_PER * 1024;
  // fast path: gradients and parameters live in registers across the norm reduction, loads are issued in batches of 32
  float g[PA_PER], p[PA_PER];
  double ss = 0.0;
  if (fast) {
#pragma unroll
    for (int k = 0; k < PA_PER; k++) {
      const int i = tid + k * 1024;
      const bool ok = i < A.n_params;
      g[k] = ok ? A.grads[i] : 0.f;
      p[k] = ok ? A.params[i] : 0.f;
    }
#pragma unroll
    for (int k = 0; k < PA_PER; k++) ss += (double)g[k] * (double)g[k];
  } else {
    for (int i = tid; i < A.n_params; i += blockDim.x) { const float gi = A.grads[i]; ss += (double)gi * (double)gi; }
  }
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if ((tid & 31) == 0) s_part[tid >> 5] = ss;
  __syncthreads();
  if (tid == 0) {
    double tot = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) tot += s_part[w];
    const float norm = (float)sqrt(tot);
    s_coef = fminf(A.max_grad_norm / (norm + 1e-6f), 1.f);
    float lr = *A.lr;
    const float inv_n = 1.f / (float)A.n_samples;
    if (A.adaptive) {
      const float kl = A.sums[2] * inv_n;
      if (kl > A.desired_kl * 2.f) lr = fmaxf(lr / 1.5f, 1e-5f);
      else if (kl < A.desired_kl * 0.5f && kl > 0.f) lr = fminf(lr * 1.5f, 1e-2f);
      *A.lr = lr;
    }
    const float t = *A.step + 1.f;
    *A.step = t;
    s_lr = lr;
    s_bc1 = 1.f - powf(A.beta1, t);
    s_bc2s = sqrtf(1.f - powf(A.beta2, t));
    if (A.loss_acc) { A.loss_acc[0] += A.sums[1] * inv_n; A.loss_acc[1] += A.sums[0] * inv_n; }
  }
  __syncthreads();
  const float coef = s_coef, step_size = s_lr / s_bc1, bc2s = s_bc2s;
  // torch.optim.Adam: exp_avg.lerp_(grad, 1 - beta1); exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2);
  //                   param.addcdiv_(exp_avg, exp_avg_sq.sqrt() / sqrt(1 - beta2^t) + eps, -lr / (1 - beta1^t))
  if (fast) {
    float m[PA_PER], v[PA_PER];                              // (64 registers per thread at 1024 threads: the moments come second)
#pragma unroll
    for (int k = 0; k < PA_PER; k++) {
      const int i = tid + k * 1024;
      const bool ok = i < A.n_params;
      m[k] = ok ? A.exp_avg[i] : 0.f;
      v[k] = ok ? A.exp_avg_sq[i] : 0.f;
    }
#pragma unroll
    for (int k = 0; k < PA_PER; k++) {
      const int i = tid + k * 1024;
      if (i < A.n_params) {
        const float gi = g[k] * coef;
        const float mi = m[k] + (gi - m[k]) * (1.f - A.beta1);
        const float vi = v[k] * A.beta2 + (1.f - A.beta2) * gi * gi;
        A.exp_avg[i] = mi;
        A.exp_avg_sq[i] = vi;
        A.grads[i] = gi;
        A.params[i] = p[k] - step_size * (mi / (sqrtf(vi) / bc2s + A.eps));
      }
    }
  } else {
    for (int i = tid; i < A.n_params; i += blockDim.x) {
      const float gi = A.grads[i] * coef;
      const float mi = A.exp_avg[i] + (gi - A.exp_avg[i]) * (1.f - A.beta1);
      const float vi = A.exp_avg_sq[i] * A.beta2 + (1.f - A.beta2) * gi * gi;
      A.exp_avg[i] = mi;
      A.exp_avg_sq[i] = vi;
      A.grads[i] = gi;
      A.params[i] -= step_size * (mi / (sqrtf(vi) / bc2s + A.eps));
    }
  }
}

extern "C" int nm_ppo_adam(const nm_ppo_adam_args* a, nm_stream stream) {
  if (!a || a->n_params <= 0 || a->n_samples <= 0 || !a->params || !a->grads || !a->exp_avg || !a->exp_avg_sq || !a->step || !a->lr || !a->sums)
    return nm_fail(NM_ERR_ARG, "nm_ppo_adam: bad argument");
  nm_ppo_adam_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(*a);
  if (cudaGetLastError() != cudaSuccess) return nm_fail(NM_ERR_CUDA, "nm_ppo_adam: launch failed");
  return NM_OK;
}
