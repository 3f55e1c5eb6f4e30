// nm_kernels.cu — the whole environment step of the Nightmare-v3 hexapod in ONE kernel launch (sm_100a).
//
// Replaces, per environment and per call of NightmareV3Env.step (reference envs/nightmare_v3_env.py:145-311):
//   action scaling/clipping + PD law (:152-188), `decimation` x mj_step (:200; MuJoCo 3.1.2 pipeline:
//   kinematics, comPos, CRBA, factorisation, plane/hull collision, pyramidal contact rows, RNE, velocity
//   actuators, PGS + noslip, touch sensors, implicitfast), state extraction (:216-232), command
//   resampling (:235,:321-333), termination (:239-258), reset (:335-361), rewards (:277-288,:399-497)
//   and observations (:291-309).
//
// Work decomposition (design, not a port): one environment per 8-lane group ("octet") of a warp,
// four environments per warp.  Lanes 0..5 each own one leg (3 hinges, 3 links, the tibia hull and
// its contacts); lane 6 owns the base hull and its contacts; lane 7 is a spare.  Everything that
// belongs to the floating base (pose, 6x6 Schur complement, base accelerations) is computed
// redundantly by all eight lanes from bit-identical inputs, so it never has to be communicated; sums
// over legs are xor-butterfly shuffles inside the octet.  The mass matrix is never formed: with the
// legs eliminated first, M^-1 splits into six private 3x3 Cholesky factors G_k, six 3x6 couplings
// E_k = M_k^-1 C_k and one replicated 6x6 factor G_S of S = M_bb - sum_k C_k^T E_k.  Contact rows are
// whitened by those factors (Y = J~ G_S^-T in base space, Z = J_k G_k^-T in leg space), which turns
// the dual PGS / noslip sweeps into updates of a 6-vector u (shared) and a 3-vector w_k (private):
// A = J M^-1 J^T + R is never materialised either.
//
// All state lives in registers for the whole step (both substeps); HBM is touched once on the way
// in and once on the way out.  Shared memory only holds the per-CTA copy of the model constants.
#include <cuda_runtime.h>
#include <math_constants.h>

#include "nm_device.hpp"

#define FULL 0xffffffffu
// The warps of a CTA are kept in step at phase boundaries so that instruction-cache lines fetched by the leading
// warp are reused by the others: the hot code (~120 KB) is ~4x the 32 KB L1.5 I-cache, and unsynchronised warps
// each stream it from L2 on their own (measured: 42 -> 83 M env-steps/s at 131072 envs; profiles/r01_notes.md).
#ifdef NM_NOSYNC_SMALL
#define PHASE_SYNC() do { if (BLOCK > 128) __syncthreads(); } while (0)
#else
#define PHASE_SYNC() __syncthreads()
#endif
// experiment switches: individual phase barriers off (NM_SKIP_SYNC_B: before CRBA, _C: before the factorisations, _E: before the contact build)
#ifdef NM_SKIP_SYNC_B
#define PHASE_SYNC_B()
#else
#define PHASE_SYNC_B() PHASE_SYNC()
#endif
#ifdef NM_SKIP_SYNC_C
#define PHASE_SYNC_C()
#else
#define PHASE_SYNC_C() PHASE_SYNC()
#endif
#ifdef NM_SKIP_SYNC_A
#define PHASE_SYNC_A()
#else
#define PHASE_SYNC_A() PHASE_SYNC()
#endif
#ifdef NM_SKIP_SYNC_D
#define PHASE_SYNC_D()
#else
// before the collision phase: kept for one-wave launches (87.1 vs 89.0 us at 4096 envs without it), dropped for the 256-thread
// CTAs of large batches (+1.2 % at 131 072 envs, +1.6 % at 16 384; gpurun_out/r02_qb21.log)
#ifndef NM_SYNC_D_MAX
#define NM_SYNC_D_MAX 128
#endif
#define PHASE_SYNC_D() do { if (BLOCK <= NM_SYNC_D_MAX) __syncthreads(); } while (0)
#endif
#ifdef NM_SKIP_SYNC_F
#define PHASE_SYNC_F()
#else
#define PHASE_SYNC_F() PHASE_SYNC()
#endif
#ifndef NM_KEEP_SYNC_E
#define PHASE_SYNC_E()       // measured (gpurun_out/r02_qb18.log): without this barrier +2.6 % at 131 072 envs, +3 % at 16 384, equal at 4096
#else
#define PHASE_SYNC_E() PHASE_SYNC()
#endif
// Experiment builds (-DNM_TIMING): every warp records clock64() at the phase boundaries into a global buffer
// (tools/phase_timing.py); compiled out of the product library.
#ifdef NM_TIMING
__device__ long long* nm_timing_buf = nullptr;
#define TSTAMP(k) do { if (nm_timing_buf && (threadIdx.x & 31) == 0) nm_timing_buf[(size_t)(gtid >> 5) * 32 + (k)] = clock64(); } while (0)
#else
#define TSTAMP(k)
#endif
#define NM_MINVAL 1e-15f
// Directed hull edge e of a geom whose vertex table starts at hv: the neighbour's coordinates (xyz) and its geom-local id (w),
// through the compact adjacency (16-bit neighbour id -> vertex table).  Generic loads: the tables live in shared memory when the
// CTA staged them, in global memory otherwise.  (Round 2 also measured a 16-byte-per-edge table with the coordinates inline -- one
// load level less, 165 KB -- and `ld.global.nc.L1::evict_last` on either layout: no difference to 0.1 us, the walk is bound by
// the latency of each level, not by hit rates.)
__device__ __forceinline__ float4 ld_edge(const unsigned short* nbr, const float4* hv, int e) {
  const int u = nbr[e];
  float4 q = hv[u];
  q.w = __int_as_float(u);
  return q;
}

#define NM_TINY 1e-30f

// ---------------------------------------------------------------------------------------------- small algebra
struct V3 { float x, y, z; };
struct M3 { float a[9]; };          // row-major
struct SV { V3 w, v; };             // spatial motion [angular; linear] or force [torque; force]
struct In { float xx, yy, zz, xy, xz, yz; V3 h; float m; };   // spatial inertia about the c-frame origin

__device__ __forceinline__ V3 mk(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3 ld3(const float* p) { return mk(p[0], p[1], p[2]); }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator*(float s, V3 a) { return mk(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ float dot(V3 a, V3 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)); }
__device__ __forceinline__ V3 cross(V3 a, V3 b) { return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ V3 fma3(float s, V3 a, V3 b) { return mk(fmaf(s, a.x, b.x), fmaf(s, a.y, b.y), fmaf(s, a.z, b.z)); }
__device__ __forceinline__ V3 mul(const M3& m, V3 v) {
  return mk(fmaf(m.a[0], v.x, fmaf(m.a[1], v.y, m.a[2] * v.z)), fmaf(m.a[3], v.x, fmaf(m.a[4], v.y, m.a[5] * v.z)),
            fmaf(m.a[6], v.x, fmaf(m.a[7], v.y, m.a[8] * v.z)));
}
__device__ __forceinline__ V3 mulT(const M3& m, V3 v) {
  return mk(fmaf(m.a[0], v.x, fmaf(m.a[3], v.y, m.a[6] * v.z)), fmaf(m.a[1], v.x, fmaf(m.a[4], v.y, m.a[7] * v.z)),
            fmaf(m.a[2], v.x, fmaf(m.a[5], v.y, m.a[8] * v.z)));
}
__device__ __forceinline__ V3 col(const M3& m, int i) { return mk(m.a[i], m.a[3 + i], m.a[6 + i]); }
__device__ __forceinline__ M3 matmul(const M3& A, const M3& B) {
  M3 C;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) C.a[3 * i + j] = fmaf(A.a[3 * i], B.a[j], fmaf(A.a[3 * i + 1], B.a[3 + j], A.a[3 * i + 2] * B.a[6 + j]));
  return C;
}
__device__ __forceinline__ M3 quat2mat(float w, float x, float y, float z) {
  M3 m;
  m.a[0] = w * w + x * x - y * y - z * z; m.a[1] = 2.f * (x * y - w * z); m.a[2] = 2.f * (x * z + w * y);
  m.a[3] = 2.f * (x * y + w * z); m.a[4] = w * w - x * x + y * y - z * z; m.a[5] = 2.f * (y * z - w * x);
  m.a[6] = 2.f * (x * z - w * y); m.a[7] = 2.f * (y * z + w * x); m.a[8] = w * w - x * x - y * y + z * z;
  return m;
}
// rotation by `angle` about unit `ax` (Rodrigues)
__device__ __forceinline__ M3 axis_rot(V3 ax, float angle) {
  float s, c;
  sincosf(angle, &s, &c);
  float t = 1.f - c;
  M3 m;
  m.a[0] = fmaf(t * ax.x, ax.x, c); m.a[1] = fmaf(t * ax.x, ax.y, -s * ax.z); m.a[2] = fmaf(t * ax.x, ax.z, s * ax.y);
  m.a[3] = fmaf(t * ax.x, ax.y, s * ax.z); m.a[4] = fmaf(t * ax.y, ax.y, c); m.a[5] = fmaf(t * ax.y, ax.z, -s * ax.x);
  m.a[6] = fmaf(t * ax.x, ax.z, -s * ax.y); m.a[7] = fmaf(t * ax.y, ax.z, s * ax.x); m.a[8] = fmaf(t * ax.z, ax.z, c);
  return m;
}
// X * Iloc * X^T for symmetric Iloc = (xx yy zz xy xz yz)
__device__ __forceinline__ void rot_inertia(const M3& X, const float* I, float* o) {
  float T[9];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    float a = X.a[3 * i], b = X.a[3 * i + 1], c = X.a[3 * i + 2];
    T[3 * i] = fmaf(a, I[0], fmaf(b, I[3], c * I[4]));
    T[3 * i + 1] = fmaf(a, I[3], fmaf(b, I[1], c * I[5]));
    T[3 * i + 2] = fmaf(a, I[4], fmaf(b, I[5], c * I[2]));
  }
  o[0] = fmaf(T[0], X.a[0], fmaf(T[1], X.a[1], T[2] * X.a[2]));
  o[1] = fmaf(T[3], X.a[3], fmaf(T[4], X.a[4], T[5] * X.a[5]));
  o[2] = fmaf(T[6], X.a[6], fmaf(T[7], X.a[7], T[8] * X.a[8]));
  o[3] = fmaf(T[0], X.a[3], fmaf(T[1], X.a[4], T[2] * X.a[5]));
  o[4] = fmaf(T[0], X.a[6], fmaf(T[1], X.a[7], T[2] * X.a[8]));
  o[5] = fmaf(T[3], X.a[6], fmaf(T[4], X.a[7], T[5] * X.a[8]));
}
__device__ __forceinline__ In make_inertia(const float* Iw, float m, V3 r) {
  In c;
  c.xx = fmaf(m, r.y * r.y + r.z * r.z, Iw[0]);
  c.yy = fmaf(m, r.x * r.x + r.z * r.z, Iw[1]);
  c.zz = fmaf(m, r.x * r.x + r.y * r.y, Iw[2]);
  c.xy = fmaf(-m, r.x * r.y, Iw[3]);
  c.xz = fmaf(-m, r.x * r.z, Iw[4]);
  c.yz = fmaf(-m, r.y * r.z, Iw[5]);
  c.h = m * r;
  c.m = m;
  return c;
}
__device__ __forceinline__ In operator+(const In& a, const In& b) {
  In c;
  c.xx = a.xx + b.xx; c.yy = a.yy + b.yy; c.zz = a.zz + b.zz; c.xy = a.xy + b.xy; c.xz = a.xz + b.xz; c.yz = a.yz + b.yz;
  c.h = a.h + b.h; c.m = a.m + b.m;
  return c;
}
__device__ __forceinline__ SV imul(const In& i, const SV& s) {   // spatial inertia times motion
  SV r;
  r.w.x = fmaf(i.xx, s.w.x, fmaf(i.xy, s.w.y, i.xz * s.w.z)) + (i.h.y * s.v.z - i.h.z * s.v.y);
  r.w.y = fmaf(i.xy, s.w.x, fmaf(i.yy, s.w.y, i.yz * s.w.z)) + (i.h.z * s.v.x - i.h.x * s.v.z);
  r.w.z = fmaf(i.xz, s.w.x, fmaf(i.yz, s.w.y, i.zz * s.w.z)) + (i.h.x * s.v.y - i.h.y * s.v.x);
  r.v = fma3(i.m, s.v, cross(s.w, i.h));
  return r;
}
__device__ __forceinline__ SV cross_motion(const SV& vel, const SV& s) { SV r; r.w = cross(vel.w, s.w); r.v = cross(vel.w, s.v) + cross(vel.v, s.w); return r; }
__device__ __forceinline__ SV cross_force(const SV& vel, const SV& f) { SV r; r.w = cross(vel.w, f.w) + cross(vel.v, f.v); r.v = cross(vel.w, f.v); return r; }
__device__ __forceinline__ float sdot(const SV& a, const SV& b) { return dot(a.w, b.w) + dot(a.v, b.v); }
__device__ __forceinline__ SV operator+(const SV& a, const SV& b) { SV r; r.w = a.w + b.w; r.v = a.v + b.v; return r; }
__device__ __forceinline__ SV sfma(float s, const SV& a, const SV& b) { SV r; r.w = fma3(s, a.w, b.w); r.v = fma3(s, a.v, b.v); return r; }

// ---------------------------------------------------------------------------------------------- octet collectives
__device__ __forceinline__ float oct_sum(float v) {
  v += __shfl_xor_sync(FULL, v, 1);
  v += __shfl_xor_sync(FULL, v, 2);
  v += __shfl_xor_sync(FULL, v, 4);
  return v;
}
__device__ __forceinline__ float oct_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(FULL, v, 1));
  v = fmaxf(v, __shfl_xor_sync(FULL, v, 2));
  v = fmaxf(v, __shfl_xor_sync(FULL, v, 4));
  return v;
}
__device__ __forceinline__ int oct_sumi(int v) {
  v += __shfl_xor_sync(FULL, v, 1);
  v += __shfl_xor_sync(FULL, v, 2);
  v += __shfl_xor_sync(FULL, v, 4);
  return v;
}
__device__ __forceinline__ V3 oct_sum3(V3 v) { return mk(oct_sum(v.x), oct_sum(v.y), oct_sum(v.z)); }
__device__ __forceinline__ float oct_bcast(float v, int src_lane) { return __shfl_sync(FULL, v, src_lane); }

// ---------------------------------------------------------------------------------------------- factorisations
// 3x3 Cholesky of (m00 m10 m11 m20 m21 m22): g = (l00 l10 l11 l20 l21 l22), gi = inverse diagonal
__device__ __forceinline__ void chol3(const float* m, float* g, float* gi) {
  gi[0] = rsqrtf(fmaxf(m[0], NM_TINY)); g[0] = m[0] * gi[0];
  g[1] = m[1] * gi[0];
  float d1 = fmaxf(fmaf(-g[1], g[1], m[2]), NM_TINY);
  gi[1] = rsqrtf(d1); g[2] = d1 * gi[1];
  g[3] = m[3] * gi[0];
  g[4] = fmaf(-g[3], g[1], m[4]) * gi[1];
  float d2 = fmaxf(fmaf(-g[4], g[4], fmaf(-g[3], g[3], m[5])), NM_TINY);
  gi[2] = rsqrtf(d2); g[5] = d2 * gi[2];
}
__device__ __forceinline__ void fwd3(const float* g, const float* gi, const float* b, float* y) {
  y[0] = b[0] * gi[0];
  y[1] = fmaf(-g[1], y[0], b[1]) * gi[1];
  y[2] = fmaf(-g[4], y[1], fmaf(-g[3], y[0], b[2])) * gi[2];
}
__device__ __forceinline__ void bwd3(const float* g, const float* gi, const float* y, float* x) {
  x[2] = y[2] * gi[2];
  x[1] = fmaf(-g[4], x[2], y[1]) * gi[1];
  x[0] = fmaf(-g[3], x[2], fmaf(-g[1], x[1], y[0])) * gi[0];
}
// 6x6 lower-triangular packed row-wise: index(i,j) = i(i+1)/2 + j
#define TRI(i, j) ((i) * ((i) + 1) / 2 + (j))
__device__ __forceinline__ void chol6(float* s, float* si) {   // in place
#pragma unroll
  for (int j = 0; j < 6; j++) {
    float d = s[TRI(j, j)];
#pragma unroll
    for (int k = 0; k < j; k++) d = fmaf(-s[TRI(j, k)], s[TRI(j, k)], d);
    d = fmaxf(d, NM_TINY);
    si[j] = rsqrtf(d);
    s[TRI(j, j)] = d * si[j];
#pragma unroll
    for (int i = j + 1; i < 6; i++) {
      float t = s[TRI(i, j)];
#pragma unroll
      for (int k = 0; k < j; k++) t = fmaf(-s[TRI(i, k)], s[TRI(j, k)], t);
      s[TRI(i, j)] = t * si[j];
    }
  }
}
__device__ __forceinline__ void fwd6(const float* s, const float* si, const float* b, float* y) {
#pragma unroll
  for (int i = 0; i < 6; i++) {
    float t = b[i];
#pragma unroll
    for (int k = 0; k < i; k++) t = fmaf(-s[TRI(i, k)], y[k], t);
    y[i] = t * si[i];
  }
}
__device__ __forceinline__ void bwd6(const float* s, const float* si, const float* y, float* x) {
#pragma unroll
  for (int i = 5; i >= 0; i--) {
    float t = y[i];
#pragma unroll
    for (int k = i + 1; k < 6; k++) t = fmaf(-s[TRI(k, i)], x[k], t);
    x[i] = t * si[i];
  }
}

// Block elimination of the legs: G_k, E_k = M_k^-1 C_k, G_S = chol(M_bb - sum_k C_k^T E_k)
struct Factor { float g[6], gi[3], E[3][6], S[21], Si[6]; };
__device__ __forceinline__ void factor_system(const float* Mk, const float (*C)[6], const float* Mbb, const float* dadd, Factor& F) {
  float m[6] = {Mk[0] + dadd[0], Mk[1], Mk[2] + dadd[1], Mk[3], Mk[4], Mk[5] + dadd[2]};
  chol3(m, F.g, F.gi);
#pragma unroll
  for (int c = 0; c < 6; c++) {
    float b[3] = {C[0][c], C[1][c], C[2][c]}, y[3], x[3];
    fwd3(F.g, F.gi, b, y);
    bwd3(F.g, F.gi, y, x);
    F.E[0][c] = x[0]; F.E[1][c] = x[1]; F.E[2][c] = x[2];
  }
#pragma unroll
  for (int a = 0; a < 6; a++)
#pragma unroll
    for (int b = 0; b <= a; b++) {
      float t = fmaf(C[0][a], F.E[0][b], fmaf(C[1][a], F.E[1][b], C[2][a] * F.E[2][b]));
      F.S[TRI(a, b)] = Mbb[TRI(a, b)] - oct_sum(t);
    }
  chol6(F.S, F.Si);
}
// x = M^-1 [rb; rk]  (rb replicated across the octet, rk private)
__device__ __forceinline__ void solve_system(const Factor& F, const float* rb, const float* rk, float* xb, float* xk) {
  float y[3], yk[3];
  fwd3(F.g, F.gi, rk, y);
  bwd3(F.g, F.gi, y, yk);
  float rt[6], t6[6];
#pragma unroll
  for (int a = 0; a < 6; a++) {
    float t = fmaf(F.E[0][a], rk[0], fmaf(F.E[1][a], rk[1], F.E[2][a] * rk[2]));
    rt[a] = rb[a] - oct_sum(t);
  }
  fwd6(F.S, F.Si, rt, t6);
  bwd6(F.S, F.Si, t6, xb);
#pragma unroll
  for (int i = 0; i < 3; i++) {
    float t = yk[i];
#pragma unroll
    for (int a = 0; a < 6; a++) t = fmaf(-F.E[i][a], xb[a], t);
    xk[i] = t;
  }
}

// ---------------------------------------------------------------------------------------------- misc
__device__ __noinline__ float impedance_pow(float x, float mid, float power) {   // general solimp power: cold path, kept out of line
  return (x <= mid) ? powf(x, power) / powf(mid, power - 1.f) : 1.f - powf(1.f - x, power) / powf(1.f - mid, power - 1.f);
}
__device__ __forceinline__ float impedance(const NmGeom& g, float pos) {
  if (g.dmin == g.dmax || g.width <= NM_MINVAL) return 0.5f * (g.dmin + g.dmax);
  float x = fabsf(pos / g.width);
  if (x >= 1.f) return g.dmax;
  if (x <= 0.f) return g.dmin;
  float y;
  if (g.power == 1.f) y = x;
  else if (g.power == 2.f) y = (x <= g.mid) ? x * x / g.mid : 1.f - (1.f - x) * (1.f - x) / (1.f - g.mid);
  else y = impedance_pow(x, g.mid, g.power);
  return fmaf(y, g.dmax - g.dmin, g.dmin);
}

__device__ __forceinline__ float ray_sphere(V3 center, float radius, V3 pnt, V3 vec) {
  V3 dif = pnt - center;
  float a = dot(vec, vec), b = dot(vec, dif), c = dot(dif, dif) - radius * radius;
  float det = b * b - a * c;
  if (det < NM_MINVAL || a < NM_MINVAL) return -1.f;
  det = sqrtf(det);
  float x0 = (-b - det) / a, x1 = (-b + det) / a;
  if (x0 >= 0.f) return x0;
  if (x1 >= 0.f) return x1;
  return -1.f;
}

__device__ __noinline__ void philox4x32(unsigned k0, unsigned k1, unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned* out) {
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
__device__ __forceinline__ float u01(unsigned x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }
// uniform noise in [-1,1) for observation entry k: Philox block (phase 2 + k/4), word k%4 (same definition as the oracle)
__device__ __noinline__ float obs_noise(unsigned long long seed, long long genv, long long step, int k) {
  unsigned nz[4];
  philox4x32((unsigned)seed, (unsigned)genv, (unsigned)step, (unsigned)((unsigned long long)step >> 32), (unsigned)(2 + (k >> 2)), (unsigned)(seed >> 32), nz);
  return 2.f * u01(nz[k & 3]) - 1.f;
}

// ≙ _resample_commands (env.py:321-333); RNG keyed by (seed, GLOBAL env id), counter (step, phase)
__device__ __forceinline__ void resample_commands(const NmDevCfg& c, unsigned long long seed, long long genv, long long step, int phase, float* cmd) {
  unsigned r[4];
  philox4x32((unsigned)seed, (unsigned)genv, (unsigned)step, (unsigned)((unsigned long long)step >> 32), (unsigned)phase, (unsigned)(seed >> 32), r);
  float cx = u01(r[0]) * 2.f * c.max_lin_vel_x - c.max_lin_vel_x;
  float cy = 0.f;
  float cw = u01(r[1]) * 2.f * c.max_ang_vel - c.max_ang_vel;
  float keep = sqrtf(cx * cx + cy * cy) > 0.02f ? 1.f : 0.f;
  cmd[0] = cx * keep; cmd[1] = cy * keep; cmd[2] = cw;
}

// per-lane contact blocks (local memory, L1-resident): 55 floats per contact.  The three whitened contact-frame
// rows couple a contact to the rest of the system; the 4x4 Gram matrix of its own pyramid edges Jn +- mu*Jt
// (computed from the edge vectors themselves, no cancellation) carries the coupling among its four rows.
struct ConBlk {
  float Y[NM_MAXC][3][6];   // whitened base-space image of the contact-frame rows (normal, tangent 1, tangent 2)
  float Z[NM_MAXC][3][3];   // whitened leg-space image
  float b[NM_MAXC][4];      // J_edge qacc_smooth - aref_edge
  float adi[NM_MAXC][4];    // 1 / (|edge|^2 + R)
  float G[NM_MAXC][10];     // Gram matrix of the 4 pyramid edges (= A restricted to this contact, without R):
                            //   00 11 22 33 | 01 23 (opposing pairs) | 02 03 12 13 (across the two tangents)
  float ik[NM_MAXC][2];     // 1 / (G_aa + G_bb - 2 G_ab) of the two opposing pairs (noslip curvature), 0 if degenerate
  float f[NM_MAXC][4];      // pyramid-edge forces
  float R[NM_MAXC];
  V3 pos[NM_MAXC];
};

// register image of one contact block, and one Gauss-Seidel visit of it (PGS: 4 single edges with R; noslip: 2 pairs)
struct ConRegs { float Y[3][6], Z[3][3], G[10], b[4], adi[4], R, f[4], ik[2]; };
__device__ __forceinline__ void con_load(const ConBlk& cb, int c, ConRegs& k) {
#pragma unroll
  for (int f = 0; f < 3; f++) {
#pragma unroll
    for (int a = 0; a < 6; a++) k.Y[f][a] = cb.Y[c][f][a];
#pragma unroll
    for (int j = 0; j < 3; j++) k.Z[f][j] = cb.Z[c][f][j];
  }
#pragma unroll
  for (int i = 0; i < 10; i++) k.G[i] = cb.G[c][i];
#pragma unroll
  for (int i = 0; i < 4; i++) { k.b[i] = cb.b[c][i]; k.adi[i] = cb.adi[c][i]; k.f[i] = cb.f[c][i]; }
  k.R = cb.R[c];
  k.ik[0] = cb.ik[c][0]; k.ik[1] = cb.ik[c][1];
}
// Residuals of the 4 edges against the CURRENT dual state, all at once: r_e = b_e + (p0 +- mu*p_t), p_a = Y_a.u + Z_a.w
// (three independent dot products); the Gauss-Seidel coupling among the 4 rows of the contact is then applied through its
// edge Gram matrix instead of re-walking u after every row.
__device__ __forceinline__ void con_sweep(ConRegs& k, float* u, float* wv, float mu, bool in_noslip, float& improvement,
                                          float pe0 = 0.f, float pe1 = 0.f, float pe2 = 0.f, float* cout = nullptr) {
  float p0 = pe0, p1 = pe1, p2 = pe2;          // (pair contacts: the other leg's Z.w, plane contacts: 0)
#pragma unroll
  for (int a = 0; a < 6; a++) { const float ua = u[a]; p0 = fmaf(k.Y[0][a], ua, p0); p1 = fmaf(k.Y[1][a], ua, p1); p2 = fmaf(k.Y[2][a], ua, p2); }
#pragma unroll
  for (int j = 0; j < 3; j++) { const float wj = wv[j]; p0 = fmaf(k.Z[0][j], wj, p0); p1 = fmaf(k.Z[1][j], wj, p1); p2 = fmaf(k.Z[2][j], wj, p2); }
  float r[4] = {k.b[0] + fmaf(mu, p1, p0), k.b[1] + fmaf(-mu, p1, p0), k.b[2] + fmaf(mu, p2, p0), k.b[3] + fmaf(-mu, p2, p0)};
  float* o = k.f;
  float d[4];
  const float g00 = k.G[0], g11 = k.G[1], g22 = k.G[2], g33 = k.G[3], g01 = k.G[4], g23 = k.G[5];
  const float g02 = k.G[6], g03 = k.G[7], g12 = k.G[8], g13 = k.G[9];
  if (!in_noslip) {
    const float R = k.R;
    const float gd[4] = {g00, g11, g22, g33};
#pragma unroll
    for (int e = 0; e < 4; e++) {
      const float res = fmaf(R, o[e], r[e]);
      float fnew = fmaxf(0.f, fmaf(-res, k.adi[e], o[e]));
      float de = fnew - o[e];
      float change = de * fmaf(0.5f * de, gd[e] + R, res);
      if (change > 1e-10f) { de = 0.f; fnew = o[e]; change = 0.f; }
      improvement -= change;
      o[e] = fnew; d[e] = de;
      if (e == 0) { r[1] = fmaf(de, g01, r[1]); r[2] = fmaf(de, g02, r[2]); r[3] = fmaf(de, g03, r[3]); }
      if (e == 1) { r[2] = fmaf(de, g12, r[2]); r[3] = fmaf(de, g13, r[3]); }
      if (e == 2) { r[3] = fmaf(de, g23, r[3]); }
    }
  } else {
    // noslip on the two opposing edge pairs, their sum 2*mid held fixed: f = (mid + x, mid - x).  With o = (mid + s, mid - s) the
    // stationarity condition of the oracle's pair problem (K0 + K1 x = 0, K0 = mid (a00 - a11) + bc0 - bc1, bc = res - A o)
    // reduces to x = s - (res0 - res1) / K1, K1 = a00 + a11 - 2 a01, and the cost change to d0 (K1 d0 / 2 + res0 - res1):
    // half the instructions of evaluating K0, bc and the 2x2 quadratic form literally.
#pragma unroll
    for (int t = 0; t < 2; t++) {
      const float a00 = t ? g22 : g00, a11 = t ? g33 : g11, a01 = t ? g23 : g01;
      const float o0 = o[2 * t], o1 = o[2 * t + 1], dr = r[2 * t] - r[2 * t + 1];
      const float mid = 0.5f * (o0 + o1);
      const float x = k.ik[t] == 0.f ? 0.f : fmaf(-dr, k.ik[t], 0.5f * (o0 - o1));
      float f0 = mid + x, f1 = mid - x;
      if (x < -mid) { f0 = 0.f; f1 = 2.f * mid; }
      else if (x > mid) { f0 = 2.f * mid; f1 = 0.f; }
      float d0 = f0 - o0, d1 = f1 - o1;
      float change = d0 * fmaf(0.5f * d0, (a00 + a11) - 2.f * a01, dr);
      if (change > 1e-10f) { f0 = o0; f1 = o1; d0 = 0.f; d1 = 0.f; change = 0.f; }
      improvement -= change;
      o[2 * t] = f0; o[2 * t + 1] = f1; d[2 * t] = d0; d[2 * t + 1] = d1;
      if (t == 0) { r[2] = fmaf(d0, g02, fmaf(d1, g12, r[2])); r[3] = fmaf(d0, g03, fmaf(d1, g13, r[3])); }
    }
  }
  const float c1 = mu * (d[0] - d[1]), c2 = mu * (d[2] - d[3]);
  if (in_noslip) {
    // the normal component (d0 + d1) + (d2 + d3) is zero by construction (fp32 rounding aside): only the tangential rows move u, w
    if (cout != nullptr) { cout[0] = 0.f; cout[1] = c1; cout[2] = c2; }
#pragma unroll
    for (int a = 0; a < 6; a++) u[a] = fmaf(k.Y[1][a], c1, fmaf(k.Y[2][a], c2, u[a]));
#pragma unroll
    for (int j = 0; j < 3; j++) wv[j] = fmaf(k.Z[1][j], c1, fmaf(k.Z[2][j], c2, wv[j]));
    return;
  }
  const float c0 = (d[0] + d[1]) + (d[2] + d[3]);
  if (cout != nullptr) { cout[0] = c0; cout[1] = c1; cout[2] = c2; }
#pragma unroll
  for (int a = 0; a < 6; a++) u[a] = fmaf(k.Y[0][a], c0, fmaf(k.Y[1][a], c1, fmaf(k.Y[2][a], c2, u[a])));
#pragma unroll
  for (int j = 0; j < 3; j++) wv[j] = fmaf(k.Z[0][j], c0, fmaf(k.Z[1][j], c1, fmaf(k.Z[2][j], c2, wv[j])));
}


// ================================================================================================ convex-convex pairs
// Tibia-tibia contacts (reference models/nightmare_v3/mjmodel.xml:47).  MuJoCo collides two convex meshes with libccd's
// Minkowski Portal Refinement; the same algorithm is restated here in fp32, run
// by the whole octet: the portal is replicated in all eight lanes, the support query (the only loop over hull vertices) is
// split across them.  Coordinates are relative to the base origin, so that fp32 resolution does not depend on how far the
// robot has walked.  Cold code: it runs only for pairs whose bounding capsules overlap.
struct HullPose { float X[9], p[3], c[3]; int adr, num; };
struct PairBlk {                 // one pair contact, shared memory; everything but Zi/Zj is the same for the whole octet
  float Y[3][6], Zi[3][3], Zj[3][3], G[10], b[4], adi[4], f[4], ik[2], R, dist;
  float pos[3], nrm[3], mu;
  int li, lj;
};
struct Supp { V3 v, v1, v2; };
#define NM_CCD_EPS 1.1920929e-07f
__device__ __forceinline__ bool ccd_zero(float x) { return fabsf(x) < NM_CCD_EPS; }
__device__ __forceinline__ bool ccd_eq(float a_, float b_) {
  const float ab = fabsf(a_ - b_);
  if (ab < NM_CCD_EPS) return true;
  const float a = fabsf(a_), b = fabsf(b_);
  return b > a ? ab < NM_CCD_EPS * b : ab < NM_CCD_EPS * a;
}
__device__ __forceinline__ float dot_plain(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 normalized(V3 v) { const float n = sqrtf(dot_plain(v, v)); return mk(v.x / n, v.y / n, v.z / n); }

// support vertex of a hull along `dir` (world): argmax over all hull vertices, 1/8 of them per lane, lowest index wins ties
__device__ __noinline__ V3 hull_support_oct(const float4* __restrict__ hull_vert, const HullPose* h, V3 dir, int l, unsigned omask) {
  const float* X = h->X;
  const V3 dl = mk(X[0] * dir.x + X[3] * dir.y + X[6] * dir.z, X[1] * dir.x + X[4] * dir.y + X[7] * dir.z, X[2] * dir.x + X[5] * dir.y + X[8] * dir.z);
  const float4* hv = hull_vert + h->adr;
  float best = -CUDART_INF_F;
  int bi = 0x7fffffff;
  for (int v = l; v < h->num; v += 8) {
    const float4 q = __ldg(hv + v);
    const float val = dl.x * q.x + dl.y * q.y + dl.z * q.z;
    if (val > best) { best = val; bi = v; }
  }
#pragma unroll
  for (int o = 1; o < 8; o <<= 1) {
    const float ov = __shfl_xor_sync(omask, best, o);
    const int oi = __shfl_xor_sync(omask, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  const float4 q = __ldg(hv + bi);
  return mk(X[0] * q.x + X[1] * q.y + X[2] * q.z + h->p[0], X[3] * q.x + X[4] * q.y + X[5] * q.z + h->p[1], X[6] * q.x + X[7] * q.y + X[8] * q.z + h->p[2]);
}
__device__ __forceinline__ void mpr_support(const float4* hv, const HullPose* A, const HullPose* B, V3 dir, int l, unsigned om, Supp& s) {
  s.v1 = hull_support_oct(hv, A, dir, l, om);
  s.v2 = hull_support_oct(hv, B, mk(-dir.x, -dir.y, -dir.z), l, om);
  s.v = s.v1 - s.v2;
}
__device__ __forceinline__ V3 portal_dir(const Supp* P) { return normalized(cross(P[2].v - P[1].v, P[3].v - P[1].v)); }
__device__ __forceinline__ bool portal_reach_tolerance(const Supp* P, const Supp& v4, V3 dir, float tol) {
  const float dv4 = dot_plain(v4.v, dir);
  const float d1 = fminf(fminf(dv4 - dot_plain(P[1].v, dir), dv4 - dot_plain(P[2].v, dir)), dv4 - dot_plain(P[3].v, dir));
  return ccd_eq(d1, tol) || d1 < tol;
}
__device__ __forceinline__ void expand_portal(Supp* P, const Supp& v4) {
  const V3 v4v0 = cross(v4.v, P[0].v);
  if (dot_plain(P[1].v, v4v0) > 0.f) {
    if (dot_plain(P[2].v, v4v0) > 0.f) P[1] = v4; else P[3] = v4;
  } else {
    if (dot_plain(P[3].v, v4v0) > 0.f) P[2] = v4; else P[1] = v4;
  }
}
__device__ __forceinline__ float point_seg_dist2(V3 x0, V3 b, V3& wit) {       // distance of the origin from the segment (x0, b)
  const V3 d = b - x0;
  const float t = -dot_plain(x0, d) / dot_plain(d, d);
  if (t < 0.f || ccd_zero(t)) { wit = x0; return dot_plain(x0, x0); }
  if (t > 1.f || ccd_eq(t, 1.f)) { wit = b; return dot_plain(b, b); }
  wit = mk(d.x * t + x0.x, d.y * t + x0.y, d.z * t + x0.z);
  return dot_plain(wit, wit);
}
__device__ __forceinline__ float origin_tri_dist2(V3 x0, V3 B, V3 C, V3& wit) {
  const V3 d1 = B - x0, d2 = C - x0;
  const float v = dot_plain(d1, d1), w = dot_plain(d2, d2), p = dot_plain(x0, d1), q = dot_plain(x0, d2), r = dot_plain(d1, d2);
  const float den = w * v - r * r;
  float s, t;
  if (ccd_zero(den)) { s = t = -1.f; }
  else { s = (q * r - w * p) / den; t = (-s * r - q) / w; }
  if ((ccd_zero(s) || s > 0.f) && (ccd_eq(s, 1.f) || s < 1.f) && (ccd_zero(t) || t > 0.f) && (ccd_eq(t, 1.f) || t < 1.f) && (ccd_eq(t + s, 1.f) || t + s < 1.f)) {
    wit = mk(x0.x + s * d1.x + t * d2.x, x0.y + s * d1.y + t * d2.y, x0.z + s * d1.z + t * d2.z);
    return dot_plain(wit, wit);
  }
  V3 w2;
  float dist = point_seg_dist2(x0, B, wit);
  float dd = point_seg_dist2(x0, C, w2);
  if (dd < dist) { dist = dd; wit = w2; }
  dd = point_seg_dist2(B, C, w2);
  if (dd < dist) { dist = dd; wit = w2; }
  return dist;
}
// out = {depth, normal(3), position(3)}; true when the hulls intersect (≙ ccdMPRPenetration == 0 with a defined direction)
__device__ __noinline__ bool mpr_penetration_oct(const float4* __restrict__ hv, const HullPose* A, const HullPose* B, float tol, int maxit, int l,
                                                 unsigned om, float* out) {
  Supp P[4], v4;
  V3 dir, va;
  float d;
  P[0].v1 = ld3(A->c); P[0].v2 = ld3(B->c); P[0].v = P[0].v1 - P[0].v2;
  if (ccd_eq(P[0].v.x, 0.f) && ccd_eq(P[0].v.y, 0.f) && ccd_eq(P[0].v.z, 0.f)) P[0].v.x += NM_CCD_EPS * 10.f;
  dir = normalized(mk(-P[0].v.x, -P[0].v.y, -P[0].v.z));
  mpr_support(hv, A, B, dir, l, om, P[1]);
  d = dot_plain(P[1].v, dir);
  if (ccd_zero(d) || d < 0.f) return false;
  dir = cross(P[0].v, P[1].v);
  if (ccd_zero(dot_plain(dir, dir))) {
    if (ccd_eq(P[1].v.x, 0.f) && ccd_eq(P[1].v.y, 0.f) && ccd_eq(P[1].v.z, 0.f)) return false;     // touching on v1: direction undefined
    const V3 n = normalized(P[1].v);                                                                // origin on the v0-v1 segment
    out[0] = sqrtf(dot_plain(P[1].v, P[1].v)); out[1] = n.x; out[2] = n.y; out[3] = n.z;
    out[4] = 0.5f * (P[1].v1.x + P[1].v2.x); out[5] = 0.5f * (P[1].v1.y + P[1].v2.y); out[6] = 0.5f * (P[1].v1.z + P[1].v2.z);
    return true;
  }
  dir = normalized(dir);
  mpr_support(hv, A, B, dir, l, om, P[2]);
  d = dot_plain(P[2].v, dir);
  if (ccd_zero(d) || d < 0.f) return false;
  dir = normalized(cross(P[1].v - P[0].v, P[2].v - P[0].v));
  if (dot_plain(dir, P[0].v) > 0.f) {
    const Supp t = P[1]; P[1] = P[2]; P[2] = t;
    dir = mk(-dir.x, -dir.y, -dir.z);
  }
  for (int guard = 0; guard < 64; guard++) {                       // portal discovery (libccd has no bound here; 64 is never reached)
    mpr_support(hv, A, B, dir, l, om, P[3]);
    d = dot_plain(P[3].v, dir);
    if (ccd_zero(d) || d < 0.f) return false;
    bool cont = false;
    va = cross(P[1].v, P[3].v);
    d = dot_plain(va, P[0].v);
    if (d < 0.f && !ccd_zero(d)) { P[2] = P[3]; cont = true; }
    if (!cont) {
      va = cross(P[3].v, P[2].v);
      d = dot_plain(va, P[0].v);
      if (d < 0.f && !ccd_zero(d)) { P[1] = P[3]; cont = true; }
    }
    if (!cont) break;
    dir = normalized(cross(P[1].v - P[0].v, P[2].v - P[0].v));
  }
  for (int guard = 0; guard < 256; guard++) {                      // portal refinement
    dir = portal_dir(P);
    d = dot_plain(dir, P[1].v);
    if (ccd_zero(d) || d > 0.f) break;
    mpr_support(hv, A, B, dir, l, om, v4);
    d = dot_plain(v4.v, dir);
    if (!(ccd_zero(d) || d > 0.f) || portal_reach_tolerance(P, v4, dir, tol)) return false;
    expand_portal(P, v4);
  }
  for (int it = 0;; it++) {                                        // penetration from the refined portal
    dir = portal_dir(P);
    mpr_support(hv, A, B, dir, l, om, v4);
    if (portal_reach_tolerance(P, v4, dir, tol) || it > maxit) {
      V3 wit;
      const float depth = sqrtf(origin_tri_dist2(P[1].v, P[2].v, P[3].v, wit));
      if (ccd_zero(depth)) return false;
      const V3 n = normalized(wit);
      // position: barycentric coordinates of the origin ray in the portal (≙ findPos)
      float b[4];
      b[0] = dot_plain(cross(P[1].v, P[2].v), P[3].v);
      b[1] = dot_plain(cross(P[3].v, P[2].v), P[0].v);
      b[2] = dot_plain(cross(P[0].v, P[1].v), P[3].v);
      b[3] = dot_plain(cross(P[2].v, P[1].v), P[0].v);
      float sum = b[0] + b[1] + b[2] + b[3];
      if (ccd_zero(sum) || sum < 0.f) {
        b[0] = 0.f;
        b[1] = dot_plain(cross(P[2].v, P[3].v), dir);
        b[2] = dot_plain(cross(P[3].v, P[1].v), dir);
        b[3] = dot_plain(cross(P[1].v, P[2].v), dir);
        sum = b[1] + b[2] + b[3];
      }
      const float inv = 1.f / sum;
      V3 p1 = mk(0, 0, 0), p2 = mk(0, 0, 0);
#pragma unroll
      for (int i = 0; i < 4; i++) { p1 = p1 + b[i] * P[i].v1; p2 = p2 + b[i] * P[i].v2; }
      out[0] = depth; out[1] = n.x; out[2] = n.y; out[3] = n.z;
      out[4] = 0.5f * (p1.x * inv + p2.x * inv); out[5] = 0.5f * (p1.y * inv + p2.y * inv); out[6] = 0.5f * (p1.z * inv + p2.z * inv);
      return true;
    }
    expand_portal(P, v4);
  }
}
// Upper bound of the support function max_v d.v of a hull from its support map (nm_device.hpp, NM_SMAP_N); d in the hull frame.
__device__ __forceinline__ float smap_support(const float* __restrict__ T, V3 d) {
  const float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
  int f; float m, u, v;
  if (ax >= ay && ax >= az) { f = 0; m = d.x; u = d.y; v = d.z; }
  else if (ay >= az) { f = 2; m = d.y; u = d.z; v = d.x; }
  else { f = 4; m = d.z; u = d.x; v = d.y; }
  const float am = fmaxf(fabsf(m), 1e-30f), s = 1.f / am;
  f += m < 0.f;
  const float gu = fminf(fmaxf(fmaf(u, s, 1.f) * (0.5f * NM_SMAP_N), 0.f), (float)NM_SMAP_N);
  const float gv = fminf(fmaxf(fmaf(v, s, 1.f) * (0.5f * NM_SMAP_N), 0.f), (float)NM_SMAP_N);
  const int iu = min((int)gu, NM_SMAP_N - 1), iv = min((int)gv, NM_SMAP_N - 1);
  const float fu = gu - (float)iu, fv = gv - (float)iv;
  const float* p = T + (f * (NM_SMAP_N + 1) + iv) * (NM_SMAP_N + 1) + iu;
  const float t00 = __ldg(p), t10 = __ldg(p + 1), t01 = __ldg(p + NM_SMAP_N + 1), t11 = __ldg(p + NM_SMAP_N + 2);
  const float lo = fmaf(fu, t10 - t00, t00), hi = fmaf(fu, t11 - t01, t01);
  return fmaf(fv, hi - lo, lo) * am;
}
// squared distance between two segments (bounding-capsule broad phase); ia / ie = 1 / squared lengths (model constants).
// Approximate division: the caller compares against a threshold with 0.1 mm of slack.
__device__ __forceinline__ float segseg_dist2(V3 p1, V3 d1, float ia, V3 p2, V3 d2, float ie) {
  const V3 r = p1 - p2;
  const float f = dot(d2, r), c = dot(d1, r), b = dot(d1, d2);
  const float den = fmaf(-b, b, __fdividef(1.f, ia * ie));
  float s = den > 1e-12f ? __saturatef(__fdividef(fmaf(b, f, -c * __fdividef(1.f, ie)), den)) : 0.f;
  float t = fmaf(b, s, f) * ie;
  if (t < 0.f) { t = 0.f; s = __saturatef(-c * ia); }
  else if (t > 1.f) { t = 1.f; s = __saturatef((b - c) * ia); }
  const V3 dd = fma3(s, d1, p1) - fma3(t, d2, p2);
  return dot(dd, dd);
}
__device__ __forceinline__ float oct_sum_m(unsigned om, float v) {
  v += __shfl_xor_sync(om, v, 1);
  v += __shfl_xor_sync(om, v, 2);
  v += __shfl_xor_sync(om, v, 4);
  return v;
}

// Everything the cold pair path needs from the step kernel's registers, handed over through (local) memory so that the cold
// code lives in its own function and leaves the register allocation of the hot kernel alone (measured: inlined it cost 6 us
// per 4096-env step without ever executing).
struct PairIn {
  float E[3][6], S[21], Si[6], g[6], gi[3];     // block factors of M
  SV cd[3];                                      // this leg's c-frame dofs
  float thd[3], xsk[3], awk[3];                  // joint velocities, smooth accelerations, warm-start accelerations
  V3 comr, p;                                    // c-frame origin relative to the base origin; base origin
  float X[9];                                    // hull frame
  V3 prel;                                       // hull frame origin relative to the base origin
  float dr_mu;
};
__device__ __noinline__ int pair_contacts_cold(const PairIn& in, const NmGeom& G, const float4* __restrict__ hull_vert, HullPose* pose_s, PairBlk* pblk,
                                               unsigned cand, float mpr_tol, int mpr_iter, int l, unsigned omask) {
  int npair = 0;
  cand |= __shfl_xor_sync(omask, cand, 1); cand |= __shfl_xor_sync(omask, cand, 2); cand |= __shfl_xor_sync(omask, cand, 4);
  if (l < 6) {                                // stage the six hull poses of the environment
    HullPose& hp = pose_s[l];
#pragma unroll
    for (int k = 0; k < 9; k++) hp.X[k] = in.X[k];
    hp.p[0] = in.prel.x; hp.p[1] = in.prel.y; hp.p[2] = in.prel.z;
    const float* X = in.X;
    hp.c[0] = in.prel.x + X[0] * G.center[0] + X[1] * G.center[1] + X[2] * G.center[2];
    hp.c[1] = in.prel.y + X[3] * G.center[0] + X[4] * G.center[1] + X[5] * G.center[2];
    hp.c[2] = in.prel.z + X[6] * G.center[0] + X[7] * G.center[1] + X[8] * G.center[2];
    hp.adr = G.hull_adr; hp.num = G.hull_num;
  }
  __syncwarp(omask);
  int idx = 0;
#pragma unroll 1
  for (int i = 0; i < 5; i++)
#pragma unroll 1
    for (int j = i + 1; j < 6; j++, idx++) {
      if (!((cand >> idx) & 1u) || npair >= NM_MAXPAIR) continue;          // (uniform within the octet)
      float mo[7];
      if (!mpr_penetration_oct(hull_vert, &pose_s[i], &pose_s[j], mpr_tol, mpr_iter, l, omask, mo)) continue;
      // ---- contact frame (≙ mju_makeFrame on the MPR direction), rows of body j minus rows of body i
      PairBlk& P = pblk[npair];
      const V3 n = normalized(mk(mo[1], mo[2], mo[3]));
      V3 t1 = (n.y < 0.5f && n.y > -0.5f) ? mk(0.f, 1.f, 0.f) : mk(0.f, 0.f, 1.f);
      { const float dd = dot_plain(n, t1); t1 = normalized(mk(t1.x - dd * n.x, t1.y - dd * n.y, t1.z - dd * n.z)); }
      const V3 t2 = cross(n, t1);
      const V3 frm[3] = {n, t1, t2};
      const V3 r = mk(mo[4], mo[5], mo[6]) - in.comr;       // contact point relative to the c-frame origin
      const float sgn = l == i ? -1.f : (l == j ? 1.f : 0.f);
      V3 colk[3];
#pragma unroll
      for (int q = 0; q < 3; q++) colk[q] = sgn * (in.cd[q].v + cross(in.cd[q].w, r));
      const float mu = G.mu * in.dr_mu;                      // (lanes i and j agree: shared contact parameters)
      float Y[3][6], Z[3][3], vb[3], as[3], aw[3];
#pragma unroll 1
      for (int f = 0; f < 3; f++) {
        float Jk[3], Jt[6];
#pragma unroll
        for (int q = 0; q < 3; q++) Jk[q] = dot(frm[f], colk[q]);
#pragma unroll
        for (int a = 0; a < 6; a++) Jt[a] = oct_sum_m(omask, -fmaf(Jk[0], in.E[0][a], fmaf(Jk[1], in.E[1][a], Jk[2] * in.E[2][a])));
        fwd6(in.S, in.Si, Jt, Y[f]);
        fwd3(in.g, in.gi, Jk, Z[f]);
        vb[f] = oct_sum_m(omask, Jk[0] * in.thd[0] + Jk[1] * in.thd[1] + Jk[2] * in.thd[2]);
        as[f] = oct_sum_m(omask, fmaf(Jk[0], in.xsk[0], fmaf(Jk[1], in.xsk[1], Jk[2] * in.xsk[2])));
        aw[f] = oct_sum_m(omask, fmaf(Jk[0], in.awk[0], fmaf(Jk[1], in.awk[1], Jk[2] * in.awk[2])));
      }
      const float dist = G.margin - mo[0];
      const float pos = dist - G.margin;
      const float imp = impedance(G, pos);
      const float rself = oct_sum_m(omask, (l == i || l == j) ? G.rfac_self * (in.dr_mu * in.dr_mu) * (1.f + mu * mu) / (1.f + G.mu * G.mu) : 0.f);
      const float R = fmaxf(rself * (1.f - imp) / imp, NM_MINVAL);
      const float kd = G.K * imp * pos, rinv = 1.f / R;
      float ey[4][6], ez[4][3];
#pragma unroll
      for (int e = 0; e < 4; e++) {
        const float sg = (e & 1) ? -mu : mu;
#pragma unroll
        for (int a = 0; a < 6; a++) ey[e][a] = fmaf(sg, Y[1 + (e >> 1)][a], Y[0][a]);
#pragma unroll
        for (int q = 0; q < 3; q++) ez[e][q] = fmaf(sg, Z[1 + (e >> 1)][q], Z[0][q]);
      }
      float Gm[10];
      const int gi_[10] = {0, 1, 2, 3, 0, 2, 0, 0, 1, 1}, gj_[10] = {0, 1, 2, 3, 1, 3, 2, 3, 2, 3};
#pragma unroll 1
      for (int k = 0; k < 10; k++) {
        float t = 0.f;
        for (int q = 0; q < 3; q++) t = fmaf(ez[gi_[k]][q], ez[gj_[k]][q], t);
        t = oct_sum_m(omask, t);                           // both legs' parts
        for (int a = 0; a < 6; a++) t = fmaf(ey[gi_[k]][a], ey[gj_[k]][a], t);
        Gm[k] = t;
      }
      if (l == i || l == j) {
        float (*Zs)[3] = l == i ? P.Zi : P.Zj;
#pragma unroll
        for (int f = 0; f < 3; f++)
#pragma unroll
          for (int q = 0; q < 3; q++) Zs[f][q] = Z[f][q];
      }
      if (l == i) {                                        // the octet-uniform part is written once, by the lane whose geom parameters apply
#pragma unroll
        for (int f = 0; f < 3; f++)
#pragma unroll
          for (int a = 0; a < 6; a++) P.Y[f][a] = Y[f][a];
#pragma unroll
        for (int k = 0; k < 10; k++) P.G[k] = Gm[k];
#pragma unroll
        for (int e = 0; e < 4; e++) {
          const float sg = (e & 1) ? -mu : mu;
          const int t = 1 + (e >> 1);
          const float aref = -G.B * fmaf(sg, vb[t], vb[0]) - kd;
          P.b[e] = fmaf(sg, as[t], as[0]) - aref;
          P.adi[e] = 1.f / (Gm[e] + R);
          const float jar = fmaf(sg, aw[t], aw[0]) - aref;
          P.f[e] = jar < 0.f ? -jar * rinv : 0.f;
        }
#pragma unroll
        for (int t = 0; t < 2; t++) {
          const float K1 = Gm[2 * t] + Gm[2 * t + 1] - 2.f * Gm[4 + t];
          P.ik[t] = K1 < NM_MINVAL ? 0.f : 1.f / K1;
        }
        P.R = R; P.dist = dist; P.mu = mu;
        P.pos[0] = mo[4] + in.p.x; P.pos[1] = mo[5] + in.p.y; P.pos[2] = mo[6] + in.p.z;
        P.nrm[0] = n.x; P.nrm[1] = n.y; P.nrm[2] = n.z;
        P.li = i; P.lj = j;
      }
      npair++;
    }
  __syncwarp(omask);
  return npair;
}

enum { RW_ACTION_RATE = 0, RW_ANG_VEL_XY, RW_BASE_HEIGHT, RW_BODY_CONTACT_FORCES, RW_COLLISION, RW_DEFAULT_POSITION,
       RW_DOF_ACC, RW_DOF_VEL, RW_FEET_AIR_TIME, RW_FEET_CONTACT_FORCES, RW_FEET_STUMBLE, RW_LIN_VEL_Z, RW_ORIENTATION,
       RW_STAND_STILL, RW_TERMINATION, RW_TORQUES, RW_TRACKING_ANG_VEL, RW_TRACKING_LIN_VEL };

// ================================================================================================ the step kernel
// Two launch shapes of the same code, both on the 255-register budget (8 warps/SM): <BLOCK=128, MINB=2> for batches that fit
// one wave, <BLOCK=256, MINB=1> (one CTA per SM, its eight warps marching through the code together) for large batches.
// Measured at 131 072 envs (round 2, gpurun_out/r02_qb13.log): <256,1> 74.7 M env-steps/s, <256,2> (128 registers, 16 warps/SM,
// spills) 71.3 M, <384,1> 66.8 M, <512,1> 63.2 M; at 4096 envs <64,4>, <128,2> and <256,1> all take 88.8 us -- one wave is
// as slow as its slowest warp, whatever shares the SM with it.
#ifndef NM_LARGE_BLOCK
#define NM_LARGE_BLOCK 256
#define NM_LARGE_MINB 1
#endif
#ifndef NM_MID_BLOCK
#define NM_MID_BLOCK 224      // 7 warps per SM: for batches whose last round of 8-warp CTAs would leave most SMs idle (nm_launch_step)
#define NM_MID_MINB 1
#endif
#ifndef NM_SMALL_BLOCK
#define NM_SMALL_BLOCK 128
#define NM_SMALL_MINB 2
#endif
template <bool ENV, int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB) nm_step_kernel(const NmKernelArgs A) {
  __shared__ NmDevModel sm;
  __shared__ NmDevCfg scfg;
  // observations of the CTA's environments, staged so that they leave the SM as fully coalesced rows (to HBM and,
  // for host-resident callers, straight to pinned host memory over PCIe)
  __shared__ __align__(16) float obs_tile[ENV ? (BLOCK / NM_OCT) * NM_NOBS_DEV : 4];
  // convex-convex pairs (cold): per environment 6 hull poses and up to NM_MAXPAIR contact blocks, dynamic shared memory
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  HullPose* const pose_s = reinterpret_cast<HullPose*>(dyn_smem) + (threadIdx.x >> 3) * 6;
  PairBlk* const pblk = reinterpret_cast<PairBlk*>(dyn_smem + sizeof(HullPose) * 6 * (BLOCK / NM_OCT)) + (threadIdx.x >> 3) * NM_MAXPAIR;
  // bounding capsules of the six leg hulls (hot): start point and direction, one float4 pair per leg
  float4* const cap_s = reinterpret_cast<float4*>(dyn_smem + (sizeof(HullPose) * 6 + sizeof(PairBlk) * NM_MAXPAIR) * (BLOCK / NM_OCT)) + (threadIdx.x >> 3) * 12;
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const int env_raw = gtid >> 3;
  const bool valid = env_raw < A.num_envs;
  const int env = valid ? env_raw : A.num_envs - 1;
  const int l = threadIdx.x & 7;                 // lane inside the octet
  const int lane = threadIdx.x & 31;
  const int obase = lane & 24;                   // first warp lane of this octet
  // ------------------------------------------------------------------ load state (registers for the whole step)
  // Issued BEFORE the constant tables are copied to shared memory: with the L2 flushed both are DRAM round trips, and the
  // state rows do not depend on the tables (a leg lane's dof offset is 3 l whatever the model says; lanes that turn out not
  // to be legs discard what they read).
  const float* qp = A.qpos + (size_t)env * 25;
  const float* qv = A.qvel + (size_t)env * 24;
  const float* qw = A.warm + (size_t)env * 24;
  V3 p = ld3(qp);
  float q0 = qp[3], q1 = qp[4], q2 = qp[5], q3 = qp[6];
  V3 vlin = ld3(qv), wloc = ld3(qv + 3);
  float awb[6], awk[3], th[3], thd[3], ctrl[3];
#pragma unroll
  for (int i = 0; i < 6; i++) awb[i] = qw[i];
  const int jo6 = l < 6 ? 3 * l : 0;
#pragma unroll
  for (int j = 0; j < 3; j++) { th[j] = qp[7 + jo6 + j]; thd[j] = qv[6 + jo6 + j]; awk[j] = qw[6 + jo6 + j]; }
  float in0[3], in1[3], in2[3], in3[3];          // env mode: previous actions, new actions, previous dof velocities, carried dof positions
#pragma unroll
  for (int j = 0; j < 3; j++) {
    if (ENV) {
      in0[j] = A.actions[(size_t)env * 18 + jo6 + j];
      in1[j] = A.in_actions[(size_t)env * A.act_stride + jo6 + j];
      in2[j] = A.dof_vel[(size_t)env * 18 + jo6 + j];
      in3[j] = A.dof_pos[(size_t)env * 18 + jo6 + j];
    } else {
      in0[j] = A.in_ctrl[(size_t)env * 18 + jo6 + j];
      in1[j] = in2[j] = in3[j] = 0.f;
    }
  }
  {
    static_assert(sizeof(NmDevModel) % 16 == 0 && sizeof(NmDevCfg) % 16 == 0, "constant tables are copied as int4");
    const int4* src = reinterpret_cast<const int4*>(A.model);
    int4* dst = reinterpret_cast<int4*>(&sm);
    for (int i = threadIdx.x; i < (int)(sizeof(NmDevModel) / 16); i += blockDim.x) dst[i] = __ldg(src + i);
    if (ENV) {
      const int4* s2 = reinterpret_cast<const int4*>(A.cfg);
      int4* d2 = reinterpret_cast<int4*>(&scfg);
      for (int i = threadIdx.x; i < (int)(sizeof(NmDevCfg) / 16); i += blockDim.x) d2[i] = __ldg(s2 + i);
    }
  }
  // hull tables of the support-vertex walk, optionally staged once per CTA (experiment switch NM_HULL_SMEM=1, see nm_abi.cu:
  // measured neutral -- the walk is ~3 rounds of ~120 instructions per substep and issue bound with the warps in lockstep)
  const float4* hvt = A.hull_vert;
  const unsigned short* nbt = A.hull_nbr16;
  const unsigned short* nat = A.hull_nadr16;
  if (A.hull_smem) {
    unsigned char* base = dyn_smem + (sizeof(HullPose) * 6 + sizeof(PairBlk) * NM_MAXPAIR + 12 * sizeof(float4)) * (BLOCK / NM_OCT);
    int4* hs = reinterpret_cast<int4*>(base);
    int4* ns = hs + A.hull_nv;
    int4* as = ns + A.hull_ne_pad / 8;
    const int4 *gv = reinterpret_cast<const int4*>(A.hull_vert), *gn = reinterpret_cast<const int4*>(A.hull_nbr16), *ga = reinterpret_cast<const int4*>(A.hull_nadr16);
    for (int i = threadIdx.x; i < A.hull_nv; i += BLOCK) hs[i] = __ldg(gv + i);
    for (int i = threadIdx.x; i < A.hull_ne_pad / 8; i += BLOCK) ns[i] = __ldg(gn + i);
    for (int i = threadIdx.x; i < A.hull_na_pad / 8; i += BLOCK) as[i] = __ldg(ga + i);
    hvt = reinterpret_cast<const float4*>(hs);
    nbt = reinterpret_cast<const unsigned short*>(ns);
    nat = reinterpret_cast<const unsigned short*>(as);
  }
  __syncthreads();

  TSTAMP(0);
  const NmLeg& L = sm.leg[l];
  const NmGeom& G = L.geom;
  const float isleg = L.isleg;
  const bool leg = l < sm.nleg;
  const float h = sm.timestep;

  const int jo = leg ? 3 * l : 0;
#pragma unroll
  for (int j = 0; j < 3; j++) {
    th[j] = leg ? th[j] : 0.f;
    thd[j] = leg ? thd[j] : 0.f;
    awk[j] = leg ? awk[j] : 0.f;
  }

  // ------------------------------------------------------------------ E1/E3: actions -> ctrl   (env.py:152-188)
  float act[3] = {0.f, 0.f, 0.f}, prev_act[3] = {0.f, 0.f, 0.f}, prev_dof_vel[3] = {0.f, 0.f, 0.f};
  if (ENV) {
#pragma unroll
    for (int j = 0; j < 3; j++) {
      if (leg) {
        prev_act[j] = in0[j];
        float a = in1[j] * scfg.action_scale;
        act[j] = fminf(fmaxf(a, -scfg.clip_actions), scfg.clip_actions);
        prev_dof_vel[j] = in2[j];
        float dpos = in3[j];                                        // carried buffer, stale after a reset (quirk Q2)
        ctrl[j] = ((act[j] - scfg.default_pos[jo + j]) - dpos) * scfg.p_gain;
      } else ctrl[j] = 0.f;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 3; j++) ctrl[j] = leg ? in0[j] : 0.f;
  }

  // outputs of the LAST substep's forward pass that the env layer reads (stale by one substep, quirk Q4)
  SV cvel_b; cvel_b.w = mk(0, 0, 0); cvel_b.v = mk(0, 0, 0);
  float sens0 = 0.f, sens1 = 0.f, base_height = 0.f;
  int bad = 0;
  ConBlk cb;
  // support-vertex hint of this lane's hull (pure accelerator: any start vertex gives the same support vertex up to exact ties)
  // packed: vertex (9 bits) | its degree (6 bits, 0 = unknown) | first edge of its neighbour list (from bit 15)
  int hint = G.has ? A.hull_hint[(size_t)env * NM_OCT + l] : 0;
  hint = (hint >= 0 && (hint & 0x1ff) < G.hull_num) ? hint : 0;
  // domain randomisation (opt-in, NOT in the reference): per-env scales of contact friction, actuator kv and base mass.
  // All 1 when disabled -- multiplying by 1.0f is exact, so the reference path is bit-identical with or without it.
  float dr_mu = 1.f, dr_kv = 1.f, dr_bm = 1.f;
  if (A.dr != nullptr) { const float4 d4 = __ldg(reinterpret_cast<const float4*>(A.dr) + env); dr_mu = d4.x; dr_kv = d4.y; dr_bm = d4.z; }
  const float b_mass = sm.b_mass * dr_bm, total_mass = fmaf(sm.b_mass, dr_bm - 1.f, sm.total_mass);

#pragma unroll 1
  for (int sub = 0; sub < A.nstep; sub++) {
    TSTAMP(1 + 10 * sub);
    PHASE_SYNC_A();
    TSTAMP(2 + 10 * sub);
    // ================================================================ P1 kinematics
    {
      float n2 = q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3;
      if (n2 < NM_MINVAL) { q0 = 1.f; q1 = q2 = q3 = 0.f; }
      else { float inv = rsqrtf(n2); q0 *= inv; q1 *= inv; q2 *= inv; q3 *= inv; }
    }
    const M3 Rb = quat2mat(q0, q1, q2, q3);
    M3 X = Rb;
    V3 pj = p;
    V3 axw[3], anc[3], xip[3];
    float Iw[3][6];
#pragma unroll
    for (int j = 0; j < 3; j++) {
      pj = pj + mul(X, ld3(L.pos[j]));
      M3 Rj = axis_rot(ld3(L.axis[j]), th[j] - L.qref[j]);
      if (!L.rc_ident[j]) { M3 Rc;
#pragma unroll
        for (int k = 0; k < 9; k++) Rc.a[k] = L.rc[j][k];
        Rj = matmul(Rc, Rj); }
      X = matmul(X, Rj);
      axw[j] = mul(X, ld3(L.axis[j]));
      anc[j] = pj;
      xip[j] = pj + mul(X, ld3(L.ipos[j]));
      rot_inertia(X, L.iloc[j], Iw[j]);
    }
    const M3 Xg = X;          // frame of the lane's collision geom (tibia; base for the pseudo-legs)
    const V3 pg = pj;
    const V3 xip_b = p + mul(Rb, ld3(sm.b_ipos));
    float Iwb[6];
    rot_inertia(Rb, sm.b_iloc, Iwb);
#pragma unroll
    for (int i = 0; i < 6; i++) Iwb[i] *= dr_bm;

    // ================================================================ P2 comPos: subtree COM, c-frame inertias and dofs
    V3 msum = fma3(L.mass[0], xip[0], fma3(L.mass[1], xip[1], L.mass[2] * xip[2]));
    msum = oct_sum3(msum);
    const V3 com = (1.f / total_mass) * fma3(b_mass, xip_b, msum);
    In ci[3];
    SV cd[3];
#pragma unroll
    for (int j = 0; j < 3; j++) {
      ci[j] = make_inertia(Iw[j], L.mass[j], xip[j] - com);
      cd[j].w = axw[j];
      cd[j].v = cross(axw[j], com - anc[j]);
    }
    const In cib = make_inertia(Iwb, b_mass, xip_b - com);
    const V3 offb = com - p;
    SV cdr[3];                 // base rotational dofs (body-frame axes); translational dofs are [0; e_i]
#pragma unroll
    for (int i = 0; i < 3; i++) { cdr[i].w = col(Rb, i); cdr[i].v = cross(cdr[i].w, offb); }

    // ================================================================ P7 comVel + RNE bias forces
    const V3 omega = mul(Rb, wloc);
    SV cvb; cvb.w = omega; cvb.v = vlin + cross(omega, offb);
    SV cab; cab.w = mk(0, 0, 0); cab.v = cross(vlin, omega) - ld3(sm.gravity);
    float bias_k[3], bias_b[6];
    {
      SV fb = imul(cib, cab) + cross_force(cvb, imul(cib, cvb));
      SV cv = cvb, ca = cab, f[3];
#pragma unroll
      for (int j = 0; j < 3; j++) {
        SV cdd = cross_motion(cv, cd[j]);
        ca = sfma(thd[j], cdd, ca);
        cv = sfma(thd[j], cd[j], cv);
        f[j] = imul(ci[j], ca) + cross_force(cv, imul(ci[j], cv));
      }
      f[1] = f[1] + f[2];
      f[0] = f[0] + f[1];
      bias_k[2] = sdot(cd[2], f[2]); bias_k[1] = sdot(cd[1], f[1]); bias_k[0] = sdot(cd[0], f[0]);
      SV Fb;
      Fb.w = fb.w + oct_sum3(f[0].w);
      Fb.v = fb.v + oct_sum3(f[0].v);
      bias_b[0] = Fb.v.x; bias_b[1] = Fb.v.y; bias_b[2] = Fb.v.z;
#pragma unroll
      for (int i = 0; i < 3; i++) bias_b[3 + i] = sdot(cdr[i], Fb);
    }

    TSTAMP(3 + 10 * sub);
    PHASE_SYNC_B();
    // ================================================================ P3 CRBA in block form
    float Mk[6], C[3][6], Mbb[21];
    {
      In crb2 = ci[2], crb1 = ci[1] + crb2, crb0 = ci[0] + crb1;
      In s = crb0;
      In cb;
      cb.xx = cib.xx + oct_sum(s.xx); cb.yy = cib.yy + oct_sum(s.yy); cb.zz = cib.zz + oct_sum(s.zz);
      cb.xy = cib.xy + oct_sum(s.xy); cb.xz = cib.xz + oct_sum(s.xz); cb.yz = cib.yz + oct_sum(s.yz);
      cb.h = cib.h + oct_sum3(s.h); cb.m = cib.m + oct_sum(s.m);
      SV b0 = imul(crb0, cd[0]), b1 = imul(crb1, cd[1]), b2 = imul(crb2, cd[2]);
      Mk[0] = sdot(cd[0], b0) + L.armature[0];
      Mk[1] = sdot(cd[0], b1); Mk[2] = sdot(cd[1], b1) + L.armature[1];
      Mk[3] = sdot(cd[0], b2); Mk[4] = sdot(cd[1], b2); Mk[5] = sdot(cd[2], b2) + L.armature[2];
      const SV bb[3] = {b0, b1, b2};
#pragma unroll
      for (int i = 0; i < 3; i++) {
        C[i][0] = bb[i].v.x * isleg; C[i][1] = bb[i].v.y * isleg; C[i][2] = bb[i].v.z * isleg;
#pragma unroll
        for (int r = 0; r < 3; r++) C[i][3 + r] = sdot(cdr[r], bb[i]) * isleg;
      }
      // pseudo-legs carry no mass: make their private block the identity so the factorisation is well defined
      Mk[0] = fmaf(Mk[0], isleg, 1.f - isleg); Mk[2] = fmaf(Mk[2], isleg, 1.f - isleg); Mk[5] = fmaf(Mk[5], isleg, 1.f - isleg);
      Mk[1] *= isleg; Mk[3] *= isleg; Mk[4] *= isleg;
#pragma unroll
      for (int i = 0; i < 21; i++) Mbb[i] = 0.f;
      Mbb[TRI(0, 0)] = cb.m; Mbb[TRI(1, 1)] = cb.m; Mbb[TRI(2, 2)] = cb.m;
#pragma unroll
      for (int r = 0; r < 3; r++) {
        SV br = imul(cb, cdr[r]);
        Mbb[TRI(3 + r, 0)] = br.v.x; Mbb[TRI(3 + r, 1)] = br.v.y; Mbb[TRI(3 + r, 2)] = br.v.z;
#pragma unroll
        for (int s2 = 0; s2 <= r; s2++) Mbb[TRI(3 + r, 3 + s2)] = sdot(cdr[s2], br);
      }
    }

    TSTAMP(4 + 10 * sub);
    PHASE_SYNC_C();
    // ================================================================ P8 actuation + smooth acceleration
    float rk[3], rb[6], hD[3];
#pragma unroll
    for (int j = 0; j < 3; j++) {
      float c = fminf(fmaxf(ctrl[j], L.clo[j]), L.chi[j]);
      float g = L.gear[j];
      float frc = fmaf(L.gain0[j] * dr_kv, c, L.bias0[j]) + L.bias1[j] * (th[j] * g) + (L.bias2[j] * dr_kv) * (thd[j] * g);
      frc = fminf(fmaxf(frc, L.flo[j]), L.fhi[j]);
      rk[j] = (g * frc - L.damping[j] * thd[j] - bias_k[j]) * isleg;
      hD[j] = h * (sm.imp_damp * L.damping[j] - sm.imp_act * (L.bias2[j] * dr_kv) * g * g) * isleg;   // -h * d(qfrc_smooth)/d(qvel), diagonal
    }
#pragma unroll
    for (int i = 0; i < 6; i++) rb[i] = -bias_b[i];
    Factor F, FH;
    {
      // (a single copy of the factorisation run twice in a loop was 2 us faster and produced wrong in-contact results on the
      // GPU -- gpurun_out/r02_t_default.log vs r02_t_ftwice.log -- so the two factorisations stay inlined)
      const float zero3[3] = {0.f, 0.f, 0.f};
      factor_system(Mk, C, Mbb, zero3, F);
      factor_system(Mk, C, Mbb, hD, FH);      // (M - h*qDeriv) for the implicit velocity update
    }
    float xsb[6], xsk[3];                      // qacc_smooth
    solve_system(F, rb, rk, xsb, xsk);

    TSTAMP(5 + 10 * sub);
    PHASE_SYNC_D();
    if (sub == 0) TSTAMP(23);
    // ================================================================ P4 collision: convex hull vs plane
    // Support vertex by hill-climbing the hull graph (a local minimum of a linear function on a convex hull is the
    // global one), warm-started from the previous substep's / step's support vertex.  The hint word also remembers
    // where that vertex's neighbour list lives, and the edge table stores every neighbour's COORDINATES next to its
    // id, so the usual case (support vertex unchanged) costs one level of loads instead of three dependent ones.
    const V3 pn = ld3(sm.plane_n);
    int nc = 0;
    float cdist[NM_MAXC];
    int cvert[NM_MAXC];
    if (G.has) {
      const V3 dl = mulT(Xg, pn);                        // plane normal in the geom frame; minimise dl . v
      const float4* hv = hvt + G.hull_adr;
      const unsigned short* nadr = nat + G.hull_adr;
      int best = hint & 0x1ff, deg = (hint >> 9) & 0x3f, e0 = hint >> 15;
      float4 vb = hv[best];
      float bval = fmaf(dl.x, vb.x, fmaf(dl.y, vb.y, dl.z * vb.z));
#ifdef NM_TIMING
      int dbg_rounds = 0, dbg_trips = 0;
#endif
      for (;;) {
#ifdef NM_TIMING
        dbg_rounds++;
#endif
        if (deg == 0) { e0 = nadr[best]; deg = nadr[best + 1] - e0; }
        int nb = best;
        const int el = e0 + deg - 1;
        for (int e = e0; e <= el; e += 4) {              // 4 neighbours per trip: loads issued together, compared in list order
          const float4 w0 = ld_edge(nbt, hv, e), w1 = ld_edge(nbt, hv, min(e + 1, el)), w2 = ld_edge(nbt, hv, min(e + 2, el)), w3 = ld_edge(nbt, hv, min(e + 3, el));
          const float a0 = fmaf(dl.x, w0.x, fmaf(dl.y, w0.y, dl.z * w0.z)), a1 = fmaf(dl.x, w1.x, fmaf(dl.y, w1.y, dl.z * w1.z));
          const float a2 = fmaf(dl.x, w2.x, fmaf(dl.y, w2.y, dl.z * w2.z)), a3 = fmaf(dl.x, w3.x, fmaf(dl.y, w3.y, dl.z * w3.z));
          if (a0 < bval) { bval = a0; nb = __float_as_int(w0.w); vb = w0; }
          if (a1 < bval) { bval = a1; nb = __float_as_int(w1.w); vb = w1; }
          if (a2 < bval) { bval = a2; nb = __float_as_int(w2.w); vb = w2; }
          if (a3 < bval) { bval = a3; nb = __float_as_int(w3.w); vb = w3; }
        }
        if (nb == best) break;
        best = nb; deg = 0;
      }
      hint = best | (deg << 9) | (e0 << 15);
#ifdef NM_TIMING
      if (sub == 0 && nm_timing_buf) {
        const unsigned am = __activemask();
        int r = dbg_rounds, dg = deg;
        for (int o = 16; o > 0; o >>= 1) { r = max(r, __shfl_xor_sync(am, r, o)); dg = max(dg, __shfl_xor_sync(am, dg, o)); }
        if ((threadIdx.x & 31) == __ffs(am) - 1) { nm_timing_buf[(size_t)(gtid >> 5) * 32 + 27] = r; nm_timing_buf[(size_t)(gtid >> 5) * 32 + 28] = dg; }
        (void)dbg_trips;
      }
      if (sub == 0) TSTAMP(24);
#endif
      V3 wv = pg + mul(Xg, mk(vb.x, vb.y, vb.z));
      float dist = dot(pn, wv) - sm.plane_d;
      if (dist <= G.margin) {
        cdist[0] = dist; cvert[0] = best; cb.pos[0] = fma3(-0.5f * dist, pn, wv); nc = 1;
        const float thr2 = (sm.planemesh_sep * G.rbound) * (sm.planemesh_sep * G.rbound);
        const float dpl = dot(pn, pg) - sm.plane_d;       // cheap pre-test in the geom frame: dist(u) ~= dl.v_u + dpl
        const int maxc = sm.planemesh_maxcon;              // model option, <= NM_MAXC (kept out of the loop bounds: they stay compile-time)
        // candidates (model option): the support vertex's hull-graph neighbours from the edge table, or every hull vertex in
        // index order from the vertex table (both tables hold float4 coordinates)
        const bool allv = sm.planemesh_allverts != 0;
        const int ef = allv ? 0 : e0, el = allv ? G.hull_num - 1 : e0 + deg - 1;
        const float hsep = sm.planemesh_sepvert ? 0.5f : 0.f;   // separation between contact points (0) or hull vertices (point + dist/2 n)
        for (int e = ef; e <= el && nc < NM_MAXC; e += 4) {   // up to maxc - 1 more
          // four candidates per trip (L1-resident after the walk); almost all fail the cheap depth pre-test
          float4 ww[4];
#pragma unroll
          for (int k = 0; k < 4; k++) ww[k] = allv ? hv[min(e + k, el)] : ld_edge(nbt, hv, min(e + k, el));
#pragma unroll
          for (int k = 0; k < 4; k++) {
            const float4 w4 = ww[k];
            const int vid = allv ? e + k : __float_as_int(w4.w);
            if (e + k > el || nc >= NM_MAXC || (allv && vid == best) || fmaf(dl.x, w4.x, fmaf(dl.y, w4.y, dl.z * w4.z)) + dpl > G.margin + 1e-4f) continue;
            V3 wu = pg + mul(Xg, mk(w4.x, w4.y, w4.z));
            float du = dot(pn, wu) - sm.plane_d;
            if (du > G.margin) continue;
            V3 cp = fma3(-0.5f * du, pn, wu);
            bool close = false;
            for (int q = 0; q < nc; q++) { V3 d3 = fma3(hsep * (cdist[q] - du), pn, cb.pos[q] - cp); close |= dot(d3, d3) < thr2; }
            if (close || nc >= maxc) continue;
            cdist[nc] = du; cvert[nc] = vid; cb.pos[nc] = cp; nc++;
          }
        }
      }
    }
    if (sub == 0) TSTAMP(25);
    // ================================================================ P4b convex-convex pairs between the legs' hulls
    // Broad phase on bounding capsules (hot, ~3 segment-segment tests per lane); MPR + contact block only for overlapping
    // capsules (cold: the whole warp skips it when none of its 4 environments has a candidate).
    const unsigned omask = 0xffu << obase;
    int npair = 0;                                  // pair contacts of this environment (same in all 8 lanes)
#ifdef NM_NO_PAIRS
    if (false) {
#else
    if (sm.pair_mask != 0) {
#endif
      const V3 prel = pg - p;                       // hull frame origin relative to the base origin
      const V3 ca = prel + mul(Xg, ld3(G.cap_a)), cdir = mul(Xg, ld3(G.cap_b) - ld3(G.cap_a));
      if (l < 6) {
        cap_s[2 * l] = make_float4(ca.x, ca.y, ca.z, cdir.x);
        cap_s[2 * l + 1] = make_float4(cdir.y, cdir.z, 0.f, 0.f);
      }
      __syncwarp();
      unsigned cand = 0u;                           // bit idx(i,j) of the candidate pairs this lane found
      if (l < 6) {
#pragma unroll
        for (int dlt = 1; dlt <= 3; dlt++) {
          const int other = (l + dlt) % 6;
          const int i = min(l, other), j = max(l, other);
          const int idx = i * 5 - (i * (i - 1)) / 2 + (j - i - 1);          // lexicographic index of (i, j), i < j < 6
          if ((dlt == 3 && l >= 3) || !((sm.pair_mask >> idx) & 1)) continue;
          const float4 q0 = cap_s[2 * other], q1 = cap_s[2 * other + 1];
          const NmGeom& Go = sm.leg[other].geom;
          const float thr = G.cap_r + Go.cap_r + 1e-4f;
          const V3 oa = mk(q0.x, q0.y, q0.z), od = mk(q0.w, q1.x, q1.y);
          // one separating-axis test first, along the line between the capsule midpoints (a capsule projects onto a unit axis
          // a as centre.a +- (|dir.a| / 2 + r)): tibias that stand side by side are separated by it, the segment-segment
          // distance is only needed for legs that really come close
          const V3 mm = fma3(0.5f, cdir, ca) - fma3(0.5f, od, oa);
          const float m2 = dot(mm, mm);
          if (m2 > fmaf(0.5f, fabsf(dot(cdir, mm)) + fabsf(dot(od, mm)), thr * sqrtf(m2))) continue;
          if (segseg_dist2(ca, cdir, G.cap_il2, oa, od, Go.cap_il2) < thr * thr) cand |= 1u << idx;
        }
      }
#ifdef NM_PAIRS_E3
      if (A.debug != nullptr && cand != 0u) A.debug[(size_t)env * NM_DBG + 5] = __uint_as_float(cand);   // experiment: broad phase kept alive, nothing else
      if (false) {
#else
      if (__builtin_expect(__any_sync(FULL, cand != 0u), 0)) {          // cold: at least one of the warp's 4 environments has a candidate pair
#endif
        unsigned oc = cand;
        oc |= __shfl_xor_sync(FULL, oc, 1); oc |= __shfl_xor_sync(FULL, oc, 2); oc |= __shfl_xor_sync(FULL, oc, 4);
#ifdef NM_TIMING
        int dbg_c0 = 0, dbg_c1 = 0;
        const long long dbg_t0 = clock64();
#endif
        // First step of MPR, here and without touching a vertex: the hulls are disjoint when they are separated along the line
        // between their centres, h_i(dir) + h_j(-dir) <= 0 with h(d) = max over hull vertices of d.x.  That is how MPR itself
        // answers almost every candidate (pairs that are close but do not touch).  h comes from the hull's support map
        // (NM_SMAP_N: an upper bound, four loads), so the test is conservative; the pairs it keeps go to MPR, whose own first
        // step is the exact one.  Code size matters more than arithmetic here: one warp in ~40 enters per substep on a walking
        // batch, always with cold instruction lines (the exact scan over both hulls' vertices that stood here cost 9.3 us per
        // visit, `tools/phase_timing.py`, and a one-wave launch lasts as long as its slowest warp).
        if (oc != 0u && !A.pair_filter_off) {
          const V3 cw = prel + mul(Xg, ld3(G.center));
          unsigned rem = oc;
#ifdef NM_TIMING
          dbg_c0 = __popc(oc);
#endif
          oc = 0u;
#pragma unroll 1
          while (rem != 0u) {
            const int idx = __ffs(rem) - 1;
            rem &= rem - 1u;
            int i = 0, base = 0;
            while (idx >= base + 5 - i) { base += 5 - i; i++; }                // lexicographic (i, j) of the pair
            const int j = i + 1 + (idx - base);
            const int si = obase | i, sj = obase | j;
            const V3 ci = mk(__shfl_sync(omask, cw.x, si), __shfl_sync(omask, cw.y, si), __shfl_sync(omask, cw.z, si));
            const V3 cj = mk(__shfl_sync(omask, cw.x, sj), __shfl_sync(omask, cw.y, sj), __shfl_sync(omask, cw.z, sj));
            const V3 dir = normalized(cj - ci);
            const V3 dloc = mulT(Xg, l == j ? mk(-dir.x, -dir.y, -dir.z) : dir);   // the direction in this lane's hull frame
            const float off = dot_plain(prel, dir);
            const float hme = (l == i || l == j) ? smap_support(A.hull_smap + G.smap_adr, dloc) : 0.f;
            const float hi = __shfl_sync(omask, hme, si), hj = __shfl_sync(omask, hme, sj);
            const float oi = __shfl_sync(omask, off, si), oj = __shfl_sync(omask, off, sj);
            if ((hi + oi) + (hj - oj) > -1e-6f) oc |= 1u << idx;                // not clearly separated along this axis: MPR decides
          }
#ifdef NM_TIMING
          dbg_c1 = __popc(oc);
#endif
        }
#ifdef NM_TIMING
        const long long dbg_t1 = clock64();
        if (sub == 0 && nm_timing_buf && l == 0)     // packed event counts: capsules overlap | survive the support-map test << 20 | handed to MPR << 40
          atomicAdd((unsigned long long*)&nm_timing_buf[(size_t)(gtid >> 5) * 32 + 29],
                    (unsigned long long)dbg_c0 | ((unsigned long long)dbg_c1 << 20) | ((unsigned long long)__popc(oc) << 40));
#endif
        if (oc != 0u) {                             // (uniform within the octet)
          cand = oc;
          PairIn in;
#pragma unroll
          for (int q = 0; q < 3; q++) {
#pragma unroll
            for (int a = 0; a < 6; a++) in.E[q][a] = F.E[q][a];
            in.cd[q] = cd[q]; in.thd[q] = thd[q]; in.xsk[q] = xsk[q]; in.awk[q] = awk[q]; in.gi[q] = F.gi[q];
          }
#pragma unroll
          for (int a = 0; a < 21; a++) in.S[a] = F.S[a];
#pragma unroll
          for (int a = 0; a < 6; a++) { in.Si[a] = F.Si[a]; in.g[a] = F.g[a]; }
#pragma unroll
          for (int k = 0; k < 9; k++) in.X[k] = Xg.a[k];
          in.comr = com - p; in.p = p; in.prel = prel; in.dr_mu = dr_mu;
          npair = pair_contacts_cold(in, G, A.hull_vert, pose_s, pblk, cand, sm.mpr_tolerance, sm.mpr_iterations, l, omask);
        }
        __syncwarp();
#ifdef NM_TIMING
        if (sub == 0 && nm_timing_buf && (threadIdx.x & 31) == 0) {      // cycles of the candidate filter and of MPR + contact blocks
          nm_timing_buf[(size_t)(gtid >> 5) * 32 + 30] = dbg_t1 - dbg_t0;
          nm_timing_buf[(size_t)(gtid >> 5) * 32 + 31] = clock64() - dbg_t1;
        }
#endif
      }
    }
    int npair_max = max(npair, __shfl_xor_sync(FULL, npair, 8));
    npair_max = max(npair_max, __shfl_xor_sync(FULL, npair_max, 16));          // warp-uniform
    if (sub == 0) TSTAMP(26);
    const int ncon_env = oct_sumi(nc) + npair;
    // Sweep order inside an env is MuJoCo's row order: base geom (lane 6) first, then legs 0..5.  Each octet walks ITS
    // OWN list of contact-owning lanes: in slot k of a sweep the k-th owner of every octet works, so a sweep costs
    // max(#owners per env) slots for the warp, not |union of owner lanes over its 4 envs|.
    const unsigned has_bal = __ballot_sync(FULL, nc > 0);
    const unsigned m8 = (has_bal >> obase) & 0x7fu;
    const unsigned ord = ((m8 >> 6) & 1u) | ((m8 & 0x3fu) << 1);          // bit 0 = lane 6, bit 1+k = lane k
    const int my_bit = l == 6 ? 0 : l + 1;
    const int my_slot = __popc(ord & ((1u << my_bit) - 1u));               // my position among this octet's owners
    int nslot = __popc(ord);
    nslot = max(nslot, __shfl_xor_sync(FULL, nslot, 8));
    nslot = max(nslot, __shfl_xor_sync(FULL, nslot, 16));                   // warp-uniform: slots per sweep
    const bool any_contact = nslot > 0 || npair_max > 0;
    unsigned owner_tab = 0u;                                              // 4 bits per slot: owning lane of this octet (0 if none)
    {
      unsigned t = ord;
#pragma unroll
      for (int k = 0; k < 7; k++) {
        const int b = __ffs(t) - 1;                                       // -1 when exhausted
        const unsigned ln = b < 0 ? 0u : (b == 0 ? 6u : (unsigned)(b - 1));
        owner_tab |= ln << (4 * k);
        t &= t - 1u;
      }
    }

    TSTAMP(6 + 10 * sub);
    PHASE_SYNC_E();
    float xb[6], xk[3];          // constraint-induced acceleration M^-1 J^T f (after noslip)
#pragma unroll
    for (int i = 0; i < 6; i++) xb[i] = 0.f;
#pragma unroll
    for (int i = 0; i < 3; i++) xk[i] = 0.f;
    float wsb[6], wsk[3];        // qacc to store as next warm start (after PGS, before noslip)
#pragma unroll
    for (int i = 0; i < 6; i++) wsb[i] = xsb[i];
#pragma unroll
    for (int i = 0; i < 3; i++) wsk[i] = xsk[i];
    float fn_slot0 = 0.f, fn_slot1 = 0.f;
    int dbg_pgs = 0, dbg_noslip = 0, dbg_warm = 0;

    if (any_contact) {
      // ============================================================== P5 contact blocks, whitened
      // Per contact: the three contact-frame Jacobian rows (normal, tangent 1, tangent 2) whitened by the block
      // factors (Y = J~ G_S^-T, Z = J_k G_k^-T), their 3x3 Gram matrix Gm = J M^-1 J^T restricted to the contact,
      // and the affine term split as beta0 + s*beta_t.  The four pyramid edges are Jn +- mu*Jt: every edge
      // quantity (diagonal, residual, coupling) is a +-mu combination of these, so edges are never materialised.
      const V3 fr0 = ld3(sm.frame), fr1 = ld3(sm.frame + 3), fr2 = ld3(sm.frame + 6);
      const float mu = G.mu * dr_mu;
      const float rfac = G.rfac * (dr_mu * dr_mu) * (1.f + mu * mu) / (1.f + G.mu * G.mu);   // R scales with mu^2 (1 + mu^2)
      for (int c = 0; c < nc; c++) {
        const V3 r = cb.pos[c] - com;
        V3 colb[6], colk[3];
        colb[0] = mk(1, 0, 0); colb[1] = mk(0, 1, 0); colb[2] = mk(0, 0, 1);
#pragma unroll
        for (int i = 0; i < 3; i++) colb[3 + i] = cdr[i].v + cross(cdr[i].w, r);
#pragma unroll
        for (int j = 0; j < 3; j++) colk[j] = isleg * (cd[j].v + cross(cd[j].w, r));
        float Y[3][6], Z[3][3], vb[3], as[3], aw[3];
        const V3 frm[3] = {fr0, fr1, fr2};
#pragma unroll
        for (int f = 0; f < 3; f++) {
          float Jb[6], Jk[3], Jt[6];
#pragma unroll
          for (int a = 0; a < 6; a++) Jb[a] = dot(frm[f], colb[a]);
#pragma unroll
          for (int j = 0; j < 3; j++) Jk[j] = dot(frm[f], colk[j]);
#pragma unroll
          for (int a = 0; a < 6; a++) Jt[a] = Jb[a] - fmaf(Jk[0], F.E[0][a], fmaf(Jk[1], F.E[1][a], Jk[2] * F.E[2][a]));
          fwd6(F.S, F.Si, Jt, Y[f]);
          fwd3(F.g, F.gi, Jk, Z[f]);
          vb[f] = Jb[0] * vlin.x + Jb[1] * vlin.y + Jb[2] * vlin.z + Jb[3] * wloc.x + Jb[4] * wloc.y + Jb[5] * wloc.z +
                  Jk[0] * thd[0] + Jk[1] * thd[1] + Jk[2] * thd[2];
          float s1 = 0.f, s2 = 0.f;
#pragma unroll
          for (int a = 0; a < 6; a++) { s1 = fmaf(Jb[a], xsb[a], s1); s2 = fmaf(Jb[a], awb[a], s2); }
#pragma unroll
          for (int j = 0; j < 3; j++) { s1 = fmaf(Jk[j], xsk[j], s1); s2 = fmaf(Jk[j], awk[j], s2); }
          as[f] = s1; aw[f] = s2;
#pragma unroll
          for (int a = 0; a < 6; a++) cb.Y[c][f][a] = Y[f][a];
#pragma unroll
          for (int j = 0; j < 3; j++) cb.Z[c][f][j] = Z[f][j];
        }
        const float pos = cdist[c] - G.margin;
        const float imp = impedance(G, pos);
        const float R = fmaxf(rfac * (1.f - imp) / imp, NM_MINVAL);
        cb.R[c] = R;
        const float kd = G.K * imp * pos;
        const float rinv = 1.f / R;
        {
          float ey[4][6], ez[4][3];                         // the four edges: Jn + mu*Jt1, Jn - mu*Jt1, Jn + mu*Jt2, Jn - mu*Jt2
#pragma unroll
          for (int e = 0; e < 4; e++) {
            const float sg = (e & 1) ? -mu : mu;
#pragma unroll
            for (int a = 0; a < 6; a++) ey[e][a] = fmaf(sg, Y[1 + (e >> 1)][a], Y[0][a]);
#pragma unroll
            for (int j = 0; j < 3; j++) ez[e][j] = fmaf(sg, Z[1 + (e >> 1)][j], Z[0][j]);
          }
          const int gi_[10] = {0, 1, 2, 3, 0, 2, 0, 0, 1, 1}, gj_[10] = {0, 1, 2, 3, 1, 3, 2, 3, 2, 3};
#pragma unroll
          for (int k = 0; k < 10; k++) {
            float t = 0.f;
#pragma unroll
            for (int a = 0; a < 6; a++) t = fmaf(ey[gi_[k]][a], ey[gj_[k]][a], t);
#pragma unroll
            for (int j = 0; j < 3; j++) t = fmaf(ez[gi_[k]][j], ez[gj_[k]][j], t);
            cb.G[c][k] = t;
            if (k < 4) cb.adi[c][k] = 1.f / (t + R);
          }
          // noslip pair curvature K1 = a00 + a11 - 2 a01 is a constant of the contact: inverted once (0 = degenerate)
#pragma unroll
          for (int t = 0; t < 2; t++) {
            const float K1 = cb.G[c][2 * t] + cb.G[c][2 * t + 1] - 2.f * cb.G[c][4 + t];
            cb.ik[c][t] = K1 < NM_MINVAL ? 0.f : 1.f / K1;
          }
#pragma unroll
          for (int e = 0; e < 4; e++) {
            const float sg = (e & 1) ? -mu : mu;
            const int t = 1 + (e >> 1);
            const float aref = -G.B * fmaf(sg, vb[t], vb[0]) - kd;
            cb.b[c][e] = fmaf(sg, as[t], as[0]) - aref;
            const float jar = fmaf(sg, aw[t], aw[0]) - aref;        // warm start: edge forces implied by qacc_warmstart
            cb.f[c][e] = jar < 0.f ? -jar * rinv : 0.f;
          }
        }
      }

      TSTAMP(7 + 10 * sub);
      // ============================================================== P9 warm start, PGS, noslip
      // Dual state: u = G_S^-1 (base part of J^T f), replicated across the octet; wv = G_k^-1 (leg part), private.
      float u[6], wv[3];
#pragma unroll
      for (int a = 0; a < 6; a++) u[a] = 0.f;
#pragma unroll
      for (int j = 0; j < 3; j++) wv[j] = 0.f;
      float cl = 0.f;
      for (int c = 0; c < nc; c++) {
        const float f0 = cb.f[c][0], f1 = cb.f[c][1], f2 = cb.f[c][2], f3 = cb.f[c][3];
        const float c0 = (f0 + f1) + (f2 + f3), c1 = mu * (f0 - f1), c2 = mu * (f2 - f3);
#pragma unroll
        for (int a = 0; a < 6; a++) u[a] = fmaf(cb.Y[c][0][a], c0, fmaf(cb.Y[c][1][a], c1, fmaf(cb.Y[c][2][a], c2, u[a])));
#pragma unroll
        for (int j = 0; j < 3; j++) wv[j] = fmaf(cb.Z[c][0][j], c0, fmaf(cb.Z[c][1][j], c1, fmaf(cb.Z[c][2][j], c2, wv[j])));
        const float hR = 0.5f * cb.R[c];
        cl += f0 * fmaf(hR, f0, cb.b[c][0]) + f1 * fmaf(hR, f1, cb.b[c][1]) + f2 * fmaf(hR, f2, cb.b[c][2]) + f3 * fmaf(hR, f3, cb.b[c][3]);
      }
      if (__builtin_expect(npair_max > 0, 0))
      for (int pi = 0; pi < npair; pi++) {                    // pair contacts: Y and the cost terms enter once (lane li), Z on both legs
        const PairBlk& P = pblk[pi];
        const float f0 = P.f[0], f1 = P.f[1], f2 = P.f[2], f3 = P.f[3];
        const float c0 = (f0 + f1) + (f2 + f3), c1 = P.mu * (f0 - f1), c2 = P.mu * (f2 - f3);
        if (l == P.li) {
#pragma unroll
          for (int a = 0; a < 6; a++) u[a] = fmaf(P.Y[0][a], c0, fmaf(P.Y[1][a], c1, fmaf(P.Y[2][a], c2, u[a])));
          const float hR = 0.5f * P.R;
          cl += f0 * fmaf(hR, f0, P.b[0]) + f1 * fmaf(hR, f1, P.b[1]) + f2 * fmaf(hR, f2, P.b[2]) + f3 * fmaf(hR, f3, P.b[3]);
        }
        if (l == P.li || l == P.lj) {
          const float (*Zs)[3] = l == P.li ? P.Zi : P.Zj;
#pragma unroll
          for (int j = 0; j < 3; j++) wv[j] = fmaf(Zs[0][j], c0, fmaf(Zs[1][j], c1, fmaf(Zs[2][j], c2, wv[j])));
        }
      }
      float uu = 0.f;
#pragma unroll
      for (int a = 0; a < 6; a++) { u[a] = oct_sum(u[a]); uu = fmaf(u[a], u[a], uu); }
      cl += 0.5f * (wv[0] * wv[0] + wv[1] * wv[1] + wv[2] * wv[2]);
      const float cost = fmaf(0.5f, uu, oct_sum(cl));
      if (cost > 0.f) {        // f = 0 is cheaper than the warm start
        for (int c = 0; c < nc; c++)
#pragma unroll
          for (int rr = 0; rr < 4; rr++) cb.f[c][rr] = 0.f;
        if (__builtin_expect(npair_max > 0, 0))
        for (int pi = 0; pi < npair; pi++)
          if (l == pblk[pi].li) { pblk[pi].f[0] = 0.f; pblk[pi].f[1] = 0.f; pblk[pi].f[2] = 0.f; pblk[pi].f[3] = 0.f; }
#pragma unroll
        for (int a = 0; a < 6; a++) u[a] = 0.f;
#pragma unroll
        for (int j = 0; j < 3; j++) wv[j] = 0.f;
      } else dbg_warm = ncon_env > 0 ? 1 : 0;
      if (__builtin_expect(npair_max > 0, 0)) __syncwarp();

      // ---- sweeps: rows in contact order (base geom first, then legs 1..6), Gauss-Seidel through u.
      // sweep 0..iterations-1: PGS on single edges (with R); then noslip on opposing edge pairs (without R, sum fixed).
      // the 255-register instantiation keeps this lane's first contact block in registers for all sweeps; the
      // 128-register one (large batches, throughput bound) would only spill it, and re-reads it from local memory
      constexpr bool kCache0 = BLOCK * MINB <= 256;          // i.e. the 255-register budget
      ConRegs k0;
      if (kCache0) con_load(cb, 0, k0);
      const int npgs = sm.iterations, nsweep = sm.iterations + sm.noslip_iterations;
      bool active = ncon_env > 0;
      bool in_noslip = false;
#pragma unroll 1
      for (int sweep = 0; sweep <= nsweep; sweep++) {
        if (sweep == npgs) {
          // qacc after PGS -> next step's warm start (saved BEFORE noslip)
          float tb[6], tk[3], y3[3];
          bwd6(F.S, F.Si, u, tb);
          bwd3(F.g, F.gi, wv, y3);
#pragma unroll
          for (int i = 0; i < 3; i++) {
            float t = y3[i];
#pragma unroll
            for (int a = 0; a < 6; a++) t = fmaf(-F.E[i][a], tb[a], t);
            tk[i] = t;
          }
          if (ncon_env > 0) {
#pragma unroll
            for (int i = 0; i < 6; i++) wsb[i] = xsb[i] + tb[i];
#pragma unroll
            for (int i = 0; i < 3; i++) wsk[i] = xsk[i] + tk[i];
          }
          active = ncon_env > 0;
          in_noslip = true;
        }
        if (sweep == nsweep) break;
        if (!__any_sync(FULL, active)) continue;
        float improvement = 0.f;
#pragma unroll 1
        for (int slot = 0; slot < nslot; slot++) {
          // lane of this octet that owns slot `slot` (none: broadcast from lane 0, u is unchanged there)
          const int owner = (owner_tab >> (4 * slot)) & 0xf;
#ifdef NM_TIMING_VISIT
          const long long tv0 = clock64();
#endif
          if (active && nc > 0 && my_slot == slot) {
            if (kCache0) con_sweep(k0, u, wv, mu, in_noslip, improvement);   // first contact of the lane: cached in registers
            for (int c = kCache0 ? 1 : 0; c < nc; c++) {                      // further contacts (rare): through local memory
              ConRegs kc;
              con_load(cb, c, kc);
              con_sweep(kc, u, wv, mu, in_noslip, improvement);
              cb.f[c][0] = kc.f[0]; cb.f[c][1] = kc.f[1]; cb.f[c][2] = kc.f[2]; cb.f[c][3] = kc.f[3];
            }
          }
#ifdef NM_TIMING_VISIT
          const long long tv1 = clock64();
#endif
#pragma unroll
          for (int a = 0; a < 6; a++) u[a] = oct_bcast(u[a], obase | owner);
#ifdef NM_TIMING_VISIT
          if (sub == 0 && nm_timing_buf && (threadIdx.x & 31) == 0 && slot == 0 && (sweep == 0 || sweep == npgs)) {
            // cycles of the visit alone and of the whole slot (visit + broadcast), first PGS sweep -> cols 29/30, first noslip sweep -> col 31 (slot)
            const long long tv2 = clock64() + (long long)(__float_as_int(u[0]) & 0);      // (after the broadcast has delivered)
            if (sweep == 0) { nm_timing_buf[(size_t)(gtid >> 5) * 32 + 29] = tv1 - tv0; nm_timing_buf[(size_t)(gtid >> 5) * 32 + 30] = tv2 - tv0; }
            else nm_timing_buf[(size_t)(gtid >> 5) * 32 + 31] = tv2 - tv0;
          }
#endif
        }
        // pair contacts: rows after all plane contacts (MuJoCo orders contacts by body pair; the world body's pairs come first).
        // All eight lanes run the visit on identical (shared-memory) data, so u stays replicated without a broadcast; the
        // two legs' Z.w enter through one octet sum, and each of the two lanes applies its own Z to its w.
        if (__builtin_expect(npair_max > 0, 0))
        for (int pi = 0; pi < npair_max; pi++) {
          if (active && pi < npair) {
            PairBlk& P = pblk[pi];
            const bool isI = l == P.li, isJ = l == P.lj;
            ConRegs kp;
            float zl[3][3];
#pragma unroll
            for (int f = 0; f < 3; f++) {
#pragma unroll
              for (int a = 0; a < 6; a++) kp.Y[f][a] = P.Y[f][a];
#pragma unroll
              for (int j = 0; j < 3; j++) { kp.Z[f][j] = 0.f; zl[f][j] = isI ? P.Zi[f][j] : (isJ ? P.Zj[f][j] : 0.f); }
            }
#pragma unroll
            for (int i = 0; i < 10; i++) kp.G[i] = P.G[i];
#pragma unroll
            for (int i = 0; i < 4; i++) { kp.b[i] = P.b[i]; kp.adi[i] = P.adi[i]; kp.f[i] = P.f[i]; }
            kp.R = P.R; kp.ik[0] = P.ik[0]; kp.ik[1] = P.ik[1];
            float zs[3], cc[3], imp_other = 0.f, wz[3] = {0.f, 0.f, 0.f};
#pragma unroll
            for (int f = 0; f < 3; f++) zs[f] = oct_sum_m(omask, fmaf(zl[f][0], wv[0], fmaf(zl[f][1], wv[1], zl[f][2] * wv[2])));
            con_sweep(kp, u, wz, P.mu, in_noslip, isI ? improvement : imp_other, zs[0], zs[1], zs[2], cc);
#pragma unroll
            for (int j = 0; j < 3; j++) wv[j] = fmaf(zl[0][j], cc[0], fmaf(zl[1][j], cc[1], fmaf(zl[2][j], cc[2], wv[j])));
            if (isI) { P.f[0] = kp.f[0]; P.f[1] = kp.f[1]; P.f[2] = kp.f[2]; P.f[3] = kp.f[3]; }
          }
          __syncwarp();
        }
        improvement = oct_sum(improvement);
        if (active) { if (in_noslip) dbg_noslip++; else dbg_pgs++; }
        if (improvement * sm.solver_scale < (in_noslip ? sm.noslip_tolerance : sm.tolerance)) active = false;
      }
      if (kCache0 && nc > 0) { cb.f[0][0] = k0.f[0]; cb.f[0][1] = k0.f[1]; cb.f[0][2] = k0.f[2]; cb.f[0][3] = k0.f[3]; }
      // final constraint acceleration x = M^-1 J^T f
      {
        bwd6(F.S, F.Si, u, xb);
        float y3[3];
        bwd3(F.g, F.gi, wv, y3);
#pragma unroll
        for (int i = 0; i < 3; i++) {
          float t = y3[i];
#pragma unroll
          for (int a = 0; a < 6; a++) t = fmaf(-F.E[i][a], xb[a], t);
          xk[i] = t;
        }
        if (ncon_env == 0) {
#pragma unroll
          for (int i = 0; i < 6; i++) xb[i] = 0.f;
#pragma unroll
          for (int i = 0; i < 3; i++) xk[i] = 0.f;
        } else if (sm.warm_after_noslip && sm.noslip_iterations > 0) {     // model option: the warm start is the acceleration AFTER noslip
#pragma unroll
          for (int i = 0; i < 6; i++) wsb[i] = xsb[i] + xb[i];
#pragma unroll
          for (int i = 0; i < 3; i++) wsk[i] = xsk[i] + xk[i];
        }
      }
      // ============================================================== P10 touch sensors (sum of pyramid-edge forces)
      for (int c = 0; c < nc; c++) {
        float fn = cb.f[c][0] + cb.f[c][1] + cb.f[c][2] + cb.f[c][3];
        if (fn <= 0.f) continue;
        const V3 ray = mk(-pn.x, -pn.y, -pn.z);          // normal points plane -> body; sensor is on the body
        if (L.site_r[0] >= 0.f && ray_sphere(pg + mul(Xg, ld3(L.site_pos[0])), L.site_r[0], cb.pos[c], ray) >= 0.f) fn_slot0 += fn;
        if (L.site_r[1] >= 0.f && ray_sphere(pg + mul(Xg, ld3(L.site_pos[1])), L.site_r[1], cb.pos[c], ray) >= 0.f) fn_slot1 += fn;
      }
      if (__builtin_expect(npair_max > 0, 0))
      for (int pi = 0; pi < npair; pi++) {                   // pair contacts load both legs' sensors (ray along +-normal)
        const PairBlk& P = pblk[pi];
        if (l != P.li && l != P.lj) continue;
        const float fn = P.f[0] + P.f[1] + P.f[2] + P.f[3];
        if (fn <= 0.f) continue;
        const float sg = l == P.lj ? -1.f : 1.f;             // the normal points from body i to body j; flipped for the second body
        const V3 ray = mk(sg * P.nrm[0], sg * P.nrm[1], sg * P.nrm[2]), cp = ld3(P.pos);
        if (L.site_r[0] >= 0.f && ray_sphere(pg + mul(Xg, ld3(L.site_pos[0])), L.site_r[0], cp, ray) >= 0.f) fn_slot0 += fn;
        if (L.site_r[1] >= 0.f && ray_sphere(pg + mul(Xg, ld3(L.site_pos[1])), L.site_r[1], cp, ray) >= 0.f) fn_slot1 += fn;
      }
    }
    TSTAMP(8 + 10 * sub);
    PHASE_SYNC_F();
    TSTAMP(9 + 10 * sub);
    sens0 = fn_slot0; sens1 = fn_slot1;
    cvel_b = cvb;
    base_height = xip_b.z;

    // ================================================================ optional debug record (parity harness)
    if (A.debug != nullptr && valid && sub == A.nstep - 1) {
      float* dbg = A.debug + (size_t)env * NM_DBG;
      if (l == 0) {
        dbg[0] = (float)ncon_env; dbg[1] = (float)dbg_pgs; dbg[2] = (float)dbg_noslip; dbg[3] = (float)dbg_warm;
#pragma unroll
        for (int i = 0; i < 6; i++) { dbg[96 + i] = xsb[i]; dbg[128 + i] = xsb[i] + xb[i]; }
      }
      if (l < 7) {
        float* g = dbg + 8 + l * 12;
        g[0] = (float)nc;
        for (int c = 0; c < NM_MAXC; c++) { g[1 + 2 * c] = c < nc ? (float)cvert[c] : -1.f; g[2 + 2 * c] = c < nc ? cdist[c] : 0.f; }
      }
      if (leg) {
#pragma unroll
        for (int j = 0; j < 3; j++) { dbg[102 + jo + j] = xsk[j]; dbg[134 + jo + j] = xsk[j] + xk[j]; }
      }
      if (l == 0) {
        dbg[4] = (float)npair;
        for (int pi = 0; pi < npair; pi++) {
          const PairBlk& P = pblk[pi];
          float* g = dbg + 288 + 8 * pi;
          g[0] = (float)P.li; g[1] = (float)P.lj; g[2] = P.dist; g[3] = P.pos[0]; g[4] = P.pos[1]; g[5] = P.pos[2];
          g[6] = P.f[0] + P.f[1] + P.f[2] + P.f[3]; g[7] = P.nrm[2];
        }
      }
      if (l < 7 && any_contact) {                   // pyramid-edge forces of this lane's contacts (stage attribution)
        float* g = dbg + 160 + l * 16;
        for (int c = 0; c < nc; c++)
#pragma unroll
          for (int e = 0; e < 4; e++) g[4 * c + e] = cb.f[c][e];
      }
    }

    // ================================================================ P11 implicitfast velocity update + position integration
    {
      // (M + hD) qdd = M qacc  =>  qdd = qacc - (M + hD)^-1 [0; hD qacc_k]
      float qab[6], qak[3], zb[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, rkk[3], cb6[6], ck3[3];
#pragma unroll
      for (int i = 0; i < 6; i++) qab[i] = xsb[i] + xb[i];
#pragma unroll
      for (int j = 0; j < 3; j++) { qak[j] = xsk[j] + xk[j]; rkk[j] = hD[j] * qak[j]; }
      solve_system(FH, zb, rkk, cb6, ck3);
#pragma unroll
      for (int i = 0; i < 6; i++) qab[i] -= cb6[i];
#pragma unroll
      for (int j = 0; j < 3; j++) qak[j] -= ck3[j];
      vlin = fma3(h, mk(qab[0], qab[1], qab[2]), vlin);
      wloc = fma3(h, mk(qab[3], qab[4], qab[5]), wloc);
#pragma unroll
      for (int j = 0; j < 3; j++) { thd[j] = fmaf(h, qak[j], thd[j]); th[j] = fmaf(h, thd[j], th[j]); }
      p = fma3(h, vlin, p);
      // quaternion integration with the body-frame angular velocity
      float wn = sqrtf(dot(wloc, wloc));
      if (wn >= NM_MINVAL) {
        float s, c;
        sincosf(0.5f * h * wn, &s, &c);
        float k = s / wn;
        float rx = wloc.x * k, ry = wloc.y * k, rz = wloc.z * k;
        float nw = q0 * c - q1 * rx - q2 * ry - q3 * rz;
        float nx = q0 * rx + q1 * c + q2 * rz - q3 * ry;
        float ny = q0 * ry - q1 * rz + q2 * c + q3 * rx;
        float nz = q0 * rz + q1 * ry - q2 * rx + q3 * c;
        q0 = nw; q1 = nx; q2 = ny; q3 = nz;
      }
#pragma unroll
      for (int i = 0; i < 6; i++) awb[i] = wsb[i];
#pragma unroll
      for (int j = 0; j < 3; j++) awk[j] = wsk[j];
    }
    // divergence guard (≙ mj_checkPos / mj_checkVel auto-reset): non-finite or huge state resets the env
    {
      float mx = fmaxf(fmaxf(fabsf(p.x), fabsf(p.y)), fabsf(p.z));
      mx = fmaxf(mx, fmaxf(fmaxf(fabsf(vlin.x), fabsf(vlin.y)), fabsf(vlin.z)));
      mx = fmaxf(mx, fmaxf(fmaxf(fabsf(wloc.x), fabsf(wloc.y)), fabsf(wloc.z)));
#pragma unroll
      for (int j = 0; j < 3; j++) mx = fmaxf(mx, fmaxf(fabsf(th[j]), fabsf(thd[j])));
      int b = !(mx < 1e10f);                                  // catches NaN too
      b = oct_sumi(b);
      if (b) {
        bad = 1;
        p = ld3(sm.qpos0); q0 = sm.qpos0[3]; q1 = sm.qpos0[4]; q2 = sm.qpos0[5]; q3 = sm.qpos0[6];
        vlin = mk(0, 0, 0); wloc = mk(0, 0, 0);
#pragma unroll
        for (int j = 0; j < 3; j++) { th[j] = leg ? sm.qpos0[7 + jo + j] : 0.f; thd[j] = 0.f; awk[j] = 0.f; }
#pragma unroll
        for (int i = 0; i < 6; i++) awb[i] = 0.f;
      }
    }
  }  // substeps

  TSTAMP(21);
  // ==================================================================== physics-only mode: write state and leave
  if (!ENV) {
    if (valid) {
      if (G.has) A.hull_hint[(size_t)env * NM_OCT + l] = hint;
      float* qpo = A.qpos + (size_t)env * 25;
      float* qvo = A.qvel + (size_t)env * 24;
      float* qwo = A.warm + (size_t)env * 24;
      if (l == 7) {
        qpo[0] = p.x; qpo[1] = p.y; qpo[2] = p.z; qpo[3] = q0; qpo[4] = q1; qpo[5] = q2; qpo[6] = q3;
        qvo[0] = vlin.x; qvo[1] = vlin.y; qvo[2] = vlin.z; qvo[3] = wloc.x; qvo[4] = wloc.y; qvo[5] = wloc.z;
#pragma unroll
        for (int i = 0; i < 6; i++) qwo[i] = awb[i];
      }
      if (leg) {
#pragma unroll
        for (int j = 0; j < 3; j++) { qpo[7 + jo + j] = th[j]; qvo[6 + jo + j] = thd[j]; qwo[6 + jo + j] = awk[j]; }
        A.sensordata[(size_t)env * 13 + l] = sens0;
        A.sensordata[(size_t)env * 13 + 6 + l] = sens1;
      }
      if (l == 6) A.sensordata[(size_t)env * 13 + 12] = sens0;
    }
    return;
  }

  // ==================================================================== env epilogue (env.py:212-309)
  const NmDevCfg& c = scfg;
  long long ep_len = A.episode_length[env] + 1;                              // E6
  // E7: base-frame velocities / gravity with the POST-integration quaternion, cvel from the last forward pass
  M3 Rq;
  {
    float n2 = q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3;   // conj(q) rotates world -> body; mju_rotVecQuat does not normalise
    (void)n2;
    Rq = quat2mat(q0, q1, q2, q3);
  }
  const V3 blin = mulT(Rq, cvel_b.v), bang = mulT(Rq, cvel_b.w), pgrav = mulT(Rq, mk(0.f, 0.f, -9.81f));
  // E8/E9 per-leg buffers
  float dacc2 = 0.f, arate2 = 0.f, dpos2 = 0.f, dvel2 = 0.f, dposabs = 0.f;
  float tib = 0.f, foot = 0.f, bodyf = 0.f;
  if (leg) {
#pragma unroll
    for (int j = 0; j < 3; j++) {
      float da = (thd[j] - prev_dof_vel[j]) * c.inv_dt;
      dacc2 = fmaf(da, da, dacc2);
      float ar = prev_act[j] - act[j];
      arate2 = fmaf(ar, ar, arate2);
      float dp = th[j] - c.default_pos[jo + j];
      dpos2 = fmaf(dp, dp, dpos2);
      dposabs += fabsf(dp);
      dvel2 = fmaf(thd[j], thd[j], dvel2);
    }
    foot = sens1;
    tib = (foot == 0.f) ? sens0 : 0.f;                                         // env.py:232
  }
  if (l == 6) bodyf = sens0;
  const float tib_sum = oct_sum(tib), tib_max = oct_max(leg ? tib : -CUDART_INF_F);
  const float foot_max = oct_max(leg ? foot : -CUDART_INF_F);
  const float body_f = oct_sum(bodyf);
  dacc2 = oct_sum(dacc2); arate2 = oct_sum(arate2); dpos2 = oct_sum(dpos2); dvel2 = oct_sum(dvel2); dposabs = oct_sum(dposabs);
  float fcf = 0.f;
  if (leg && foot > c.max_contact_force) { float x = foot - c.max_contact_force; fcf = x * x; }
  fcf = oct_sum(fcf);

  // E10 command resampling
  float cmd[3] = {A.commands[(size_t)env * 3], A.commands[(size_t)env * 3 + 1], A.commands[(size_t)env * 3 + 2]};
  const long long genv = A.env_offset + env;
  if (c.resample_period > 0 && ep_len % c.resample_period == 0) resample_commands(c, A.seed, genv, A.step_counter, 0, cmd);
  // E11 termination
  const bool time_out = (float)ep_len > c.max_episode_length;
  bool reset = time_out | (foot_max > c.term_force) | (bad != 0);
  if (c.tibia_mode == 2) reset |= tib_max > c.tibia_max_force;
  if (c.body_mode == 2) reset |= body_f > c.body_max_force;
  {
    float nrm = sqrtf(dot(pgrav, pgrav));
    reset |= (-pgrav.z) < 0.5f * nrm;                                          // acos(-pg_z/|pg|) > 60 deg
  }
  // E13 reset
  float esum[18];
  const bool w7 = valid && l == 7;
  if (reset) {
    resample_commands(c, A.seed, genv, A.step_counter, 1, cmd);
    ep_len = 0;
    if (A.dr != nullptr && A.dr_on_reset && w7) {           // new physical parameters for the new episode
      unsigned rn[4];
      philox4x32((unsigned)A.seed, (unsigned)genv, (unsigned)A.step_counter, (unsigned)((unsigned long long)A.step_counter >> 32), 3u, (unsigned)(A.seed >> 32), rn);
      float4 d4;
      d4.x = fmaf(u01(rn[0]), A.dr_range[1] - A.dr_range[0], A.dr_range[0]);
      d4.y = fmaf(u01(rn[1]), A.dr_range[3] - A.dr_range[2], A.dr_range[2]);
      d4.z = fmaf(u01(rn[2]), A.dr_range[5] - A.dr_range[4], A.dr_range[4]);
      d4.w = 0.f;
      reinterpret_cast<float4*>(A.dr)[env] = d4;
    }
  }
  // E14 rewards (alphabetical accumulation, termination last)
  float fat_term = 0.f;
  if (c.rew_scale[RW_FEET_AIR_TIME] != 0.f) {                                   // stateful term (env.py:447-477)
    int bits = A.contact_bits[env];
    float t = 0.f;
    int nb_last = 0, nb_filt = 0;
    if (leg) {
      float air = reset ? 0.f : A.feet_air_time[(size_t)env * 6 + l];
      int last = (bits >> l) & 1, lastf = (bits >> (8 + l)) & 1;
      int contact = foot > 1.0f, filt = contact | last;
      air += c.dt;
      air = (filt == lastf) ? air : 0.f;
      float r = (air > 1.f ? air - 1.f : 0.f) + (air < 0.5f ? 0.5f - air : 0.f);
      t = r * r;
      nb_last = contact << l; nb_filt = filt << (8 + l);
      if (valid) A.feet_air_time[(size_t)env * 6 + l] = air;
    }
    fat_term = oct_sum(t);
    int nbits = oct_sumi(nb_last | nb_filt);
    if (w7) A.contact_bits[env] = nbits;
  } else if (reset && leg && valid) A.feet_air_time[(size_t)env * 6 + l] = 0.f;

  float total = 0.f;
  {
    float term[18];
    term[RW_ACTION_RATE] = arate2;
    term[RW_ANG_VEL_XY] = bang.x * bang.x + bang.y * bang.y;
    { float x = base_height - c.base_height_target; term[RW_BASE_HEIGHT] = x * x; }
    term[RW_BODY_CONTACT_FORCES] = (c.tibia_mode == 1 ? tib_sum : 0.f) + (c.body_mode == 1 ? body_f : 0.f);
    term[RW_COLLISION] = 0.f;
    term[RW_DEFAULT_POSITION] = dpos2;
    term[RW_DOF_ACC] = dacc2;
    term[RW_DOF_VEL] = dvel2;
    term[RW_FEET_AIR_TIME] = fat_term;
    term[RW_FEET_CONTACT_FORCES] = fcf;
    term[RW_FEET_STUMBLE] = 0.f;
    term[RW_LIN_VEL_Z] = blin.z * blin.z;
    term[RW_ORIENTATION] = pgrav.x * pgrav.x + pgrav.y * pgrav.y;
    term[RW_STAND_STILL] = dposabs * (sqrtf(cmd[0] * cmd[0] + cmd[1] * cmd[1]) < 0.01f ? 1.f : 0.f);
    term[RW_TERMINATION] = 0.f;
    term[RW_TORQUES] = 0.f;
    { float x = cmd[2] - bang.z; term[RW_TRACKING_ANG_VEL] = expf(-x * x * c.inv_tracking_sigma); }
    { float x = cmd[0] - blin.x, y = cmd[1] - blin.y; term[RW_TRACKING_LIN_VEL] = expf(-(x * x + y * y) * c.inv_tracking_sigma); }
#pragma unroll
    for (int k = 0; k < 18; k++) {
      float prev = reset ? 0.f : A.episode_sums[(size_t)env * 18 + k];
      float r = 0.f;
      if (k != RW_TERMINATION && c.rew_scale[k] != 0.f) { r = term[k] * c.rew_scale[k]; total += r; }
      esum[k] = prev + r;
    }
    if (c.rew_scale[RW_TERMINATION] != 0.f) {
      float r = ((reset && !time_out) ? 1.f : 0.f) * c.rew_scale[RW_TERMINATION];
      total += r;
      esum[RW_TERMINATION] += r;
    }
  }
  // episode statistics of envs that reset this step (sums BEFORE this step's rewards, env.py:363-367)
  if (w7 && reset) {
#pragma unroll
    for (int k = 0; k < 18; k++) {
      float s = A.episode_sums[(size_t)env * 18 + k];
      if (c.rew_scale[k] != 0.f) atomicAdd(A.acc_cur + k, s);
    }
    atomicAdd(A.acc_cur + 18, 1.f);
  }

  // E12 env-0 recorder (env.py:261-272): the reference appends data[0].qpos/qvel BEFORE reset_idx (:274), so on a reset
  // step the recorded row is the terminal state, not qpos0
  if (A.rec_row != nullptr && valid && env == 0) {
    float* rr = A.rec_row;
    if (leg) {
#pragma unroll
      for (int j = 0; j < 3; j++) { rr[1 + 7 + jo + j] = th[j]; rr[26 + 6 + jo + j] = thd[j]; }
    }
    if (l == 7) {
      rr[0] = reset ? 1.f : 0.f;
      rr[1] = p.x; rr[2] = p.y; rr[3] = p.z; rr[4] = q0; rr[5] = q1; rr[6] = q2; rr[7] = q3;
      rr[26] = vlin.x; rr[27] = vlin.y; rr[28] = vlin.z; rr[29] = wloc.x; rr[30] = wloc.y; rr[31] = wloc.z;
    }
  }

  // ------------------------------------------------------------------ write back
  if (valid) {
    float* qpo = A.qpos + (size_t)env * 25;
    float* qvo = A.qvel + (size_t)env * 24;
    float* qwo = A.warm + (size_t)env * 24;
    float* ob = obs_tile + (threadIdx.x >> 3) * NM_NOBS_DEV;
    const float co = c.clip_obs;
    if (G.has) A.hull_hint[(size_t)env * NM_OCT + l] = hint;
    if (leg) {
#pragma unroll
      for (int j = 0; j < 3; j++) {
        qpo[7 + jo + j] = reset ? sm.qpos0[7 + jo + j] : th[j];
        qvo[6 + jo + j] = reset ? 0.f : thd[j];
        qwo[6 + jo + j] = awk[j];                                               // warm start survives resets (quirk Q3)
        A.actions[(size_t)env * 18 + jo + j] = act[j];
        A.dof_pos[(size_t)env * 18 + jo + j] = th[j];                           // terminal pose stays in the buffer (quirk Q2)
        A.dof_vel[(size_t)env * 18 + jo + j] = thd[j];
        float o0 = (th[j] - c.default_pos[jo + j]) * c.obs_dof_pos, o1 = thd[j] * c.obs_dof_vel, o2 = act[j];
        if (c.add_noise) {
          const int k0 = 12 + jo + j, k1 = 30 + jo + j, k2 = 48 + jo + j;
          o0 = fmaf(obs_noise(A.seed, genv, A.step_counter, k0), c.noise_vec[k0], o0);
          o1 = fmaf(obs_noise(A.seed, genv, A.step_counter, k1), c.noise_vec[k1], o1);
          o2 = fmaf(obs_noise(A.seed, genv, A.step_counter, k2), c.noise_vec[k2], o2);
        }
        ob[12 + jo + j] = fminf(fmaxf(o0, -co), co);
        ob[30 + jo + j] = fminf(fmaxf(o1, -co), co);
        ob[48 + jo + j] = fminf(fmaxf(o2, -co), co);
      }
      A.sensordata[(size_t)env * 13 + l] = sens0;
      A.sensordata[(size_t)env * 13 + 6 + l] = sens1;
    }
    if (l == 6) {
      A.sensordata[(size_t)env * 13 + 12] = sens0;
      float o[12] = {blin.x * c.obs_lin_vel, blin.y * c.obs_lin_vel, blin.z * c.obs_lin_vel, bang.x * c.obs_ang_vel, bang.y * c.obs_ang_vel,
                     bang.z * c.obs_ang_vel, pgrav.x, pgrav.y, pgrav.z, cmd[0] * c.obs_lin_vel, cmd[1] * c.obs_lin_vel, cmd[2] * c.obs_ang_vel};
#pragma unroll
      for (int k = 0; k < 12; k++) {
        float v = o[k];
        if (c.add_noise) v = fmaf(obs_noise(A.seed, genv, A.step_counter, k), c.noise_vec[k], v);
        ob[k] = fminf(fmaxf(v, -co), co);
      }
    }
    if (l == 7) {
      if (reset) {
#pragma unroll
        for (int i = 0; i < 7; i++) qpo[i] = sm.qpos0[i];
#pragma unroll
        for (int i = 0; i < 6; i++) qvo[i] = 0.f;
      } else {
        qpo[0] = p.x; qpo[1] = p.y; qpo[2] = p.z; qpo[3] = q0; qpo[4] = q1; qpo[5] = q2; qpo[6] = q3;
        qvo[0] = vlin.x; qvo[1] = vlin.y; qvo[2] = vlin.z; qvo[3] = wloc.x; qvo[4] = wloc.y; qvo[5] = wloc.z;
      }
#pragma unroll
      for (int i = 0; i < 6; i++) qwo[i] = awb[i];
      A.commands[(size_t)env * 3] = cmd[0]; A.commands[(size_t)env * 3 + 1] = cmd[1]; A.commands[(size_t)env * 3 + 2] = cmd[2];
      A.episode_length[env] = ep_len;
#pragma unroll
      for (int k = 0; k < 18; k++) A.episode_sums[(size_t)env * 18 + k] = esum[k];
      A.rew[env] = total;
      A.done[env] = reset ? 1 : 0;
      A.time_outs[env] = time_out ? 1.f : 0.f;
      if (A.host_obs != nullptr) { A.host_rew[env] = total; A.host_done[env] = reset ? 1 : 0; }
    }
  }
  __syncthreads();
  {
    const int env0 = blockIdx.x * (BLOCK / NM_OCT);
    const int nrow = min(BLOCK / NM_OCT, A.num_envs - env0);
    const int nflt = nrow * NM_NOBS_DEV;
    float* dst = A.obs + (size_t)env0 * NM_NOBS_DEV;
    for (int i = threadIdx.x; i < nflt; i += BLOCK) dst[i] = obs_tile[i];
    if (A.host_obs != nullptr) {
      float* hdst = A.host_obs + (size_t)env0 * NM_NOBS_DEV;
      for (int i = threadIdx.x; i < nflt; i += BLOCK) hdst[i] = obs_tile[i];
    }
  }
  TSTAMP(22);
}

#ifdef NM_TIMING
extern "C" int nm_debug_set_timing_buffer(long long* dev_ptr) { return (int)cudaMemcpyToSymbol(nm_timing_buf, &dev_ptr, sizeof(dev_ptr)); }
#endif

// ================================================================================================ reset_idx kernel
// ≙ reset_idx (env.py:335-361) for an explicit id list: qpos<-qpos0, qvel<-0, commands resampled (phase 1),
// feet_air_time/episode_length/episode_sums zeroed, reset_buf<-1.  One thread per listed env.
__global__ void nm_reset_kernel(const NmKernelArgs A, const long long* ids, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  long long e = ids[i];
  if (e < 0 || e >= A.num_envs) return;
  const NmDevModel& m = *A.model;
  for (int k = 0; k < 25; k++) A.qpos[e * 25 + k] = m.qpos0[k];
  for (int k = 0; k < 24; k++) A.qvel[e * 24 + k] = 0.f;
  float cmd[3];
  resample_commands(*A.cfg, A.seed, A.env_offset + e, A.step_counter, 1, cmd);
  for (int k = 0; k < 3; k++) A.commands[e * 3 + k] = cmd[k];
  for (int k = 0; k < 6; k++) A.feet_air_time[e * 6 + k] = 0.f;
  for (int k = 0; k < 18; k++) A.episode_sums[e * 18 + k] = 0.f;
  A.episode_length[e] = 0;
  A.done[e] = 1;
}

// ================================================================================================ extras kernel
// Second (tiny) launch of every env step: publishes this step's episode accumulators (episode_acc), and -- only when
// at least one env reset, which is how the reference behaves (env.py:344,:363-371, quirk Q10) -- refreshes the
// mean episode sums / episode_length_s and latches time_outs.  Also clears the accumulator half the NEXT step
// will add into, so no memset is needed between steps.
__global__ void nm_finalize_kernel(const NmKernelArgs A) {
  // Launched with programmatic stream serialization (nm_launch_finalize): its CTAs are placed while the step kernel is still
  // running and wait here until that grid has completed and its writes are visible -- the launch latency of this second
  // kernel disappears behind the first.  A no-op when launched the ordinary way.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const float cnt = A.acc_cur[18];
  if (i < 19) {
    const float v = A.acc_cur[i];
    A.episode_acc[i] = v;
    A.acc_next[i] = 0.f;
    if (i < 18 && cnt > 0.f) A.ep_means[i] = v / cnt * A.cfg->inv_episode_length_s;
  }
  if ((cnt > 0.f || A.cfg->strict == 0) && i < A.num_envs) A.time_outs_latched[i] = A.time_outs[i];
}

void nm_launch_finalize(const NmKernelArgs& a, void* stream) {
  static const bool pdl = getenv("NM_NO_PDL") == nullptr;
  if (pdl) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((a.num_envs + 255) / 256); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = static_cast<cudaStream_t>(stream);
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, nm_finalize_kernel, a) == cudaSuccess) return;
    cudaGetLastError();                                  // (not supported in this context: fall through to the ordinary launch)
  }
  nm_finalize_kernel<<<(a.num_envs + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
}

static size_t step_dyn_smem(int block) { return (size_t)(block / NM_OCT) * (6 * sizeof(HullPose) + NM_MAXPAIR * sizeof(PairBlk) + 12 * sizeof(float4)); }
static size_t hull_smem_bytes(const NmKernelArgs& a) { return a.hull_smem ? 16 * (size_t)a.hull_nv + 2 * (size_t)a.hull_ne_pad + 2 * (size_t)a.hull_na_pad : 0; }
#define NM_HULL_SMEM_MAX (96 * 1024)      // hull tables larger than this stay in global memory (nm_abi.cu clears hull_smem)

void nm_launch_step(const NmKernelArgs& a, bool env_mode, void* stream) {
  const int threads = a.num_envs * NM_OCT;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  static int sms_of[64] = {0};                          // per device (the ABI makes the batch's device current before launching)
  int dev = 0;
  cudaGetDevice(&dev);
  int& sms = sms_of[dev & 63];
  if (sms == 0 && (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)) sms = 148;
  // with the hull tables staged in shared memory a CTA needs ~120 KB: one 256-thread CTA per SM for every batch size (measured
  // equal to two 128-thread CTAs at one-wave sizes without the staging)
  const bool one_wave = threads <= sms * 8 * 32 && !a.hull_smem;
  // Beyond one wave the CTAs run in rounds of one per SM.  A round of 7-warp CTAs is ~5 % shorter than a round of 8-warp CTAs
  // (56.5 vs 59.5 us per round at 131 072 envs, `gpurun_out/r02_qb22.log`), so when both shapes need the same number of rounds
  // the smaller one wins (16 384 envs: 3.46 rounds of 8 warps or 3.96 of 7: 261.2 -> 253.9 us); otherwise 8 warps per SM.
  // Same arithmetic in every instantiation (tests/test_gpu_physics.py::test_large_batch_kernel_variant: bit-identical).
  static const bool mid_ok = getenv("NM_NO_MID_BLOCK") == nullptr;
  const int rounds8 = ((threads + NM_LARGE_BLOCK - 1) / NM_LARGE_BLOCK + sms - 1) / sms;
  const int rounds7 = ((threads + NM_MID_BLOCK - 1) / NM_MID_BLOCK + sms - 1) / sms;
  // One-wave batches get one CTA per SM sized to the warps an SM has to take anyway (4096 envs on 148 SMs: 6.9 warps per SM ->
  // 147 CTAs of 7 warps instead of 256 CTAs of 4, two of them on 108 of the SMs: 87.1 -> 83.0 us, `gpurun_out/r02_qb25.log`);
  // up to 4 warps per SM the small shape stays.
  const int warps_per_sm = ((threads + 31) / 32 + sms - 1) / sms;
  const bool mid = mid_ok && (one_wave ? (warps_per_sm > 4 && warps_per_sm <= NM_MID_BLOCK / 32) : rounds7 == rounds8);
  const bool small = one_wave && warps_per_sm <= 4;
  static bool attr_set[64] = {false};                      // the large-block build needs > 48 KB of shared memory in total
  if (!attr_set[dev & 63]) {
    cudaFuncSetAttribute(nm_step_kernel<true, NM_LARGE_BLOCK, NM_LARGE_MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(step_dyn_smem(NM_LARGE_BLOCK) + NM_HULL_SMEM_MAX));
    cudaFuncSetAttribute(nm_step_kernel<false, NM_LARGE_BLOCK, NM_LARGE_MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(step_dyn_smem(NM_LARGE_BLOCK) + NM_HULL_SMEM_MAX));
    cudaFuncSetAttribute(nm_step_kernel<true, NM_MID_BLOCK, NM_MID_MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(step_dyn_smem(NM_MID_BLOCK) + NM_HULL_SMEM_MAX));
    cudaFuncSetAttribute(nm_step_kernel<false, NM_MID_BLOCK, NM_MID_MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(step_dyn_smem(NM_MID_BLOCK) + NM_HULL_SMEM_MAX));
    cudaFuncSetAttribute(nm_step_kernel<true, NM_SMALL_BLOCK, NM_SMALL_MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)step_dyn_smem(NM_SMALL_BLOCK));
    cudaFuncSetAttribute(nm_step_kernel<false, NM_SMALL_BLOCK, NM_SMALL_MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)step_dyn_smem(NM_SMALL_BLOCK));
    attr_set[dev & 63] = true;
  }
  if (small || (one_wave && !mid_ok)) {
    const int blocks = (threads + NM_SMALL_BLOCK - 1) / NM_SMALL_BLOCK;
    if (env_mode) nm_step_kernel<true, NM_SMALL_BLOCK, NM_SMALL_MINB><<<blocks, NM_SMALL_BLOCK, step_dyn_smem(NM_SMALL_BLOCK), st>>>(a);
    else nm_step_kernel<false, NM_SMALL_BLOCK, NM_SMALL_MINB><<<blocks, NM_SMALL_BLOCK, step_dyn_smem(NM_SMALL_BLOCK), st>>>(a);
  } else if (mid) {
    const int blocks = (threads + NM_MID_BLOCK - 1) / NM_MID_BLOCK;
    const size_t smem = step_dyn_smem(NM_MID_BLOCK) + hull_smem_bytes(a);
    if (env_mode) nm_step_kernel<true, NM_MID_BLOCK, NM_MID_MINB><<<blocks, NM_MID_BLOCK, smem, st>>>(a);
    else nm_step_kernel<false, NM_MID_BLOCK, NM_MID_MINB><<<blocks, NM_MID_BLOCK, smem, st>>>(a);
  } else {
    const int blocks = (threads + NM_LARGE_BLOCK - 1) / NM_LARGE_BLOCK;
    const size_t smem = step_dyn_smem(NM_LARGE_BLOCK) + hull_smem_bytes(a);
    if (env_mode) nm_step_kernel<true, NM_LARGE_BLOCK, NM_LARGE_MINB><<<blocks, NM_LARGE_BLOCK, smem, st>>>(a);
    else nm_step_kernel<false, NM_LARGE_BLOCK, NM_LARGE_MINB><<<blocks, NM_LARGE_BLOCK, smem, st>>>(a);
  }
}

void nm_launch_reset(const NmKernelArgs& a, const long long* env_ids, int n, void* stream) {
  if (n <= 0) return;
  nm_reset_kernel<<<(n + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(a, env_ids, n);
}

// ================================================================================================ FP32 pipe peak
// FFMA micro-benchmark used as the roofline denominator of the step kernel (MEASURED_PEAKS.json has no
// CUDA-core FP32 figure): 8 independent FMA chains per thread, 2 flops per FFMA.
__global__ void nm_ffma_kernel(float* out, int iters) {
  float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f, a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
  const float b = 1.0000001f, c = 1e-7f;
#pragma unroll 1
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 16; k++) {
      a0 = fmaf(a0, b, c); a1 = fmaf(a1, b, c); a2 = fmaf(a2, b, c); a3 = fmaf(a3, b, c);
      a4 = fmaf(a4, b, c); a5 = fmaf(a5, b, c); a6 = fmaf(a6, b, c); a7 = fmaf(a7, b, c);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

double nm_run_ffma_peak(void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int blocks = sms * 8, threads = 256, iters = 4096;
  float* buf = nullptr;
  if (cudaMalloc(&buf, sizeof(float) * blocks * threads) != cudaSuccess) return -1.0;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 0.0;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(e0, st);
    nm_ffma_kernel<<<blocks, threads, 0, st>>>(buf, iters);
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    double tf = 2.0 * 8 * 16 * (double)iters * blocks * threads / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(buf);
  return best;
}
