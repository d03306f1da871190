/*
 * rsb_dev.h -- the batched env.step hot path as lane-cooperative device code (fp32).
 *
 * One "group" of RSB_LANES lanes (a warp by default) owns one environment; all per-env data of the 25 physics
 * substeps of a control step live in that group's slice of shared memory (layout: DevModel::o_*).  The stages follow
 * MuJoCo's mj_step / robosuite's controllers as restated in SURVEY.md Appendix A (cited per function); the CPU
 * oracle (oracle/rsb_oracle.c) restates the same path in fp64 with different formulations.
 *
 * Differences in formulation from the oracle (deliberate, so parity is a real check):
 *   - spatial algebra about a per-tree reference point (the root body's frame origin), not the world origin;
 *   - Featherstone-style chain walks (parallel over bodies/dofs) instead of serial recursions;
 *   - Cholesky solves instead of explicit inverses / eigen pseudo-inverses in the OSC law;
 *   - Newton solver state in shared memory, reductions by lane shuffles; packed symmetric M / Hessian, column-per-lane Hessian build,
 *     register-resident fused Cholesky factor+solve, residuals carried along the iterations, register line search.
 *
 * What shaped the code (profiles/): the kernel is bound by the dependent-instruction latency of its slowest environment
 * and, before that, by instruction fetch -- so: warp-uniform control flow (constant shuffle masks), single-copy non-inlined stage
 * functions, CTA lockstep between stages (one instruction stream per SM), per-row arrays indexed only by unrolled loops (registers,
 * never local memory), reciprocal square roots instead of IEEE sqrt/division chains where the result does not feed the state directly.
 *
 * The same source compiles for the host (RSB_EMU) where lanes are fibers; that build is TEST infrastructure only.
 */
#ifndef RSB_DEV_H
#define RSB_DEV_H

#include "rsb_devmodel.h"

#ifndef RSB_LANES
#define RSB_LANES 32
#endif

typedef float real;

#ifdef RSB_EMU
/* ---- host emulation of a lane group (tests/emu): provided by the harness */
#include <math.h>
#include <stdint.h>
struct Grp { int lane; unsigned mask; };
void emu_sync();
float emu_shfl_f(float v, int src);
int emu_shfl_i(int v, int src);
#define RSB_D static inline
#define RSB_DN static
#define RSB_DNOINL static
extern DevModel emu_model;                 /* the harness sets these before running a group */
extern float *emu_smem;
#define MDL emu_model
#define RSB_SMEM emu_smem
#define RSB_CTA_SYNC(k) ((void)0)
RSB_D void gsync(Grp) { emu_sync(); }
RSB_D real gshfl(Grp, real v, int src) { return emu_shfl_f(v, src); }
RSB_D int gshfl_i(Grp, int v, int src) { return emu_shfl_i(v, src); }
RSB_D real gshfl_xor(Grp g, real v, int x) { return emu_shfl_f(v, g.lane ^ x); }
RSB_D int gshfl_xor_i(Grp g, int v, int x) { return emu_shfl_i(v, g.lane ^ x); }
RSB_D real gshfl_up(Grp g, real v, int d) { return emu_shfl_f(v, g.lane >= d ? g.lane - d : g.lane); }
RSB_D int gshfl_up_i(Grp g, int v, int d) { return emu_shfl_i(v, g.lane >= d ? g.lane - d : g.lane); }
RSB_D void rsb_sincos(real x, real *s, real *c) { *s = sinf(x); *c = cosf(x); }
RSB_D real rsb_rsqrt(real x) { return 1.0f / sqrtf(x); }
RSB_D bool wany(bool p) { return p; }                          /* the emulator runs one group at a time */
RSB_D bool sany(bool p) { return p; }
RSB_D int f2i(real f) { int i; memcpy(&i, &f, 4); return i; }
RSB_D real i2f(int i) { real f; memcpy(&f, &i, 4); return f; }
RSB_D uint32_t mulhi32(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
RSB_D void count_event(int k, unsigned int n) { if (MDL.counters) MDL.counters[k] += n; }
#else
#include <stdint.h>
struct Grp { int lane; unsigned mask; };
#define RSB_D __device__ __forceinline__
#define RSB_DN static __device__ __noinline__
#define RSB_DNOINL static __device__ __noinline__
/* the compiled model lives in constant memory (set by the host before a launch); every env's working set is a slice of
   the CTA's dynamic shared memory, addressed by WORD OFFSET so that non-inlined stage functions still emit LDS/STS */
static __constant__ DevModel c_model;
extern __shared__ float rsb_smem[];
#define MDL c_model
#define RSB_SMEM rsb_smem
#ifdef RSB_LOCKSTEP
#define RSB_CTA_SYNC(k) do { if (MDL.lockstep & (1 << (k))) __syncthreads(); } while (0)     /* keep the CTA's warps in the same stage: they share instruction fetches */
#else
#define RSB_CTA_SYNC(k) ((void)0)
#endif
/* WARP-UNIFORM CONTROL FLOW: the groups of a warp (two 16-lane groups, or one 32-lane group) always execute every shuffle / warp
   barrier together, so the member mask is the compile-time constant 0xffffffff and a shuffle is ONE instruction (a run-time mask costs
   MATCH + REDUX + VOTE + branch per shuffle).  Data-dependent loops that contain shuffles run until no group of the warp is active
   (`wany`), with the finished group's updates predicated off. */
#define RSB_FULL 0xffffffffu
RSB_D void gsync(Grp) { __syncwarp(); }
RSB_D real gshfl(Grp, real v, int src) { return __shfl_sync(RSB_FULL, v, src, RSB_LANES); }
RSB_D int gshfl_i(Grp, int v, int src) { return __shfl_sync(RSB_FULL, v, src, RSB_LANES); }
RSB_D real gshfl_xor(Grp, real v, int x) { return __shfl_xor_sync(RSB_FULL, v, x, RSB_LANES); }
RSB_D int gshfl_xor_i(Grp, int v, int x) { return __shfl_xor_sync(RSB_FULL, v, x, RSB_LANES); }
RSB_D real gshfl_up(Grp, real v, int d) { return __shfl_up_sync(RSB_FULL, v, d, RSB_LANES); }
RSB_D int gshfl_up_i(Grp, int v, int d) { return __shfl_up_sync(RSB_FULL, v, d, RSB_LANES); }
RSB_D bool wany(bool p) { return __any_sync(RSB_FULL, p) != 0; }
/* solver vote: with lockstep bit 9 the whole CTA iterates together (a finished warp would wait at the stage barrier anyway, and
   iterating together keeps the CTA on one instruction stream); otherwise a warp vote */
RSB_D bool sany(bool p) { return (c_model.lockstep & 512) ? (__syncthreads_or(p) != 0) : (__any_sync(RSB_FULL, p) != 0); }
RSB_D void rsb_sincos(real x, real *s, real *c) { sincosf(x, s, c); }
RSB_D real rsb_rsqrt(real x) { real r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
RSB_D int f2i(real f) { return __float_as_int(f); }
RSB_D real i2f(int i) { return __int_as_float(i); }
RSB_D uint32_t mulhi32(uint32_t a, uint32_t b) { return __umulhi(a, b); }
RSB_D void count_event(int k, unsigned int n) { if (c_model.counters) atomicAdd(c_model.counters + k, n); }
#endif

#define SOFF(p) ((int)((p) - RSB_SMEM))
#define RSB_MINVAL 1e-15f
#define RSB_PI 3.14159265358979323846f

/* ------------------------------------------------------------------ group reductions */
RSB_D real gsum(Grp g, real v) {
#pragma unroll
  for (int o = RSB_LANES / 2; o > 0; o >>= 1) v += gshfl_xor(g, v, o);
  return v;
}
/* two / three sums at once: the shuffles of one level are independent, so the latency is that of ONE reduction */
RSB_D void gsum2(Grp g, real &a, real &b) {
#pragma unroll
  for (int o = RSB_LANES / 2; o > 0; o >>= 1) { real ta = gshfl_xor(g, a, o), tb = gshfl_xor(g, b, o); a += ta; b += tb; }
}
RSB_D void gsum3(Grp g, real &a, real &b, real &c) {
#pragma unroll
  for (int o = RSB_LANES / 2; o > 0; o >>= 1) { real ta = gshfl_xor(g, a, o), tb = gshfl_xor(g, b, o), tc = gshfl_xor(g, c, o); a += ta; b += tb; c += tc; }
}
RSB_D real gmaxf(Grp g, real v) {
#pragma unroll
  for (int o = RSB_LANES / 2; o > 0; o >>= 1) v = fmaxf(v, gshfl_xor(g, v, o));
  return v;
}
RSB_D int gsum_i(Grp g, int v) {
#pragma unroll
  for (int o = RSB_LANES / 2; o > 0; o >>= 1) v += gshfl_xor_i(g, v, o);
  return v;
}
/* inclusive prefix sum over lanes */
RSB_D int gscan_incl(Grp g, int v) {
#pragma unroll
  for (int o = 1; o < RSB_LANES; o <<= 1) { int t = gshfl_up_i(g, v, o); if (g.lane >= o) v += t; }
  return v;
}

/* ------------------------------------------------------------------ small vector helpers */
RSB_D real dot3(const real *a, const real *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
RSB_D void cross3(real *o, const real *a, const real *b) {
  real x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  o[0] = x; o[1] = y; o[2] = z;
}
RSB_D real dot6(const real *a, const real *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3] + a[4] * b[4] + a[5] * b[5]; }
RSB_D void matvec3(real *o, const real *R, const real *v) {
  real x = R[0] * v[0] + R[1] * v[1] + R[2] * v[2], y = R[3] * v[0] + R[4] * v[1] + R[5] * v[2], z = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
  o[0] = x; o[1] = y; o[2] = z;
}
RSB_D void matTvec3(real *o, const real *R, const real *v) {
  real x = R[0] * v[0] + R[3] * v[1] + R[6] * v[2], y = R[1] * v[0] + R[4] * v[1] + R[7] * v[2], z = R[2] * v[0] + R[5] * v[1] + R[8] * v[2];
  o[0] = x; o[1] = y; o[2] = z;
}
RSB_D void matmul3(real *o, const real *A, const real *B) {
  real t[9];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) t[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
#pragma unroll
  for (int k = 0; k < 9; k++) o[k] = t[k];
}
RSB_D void quatmul(real *o, const real *a, const real *b) {
  real w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  real x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  real y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  real z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  o[0] = w; o[1] = x; o[2] = y; o[3] = z;
}
RSB_D void quatnorm(real *q) {
  real n2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
  if (n2 < 1e-30f) { q[0] = 1; q[1] = q[2] = q[3] = 0; } else { real r = rsb_rsqrt(n2); r = r * (1.5f - 0.5f * n2 * r * r); q[0] *= r; q[1] *= r; q[2] *= r; q[3] *= r; }   /* MUFU.RSQ + one Newton step */
}
RSB_D void quat2mat(real *R, const real *q) {
  real w = q[0], x = q[1], y = q[2], z = q[3];
  R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - w * z); R[2] = 2 * (x * z + w * y);
  R[3] = 2 * (x * y + w * z); R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - w * x);
  R[6] = 2 * (x * z - w * y); R[7] = 2 * (y * z + w * x); R[8] = 1 - 2 * (x * x + y * y);
}
/* v' = q v q*  without forming the matrix */
RSB_D void quatrot(real *o, const real *q, const real *v) {
  real t[3], u[3]; cross3(t, q + 1, v); t[0] *= 2; t[1] *= 2; t[2] *= 2; cross3(u, q + 1, t);
  o[0] = v[0] + q[0] * t[0] + u[0]; o[1] = v[1] + q[0] * t[1] + u[1]; o[2] = v[2] + q[0] * t[2] + u[2];
}
RSB_D real clampf(real x, real lo, real hi) { return fminf(fmaxf(x, lo), hi); }
/* dot products over shared memory with four independent accumulators and the loads of four terms issued together: the inner
   loops of this kernel are latency-bound (LDS ~29 cycles, dependent FMA 4 cycles), not throughput-bound */
RSB_D real sdot(const real *a, const real *b, int n) {
  real s0 = 0, s1 = 0, s2 = 0, s3 = 0; int k = 0;
  for (; k + 4 <= n; k += 4) { real a0 = a[k], a1 = a[k + 1], a2 = a[k + 2], a3 = a[k + 3], b0 = b[k], b1 = b[k + 1], b2 = b[k + 2], b3 = b[k + 3];
    s0 += a0 * b0; s1 += a1 * b1; s2 += a2 * b2; s3 += a3 * b3; }
  for (; k < n; k++) s0 += a[k] * b[k];
  return (s0 + s1) + (s2 + s3);
}
RSB_D real sdot_strided(const real *a, int stride, const real *b, int n) {         /* sum_k a[k*stride] * b[k] */
  real s0 = 0, s1 = 0, s2 = 0, s3 = 0; int k = 0;
  for (; k + 4 <= n; k += 4) { real a0 = a[k * stride], a1 = a[(k + 1) * stride], a2 = a[(k + 2) * stride], a3 = a[(k + 3) * stride], b0 = b[k], b1 = b[k + 1], b2 = b[k + 2], b3 = b[k + 3];
    s0 += a0 * b0; s1 += a1 * b1; s2 += a2 * b2; s3 += a3 * b3; }
  for (; k < n; k++) s0 += a[k * stride] * b[k];
  return (s0 + s1) + (s2 + s3);
}

/* M, the Newton Hessian and their factors are PACKED lower triangles: entry (i, j), j <= i, at i(i+1)/2 + j */
RSB_D int tri_off(int i) { return (i * (i + 1)) >> 1; }
RSB_D real msym(const real *P, int i, int j) { return i >= j ? P[tri_off(i) + j] : P[tri_off(j) + i]; }
/* row i of (symmetric packed P) times v: one uniform loop over j, the index switches from row i to column i at the diagonal */
RSB_D real symv_row(const real *P, int i, const real *v, int n) {
  real s0 = 0, s1 = 0; const int ri = tri_off(i); int tj = 0, j = 0;                      /* tj = tri_off(j) */
  for (; j + 2 <= n; j += 2) {
    const int i0 = (j <= i) ? ri + j : tj + i, t1 = tj + j + 1, i1 = (j + 1 <= i) ? ri + j + 1 : t1 + i;
    s0 += P[i0] * v[j]; s1 += P[i1] * v[j + 1]; tj = t1 + j + 2;
  }
  if (j < n) { const int i0 = (j <= i) ? ri + j : tj + i; s0 += P[i0] * v[j]; }
  return s0 + s1;
}

/* ------------------------------------------------------------------ Philox4x32-10 (same counters/keys as the oracle) */
RSB_D void philox4x32(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    uint32_t h0 = mulhi32(0xD2511F53u, c[0]), l0 = 0xD2511F53u * c[0], h1 = mulhi32(0xCD9E8D57u, c[2]), l1 = 0xCD9E8D57u * c[2];
    uint32_t n0 = h1 ^ c[1] ^ k0, n1 = l1, n2 = h0 ^ c[3] ^ k1, n3 = l0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3; k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}
RSB_D void rsb_philox(uint64_t seed, uint64_t env_id, uint32_t stream, uint32_t index, uint32_t out[4]) {
  out[0] = (uint32_t)env_id; out[1] = (uint32_t)(env_id >> 32); out[2] = stream; out[3] = index;
  philox4x32(out, (uint32_t)seed, (uint32_t)(seed >> 32));
}
/* fp32 Box-Muller on the oracle's uniforms needs the full 32 bits for u1 near 0; use double for the two transcendental
   arguments only (8 values per reset, 8 per synthetic action -- negligible) */
RSB_D void box_muller(uint32_t a, uint32_t b, real *z0, real *z1) {
  double u1 = ((double)a + 0.5) * (1.0 / 4294967296.0), u2 = ((double)b + 0.5) * (1.0 / 4294967296.0);
  double rad = sqrt(-2.0 * log(u1)), ang = 2.0 * 3.14159265358979323846 * u2;
  *z0 = (real)(rad * cos(ang)); *z1 = (real)(rad * sin(ang));
}

/* ================================================================== A.3.1 kinematics */
RSB_DN void st_kinematics(int so, Grp g) { real *s = RSB_SMEM + so;
  real *qpos = s + MDL.o_qpos, *lt = s + MDL.o_jq;               /* lt: per-body LOCAL transform (pos3, quat4), 7 words per body */
  real *xpos = s + MDL.o_xpos, *xquat = s + MDL.o_xquat, *xmat = s + MDL.o_xmat, *xanchor = s + MDL.o_xanchor, *xaxis = s + MDL.o_xaxis;
  /* phase A, lane per body: fold the body's joints into its parent-relative transform (all the trigonometry happens here) */
  for (int b = g.lane; b < MDL.nbody; b += RSB_LANES) {
    const int jn = MDL.body_jntnum[b], ja = MDL.body_jntadr[b];
    real p[3], q[4];
    if (jn == 1 && MDL.jnt_type[ja] == RSB_JNT_FREE) {           /* free body: world pose straight from qpos (normalised in place) */
      const int a = MDL.jnt_qadr[ja]; quatnorm(qpos + a + 3);
      p[0] = qpos[a]; p[1] = qpos[a + 1]; p[2] = qpos[a + 2]; q[0] = qpos[a + 3]; q[1] = qpos[a + 4]; q[2] = qpos[a + 5]; q[3] = qpos[a + 6];
    } else {
      const real *bp = (b == MDL.override_body) ? s + MDL.o_bpose : MDL.body_pos + 3 * b;
      const real *bq = (b == MDL.override_body) ? s + MDL.o_bpose + 3 : MDL.body_quat + 4 * b;
      p[0] = bp[0]; p[1] = bp[1]; p[2] = bp[2]; q[0] = bq[0]; q[1] = bq[1]; q[2] = bq[2]; q[3] = bq[3];
      for (int k = 0; k < jn; k++) {
        const int j = ja + k; const real dq = qpos[MDL.jnt_qadr[j]] - MDL.qpos0[MDL.jnt_qadr[j]]; real t3[3];
        if (MDL.jnt_type[j] == RSB_JNT_SLIDE) { quatrot(t3, q, MDL.jnt_axis + 3 * j); p[0] += t3[0] * dq; p[1] += t3[1] * dq; p[2] += t3[2] * dq; }
        else {                                                     /* hinge: rotate about the anchor, which stays fixed */
          real sn, cs; rsb_sincos(0.5f * dq, &sn, &cs);
          real jq4[4] = {cs, sn * MDL.jnt_axis[3 * j], sn * MDL.jnt_axis[3 * j + 1], sn * MDL.jnt_axis[3 * j + 2]}, qn[4], a3[3];
          quatrot(a3, q, MDL.jnt_pos + 3 * j); quatmul(qn, q, jq4); quatrot(t3, qn, MDL.jnt_pos + 3 * j);
          p[0] += a3[0] - t3[0]; p[1] += a3[1] - t3[1]; p[2] += a3[2] - t3[2]; q[0] = qn[0]; q[1] = qn[1]; q[2] = qn[2]; q[3] = qn[3];
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 3; k++) lt[7 * b + k] = p[k];
#pragma unroll
    for (int k = 0; k < 4; k++) lt[7 * b + 3 + k] = q[k];
  }
  gsync(g);
  /* phase B: compose down the tree (inherently sequential; one lane, ~45 instructions per body) */
  if (g.lane == 0) {
    xpos[0] = xpos[1] = xpos[2] = 0; xquat[0] = 1; xquat[1] = xquat[2] = xquat[3] = 0;
    for (int b = 1; b < MDL.nbody; b++) {
      const int p = MDL.body_parent[b], jn = MDL.body_jntnum[b]; real pos[3], quat[4];
      if (jn == 1 && MDL.jnt_type[MDL.body_jntadr[b]] == RSB_JNT_FREE) {
#pragma unroll
        for (int k = 0; k < 3; k++) pos[k] = lt[7 * b + k];
#pragma unroll
        for (int k = 0; k < 4; k++) quat[k] = lt[7 * b + 3 + k];
      } else {
        real t3[3]; quatrot(t3, xquat + 4 * p, lt + 7 * b);
        pos[0] = xpos[3 * p] + t3[0]; pos[1] = xpos[3 * p + 1] + t3[1]; pos[2] = xpos[3 * p + 2] + t3[2];
        quatmul(quat, xquat + 4 * p, lt + 7 * b + 3); quatnorm(quat);
      }
      xpos[3 * b] = pos[0]; xpos[3 * b + 1] = pos[1]; xpos[3 * b + 2] = pos[2];
      xquat[4 * b] = quat[0]; xquat[4 * b + 1] = quat[1]; xquat[4 * b + 2] = quat[2]; xquat[4 * b + 3] = quat[3];
    }
  }
  gsync(g);
  /* phase C, lane per body: rotation matrices; lane per joint: world anchors and axes (frame of the body BEFORE the joint moves it) */
  for (int b = g.lane; b < MDL.nbody; b += RSB_LANES) { real R[9]; quat2mat(R, xquat + 4 * b);
#pragma unroll
    for (int k = 0; k < 9; k++) xmat[9 * b + k] = R[k]; }
  for (int j = g.lane; j < MDL.njnt; j += RSB_LANES) {
    const int b = MDL.jnt_body[j], p = MDL.body_parent[b], ja = MDL.body_jntadr[b]; real anchor[3], axis[3];
    if (MDL.jnt_type[j] == RSB_JNT_FREE) { anchor[0] = xpos[3 * b]; anchor[1] = xpos[3 * b + 1]; anchor[2] = xpos[3 * b + 2]; axis[0] = 0; axis[1] = 0; axis[2] = 1; }
    else {
      const real *bp = (b == MDL.override_body) ? s + MDL.o_bpose : MDL.body_pos + 3 * b;
      const real *bq = (b == MDL.override_body) ? s + MDL.o_bpose + 3 : MDL.body_quat + 4 * b;
      real pl[3] = {bp[0], bp[1], bp[2]}, ql[4] = {bq[0], bq[1], bq[2], bq[3]}, t3[3];
      for (int jj = ja; jj < j; jj++) {                           /* earlier joints of the same body (rare): replay their local motion */
        const real dq = qpos[MDL.jnt_qadr[jj]] - MDL.qpos0[MDL.jnt_qadr[jj]];
        if (MDL.jnt_type[jj] == RSB_JNT_SLIDE) { quatrot(t3, ql, MDL.jnt_axis + 3 * jj); pl[0] += t3[0] * dq; pl[1] += t3[1] * dq; pl[2] += t3[2] * dq; }
        else { real sn, cs; rsb_sincos(0.5f * dq, &sn, &cs);
          real jq4[4] = {cs, sn * MDL.jnt_axis[3 * jj], sn * MDL.jnt_axis[3 * jj + 1], sn * MDL.jnt_axis[3 * jj + 2]}, qn[4], a3[3];
          quatrot(a3, ql, MDL.jnt_pos + 3 * jj); quatmul(qn, ql, jq4); quatrot(t3, qn, MDL.jnt_pos + 3 * jj);
          pl[0] += a3[0] - t3[0]; pl[1] += a3[1] - t3[1]; pl[2] += a3[2] - t3[2]; ql[0] = qn[0]; ql[1] = qn[1]; ql[2] = qn[2]; ql[3] = qn[3]; }
      }
      real al[3], xl[3]; quatrot(t3, ql, MDL.jnt_pos + 3 * j); al[0] = pl[0] + t3[0]; al[1] = pl[1] + t3[1]; al[2] = pl[2] + t3[2];
      quatrot(xl, ql, MDL.jnt_axis + 3 * j);
      quatrot(t3, xquat + 4 * p, al); anchor[0] = xpos[3 * p] + t3[0]; anchor[1] = xpos[3 * p + 1] + t3[1]; anchor[2] = xpos[3 * p + 2] + t3[2];
      quatrot(axis, xquat + 4 * p, xl);
    }
    xanchor[3 * j] = anchor[0]; xanchor[3 * j + 1] = anchor[1]; xanchor[3 * j + 2] = anchor[2];
    xaxis[3 * j] = axis[0]; xaxis[3 * j + 1] = axis[1]; xaxis[3 * j + 2] = axis[2];
  }
  gsync(g);
  /* geoms and sites, lane-parallel */
  real *gxpos = s + MDL.o_gxpos, *gxmat = s + MDL.o_gxmat, *sxpos = s + MDL.o_sxpos, *sxmat = s + MDL.o_sxmat;
  for (int i = g.lane; i < MDL.ngeom + MDL.nsite; i += RSB_LANES) {
    bool isg = i < MDL.ngeom; int k = isg ? i : i - MDL.ngeom; int b = isg ? MDL.geom_body[k] : MDL.site_body[k];
    const real *lp = isg ? MDL.geom_pos + 3 * k : MDL.site_pos + 3 * k, *lm = isg ? MDL.geom_mat + 9 * k : MDL.site_mat + 9 * k;
    real t3[3], R[9]; matvec3(t3, xmat + 9 * b, lp); matmul3(R, xmat + 9 * b, lm);
    real *op = isg ? gxpos + 3 * k : sxpos + 3 * k, *om = isg ? gxmat + 9 * k : sxmat + 9 * k;
    op[0] = xpos[3 * b] + t3[0]; op[1] = xpos[3 * b + 1] + t3[1]; op[2] = xpos[3 * b + 2] + t3[2];
#pragma unroll
    for (int q = 0; q < 9; q++) om[q] = R[q];
  }
}

/* ================================================================== A.3.2 inertias + motion subspaces about the tree reference point */
/* spatial inertia record: [Ixx Iyy Izz Ixy Ixz Iyz hx hy hz m], h = m c, about the reference point */
RSB_D void inert_mul(real *f, const real *I, const real *v) {
  const real *w = v, *l = v + 3; const real *h = I + 6; real t[3];
  f[0] = I[0] * w[0] + I[3] * w[1] + I[4] * w[2]; f[1] = I[3] * w[0] + I[1] * w[1] + I[5] * w[2]; f[2] = I[4] * w[0] + I[5] * w[1] + I[2] * w[2];
  cross3(t, h, l); f[0] += t[0]; f[1] += t[1]; f[2] += t[2];
  cross3(t, h, w); f[3] = I[9] * l[0] - t[0]; f[4] = I[9] * l[1] - t[1]; f[5] = I[9] * l[2] - t[2];
}

RSB_DN void st_inertia(int so, Grp g) { real *s = RSB_SMEM + so;
  const real *xpos = s + MDL.o_xpos, *xmat = s + MDL.o_xmat, *xanchor = s + MDL.o_xanchor, *xaxis = s + MDL.o_xaxis;
  real *cinert = s + MDL.o_cinert, *crb = s + MDL.o_crb, *cdof = s + MDL.o_cdof, *M = s + MDL.o_M;
  for (int b = g.lane; b < MDL.nbody; b += RSB_LANES) {
    real I[10];
    real ms = MDL.body_mass[b];
    if (b == 0 || ms <= 0) { for (int k = 0; k < 10; k++) I[k] = 0; }
    else {
      int r = MDL.body_root[b]; real Rw[9], c[3], t3[3];
      matmul3(Rw, xmat + 9 * b, MDL.body_imat + 9 * b); matvec3(t3, xmat + 9 * b, MDL.body_ipos + 3 * b);
      c[0] = xpos[3 * b] + t3[0] - xpos[3 * r]; c[1] = xpos[3 * b + 1] + t3[1] - xpos[3 * r + 1]; c[2] = xpos[3 * b + 2] + t3[2] - xpos[3 * r + 2];
      real d0 = MDL.body_inertia[3 * b], d1 = MDL.body_inertia[3 * b + 1], d2 = MDL.body_inertia[3 * b + 2], cc = dot3(c, c);
      I[0] = Rw[0] * Rw[0] * d0 + Rw[1] * Rw[1] * d1 + Rw[2] * Rw[2] * d2 + ms * (cc - c[0] * c[0]);
      I[1] = Rw[3] * Rw[3] * d0 + Rw[4] * Rw[4] * d1 + Rw[5] * Rw[5] * d2 + ms * (cc - c[1] * c[1]);
      I[2] = Rw[6] * Rw[6] * d0 + Rw[7] * Rw[7] * d1 + Rw[8] * Rw[8] * d2 + ms * (cc - c[2] * c[2]);
      I[3] = Rw[0] * Rw[3] * d0 + Rw[1] * Rw[4] * d1 + Rw[2] * Rw[5] * d2 - ms * c[0] * c[1];
      I[4] = Rw[0] * Rw[6] * d0 + Rw[1] * Rw[7] * d1 + Rw[2] * Rw[8] * d2 - ms * c[0] * c[2];
      I[5] = Rw[3] * Rw[6] * d0 + Rw[4] * Rw[7] * d1 + Rw[5] * Rw[8] * d2 - ms * c[1] * c[2];
      I[6] = ms * c[0]; I[7] = ms * c[1]; I[8] = ms * c[2]; I[9] = ms;
    }
#pragma unroll
    for (int k = 0; k < 10; k++) { cinert[10 * b + k] = I[k]; crb[10 * b + k] = I[k]; }
  }
  for (int d = g.lane; d < MDL.nv; d += RSB_LANES) {
    int kind = MDL.dof_kind[d], j = MDL.dof_jnt[d], b = MDL.dof_body[d], r = MDL.dof_root[d]; real w[3] = {0, 0, 0}, v[3] = {0, 0, 0};
    if (kind == RSB_DOF_SLIDE) { v[0] = xaxis[3 * j]; v[1] = xaxis[3 * j + 1]; v[2] = xaxis[3 * j + 2]; }
    else if (kind == RSB_DOF_FREE_T) { v[d - MDL.jnt_dadr[j]] = 1; }
    else {
      real off[3];
      if (kind == RSB_DOF_HINGE) {
        w[0] = xaxis[3 * j]; w[1] = xaxis[3 * j + 1]; w[2] = xaxis[3 * j + 2];
        off[0] = xpos[3 * r] - xanchor[3 * j]; off[1] = xpos[3 * r + 1] - xanchor[3 * j + 1]; off[2] = xpos[3 * r + 2] - xanchor[3 * j + 2];
      } else {                                      /* free rotation: body-local axis k expressed in the world */
        int k = d - MDL.jnt_dadr[j] - 3; w[0] = xmat[9 * b + k]; w[1] = xmat[9 * b + 3 + k]; w[2] = xmat[9 * b + 6 + k];
        off[0] = xpos[3 * r] - xpos[3 * b]; off[1] = xpos[3 * r + 1] - xpos[3 * b + 1]; off[2] = xpos[3 * r + 2] - xpos[3 * b + 2];
      }
      cross3(v, w, off);                            /* velocity of the reference point under unit rotation */
    }
    cdof[6 * d] = w[0]; cdof[6 * d + 1] = w[1]; cdof[6 * d + 2] = w[2]; cdof[6 * d + 3] = v[0]; cdof[6 * d + 4] = v[1]; cdof[6 * d + 5] = v[2];
  }
  for (int i = g.lane; i < MDL.ntri; i += RSB_LANES) M[i] = 0;
  gsync(g);
}

/* ================================================================== A.3.3 composite rigid bodies -> dense symmetric M (+ armature) */
RSB_DN void st_crb(int so, Grp g) { real *s = RSB_SMEM + so;
  real *crb = s + MDL.o_crb, *cdof = s + MDL.o_cdof, *fi = s + MDL.o_fi, *M = s + MDL.o_M;
  if (g.lane < 10) for (int b = MDL.nbody - 1; b > 0; b--) crb[10 * MDL.body_parent[b] + g.lane] += crb[10 * b + g.lane];
  gsync(g);
  for (int d = g.lane; d < MDL.nv; d += RSB_LANES) { real f[6]; inert_mul(f, crb + 10 * MDL.dof_body[d], cdof + 6 * d);
#pragma unroll
    for (int k = 0; k < 6; k++) fi[6 * d + k] = f[k]; }
  gsync(g);
  for (int p = g.lane; p < MDL.nmpair; p += RSB_LANES) {
    int i = MDL.mpair_i[p], j = MDL.mpair_j[p]; real v = dot6(cdof + 6 * j, fi + 6 * i);
    if (i == j) v += MDL.dof_armature[i];
    M[tri_off(i) + j] = v;                         /* mpair: j is i or one of its ancestors, j <= i */
  }
  gsync(g);
}

/* ================================================================== A.3.7 RNE(acc = 0) with gravity -> qfrc_bias; passive forces */
RSB_D void crossm(real *o, const real *v, const real *c) {          /* motion cross: v x c */
  real a[3], b[3], d[3]; cross3(a, v, c); cross3(b, v, c + 3); cross3(d, v + 3, c);
  o[0] = a[0]; o[1] = a[1]; o[2] = a[2]; o[3] = b[0] + d[0]; o[4] = b[1] + d[1]; o[5] = b[2] + d[2];
}
RSB_D void crossf(real *o, const real *v, const real *f) {          /* force cross: v x* f */
  real a[3], b[3], d[3]; cross3(a, v, f); cross3(b, v + 3, f + 3); cross3(d, v, f + 3);
  o[0] = a[0] + b[0]; o[1] = a[1] + b[1]; o[2] = a[2] + b[2]; o[3] = d[0]; o[4] = d[1]; o[5] = d[2];
}

RSB_DN void st_bias(int so, Grp g) { real *s = RSB_SMEM + so;
  const real *qvel = s + MDL.o_qvel, *qpos = s + MDL.o_qpos, *cdof = s + MDL.o_cdof, *cinert = s + MDL.o_cinert;
  real *cvel = s + MDL.o_cvel, *cacc = s + MDL.o_cacc, *cdd = s + MDL.o_cdofdot, *bias = s + MDL.o_bias, *passive = s + MDL.o_passive;
  /* body velocities: sum over the dof chain (all dofs of a tree share the reference point) */
  for (int b = g.lane; b < MDL.nbody; b += RSB_LANES) {
    real v[6] = {0, 0, 0, 0, 0, 0}; const unsigned mask = (unsigned)MDL.body_dofmask[b];
#pragma unroll 4
    for (int d = 0; d < MDL.nv; d++) { real q = ((mask >> d) & 1u) ? qvel[d] : 0.0f;      /* predicated, no pointer chasing up the chain */
#pragma unroll
      for (int k = 0; k < 6; k++) v[k] += cdof[6 * d + k] * q; }
#pragma unroll
    for (int k = 0; k < 6; k++) cvel[6 * b + k] = v[k];
  }
  /* cdof_dot = (velocity accumulated before this dof) x cdof */
  for (int d = g.lane; d < MDL.nv; d += RSB_LANES) {
    real o[6] = {0, 0, 0, 0, 0, 0}; int st = MDL.dof_velstart[d];
    if (st != -2) {
      real v[6] = {0, 0, 0, 0, 0, 0}; const unsigned mask = (unsigned)MDL.dof_velmask[d];
#pragma unroll 4
      for (int e = 0; e < MDL.nv; e++) { real q = ((mask >> e) & 1u) ? qvel[e] : 0.0f;
#pragma unroll
        for (int k = 0; k < 6; k++) v[k] += cdof[6 * e + k] * q; }
      crossm(o, v, cdof + 6 * d);
    }
#pragma unroll
    for (int k = 0; k < 6; k++) cdd[6 * d + k] = o[k];
  }
  gsync(g);
  /* body accelerations (acc = 0, gravity as base acceleration) and body forces */
  for (int b = g.lane; b < MDL.nbody; b += RSB_LANES) {
    real a[6] = {0, 0, 0, -MDL.gravity[0], -MDL.gravity[1], -MDL.gravity[2]}; const unsigned mask = (unsigned)MDL.body_dofmask[b];
#pragma unroll 4
    for (int d = 0; d < MDL.nv; d++) { real q = ((mask >> d) & 1u) ? qvel[d] : 0.0f;
#pragma unroll
      for (int k = 0; k < 6; k++) a[k] += cdd[6 * d + k] * q; }
    real f[6] = {0, 0, 0, 0, 0, 0};
    if (b > 0) { real Ia[6], Iv[6], t[6]; inert_mul(Ia, cinert + 10 * b, a); inert_mul(Iv, cinert + 10 * b, cvel + 6 * b); crossf(t, cvel + 6 * b, Iv);
#pragma unroll
      for (int k = 0; k < 6; k++) f[k] = Ia[k] + t[k]; }
#pragma unroll
    for (int k = 0; k < 6; k++) cacc[6 * b + k] = f[k];          /* cacc now holds cfrc_body */
  }
  gsync(g);
  if (g.lane < 6) for (int b = MDL.nbody - 1; b > 0; b--) cacc[6 * MDL.body_parent[b] + g.lane] += cacc[6 * b + g.lane];
  gsync(g);
  for (int d = g.lane; d < MDL.nv; d += RSB_LANES) {
    bias[d] = dot6(cdof + 6 * d, cacc + 6 * MDL.dof_body[d]);
    real p = -MDL.dof_damping[d] * qvel[d]; int j = MDL.dof_jnt[d];
    if (MDL.dof_kind[d] <= RSB_DOF_SLIDE && MDL.jnt_stiffness[j] != 0) p -= MDL.jnt_stiffness[j] * (qpos[MDL.jnt_qadr[j]] - MDL.qpos_spring[MDL.jnt_qadr[j]]);
    passive[d] = p;
  }
  gsync(g);
}

/* ================================================================== A.3.5 collision */
struct RawCon { real pos[3], normal[3], dist; };

RSB_D void make_frame(real *frame) {              /* frame[0..2] = normal; fills the two tangents (mju_makeFrame) */
  real *x = frame, *y = frame + 3, *z = frame + 6;
  if (x[1] < 0.5f && x[1] > -0.5f) { y[0] = 0; y[1] = 1; y[2] = 0; } else { y[0] = 0; y[1] = 0; y[2] = 1; }
  real d = dot3(x, y); y[0] -= x[0] * d; y[1] -= x[1] * d; y[2] -= x[2] * d;
  real n2 = dot3(y, y); if (n2 < RSB_MINVAL * RSB_MINVAL) { y[0] = 1; y[1] = 0; y[2] = 0; } else { real r = rsb_rsqrt(n2); r = r * (1.5f - 0.5f * n2 * r * r); y[0] *= r; y[1] *= r; y[2] *= r; }
  cross3(z, x, y);
}

RSB_D int col_plane_sphere(const real *ppos, const real *pmat, const real *spos, real r, real margin, RawCon *out) {
  real n[3] = {pmat[2], pmat[5], pmat[8]}, dif[3] = {spos[0] - ppos[0], spos[1] - ppos[1], spos[2] - ppos[2]};
  real d = dot3(dif, n) - r; if (d > margin) return 0;
  out->dist = d;
  for (int k = 0; k < 3; k++) { out->normal[k] = n[k]; out->pos[k] = spos[k] - n[k] * (r + 0.5f * d); }
  return 1;
}
RSB_D int col_plane_box(const real *ppos, const real *pmat, const real *bpos, const real *bmat, const real *size, real margin, RawCon *out) {
  real n[3] = {pmat[2], pmat[5], pmat[8]}; int cnt = 0;
  real dif[3] = {bpos[0] - ppos[0], bpos[1] - ppos[1], bpos[2] - ppos[2]}; real d0 = dot3(dif, n);
  for (int i = 0; i < 8 && cnt < 4; i++) {
    real loc[3] = {(i & 1 ? size[0] : -size[0]), (i & 2 ? size[1] : -size[1]), (i & 4 ? size[2] : -size[2])}, w[3];
    matvec3(w, bmat, loc);
    real ld = d0 + dot3(w, n);
    if (ld > margin) continue;
    RawCon *c = &out[cnt++]; c->dist = ld;
    for (int k = 0; k < 3; k++) { c->normal[k] = n[k]; c->pos[k] = bpos[k] + w[k] - n[k] * ld * 0.5f; }
  }
  return cnt;
}
RSB_D int col_plane_capsule(const real *ppos, const real *pmat, const real *cpos, const real *cmat, const real *size, real margin, RawCon *out) {
  real ax[3] = {cmat[2], cmat[5], cmat[8]}; int cnt = 0;
  for (int sg = -1; sg <= 1; sg += 2) { real p[3] = {cpos[0] + ax[0] * sg * size[1], cpos[1] + ax[1] * sg * size[1], cpos[2] + ax[2] * sg * size[1]}; cnt += col_plane_sphere(ppos, pmat, p, size[0], margin, out + cnt); }
  return cnt;
}
RSB_D int col_sphere_sphere(const real *p1, real r1, const real *p2, real r2, real margin, RawCon *out) {
  real d[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]}; real len = sqrtf(dot3(d, d)), dist = len - r1 - r2;
  if (dist > margin) return 0;
  if (len < RSB_MINVAL) { d[0] = 1; d[1] = 0; d[2] = 0; } else { d[0] /= len; d[1] /= len; d[2] /= len; }
  out->dist = dist;
  for (int k = 0; k < 3; k++) { out->normal[k] = d[k]; out->pos[k] = p1[k] + d[k] * (r1 + 0.5f * dist); }
  return 1;
}
RSB_D int col_capsule_capsule(const real *p1, const real *m1, const real *s1, const real *p2, const real *m2, const real *s2, real margin, RawCon *out) {
  real a1[3] = {m1[2], m1[5], m1[8]}, a2[3] = {m2[2], m2[5], m2[8]}, d[3] = {p1[0] - p2[0], p1[1] - p2[1], p1[2] - p2[2]};
  real b = dot3(a1, a2), c1 = dot3(a1, d), c2 = dot3(a2, d), den = 1 - b * b, t1, t2;
  if (den < 1e-12f) { t1 = 0; t2 = c2; } else { t1 = (b * c2 - c1) / den; t2 = (c2 - b * c1) / den; }
  t1 = clampf(t1, -s1[1], s1[1]);
  t2 = clampf(c2 + b * t1, -s2[1], s2[1]);
  t1 = clampf(-c1 + b * t2, -s1[1], s1[1]);
  real q1[3] = {p1[0] + a1[0] * t1, p1[1] + a1[1] * t1, p1[2] + a1[2] * t1}, q2[3] = {p2[0] + a2[0] * t2, p2[1] + a2[1] * t2, p2[2] + a2[2] * t2};
  return col_sphere_sphere(q1, s1[0], q2, s2[0], margin, out);
}
RSB_D int col_sphere_box(const real *sp, real r, const real *bp, const real *bm, const real *size, real margin, RawCon *out) {
  real d[3] = {sp[0] - bp[0], sp[1] - bp[1], sp[2] - bp[2]}, loc[3], cl[3]; matTvec3(loc, bm, d); int inside = 1;
  for (int k = 0; k < 3; k++) { cl[k] = loc[k]; if (cl[k] > size[k]) { cl[k] = size[k]; inside = 0; } if (cl[k] < -size[k]) { cl[k] = -size[k]; inside = 0; } }
  real nl[3], dist;
  if (!inside) { nl[0] = loc[0] - cl[0]; nl[1] = loc[1] - cl[1]; nl[2] = loc[2] - cl[2]; real len = sqrtf(dot3(nl, nl));
    if (len < RSB_MINVAL) { nl[0] = 1; nl[1] = 0; nl[2] = 0; } else { nl[0] /= len; nl[1] /= len; nl[2] /= len; } dist = len - r; }
  else {
    int bk = 0; real best = 1e30f;
    for (int k = 0; k < 3; k++) { real f = size[k] - fabsf(loc[k]); if (f < best) { best = f; bk = k; } }
    nl[0] = nl[1] = nl[2] = 0; nl[bk] = loc[bk] >= 0 ? 1.0f : -1.0f; cl[bk] = nl[bk] * size[bk]; dist = -best - r;
  }
  if (dist > margin) return 0;
  real nw[3], cw[3]; matvec3(nw, bm, nl); matvec3(cw, bm, cl);
  out->dist = dist;
  for (int k = 0; k < 3; k++) { out->normal[k] = -nw[k]; out->pos[k] = bp[k] + cw[k] + nw[k] * 0.5f * dist; }
  return 1;
}
RSB_D int col_capsule_box(const real *cp, const real *cm, const real *cs, const real *bp, const real *bm, const real *size, real margin, RawCon *out) {
  real ax[3] = {cm[2], cm[5], cm[8]}; int cnt = 0;
  for (int sg = -1; sg <= 1; sg++) { real p[3] = {cp[0] + ax[0] * sg * cs[1], cp[1] + ax[1] * sg * cs[1], cp[2] + ax[2] * sg * cs[1]}; cnt += col_sphere_box(p, cs[0], bp, bm, size, margin, out + cnt); }
  return cnt;
}

/* box-box: separating-axis test over 15 axes, then reference-face clipping (face case, <= 8 contacts) or the closest
   points of the two supporting edges (edge case, 1 contact).  Same contact generation rules as DESIGN.md "box-box". */
/* `scr`: 48 words of scratch for the clipped polygon (two buffers of 8 vertices).  The caller passes a slice of shared memory when it
   has one (a generic pointer): thread-local arrays with run-time indices live in local memory, which at 448 threads x 230 KB of shared
   memory per SM no longer fits L1 and costs an L2 round trip per access (60% of this function's stall samples, profiles/). */
RSB_DNOINL int col_box_box(const real *pa, const real *Ra, const real *ha, const real *pb, const real *Rb, const real *hb, real margin, RawCon *out, real *scr) {
  real R[9], Q[9], t[3], d[3] = {pb[0] - pa[0], pb[1] - pa[1], pb[2] - pa[2]};
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) { real v = Ra[i] * Rb[j] + Ra[3 + i] * Rb[3 + j] + Ra[6 + i] * Rb[6 + j]; R[3 * i + j] = v; Q[3 * i + j] = fabsf(v) + 1e-9f; }
  matTvec3(t, Ra, d);
  real best = -1e30f; int code = -1; real bsign = 1;
#pragma unroll
  for (int i = 0; i < 3; i++) {
    real sp = fabsf(t[i]) - (ha[i] + hb[0] * Q[3 * i] + hb[1] * Q[3 * i + 1] + hb[2] * Q[3 * i + 2]);
    if (sp > margin) return 0;
    if (sp > best + 1e-6f) { best = sp; code = i; bsign = t[i] >= 0 ? 1.0f : -1.0f; }
  }
#pragma unroll
  for (int j = 0; j < 3; j++) {
    real tb = t[0] * R[j] + t[1] * R[3 + j] + t[2] * R[6 + j];
    real sp = fabsf(tb) - (hb[j] + ha[0] * Q[j] + ha[1] * Q[3 + j] + ha[2] * Q[6 + j]);
    if (sp > margin) return 0;
    if (sp > best + 1e-6f) { best = sp; code = 3 + j; bsign = tb >= 0 ? 1.0f : -1.0f; }
  }
  real en[3] = {0, 0, 0};
  /* the nine edge-edge axes: fully unrolled (compile-time indices keep R, Q, t in registers), one reciprocal square root per axis */
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) {
      const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      const real La = -R[3 * i2 + j], Lb = R[3 * i1 + j], len2 = La * La + Lb * Lb;        /* L[i1] = La, L[i2] = Lb, L[i] = 0 */
      if (len2 < 1e-12f) continue;
      const real inv = rsb_rsqrt(len2);
      const real tl = (t[i1] * La + t[i2] * Lb) * inv;
      const real ra = (ha[i1] * Q[3 * i2 + j] + ha[i2] * Q[3 * i1 + j]) * inv;
      const real rb = (hb[j1] * Q[3 * i + j2] + hb[j2] * Q[3 * i + j1]) * inv;
      const real sp = fabsf(tl) - (ra + rb);
      if (sp > margin) return 0;
      if (sp > best + 1e-4f) { best = sp; code = 6 + 3 * i + j; bsign = tl >= 0 ? 1.0f : -1.0f; en[i] = 0; en[i1] = La * inv; en[i2] = Lb * inv; }
    }
  if (code < 0) return 0;
  if (code >= 6) {
    int i = (code - 6) / 3, j = (code - 6) % 3;
    real nA[3] = {en[0] * bsign, en[1] * bsign, en[2] * bsign}, n[3]; matvec3(n, Ra, nA);
    real ca[3] = {pa[0], pa[1], pa[2]}, cb[3] = {pb[0], pb[1], pb[2]};
    for (int k = 0; k < 3; k++) if (k != i) { real sg = nA[k] > 0 ? 1.0f : -1.0f; for (int c = 0; c < 3; c++) ca[c] += sg * ha[k] * Ra[3 * c + k]; }
    real nB[3]; matTvec3(nB, Rb, n);
    for (int k = 0; k < 3; k++) if (k != j) { real sg = nB[k] > 0 ? -1.0f : 1.0f; for (int c = 0; c < 3; c++) cb[c] += sg * hb[k] * Rb[3 * c + k]; }
    real ua[3] = {Ra[i], Ra[3 + i], Ra[6 + i]}, ub[3] = {Rb[j], Rb[3 + j], Rb[6 + j]}, w[3] = {ca[0] - cb[0], ca[1] - cb[1], ca[2] - cb[2]};
    real bb = dot3(ua, ub), dd = dot3(ua, w), ee = dot3(ub, w), den = 1 - bb * bb;
    real sa = den > 1e-12f ? (bb * ee - dd) / den : 0, sb = den > 1e-12f ? (ee - bb * dd) / den : 0;
    sa = clampf(sa, -ha[i], ha[i]); sb = clampf(sb, -hb[j], hb[j]);
    out->dist = best;
    for (int k = 0; k < 3; k++) { out->normal[k] = n[k]; out->pos[k] = 0.5f * ((ca[k] + ua[k] * sa) + (cb[k] + ub[k] * sb)); }
    return 1;
  }
  const real *pr, *Rr, *hr, *pi, *Ri, *hi; int ax; real nsign;
  if (code < 3) { pr = pa; Rr = Ra; hr = ha; pi = pb; Ri = Rb; hi = hb; ax = code; nsign = bsign; }
  else { pr = pb; Rr = Rb; hr = hb; pi = pa; Ri = Ra; hi = ha; ax = code - 3; nsign = -bsign; }
  real nr[3] = {Rr[ax] * nsign, Rr[3 + ax] * nsign, Rr[6 + ax] * nsign};
  real nl[3]; matTvec3(nl, Ri, nr); int ia = 0; real am = fabsf(nl[0]);
  for (int k = 1; k < 3; k++) if (fabsf(nl[k]) > am) { am = fabsf(nl[k]); ia = k; }
  real isg = nl[ia] > 0 ? -1.0f : 1.0f; int u = (ia + 1) % 3, v = (ia + 2) % 3;
  real *poly = scr, *tmp = scr + 24; int np = 4;               /* a quad clipped by four half-planes has at most 8 vertices */
  const int ru = (ax + 1) % 3, rv = (ax + 2) % 3; bool inside = true;
  for (int c = 0; c < 4; c++) {
    real su = (c == 0 || c == 3) ? -1.0f : 1.0f, sv = (c < 2) ? -1.0f : 1.0f; real wv[3], pv[3];
    for (int k = 0; k < 3; k++) wv[k] = pi[k] + isg * hi[ia] * Ri[3 * k + ia] + su * hi[u] * Ri[3 * k + u] + sv * hi[v] * Ri[3 * k + v] - pr[k];
    matTvec3(pv, Rr, wv); poly[3 * c] = pv[0]; poly[3 * c + 1] = pv[1]; poly[3 * c + 2] = pv[2];
    const real pu = ru == 0 ? pv[0] : (ru == 1 ? pv[1] : pv[2]), pw = rv == 0 ? pv[0] : (rv == 1 ? pv[1] : pv[2]);
    inside = inside && fabsf(pu) <= hr[ru] && fabsf(pw) <= hr[rv];
  }
  /* the incident face lies inside the reference face (a small box resting on a large one: every vertex passes all four half-planes, so the
     clipping below would copy the quad unchanged four times) */
  for (int side = 0; side < (inside ? 0 : 4); side++) {
    int k = side < 2 ? ru : rv; real sg = (side & 1) ? -1.0f : 1.0f, lim = hr[k];
    int nn = 0;
    for (int c = 0; c < np; c++) {
      const int c1 = (c + 1 == np) ? 0 : c + 1;
      const real P0 = poly[3 * c], P1 = poly[3 * c + 1], P2 = poly[3 * c + 2], Q0 = poly[3 * c1], Q1 = poly[3 * c1 + 1], Q2 = poly[3 * c1 + 2];
      const real Pk = k == 0 ? P0 : (k == 1 ? P1 : P2), Qk = k == 0 ? Q0 : (k == 1 ? Q1 : Q2);
      real dp = sg * Pk - lim, dq = sg * Qk - lim;
      if (dp <= 0 && nn < 8) { tmp[3 * nn] = P0; tmp[3 * nn + 1] = P1; tmp[3 * nn + 2] = P2; nn++; }
      if (((dp < 0 && dq > 0) || (dp > 0 && dq < 0)) && nn < 8) { real f = dp / (dp - dq); tmp[3 * nn] = P0 + f * (Q0 - P0); tmp[3 * nn + 1] = P1 + f * (Q1 - P1); tmp[3 * nn + 2] = P2 + f * (Q2 - P2); nn++; }
    }
    np = nn; { real *t_ = poly; poly = tmp; tmp = t_; }            /* swap the buffers instead of copying back */
    if (np == 0) return 0;
  }
  int cnt = 0;
  real nout[3]; for (int k = 0; k < 3; k++) nout[k] = (code < 3) ? nr[k] : -nr[k];
  for (int c = 0; c < np && cnt < 8; c++) {
    real pl[3] = {poly[3 * c], poly[3 * c + 1], poly[3 * c + 2]};
    real depth = nsign * pl[ax] - hr[ax];
    if (depth > margin) continue;
    pl[ax] -= nsign * 0.5f * depth;
    RawCon *o = &out[cnt++]; real w[3]; matvec3(w, Rr, pl);
    o->dist = depth; for (int k = 0; k < 3; k++) { o->pos[k] = pr[k] + w[k]; o->normal[k] = nout[k]; }
  }
  return cnt;
}

/* contact record in shared memory (RSB_CONW = 10 words): pos3 normal3 dist mu | pair efc_address.  Everything else is a function of
   the candidate pair (condim, geoms, includemargin) or of the normal (the tangent frame, rebuilt where the Jacobian is formed). */
#define CON_DIST 6
#define CON_MU 7
#define CON_PAIR 8
#define CON_ADR 9
#define CON_DIM_OF(ci) (MDL.pair_dim[(ci)[CON_PAIR]])
#define CON_G1_OF(ci) (MDL.pair_g1[(ci)[CON_PAIR]])
#define CON_G2_OF(ci) (MDL.pair_g2[(ci)[CON_PAIR]])
#define MISC_NCON 0
#define MISC_NEFC 1
#define MISC_ITER 2
#define MISC_NLIMROW 3
#define MISC_OVF 4              /* sticky over the control step: bit 0 = a contact did not fit ncon_max, bit 1 = a constraint row did not fit nefc_max */
#define MISC_ITERSUM 5          /* Newton iterations of this env summed over the substeps of the control step */

/* Conservative box-pair cull (after the bounding spheres): the bounding sphere of one box against the other box itself, both ways.
   A flat table has a bounding sphere of 0.57 m that every gripper geom is always inside; its slab is what matters. */
RSB_D bool sphere_box_apart(const real *c, real r, const real *bp, const real *bm, const real *h, real margin) {
  real d[3] = {c[0] - bp[0], c[1] - bp[1], c[2] - bp[2]}, l[3]; matTvec3(l, bm, d);
  real ex = fmaxf(fabsf(l[0]) - h[0], 0.0f), ey = fmaxf(fabsf(l[1]) - h[1], 0.0f), ez = fmaxf(fabsf(l[2]) - h[2], 0.0f), rr = r + margin;
  return ex * ex + ey * ey + ez * ez > rr * rr;
}
RSB_D bool box_pair_separated(const real *p1, const real *R1, const real *h1, real r1, const real *p2, const real *R2, const real *h2, real r2, real margin) {
  real dv[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]}, bound = r1 + r2 + margin;
  if (dot3(dv, dv) > bound * bound) return true;
  return sphere_box_apart(p2, r2, p1, R1, h1, margin) || sphere_box_apart(p1, r1, p2, R2, h2, margin);
}
#define RSB_BB_SLOTS 4
RSB_DN void st_collision(int so, Grp g) { real *s = RSB_SMEM + so;
  const real *gxpos = s + MDL.o_gxpos, *gxmat = s + MDL.o_gxmat; real *con = s + MDL.o_con; int *misc = (int *)(s + MDL.o_misc);
  int base = 0;
  for (int p0 = 0; p0 < MDL.npair; p0 += RSB_LANES) {
    int p = p0 + g.lane; RawCon rc[9]; int n = 0; real inc = 0; bool bb = false;
    /* broad phase first, for every lane, so that the box-box candidates can be handed shared-memory scratch by rank */
    if (p < MDL.npair) {
      int g1 = MDL.pair_g1[p], g2 = MDL.pair_g2[p], t1 = MDL.geom_type[g1], t2 = MDL.geom_type[g2];
      bb = (t1 == RSB_GEOM_BOX && t2 == RSB_GEOM_BOX) && !box_pair_separated(gxpos + 3 * g1, gxmat + 9 * g1, MDL.geom_size + 3 * g1, MDL.geom_rbound[g1],
                                                                               gxpos + 3 * g2, gxmat + 9 * g2, MDL.geom_size + 3 * g2, MDL.geom_rbound[g2], MDL.pair_margin[p]);
    }
    const int rank = gscan_incl(g, bb ? 1 : 0) - 1;                /* rank among this group's box-box candidates */
    real lscr[48]; real *scr = (bb && rank < RSB_BB_SLOTS) ? s + MDL.o_cscr + 48 * rank : lscr;       /* o_cscr is free between the inertia and the controller stages */
    if (p < MDL.npair) {
      int g1 = MDL.pair_g1[p], g2 = MDL.pair_g2[p], t1 = MDL.geom_type[g1], t2 = MDL.geom_type[g2];
      real margin = MDL.pair_margin[p]; const real *p1 = gxpos + 3 * g1, *p2 = gxpos + 3 * g2, *R1 = gxmat + 9 * g1, *R2 = gxmat + 9 * g2;
      const real *s1 = MDL.geom_size + 3 * g1, *s2 = MDL.geom_size + 3 * g2;
      bool reject;
      real dv[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
      if (t1 == RSB_GEOM_PLANE) { real nn[3] = {R1[2], R1[5], R1[8]}; reject = dot3(dv, nn) > MDL.geom_rbound[g2] + margin; }
      else { real bound = MDL.geom_rbound[g1] + MDL.geom_rbound[g2] + margin; reject = dot3(dv, dv) > bound * bound; }
      if (!reject) {
        if (t1 == RSB_GEOM_PLANE && t2 == RSB_GEOM_BOX) n = col_plane_box(p1, R1, p2, R2, s2, margin, rc);
        else if (t1 == RSB_GEOM_PLANE && t2 == RSB_GEOM_SPHERE) n = col_plane_sphere(p1, R1, p2, s2[0], margin, rc);
        else if (t1 == RSB_GEOM_PLANE && t2 == RSB_GEOM_CAPSULE) n = col_plane_capsule(p1, R1, p2, R2, s2, margin, rc);
        else if (t1 == RSB_GEOM_SPHERE && t2 == RSB_GEOM_SPHERE) n = col_sphere_sphere(p1, s1[0], p2, s2[0], margin, rc);
        else if (t1 == RSB_GEOM_SPHERE && t2 == RSB_GEOM_BOX) n = col_sphere_box(p1, s1[0], p2, R2, s2, margin, rc);
        else if (t1 == RSB_GEOM_CAPSULE && t2 == RSB_GEOM_CAPSULE) n = col_capsule_capsule(p1, R1, s1, p2, R2, s2, margin, rc);
        else if (t1 == RSB_GEOM_CAPSULE && t2 == RSB_GEOM_BOX) n = col_capsule_box(p1, R1, s1, p2, R2, s2, margin, rc);
        else if (bb) n = col_box_box(p1, R1, s1, p2, R2, s2, margin, rc, scr);
      }
      inc = margin - MDL.pair_gap[p];
      int keep = 0;                                 /* active contacts only: dist < includemargin; compact in place */
      for (int k = 0; k < n; k++) if (rc[k].dist < inc) { if (keep != k) rc[keep] = rc[k]; keep++; }
      n = keep;
    }
    int incl = gscan_incl(g, n), off = base + incl - n;
    for (int k = 0; k < n; k++) {
      int c = off + k; if (c >= MDL.ncon_max) break;
      real *o = con + c * RSB_CONW; int *oi = (int *)o;
      o[0] = rc[k].pos[0]; o[1] = rc[k].pos[1]; o[2] = rc[k].pos[2];
      o[3] = rc[k].normal[0]; o[4] = rc[k].normal[1]; o[5] = rc[k].normal[2];
      o[CON_DIST] = rc[k].dist; o[CON_MU] = 0; oi[CON_PAIR] = p; oi[CON_ADR] = -1;
    }
    base += gshfl_i(g, incl, RSB_LANES - 1);
  }
  if (base > MDL.ncon_max) { base = MDL.ncon_max; if (g.lane == 0) misc[MISC_OVF] |= 1; }       /* truncation is counted, never silent (RSB_INFO_NCON_OVERFLOW) */
  if (g.lane == 0) misc[MISC_NCON] = base;
  gsync(g);
}

/* ================================================================== A.3.6 constraint rows */
enum { EFC_FRICTION = 0, EFC_LIMIT = 1, EFC_CONTACT_NORMAL = 2, EFC_CONTACT_FRICTION = 3 };
/* one word per constraint row: type in bits 0-1, 'upper limit' flag (J = -e_dof) in bit 2, id (dof / joint / contact) above */
#define ET_PACK(type, neg, id) (((id) << 3) | ((neg) << 2) | (type))
#define ET_TYPE(w) ((w) & 3)
#define ET_NEG(w) (((w) >> 2) & 1)
#define ET_ID(w) ((w) >> 3)

RSB_D real impedance_fn(const real *si, real pos, real margin) {
  real d0 = si[0], d1 = si[1], w = si[2], mid = si[3], pw = si[4];
  if (d0 == d1 || w <= RSB_MINVAL) return 0.5f * (d0 + d1);
  real x = fabsf(pos - margin) / w, y;
  if (x >= 1) return d1;
  if (x <= 0) return d0;
  if (pw == 1) y = x;
  else if (pw == 2) y = (x <= mid) ? x * x / mid : 1 - (1 - x) * (1 - x) / (1 - mid);
  else if (x <= mid) y = powf(x, pw) / powf(mid, pw - 1);
  else y = 1 - powf(1 - x, pw) / powf(1 - mid, pw - 1);
  return d0 + y * (d1 - d0);
}
RSB_DN void st_constraint(int so, Grp g) { real *s = RSB_SMEM + so;
  const real *qpos = s + MDL.o_qpos, *qvel = s + MDL.o_qvel, *cdof = s + MDL.o_cdof, *xpos = s + MDL.o_xpos;
  real *con = s + MDL.o_con, *J = s + MDL.o_J; int *misc = (int *)(s + MDL.o_misc);
  real *epos = s + MDL.o_epos, *emargin = s + MDL.o_emargin, *eR = s + MDL.o_eR, *eD = s + MDL.o_eD, *earef = s + MDL.o_earef;
  int *etid = (int *)(s + MDL.o_etype);
  const int ncon = misc[MISC_NCON], nv = MDL.nv, ldj = MDL.ldj;
  /* rows 0..nfl-1: dof friction loss (static) */
  for (int k = g.lane; k < MDL.nfl; k += RSB_LANES) { etid[k] = ET_PACK(EFC_FRICTION, 0, MDL.fl_dof[k]); epos[k] = 0; emargin[k] = 0; }
  /* joint limits: ordered compaction (joint order, lower side before upper side) */
  int nrow = MDL.nfl;
  for (int k0 = 0; k0 < MDL.nlimj; k0 += RSB_LANES) {
    int k = k0 + g.lane, cnt = 0; real dlo = 0, dhi = 0, mg = 0; int j = 0;
    if (k < MDL.nlimj) { j = MDL.lim_jnt[k]; real q = qpos[MDL.jnt_qadr[j]]; mg = MDL.jnt_margin[j]; dlo = q - MDL.jnt_range[2 * j]; dhi = MDL.jnt_range[2 * j + 1] - q; cnt = (dlo < mg) + (dhi < mg); }
    int incl = gscan_incl(g, cnt), r = nrow + incl - cnt;
    if (cnt && dlo < mg && r < MDL.nefc_max) { etid[r] = ET_PACK(EFC_LIMIT, 0, j); epos[r] = dlo; emargin[r] = mg; r++; }
    if (cnt && dhi < mg && r < MDL.nefc_max) { etid[r] = ET_PACK(EFC_LIMIT, 1, j); epos[r] = dhi; emargin[r] = mg; }
    nrow += gshfl_i(g, incl, RSB_LANES - 1);
  }
  if (nrow > MDL.nefc_max) { nrow = MDL.nefc_max; if (g.lane == 0) misc[MISC_OVF] |= 2; }
  const int nscalar = nrow;
  if (g.lane == 0) misc[MISC_NLIMROW] = nscalar;                 /* number of scalar rows (friction loss + limits): they precede the contact rows */
  /* contact row addresses: serial rule of the reference (a contact that does not fit is skipped, later ones may fit) */
  gsync(g);
  if (g.lane == 0) {
    int n = nscalar;
    for (int c = 0; c < ncon; c++) { int *ci = (int *)(con + c * RSB_CONW); int dim = CON_DIM_OF(ci);
      if (n + dim > MDL.nefc_max) { ci[CON_ADR] = -1; misc[MISC_OVF] |= 2; } else { ci[CON_ADR] = n; n += dim; } }
    misc[MISC_NEFC] = n;
  }
  gsync(g);
  const int nefc = misc[MISC_NEFC];
  /* scalar rows of J */
  const int nvp = 1 << MDL.nvsh;                                 /* items are (row, dof) with dof = item & (nvp - 1): no integer division */
  for (int i = g.lane; i < nscalar * nvp; i += RSB_LANES) {
    int r = i >> MDL.nvsh, d = i & (nvp - 1); real v = 0; if (d >= nv) continue;
    const int w = etid[r];
    if (ET_TYPE(w) == EFC_FRICTION) v = (ET_ID(w) == d) ? 1.0f : 0.0f;
    else v = (MDL.jnt_dadr[ET_ID(w)] == d) ? (ET_NEG(w) ? -1.0f : 1.0f) : 0.0f;
    J[r * ldj + d] = v;
  }
  /* first tangent of every contact frame, once per contact (the item loop below would rebuild the frame for each of a contact's nv items);
     kept in the Hessian-weight array, which the solver only writes later.  Needs 3 ncon_max <= nefc_max (checked at create). */
  real *tcache = s + MDL.o_ew;
  if (MDL.frame_cache) {
    for (int c = g.lane; c < ncon; c += RSB_LANES) { const real *cr = con + c * RSB_CONW; real fr[9] = {cr[3], cr[4], cr[5], 0, 0, 0, 0, 0, 0}; make_frame(fr);
      tcache[3 * c] = fr[3]; tcache[3 * c + 1] = fr[4]; tcache[3 * c + 2] = fr[5]; }
    gsync(g);
  }
  /* contact rows of J: item = (contact, dof) */
  for (int i = g.lane; i < ncon * nvp; i += RSB_LANES) {
    int c = i >> MDL.nvsh, d = i & (nvp - 1); if (d >= nv) continue;
    const real *cr = con + c * RSB_CONW; const int *ci = (const int *)cr;
    int adr = ci[CON_ADR]; if (adr < 0) continue;
    const int pr_ = ci[CON_PAIR], dim = MDL.pair_dim[pr_];
    real fr[9] = {cr[3], cr[4], cr[5], 0, 0, 0, 0, 0, 0};
    if (MDL.frame_cache) { fr[3] = tcache[3 * c]; fr[4] = tcache[3 * c + 1]; fr[5] = tcache[3 * c + 2]; cross3(fr + 6, fr, fr + 3); } else make_frame(fr);
    int sgn = ((MDL.pair_dm2[pr_] >> d) & 1) - ((MDL.pair_dm1[pr_] >> d) & 1);
    real jp[3] = {0, 0, 0}, jr[3] = {0, 0, 0};
    if (sgn != 0) {
      int r = MDL.dof_root[d]; real off[3] = {cr[0] - xpos[3 * r], cr[1] - xpos[3 * r + 1], cr[2] - xpos[3 * r + 2]}, t[3];
      cross3(t, cdof + 6 * d, off); real sg = (real)sgn;
      jp[0] = sg * (cdof[6 * d + 3] + t[0]); jp[1] = sg * (cdof[6 * d + 4] + t[1]); jp[2] = sg * (cdof[6 * d + 5] + t[2]);
      jr[0] = sg * cdof[6 * d]; jr[1] = sg * cdof[6 * d + 1]; jr[2] = sg * cdof[6 * d + 2];
    }
    J[adr * ldj + d] = dot3(fr, jp);
    if (dim >= 3) { J[(adr + 1) * ldj + d] = dot3(fr + 3, jp); J[(adr + 2) * ldj + d] = dot3(fr + 6, jp); }
    if (dim >= 4) J[(adr + 3) * ldj + d] = dot3(fr, jr);
  }
  /* contact row bookkeeping */
  for (int c = g.lane; c < ncon; c += RSB_LANES) {
    const real *cr = con + c * RSB_CONW; const int *ci = (const int *)cr; int adr = ci[CON_ADR]; if (adr < 0) continue;
    const int pr = ci[CON_PAIR]; const real inc = MDL.pair_margin[pr] - MDL.pair_gap[pr];
    for (int r = 0; r < CON_DIM_OF(ci); r++) { etid[adr + r] = ET_PACK(r == 0 ? EFC_CONTACT_NORMAL : EFC_CONTACT_FRICTION, 0, c);
      epos[adr + r] = r == 0 ? cr[CON_DIST] : 0; emargin[adr + r] = r == 0 ? inc : 0; }
  }
  gsync(g);
  /* impedance, regulariser, reference acceleration (mj_makeImpedance), lane per row */
  for (int r = g.lane; r < nefc; r += RSB_LANES) {
    const real *solimp, *kb; real diag; int type = ET_TYPE(etid[r]), id = ET_ID(etid[r]);       /* (K, B) and the inverse weights are host-side tables */
    if (type == EFC_FRICTION) { kb = MDL.dof_kb + 2 * id; solimp = MDL.dof_solimp + 5 * id; diag = MDL.dof_invw[id]; }
    else if (type == EFC_LIMIT) { kb = MDL.jnt_kb + 2 * id; solimp = MDL.jnt_solimp + 5 * id; diag = MDL.dof_invw[MDL.jnt_dadr[id]]; }
    else {
      const int *ci = (const int *)(con + id * RSB_CONW); int p = ci[CON_PAIR], rr = r - ci[CON_ADR];
      kb = MDL.pair_kb + 2 * p; solimp = MDL.pair_solimp + 5 * p; diag = MDL.pair_invw[2 * p + (rr < 3 ? 0 : 1)];
    }
    real K = kb[0]; const real B = kb[1];
    if (type == EFC_FRICTION || type == EFC_CONTACT_FRICTION) K = 0;
    real imp = impedance_fn(solimp, epos[r], emargin[r]);
    real Rr = fmaxf((1 - imp) / imp * diag, RSB_MINVAL); eR[r] = Rr;
    real vel = sdot(J + r * ldj, qvel, nv);
    earef[r] = -B * vel - K * imp * (epos[r] - emargin[r]);
  }
  gsync(g);
  /* elliptic friction rows: R from the normal row and impratio; regularised cone slope mu */
  for (int c = g.lane; c < ncon; c += RSB_LANES) {
    real *cr = con + c * RSB_CONW; const int *ci = (const int *)cr; int i = ci[CON_ADR], dim = CON_DIM_OF(ci); const real *fr = MDL.pair_friction + 5 * ci[CON_PAIR];
    if (i < 0) continue;
    if (dim < 2) { cr[CON_MU] = fr[0]; continue; }
    real ir = fmaxf(MDL.impratio, RSB_MINVAL);
    eR[i + 1] = eR[i] / ir; cr[CON_MU] = fr[0] * sqrtf(eR[i + 1] / eR[i]);
    for (int j = 1; j < dim - 1; j++) eR[i + 1 + j] = eR[i + 1] * fr[0] * fr[0] / (fr[j] * fr[j]);
  }
  gsync(g);
  for (int r = g.lane; r < nefc; r += RSB_LANES) eD[r] = 1.0f / eR[r];
  gsync(g);
}

/* ================================================================== dense Cholesky, one matrix row per lane (n <= RSB_LANES) */
/* Every factorisation of this kernel is followed by a solve with the fresh factor (qacc_smooth, the Newton direction, the implicit
   Euler step, the task-space systems of the OSC law), and one such pair runs 5-13 times per physics substep on the critical path of an
   environment: it is written for LATENCY.
   chol_fs_t<N>: FUSED factor + solve with the matrix row in REGISTERS.  Lane i owns row i (identity-padded to N, so no size tests).
     factor  (right-looking): at column j the pivot is broadcast by one shuffle, every trailing entry takes one shuffle + one FMA;
             dependent chain per column = shuffle -> rsqrt -> mul -> (shuffle, FMA): ~110 cycles, no shared-memory traffic.
     forward sweep straight from the row registers; the factor is then stored (inverse pivots on the diagonal) because the backward
     sweep needs COLUMN i of L on lane i (loaded up front, off the dependent chain); each sweep step is one shuffle + one FMA.
   Measured against a rolled shared-memory variant (150 instructions, dot-product loops): 4x fewer cycles per call, which outweighs the
   larger footprint in the instruction cache (profiles/, stage profile of round 1).
   Pivots <= 1e-30 are dropped directions (pinv-like): inverse pivot 0.  ld == 0 selects packed lower-triangular storage.  xo < 0:
   factor only (the factor is still stored). */
template <int N> RSB_D void chol_fs_t(real *A, int n, int ld, real *x, bool solve, Grp g) {
  const int i = g.lane; real a[N]; real *Ai = A + (ld ? i * ld : tri_off(i)); const bool act = i < n;
#pragma unroll
  for (int k = 0; k < N; k++) a[k] = (act && k <= i) ? Ai[k] : ((k == i) ? 1.0f : 0.0f);
  real b = (solve && act) ? x[i] : 0.0f, dinv = 0.0f;
#pragma unroll
  for (int j = 0; j < N; j++) {
    const real sj = gshfl(g, a[j], j), inv = sj > 1e-30f ? rsb_rsqrt(sj) : 0.0f;
    const real lij = (i == j) ? inv : a[j] * inv;                 /* lane j keeps 1/L_jj, lanes i > j keep L_ij */
    a[j] = lij; if (i == j) dinv = inv;
#pragma unroll
    for (int k = j + 1; k < N; k++) { const real lkj = gshfl(g, a[j], k); a[k] -= lij * lkj; }   /* lane k supplies L_kj; rows i >= k use it */
  }
#pragma unroll
  for (int k = 0; k < N; k++) if (act && k <= i) Ai[k] = a[k];
  gsync(g);
  if (!solve) return;
  real col[N];
#pragma unroll
  for (int k = 0; k < N; k++) col[k] = (act && k > i && k < n) ? A[(ld ? k * ld : (k * (k + 1)) / 2) + i] : 0.0f;
#pragma unroll
  for (int k = 0; k < N; k++) { const real xk = gshfl(g, b * dinv, k); b = (i == k) ? xk : ((k < i) ? b - a[k] * xk : b); }      /* forward: L y = b */
#pragma unroll
  for (int k = N - 1; k >= 0; k--) { const real xk = gshfl(g, b * dinv, k); b = (i == k) ? xk : b - col[k] * xk; }               /* backward: L^T x = y */
  if (act) x[i] = b;
  gsync(g);
}
/* A <- chol(A), then (xo >= 0) x <- A^-1 x */
RSB_DN void chol_factor_solve(int ao, int n, int ld, int xo, Grp g) { real *A = RSB_SMEM + ao; real *x = RSB_SMEM + (xo < 0 ? 0 : xo); const bool sv = xo >= 0;
  if (n <= 8) chol_fs_t<8>(A, n, ld, x, sv, g);
  else if (n <= 12) chol_fs_t<12>(A, n, ld, x, sv, g);
#if RSB_LANES >= 32
  else if (n <= 16) chol_fs_t<16>(A, n, ld, x, sv, g);
  else if (n <= 24) chol_fs_t<24>(A, n, ld, x, sv, g);
  else chol_fs_t<32>(A, n, ld, x, sv, g);
#else
  else chol_fs_t<16>(A, n, ld, x, sv, g);                         /* 16-lane groups serve models with nv <= 16 only (checked at create) */
#endif
}
/* x <- A^-1 x with an already stored factor (a second right-hand side: rare).  Rolled: lane i walks row i, then column i. */
RSB_DN void chol_solve(int lo_, int n, int ld, int xo, Grp g) { const real *L = RSB_SMEM + lo_; real *x = RSB_SMEM + xo;
  const int i = g.lane; const bool act = i < n; const real *Li = L + (ld ? i * ld : tri_off(i));
  real b = act ? x[i] : 0.0f; const real dinv = act ? Li[i] : 0.0f;
  for (int k = 0; k < n; k++) {                                   /* forward: L y = b */
    const real lik = (act && k < i) ? Li[k] : 0.0f, xk = gshfl(g, b * dinv, k);
    b = (i == k) ? xk : b - lik * xk;
  }
  const real *Lk = L + (ld ? (n - 1) * ld : tri_off(n - 1)) + i;  /* L[k][i], k = n-1 .. 0 */
  for (int k = n - 1; k >= 0; k--) {                              /* backward: L^T x = y */
    const real lki = (act && k > i) ? *Lk : 0.0f, xk = gshfl(g, b * dinv, k);
    b = (i == k) ? xk : b - lki * xk;
    Lk -= ld ? ld : k;
  }
  if (act) x[i] = b;
  gsync(g);
}

/* ================================================================== A.2 controllers */
/* controller state words per robot (RSB_CS_WORDS): goal_pos3 goal_ori9 initial_joint7 grip_cur2 goal_vel7 summed_err7 last_err7
   derr_buf35 derr_n derr_ptr saturated */
#define CS_GOALPOS 0
#define CS_GOALORI 3
#define CS_INITJ 12
#define CS_GRIP 19
#define CS_GOALVEL 21
#define CS_SUMERR 28
#define CS_LASTERR 35
#define CS_DERR 42
#define CS_DERRN 77
#define CS_DERRPTR 78
#define CS_SAT 79

RSB_D real scale_action(const DevRobot &rb, int k, real a) {
  real lo = rb.in_min[k], hi = rb.in_max[k]; a = clampf(a, lo, hi);
  real sc = fabsf(rb.out_max[k] - rb.out_min[k]) / fabsf(hi - lo);
  return (a - 0.5f * (hi + lo)) * sc + 0.5f * (rb.out_max[k] + rb.out_min[k]);
}

RSB_DN void ctrl_reset(int so, Grp g) { real *s = RSB_SMEM + so;
  const real *qpos = s + MDL.o_qpos, *sxpos = s + MDL.o_sxpos, *sxmat = s + MDL.o_sxmat;
  for (int ri = 0; ri < MDL.nrobot; ri++) {
    const DevRobot &rb = MDL.robot[ri]; real *cs = s + MDL.o_cs + ri * MDL.cs_words;
    for (int k = g.lane; k < MDL.cs_words; k += RSB_LANES) {
      real v = 0;
      if (k < 3) v = sxpos[3 * rb.eef_site + k];
      else if (k < 12) v = sxmat[9 * rb.eef_site + k - 3];
      else if (k < 19) v = qpos[rb.arm_qadr[k - 12]];
      cs[k] = v;
    }
  }
  gsync(g);
}

RSB_DN void ctrl_set_goal(int so, Grp g) { real *s = RSB_SMEM + so;
  const real *act = s + MDL.o_act, *sxpos = s + MDL.o_sxpos, *sxmat = s + MDL.o_sxmat;
  if (g.lane < MDL.nrobot) {
    int ri = g.lane; const DevRobot &rb = MDL.robot[ri]; real *cs = s + MDL.o_cs + ri * MDL.cs_words; const real *a = act + rb.act_off;
    if (rb.ctrl_type == RSB_CTRL_OSC_POSE || rb.ctrl_type == RSB_CTRL_OSC_POSITION) {
      real d[6] = {0, 0, 0, 0, 0, 0}; for (int k = 0; k < rb.control_dim; k++) d[k] = scale_action(rb, k, a[k]);
      if (rb.ctrl_type == RSB_CTRL_OSC_POSE && (d[3] != 0 || d[4] != 0 || d[5] != 0)) {
        real Rm[9], Rg[9];
        if (rb.ori_mode == RSB_ORI_DELTA_AXIS_ANGLE) {             /* robosuite >= 1.1: rotation by the axis-angle vector d */
          real ang = sqrtf(d[3] * d[3] + d[4] * d[4] + d[5] * d[5]), q[4];
          if (ang < RSB_MINVAL) { q[0] = 1; q[1] = q[2] = q[3] = 0; }
          else { real sn, c; rsb_sincos(0.5f * ang, &sn, &c); real f = sn / ang; q[0] = c; q[1] = f * d[3]; q[2] = f * d[4]; q[3] = f * d[5]; }
          quat2mat(Rm, q);
        } else {                                                   /* euler2mat(d)^T with mujoco-py's euler2mat: the convention of the committed 2020 policies */
          real si, ci, sj, cj, sk, ck; rsb_sincos(-d[5], &si, &ci); rsb_sincos(-d[4], &sj, &cj); rsb_sincos(-d[3], &sk, &ck);
          const real cick = ci * ck, cisk = ci * sk, sick = si * ck, sisk = si * sk;
          /* E[r][c] of euler2mat, stored transposed: Rm[c][r] = E[r][c] */
          Rm[0] = cj * ci;          Rm[3] = cj * si;          Rm[6] = -sj;
          Rm[1] = sj * cisk - sick; Rm[4] = sj * sisk + cick; Rm[7] = cj * sk;
          Rm[2] = sj * cick + sisk; Rm[5] = sj * sick - cisk; Rm[8] = cj * ck;
        }
        matmul3(Rg, Rm, sxmat + 9 * rb.eef_site);
        for (int k = 0; k < 9; k++) cs[CS_GOALORI + k] = Rg[k];
      }
      for (int k = 0; k < 3; k++) cs[CS_GOALPOS + k] = sxpos[3 * rb.eef_site + k] + d[k];
    } else if (rb.ctrl_type == RSB_CTRL_JOINT_VELOCITY) {
      for (int k = 0; k < RSB_ARM_DOF; k++) { real v = scale_action(rb, k, a[k]); if (rb.has_vl) v = clampf(v, rb.vl_lo[k], rb.vl_hi[k]); cs[CS_GOALVEL + k] = v; }
    } else if (rb.ctrl_type == RSB_CTRL_JOINT_POSITION) {        /* robosuite JointPositionController.set_goal: goal_qpos = joint_pos + scaled delta (held in the goal-velocity slot) */
      const real *qpos = s + MDL.o_qpos;
      for (int k = 0; k < RSB_ARM_DOF; k++) cs[CS_GOALVEL + k] = qpos[rb.arm_qadr[k]] + scale_action(rb, k, a[k]);
    } else { for (int k = 0; k < RSB_ARM_DOF; k++) cs[CS_GOALVEL + k] = scale_action(rb, k, a[k]); }
  }
  gsync(g);
}

/* serial 7x7 triangular solves on one lane (fully unrolled: x stays in registers); L carries inverse pivots on its diagonal */
RSB_D void chol7_solve_reg(const real *L, real *x) {
#pragma unroll
  for (int i = 0; i < 7; i++) { real v = x[i];
#pragma unroll
    for (int k = 0; k < i; k++) v -= L[i * 7 + k] * x[k]; x[i] = v * L[i * 7 + i]; }
#pragma unroll
  for (int i = 6; i >= 0; i--) { real v = x[i];
#pragma unroll
    for (int k = i + 1; k < 7; k++) v -= L[k * 7 + i] * x[k]; x[i] = v * L[i * 7 + i]; }
}
/* solve the symmetric 3x3 system A x = b (A given with leading dim ld); x = 0 when A is singular */
RSB_D void sym3_solve(const real *A, int ld, const real *b, real *x) {
  real a = A[0], bq = A[1], c = A[2], d = A[ld + 1], e = A[ld + 2], f = A[2 * ld + 2];
  real c00 = d * f - e * e, c01 = c * e - bq * f, c02 = bq * e - c * d, det = a * c00 + bq * c01 + c * c02;
  real sc = fmaxf(fmaxf(fabsf(a), fabsf(d)), fabsf(f));
  if (fabsf(det) <= 1e-18f * sc * sc * sc || !(det == det)) { x[0] = x[1] = x[2] = 0; return; }
  real id = 1.0f / det, c11 = a * f - c * c, c12 = bq * c - a * e, c22 = a * d - bq * bq;
  x[0] = (c00 * b[0] + c01 * b[1] + c02 * b[2]) * id; x[1] = (c01 * b[0] + c11 * b[1] + c12 * b[2]) * id; x[2] = (c02 * b[0] + c12 * b[1] + c22 * b[2]) * id;
}

/* scratch layout for the OSC law (floats, in o_cscr): Jee[42] Lm[49] X[42] A[36](ld 6) F[6] pose[7] y[6] w[6] v6[6] tau[7] */
/* robosuite JointPositionController.run_controller for robot ri: torques = M_arm (kp (goal_qpos - q) - kd qd) + torque_compensation.  Its own function so that
   the operational-space law's register allocation in ctrl_run does not change. */
RSB_DN void ctrl_run_jpos(int so, Grp g, int ri) { real *s = RSB_SMEM + so;
  const DevRobot &rb = MDL.robot[ri]; const real *cs = s + MDL.o_cs + ri * MDL.cs_words, *qpos = s + MDL.o_qpos, *qvel = s + MDL.o_qvel, *M = s + MDL.o_M, *bias = s + MDL.o_bias;
  real *tau = s + MDL.o_cscr + 200;                    /* ctrl_run's raw-torque slot: its common tail clamps, writes ctrl and keeps the torques */
  if (g.lane < RSB_ARM_DOF) {
    const int c = g.lane; real t = bias[rb.arm_dadr[c]];
    for (int k = 0; k < RSB_ARM_DOF; k++) t += msym(M, rb.arm_dadr[c], rb.arm_dadr[k]) * (rb.kp[k] * (cs[CS_GOALVEL + k] - qpos[rb.arm_qadr[k]]) - rb.kd[k] * qvel[rb.arm_dadr[k]]);
    tau[c] = t;
  }
}

RSB_DN void ctrl_run(int so, Grp g) { real *s = RSB_SMEM + so;
  const real *qpos = s + MDL.o_qpos, *qvel = s + MDL.o_qvel, *cdof = s + MDL.o_cdof, *M = s + MDL.o_M, *bias = s + MDL.o_bias;
  const real *sxpos = s + MDL.o_sxpos, *sxmat = s + MDL.o_sxmat, *xpos = s + MDL.o_xpos, *cvel = s + MDL.o_cvel;
  real *ctrl = s + MDL.o_ctrl; real *scr = s + MDL.o_cscr;
  real *Jee = scr, *Lm = scr + 42, *X = scr + 91, *A = scr + 133, *F = scr + 169, *pose = scr + 175, *y = scr + 182, *w = scr + 188, *v6 = scr + 194, *tau = scr + 200;
  for (int ri = 0; ri < MDL.nrobot; ri++) {
    const DevRobot &rb = MDL.robot[ri]; real *cs = s + MDL.o_cs + ri * MDL.cs_words;
    if (rb.ctrl_type == RSB_CTRL_OSC_POSE || rb.ctrl_type == RSB_CTRL_OSC_POSITION) {
      int sb = MDL.site_body[rb.eef_site], root = MDL.body_root[sb]; const real *ep = sxpos + 3 * rb.eef_site;
      real off[3] = {ep[0] - xpos[3 * root], ep[1] - xpos[3 * root + 1], ep[2] - xpos[3 * root + 2]};
      if (g.lane < RSB_ARM_DOF) {                    /* site Jacobian, arm columns */
        int c = g.lane, d = rb.arm_dadr[c]; real t[3] = {0, 0, 0}, wv[3] = {0, 0, 0}, lv[3] = {0, 0, 0};
        if ((MDL.body_dofmask[sb] >> d) & 1) { cross3(t, cdof + 6 * d, off); wv[0] = cdof[6 * d]; wv[1] = cdof[6 * d + 1]; wv[2] = cdof[6 * d + 2];
          lv[0] = cdof[6 * d + 3] + t[0]; lv[1] = cdof[6 * d + 4] + t[1]; lv[2] = cdof[6 * d + 5] + t[2]; }
        for (int r = 0; r < 3; r++) { Jee[r * 7 + c] = lv[r]; Jee[(3 + r) * 7 + c] = wv[r]; }
      }
      for (int i = g.lane; i < 49; i += RSB_LANES) { int r = i / 7, c = i - 7 * r; Lm[i] = msym(M, rb.arm_dadr[r], rb.arm_dadr[c]); }
      if (g.lane == 8) {                             /* site velocity = full Jacobian x qvel = body twist moved to the site */
        const real *cv = cvel + 6 * sb; real t[3]; cross3(t, cv, off);
        v6[0] = cv[3] + t[0]; v6[1] = cv[4] + t[1]; v6[2] = cv[5] + t[2]; v6[3] = cv[0]; v6[4] = cv[1]; v6[5] = cv[2];
      }
      if (g.lane >= 9 && g.lane < 16) { int k = g.lane - 9; real kn = rb.null_kp; pose[k] = kn * (cs[CS_INITJ + k] - qpos[rb.arm_qadr[k]]) - 2 * sqrtf(kn) * qvel[rb.arm_dadr[k]]; }
      gsync(g);
      chol_factor_solve(SOFF(Lm), 7, 7, -1, g);
      if (g.lane < 6) {                              /* X[r][:] = M^-1 J[r][:]^T */
        real x[7];
#pragma unroll
        for (int k = 0; k < 7; k++) x[k] = Jee[g.lane * 7 + k];
        chol7_solve_reg(Lm, x);
#pragma unroll
        for (int k = 0; k < 7; k++) X[g.lane * 7 + k] = x[k];
      } else if (g.lane == 6) {                      /* desired wrench */
        const real *Rc = sxmat + 9 * rb.eef_site, *Rd = cs + CS_GOALORI; real eo[3] = {0, 0, 0};
        for (int k = 0; k < 3; k++) { real a[3] = {Rc[k], Rc[3 + k], Rc[6 + k]}, b[3] = {Rd[k], Rd[3 + k], Rd[6 + k]}, x3[3]; cross3(x3, a, b); eo[0] += 0.5f * x3[0]; eo[1] += 0.5f * x3[1]; eo[2] += 0.5f * x3[2]; }
        for (int k = 0; k < 3; k++) { F[k] = rb.kp[k] * (cs[CS_GOALPOS + k] - ep[k]) - rb.kd[k] * v6[k]; F[3 + k] = rb.kp[3 + k] * eo[k] - rb.kd[3 + k] * v6[3 + k]; }
      } else if (g.lane == 7) {                      /* y = J pose */
        for (int r = 0; r < 6; r++) { real acc = 0; for (int k = 0; k < 7; k++) acc += Jee[r * 7 + k] * pose[k]; y[r] = acc; }
      }
      gsync(g);
      for (int i = g.lane; i < 36; i += RSB_LANES) { int r = i / 6, c = i - 6 * r; real acc = 0;
#pragma unroll
        for (int k = 0; k < 7; k++) acc += X[r * 7 + k] * Jee[c * 7 + k]; A[i] = acc; }
      gsync(g);
      if (g.lane == 0) {                             /* decoupled task-space inertia applied to the wrench */
        if (rb.uncouple) { sym3_solve(A, 6, F, w); sym3_solve(A + 21, 6, F + 3, w + 3); }
      }
      gsync(g);
      chol_factor_solve(SOFF(A), 6, 6, SOFF(y), g);              /* y = Lambda_full J pose */
      if (!rb.uncouple) { chol_solve(SOFF(A), 6, 6, SOFF(F), g); if (g.lane < 6) w[g.lane] = F[g.lane]; gsync(g); }
      if (g.lane < RSB_ARM_DOF) {
        int c = g.lane; real t = bias[rb.arm_dadr[c]];
        for (int r = 0; r < 6; r++) t += Jee[r * 7 + c] * (w[r] - y[r]);
        for (int k = 0; k < 7; k++) t += msym(M, rb.arm_dadr[c], rb.arm_dadr[k]) * pose[k];
        tau[c] = t;
      }
    } else if (rb.ctrl_type == RSB_CTRL_JOINT_VELOCITY) {
      int ptr = (int)cs[CS_DERRPTR], nfill = (int)cs[CS_DERRN], satur = (int)cs[CS_SAT]; int nn = nfill < 5 ? nfill + 1 : 5;
      real raw = 0, t = 0; int sat = 0;
      if (g.lane < RSB_ARM_DOF) {
        int k = g.lane; real err = cs[CS_GOALVEL + k] - qvel[rb.arm_dadr[k]];
        cs[CS_DERR + ptr * 7 + k] = err - cs[CS_LASTERR + k]; cs[CS_LASTERR + k] = err;
        real avg = 0; for (int i = 0; i < 5; i++) avg += cs[CS_DERR + i * 7 + k] / (real)nn;
        real se = cs[CS_SUMERR + k]; if (!satur) { se += err; cs[CS_SUMERR + k] = se; }
        raw = rb.kp[k] * err + rb.ki[k] * se + rb.kd[k] * avg + bias[rb.arm_dadr[k]];
        t = clampf(raw, rb.tl_lo[k], rb.tl_hi[k]); sat = (t != raw); tau[k] = t;
      }
      sat = gsum_i(g, sat);
      gsync(g);
      if (g.lane == 0) { cs[CS_DERRPTR] = (real)((ptr + 1) % 5); cs[CS_DERRN] = (real)nn; cs[CS_SAT] = sat ? 1.0f : 0.0f; }
    } else if (rb.ctrl_type == RSB_CTRL_JOINT_POSITION) {
      ctrl_run_jpos(so, g, ri);
    } else {
      if (g.lane < RSB_ARM_DOF) tau[g.lane] = cs[CS_GOALVEL + g.lane] + bias[rb.arm_dadr[g.lane]];
    }
    gsync(g);
    if (g.lane < RSB_ARM_DOF) { int k = g.lane; real t = clampf(tau[k], rb.tl_lo[k], rb.tl_hi[k]); tau[k] = t; ctrl[rb.arm_act[k]] = t; }
    /* gripper: integrate the binary open/close command (robosuite Gripper.format_action) */
    if (rb.grip_action_dim > 0 && g.lane >= 8 && g.lane < 8 + rb.grip_ndof) {
      int k = g.lane - 8; real ga = s[MDL.o_act + rb.act_off + rb.control_dim]; real sg = ga > 0 ? 1.0f : (ga < 0 ? -1.0f : 0.0f);
      real v = clampf(cs[CS_GRIP + k] + rb.grip_sign[k] * rb.grip_speed * sg, -1.0f, 1.0f); cs[CS_GRIP + k] = v;
      int a = rb.grip_act[k]; real lo = MDL.act_crange[2 * a], hi = MDL.act_crange[2 * a + 1];
      ctrl[a] = 0.5f * (hi + lo) + 0.5f * (hi - lo) * v;
    }
    gsync(g);
    /* keep the last torques for debugging/parity: cscr[208 + 7*ri ...] */
    if (g.lane < RSB_ARM_DOF) s[MDL.o_tau + 7 * ri + g.lane] = tau[g.lane];
    gsync(g);
  }
}

/* ================================================================== actuation + smooth acceleration */
RSB_DN void st_actuation(int so, Grp g) { real *s = RSB_SMEM + so;
  const real *qpos = s + MDL.o_qpos, *qvel = s + MDL.o_qvel, *ctrl = s + MDL.o_ctrl, *bias = s + MDL.o_bias, *passive = s + MDL.o_passive;
  real *actf = s + MDL.o_actuator, *smooth = s + MDL.o_smooth, *qas = s + MDL.o_qacc_smooth, *af = s + MDL.o_cscr;
  /* lane per ACTUATOR: its ~12 model loads are independent (one latency), the force goes to scratch (the controller scratch is dead here);
     then lane per dof sums the actuators attached to it in actuator order (deterministic) */
  for (int a = g.lane; a < MDL.nu; a += RSB_LANES) {
    const int d = MDL.act_dof[a], qa = MDL.jnt_qadr[MDL.dof_jnt[d]];
    real c = ctrl[a]; if (MDL.act_climited[a]) c = clampf(c, MDL.act_crange[2 * a], MDL.act_crange[2 * a + 1]);
    const real gear = MDL.act_gear[a], len = qpos[qa] * gear, vel = qvel[d] * gear;
    real fa = MDL.act_gain[a] * c + MDL.act_bias[3 * a] + MDL.act_bias[3 * a + 1] * len + MDL.act_bias[3 * a + 2] * vel;
    if (MDL.act_flimited[a]) fa = clampf(fa, MDL.act_frange[2 * a], MDL.act_frange[2 * a + 1]);
    af[a] = gear * fa;
  }
  gsync(g);
  for (int d = g.lane; d < MDL.nv; d += RSB_LANES) {
    real f = 0;
    for (int a = 0; a < MDL.nu; a++) f += (MDL.act_dof[a] == d) ? af[a] : 0.0f;
    actf[d] = f; real sm = passive[d] - bias[d] + f; smooth[d] = sm; qas[d] = sm;
  }
  gsync(g);
}
/* qacc_smooth = M^-1 qfrc_smooth.  Runs AFTER the constraint stage: the factor workspace overlays the poses / cdof that stage reads. */
RSB_DN void st_smooth_acc(int so, Grp g) { real *s = RSB_SMEM + so; const real *M = s + MDL.o_M; real *L = s + MDL.o_L;
  for (int i = g.lane; i < MDL.ntri; i += RSB_LANES) L[i] = M[i];
  gsync(g);
  chol_factor_solve(so + MDL.o_L, MDL.nv, 0, so + MDL.o_qacc_smooth, g);
}

/* ================================================================== A.3.7 constraint solver (Newton, exact line search) */
/* Constraint cost of this lane's work items at the residuals jar (x = J qacc - aref); with_forces: also the row forces and the Hessian
   weights (ew = D on quadratic rows, 0 on inactive ones, < 0 on the rows of a contact in the middle zone of its cone).
   Work ITEMS, one per lane: the scalar rows (they come first) and the contacts (a contact's rows are handled from its first row); a
   lane-per-row loop would leave the friction rows' lanes idle and need two passes from 17 rows on.  The line search has its own
   evaluator (newton_linesearch). */
RSB_DN real efc_eval(int so, Grp g, bool with_forces, int jo) { real *s = RSB_SMEM + so;
  const real *con = s + MDL.o_con, *eD = s + MDL.o_eD, *jar = RSB_SMEM + jo;      /* jo: the residual vector to evaluate (o_ejar; o_eJv holds a second candidate during warm-start selection) */
  real *force = s + MDL.o_eforce, *ew = s + MDL.o_ew; const int *etid = (const int *)(s + MDL.o_etype);
  real cost = 0;
  const int *misc = (const int *)(s + MDL.o_misc); const int nscalar = misc[MISC_NLIMROW], nitem = nscalar + misc[MISC_NCON];
  for (int it = g.lane; it < nitem; it += RSB_LANES) {
    const int r = it < nscalar ? it : ((const int *)(con + (it - nscalar) * RSB_CONW))[CON_ADR];
    if (r < 0) continue;                                          /* contact without rows (row limit reached) */
    const int wd = etid[r], type = ET_TYPE(wd); const real D = eD[r], x = jar[r];
    if (type == EFC_FRICTION) {
      const real fl = MDL.dof_floss[ET_ID(wd)], rf = fl / D;    /* R * frictionloss */
      if (x <= -rf) { cost += -0.5f * rf * fl - fl * x; if (with_forces) { force[r] = fl; ew[r] = 0; } }
      else if (x >= rf) { cost += -0.5f * rf * fl + fl * x; if (with_forces) { force[r] = -fl; ew[r] = 0; } }
      else { cost += 0.5f * D * x * x; if (with_forces) { force[r] = -D * x; ew[r] = D; } }
      continue;
    }
    const real *cr = con + ET_ID(wd) * RSB_CONW; const int dim = (type == EFC_LIMIT) ? 1 : CON_DIM_OF((const int *)cr);
    if (dim == 1) {
      if (x < 0) { cost += 0.5f * D * x * x; if (with_forces) { force[r] = -D * x; ew[r] = D; } }
      else if (with_forces) { force[r] = 0; ew[r] = 0; }
      continue;
    }
    /* elliptic cone, dim in {3, 4}; scaled coordinates U = sc x = (N, t), T = |t|.  All per-row arrays are indexed by unrolled loops only
       (registers: a run-time index would put them in local memory, an L2 round trip per access at this kernel's shared-memory footprint) */
    const real *fr = MDL.pair_friction + 5 * ((const int *)cr)[CON_PAIR]; const real mu = cr[CON_MU];
    real sc[RSB_MAXDIM], U[RSB_MAXDIM], xs[RSB_MAXDIM]; real T2 = 0;
#pragma unroll
    for (int j = 0; j < RSB_MAXDIM; j++) { sc[j] = 0; U[j] = 0; xs[j] = 0;
      if (j < dim) { sc[j] = (j == 0) ? mu : fr[j - 1]; xs[j] = jar[r + j]; U[j] = xs[j] * sc[j]; if (j > 0) T2 += U[j] * U[j]; } }
    const real N = U[0], invT = T2 > 0 ? rsb_rsqrt(T2) : 0.0f, T = T2 * invT;
    if (N >= mu * T || (T <= 0 && N >= 0)) {                         /* top zone: separated / inside the dual cone */
      if (with_forces) {
#pragma unroll
        for (int j = 0; j < RSB_MAXDIM; j++) if (j < dim) { force[r + j] = 0; ew[r + j] = 0; } }
    } else if (mu * N + T <= 0 || (T <= 0 && N < 0)) {              /* bottom zone: plain quadratic */
#pragma unroll
      for (int j = 0; j < RSB_MAXDIM; j++) if (j < dim) {
        const real Dj = eD[r + j]; cost += 0.5f * Dj * xs[j] * xs[j];
        if (with_forces) { force[r + j] = -Dj * xs[j]; ew[r + j] = Dj; }
      }
    } else {                                                         /* middle zone: cone surface, cost = Dm (N - mu T)^2 / 2 */
      const real Dm = D / (mu * mu * (1 + mu * mu)), NmT = N - mu * T;
      cost += 0.5f * Dm * NmT * NmT;
      if (with_forces) {                                             /* force = -dcost/dx; ew < 0 marks "use the cone block" */
        const real gt = Dm * mu * NmT * invT;
#pragma unroll
        for (int j = 0; j < RSB_MAXDIM; j++) if (j < dim) { force[r + j] = (j == 0) ? -Dm * NmT * mu : gt * U[j] * sc[j]; ew[r + j] = -1.0f; }
      }
    }
  }
  return cost;
}

/* developer build (-DRSB_PROFILE, tools/stage_profile.py): cycles per stage and per barrier wait, accumulated per warp */
#if defined(RSB_PROFILE) && !defined(RSB_EMU)
#define RSB_PROF_SLOTS 16
__device__ unsigned long long g_prof[8192 * RSB_PROF_SLOTS];
#define PROF(k) do { long long n_ = clock64(); if ((threadIdx.x & 31) == 0) g_prof[(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RSB_PROF_SLOTS + (k)] += (unsigned long long)(n_ - pt_); pt_ = n_; } while (0)
#define PROF_DECL long long pt_ = clock64()
#define PROF_LOCAL long long pt_ = clock64()
#define PROF_COUNT(k) do { if ((threadIdx.x & 31) == 0) g_prof[(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RSB_PROF_SLOTS + (k)] += (1ull << 40); } while (0)
#else
#define PROF(k) ((void)0)
#define PROF_DECL long long pt_ = 0; (void)pt_
#define PROF_LOCAL ((void)0)
#define PROF_COUNT(k) ((void)0)
#endif
/* ---- shared building blocks of the Newton solver.  Each is ONE non-inlined function: the solver loop then is ~1.5k instructions in
   total and stays in the instruction cache across iterations (the envs that need 5-9 iterations are the kernel's critical path). */
/* y[r] = J[r,:] . x (- aref[r]) for every constraint row (lane per row) */
RSB_DN void efc_mulJ(int so, Grp g, int nefc, int xo, int yo, int sub_aref) { real *s = RSB_SMEM + so;
  const real *J = s + MDL.o_J, *x = RSB_SMEM + xo, *earef = s + MDL.o_earef; real *y = RSB_SMEM + yo; const int nv = MDL.nv, ldj = MDL.ldj;
#pragma unroll 1
  for (int r = g.lane; r < nefc; r += RSB_LANES) { real v = sdot(J + r * ldj, x, nv); if (sub_aref) v -= earef[r]; y[r] = v; }
  gsync(g);
}
/* lane d < nv: row d of the packed mass matrix times the shared-memory vector at vo (0 on the other lanes) */
RSB_DN real mulM_lane(int so, Grp g, int vo) { const real *s = RSB_SMEM + so;
  return g.lane < MDL.nv ? symv_row(s + MDL.o_M, g.lane, RSB_SMEM + vo, MDL.nv) : 0.0f;
}
/* lane d < nv: (J^T force)[d] */
RSB_DN real mulJT_lane(int so, Grp g, int nefc) { const real *s = RSB_SMEM + so;
  return g.lane < MDL.nv ? sdot_strided(s + MDL.o_J + g.lane, MDL.ldj, s + MDL.o_eforce, nefc) : 0.0f;
}
/* pass A of newton_hessian for a compile-time number of dofs (NV = 0: run-time nv with per-term predicates) */
template <int NV> RSB_D void hess_rows(const real *J, const real *ew, int ldj, int nv, int nefc, int j, real (&h)[RSB_LANES]) {
#pragma unroll 2
  for (int r = 0; r < nefc; r++) {
    const real *Jr = J + r * ldj; const real t = fmaxf(ew[r], 0.0f) * Jr[j];
#pragma unroll
    for (int i = 0; i < RSB_LANES; i++) if (NV ? (i < NV) : (i < nv)) h[i] += Jr[i] * t;
  }
}
RSB_DN void newton_hessian(int so, Grp g, int nefc) { real *s = RSB_SMEM + so;
  const real *M = s + MDL.o_M, *J = s + MDL.o_J, *ew = s + MDL.o_ew, *con = s + MDL.o_con, *jar = s + MDL.o_ejar; real *H = s + MDL.o_L;
  const int ldj = MDL.ldj, nv = MDL.nv, j = g.lane < nv ? g.lane : 0;
  real h[RSB_LANES];
#pragma unroll
  for (int i = 0; i < RSB_LANES; i++) h[i] = 0.0f;
  /* pass A, branch-free: every row with its clamped weight (inactive rows and cone rows carry w <= 0 -> 0); two rows per trip so that
     the loads of one row overlap the FMAs of the other */
  if (nv == 15) hess_rows<15>(J, ew, ldj, nv, nefc, j, h);          /* Lift-Panda / Lift-Sawyer (7 + 2 + 6 dofs) */
  else if (nv == 11) hess_rows<11>(J, ew, ldj, nv, nefc, j, h);     /* Door-Panda */
  else hess_rows<0>(J, ew, ldj, nv, nefc, j, h);
  /* pass B: contacts in the middle zone of their cone (ew < 0 on their rows) */
  const int ncon = ((const int *)(s + MDL.o_misc))[MISC_NCON];
#pragma unroll 1
  for (int c = 0; c < ncon; c++) {
    const real *cr = con + c * RSB_CONW; const int *ci = (const int *)cr; const int r = ci[CON_ADR];
    if (r >= 0 && ew[r] < 0) { const real *Jr = J + r * ldj;
      /* cone Hessian in scaled coordinates U = sc * x, U = (N, t), T = |t|, that = t / T:  Hu = Dm u u^T - kap (I_t - that that^T) with
         u = (1, -mu that), kap = Dm mu (N - mu T) / T.  Applied to column j without forming the dim x dim block:
         y = sc * J_c[:, j],  (Hu y)_a = Dm u_a (u.y) - kap (y_a - that_a (that.y)) [a >= 1],  t_a = sc_a (Hu y)_a. */
      const int dim = CON_DIM_OF(ci);
      const real *fr = MDL.pair_friction + 5 * ci[CON_PAIR]; const real mu = cr[CON_MU], D = (s + MDL.o_eD)[r];
      real sc[RSB_MAXDIM], u[RSB_MAXDIM], y[RSB_MAXDIM], t[RSB_MAXDIM]; real T2 = 0, N = 0;
#pragma unroll
      for (int a = 0; a < RSB_MAXDIM; a++) { sc[a] = 0; u[a] = 0; y[a] = 0;
        if (a < dim) { sc[a] = (a == 0) ? mu : fr[a - 1]; const real Ua = jar[r + a] * sc[a]; y[a] = sc[a] * Jr[a * ldj + j];
          if (a == 0) N = Ua; else { u[a] = Ua; T2 += Ua * Ua; } } }
      const real invT = rsb_rsqrt(T2), T = T2 * invT, Dm = D / (mu * mu * (1 + mu * mu)), kap = Dm * mu * (N - mu * T) * invT;     /* middle zone: T > 0 */
      real uy = y[0], ty = 0;
#pragma unroll
      for (int a = 1; a < RSB_MAXDIM; a++) { u[a] *= invT; ty += u[a] * y[a]; }      /* u[a >= 1] holds that_a for now */
      uy -= mu * ty;
      t[0] = sc[0] * Dm * uy;
#pragma unroll
      for (int a = 1; a < RSB_MAXDIM; a++) t[a] = sc[a] * (-Dm * mu * u[a] * uy - kap * (y[a] - u[a] * ty));
#pragma unroll
      for (int a = 0; a < RSB_MAXDIM; a++) if (a < dim) { const real *Ja = Jr + a * ldj;
#pragma unroll
        for (int i = 0; i < RSB_LANES; i++) if (i < nv) h[i] += Ja[i] * t[a]; }
    }
  }
  if (g.lane < nv) {
#pragma unroll
    for (int i = 0; i < RSB_LANES; i++) if (i < nv && i >= j) { const int e = tri_off(i) + j; H[e] = M[e] + h[i]; }
  }
  gsync(g);
}

/* ---- exact line search with per-item polynomial data in REGISTERS.
   Along the search line every row residual is x(alpha) = jar + alpha Jv.  What the slope / curvature of the constraint cost need per
   work item (scalar row or contact) are a handful of numbers that do not depend on alpha: they are loaded ONCE per Newton iteration
   (two items per lane; items beyond 2 x lanes are re-loaded on every evaluation), so that one evaluation is ~30 flops per lane plus one
   fused reduction instead of a pass over shared memory (mj's line search precomputes its `quad` coefficients for the same reason).
     friction loss : x0 dx D fl rf            unilateral (limit, condim 1) : x0 dx D
     elliptic cone : N0 dN tt tdt dtdt (scaled coordinates U = sc x: N = N0 + a dN, T^2 = tt + 2 a tdt + a^2 dtdt), mu, Dm,
                     B1 B2 (bottom zone: slope B1 + a B2, curvature B2) */
struct LsItem { int type; real a0, a1, a2, a3, a4, a5, a6, a7, a8; };
RSB_D void ls_load(int so, int it, int nscalar, int nitem, LsItem &I) { const real *s = RSB_SMEM + so;
  const real *con = s + MDL.o_con, *eD = s + MDL.o_eD, *jar = s + MDL.o_ejar, *Jv = s + MDL.o_eJv; const int *etid = (const int *)(s + MDL.o_etype);
  I.type = 0; I.a0 = I.a1 = I.a2 = I.a3 = I.a4 = I.a5 = I.a6 = I.a7 = I.a8 = 0;
  if (it >= nitem) return;
  const int r = it < nscalar ? it : ((const int *)(con + (it - nscalar) * RSB_CONW))[CON_ADR];
  if (r < 0) return;
  const int wd = etid[r], type = ET_TYPE(wd); const real D = eD[r];
  I.a0 = jar[r]; I.a1 = Jv[r]; I.a2 = D;
  if (type == EFC_FRICTION) { const real fl = MDL.dof_floss[ET_ID(wd)]; I.type = 1; I.a3 = fl; I.a4 = fl / D; return; }
  const real *cr = con + ET_ID(wd) * RSB_CONW; const int dim = (type == EFC_LIMIT) ? 1 : CON_DIM_OF((const int *)cr);
  if (dim == 1) { I.type = 2; return; }
  const real *fr = MDL.pair_friction + 5 * ((const int *)cr)[CON_PAIR]; const real mu = cr[CON_MU];
  real tt = 0, tdt = 0, dtdt = 0, B1 = D * I.a0 * I.a1, B2 = D * I.a1 * I.a1;
#pragma unroll
  for (int j = 1; j < RSB_MAXDIM; j++) if (j < dim) {
    const real sc = fr[j - 1], xj = jar[r + j], dxj = Jv[r + j], Dj = eD[r + j], U = xj * sc, dU = dxj * sc;
    tt += U * U; tdt += U * dU; dtdt += dU * dU; B1 += Dj * xj * dxj; B2 += Dj * dxj * dxj; }
  I.type = 3; I.a0 *= mu; I.a1 *= mu;                               /* N0, dN */
  I.a2 = tt; I.a3 = tdt; I.a4 = dtdt; I.a5 = mu; I.a6 = D / (mu * mu * (1 + mu * mu)); I.a7 = B1; I.a8 = B2;
}
RSB_D void ls_eval(const LsItem &I, real alpha, real &d1, real &d2) {
  if (I.type == 0) return;
  if (I.type == 1) { const real x = I.a0 + alpha * I.a1;
    if (x <= -I.a4) d1 -= I.a3 * I.a1; else if (x >= I.a4) d1 += I.a3 * I.a1; else { d1 += I.a2 * x * I.a1; d2 += I.a2 * I.a1 * I.a1; }
    return; }
  if (I.type == 2) { const real x = I.a0 + alpha * I.a1; if (x < 0) { d1 += I.a2 * x * I.a1; d2 += I.a2 * I.a1 * I.a1; } return; }
  const real N = I.a0 + alpha * I.a1, mu = I.a5; real T2 = I.a2 + alpha * (2 * I.a3 + alpha * I.a4); T2 = fmaxf(T2, 0.0f);
  const real invT = T2 > 0 ? rsb_rsqrt(T2) : 0.0f, T = T2 * invT;
  if (N >= mu * T || (T <= 0 && N >= 0)) return;                    /* top zone */
  if (mu * N + T <= 0 || (T <= 0 && N < 0)) { d1 += I.a7 + alpha * I.a8; d2 += I.a8; return; }       /* bottom zone */
  const real sdt = (I.a3 + alpha * I.a4) * invT, q = I.a1 - mu * sdt, NmT = N - mu * T, Dm = I.a6;
  d1 += Dm * NmT * q; d2 += Dm * q * q - Dm * mu * NmT * (I.a4 - sdt * sdt) * invT;
}
/* returns the step (0: the Newton decrement is below tolerance -- converged -- or the group is not active).  gq1, gq2: slope at 0 and
   curvature of the Gauss term along the line; sg = search.grad: slope of the total cost at 0 (its negative is the curvature there). */
RSB_DN real newton_linesearch(int so, Grp g, real gq1, real gq2, real sg, real scale, bool active) { const real *s = RSB_SMEM + so;
  const int *misc = (const int *)(s + MDL.o_misc); const int nscalar = misc[MISC_NLIMROW], nitem = nscalar + misc[MISC_NCON];
  LsItem I0, I1; ls_load(so, g.lane, nscalar, nitem, I0); ls_load(so, g.lane + RSB_LANES, nscalar, nitem, I1);
  real lo = 0, hi = -1, alpha = 0; const real d1_0 = fabsf(sg);
  bool ls = active && (sg < -1e-10f / scale);                     /* else: Newton decrement below tolerance (or fp32 noise): converged */
  if (ls) alpha = 1.0f;                                           /* the Newton step: slope and curvature at 0 are sg and -sg */
#pragma unroll 1
  for (int lit = 1; lit < MDL.ls_iters; lit++) {
    if (!sany(ls)) break;
    real e1 = 0, e2 = 0; ls_eval(I0, alpha, e1, e2); ls_eval(I1, alpha, e1, e2);
#pragma unroll 1
    for (int it = g.lane + 2 * RSB_LANES; it < nitem; it += RSB_LANES) { LsItem J; ls_load(so, it, nscalar, nitem, J); ls_eval(J, alpha, e1, e2); }
    gsum2(g, e1, e2); PROF_COUNT(14);
    const real d1 = gq1 + alpha * gq2 + e1, d2 = gq2 + e2;
    if (ls) {
#ifdef RSB_EMU_TRACE
      if (g.lane == 0) printf("    ls %d alpha %.6g d1 %.3e d2 %.3e (d1_0 %.3e)\n", lit, alpha, d1, d2, d1_0);
#endif
      if (fabsf(d1) <= MDL.ls_tol * d1_0 + 1e-30f) ls = false;
      else {
        if (d1 < 0) lo = alpha; else hi = alpha;
        if (hi >= 0 && hi - lo <= 1e-4f * hi) ls = false;           /* bracket at fp32 resolution of the derivative: the sign of d1 is noise from here on */
        real an = d2 > RSB_MINVAL ? alpha - d1 / d2 : alpha;
        if (hi >= 0 && (an <= lo || an >= hi)) an = 0.5f * (lo + hi);
        else if (hi < 0 && an <= lo) an = 2 * lo;
        if (an == alpha) ls = false; else if (ls) alpha = an;
      }
    }
  }
  return alpha;
}

/* residuals of TWO candidate accelerations in one pass over J (warm-start selection): y1 = J x1 - aref, y2 = J x2 - aref */
RSB_DN void efc_mulJ2(int so, Grp g, int nefc, int x1o, int y1o, int x2o, int y2o) { real *s = RSB_SMEM + so;
  const real *J = s + MDL.o_J, *x1 = RSB_SMEM + x1o, *x2 = RSB_SMEM + x2o, *earef = s + MDL.o_earef; real *y1 = RSB_SMEM + y1o, *y2 = RSB_SMEM + y2o;
  const int nv = MDL.nv, ldj = MDL.ldj;
#pragma unroll 1
  for (int r = g.lane; r < nefc; r += RSB_LANES) {
    const real *Jr = J + r * ldj; real a0 = 0, a1 = 0, b0 = 0, b1 = 0; int k = 0;
    for (; k + 2 <= nv; k += 2) { const real j0 = Jr[k], j1 = Jr[k + 1], p0 = x1[k], p1 = x1[k + 1], q0 = x2[k], q1 = x2[k + 1]; a0 += j0 * p0; a1 += j1 * p1; b0 += j0 * q0; b1 += j1 * q1; }
    if (k < nv) { const real j0 = Jr[k]; a0 += j0 * x1[k]; b0 += j0 * x2[k]; }
    const real ar = earef[r]; y1[r] = (a0 + a1) - ar; y2[r] = (b0 + b1) - ar;
  }
  gsync(g);
}

RSB_DN void st_solve(int so, Grp g) { real *s = RSB_SMEM + so;
  int *misc = (int *)(s + MDL.o_misc); const int nefc = misc[MISC_NEFC], nv = MDL.nv; const bool dl = g.lane < nv; const int d = g.lane;
  real *qacc = s + MDL.o_qacc, *qas = s + MDL.o_qacc_smooth, *warm = s + MDL.o_warm, *qfc = s + MDL.o_qfc, *grad = s + MDL.o_grad, *search = s + MDL.o_search, *tmpv = s + MDL.o_tmpv;
  if (!sany(nefc != 0)) {                             /* no group of this warp has constraint rows */
    if (dl) { qacc[d] = qas[d]; qfc[d] = 0; }
    if (g.lane == 0) misc[MISC_ITER] = 0;
    gsync(g); return;
  }
  /* (an unconstrained group next to a constrained one runs the loop with nefc = 0: it starts from qacc_smooth, its gradient is exactly
     0 and it is inactive from iteration 0 on, with qfc = 0 -- the same result as the early exit) */
  /* warm start: the cheaper of qacc_warmstart and qacc_smooth */
  PROF_LOCAL;
  /* cost(a) = (a - a_s)^T M (a - a_s) / 2 + s(J a - aref): both candidates' residuals from one pass over J (warm -> jar, smooth -> the Jv
     array, free until the first search direction); the Gauss term of qacc_smooth is zero */
  real *jar = s + MDL.o_ejar, *jas = s + MDL.o_eJv;
  efc_mulJ2(so, g, nefc, so + MDL.o_warm, so + MDL.o_ejar, so + MDL.o_qacc_smooth, so + MDL.o_eJv);
  if (dl) tmpv[d] = warm[d] - qas[d];
  gsync(g);
  const real mw = mulM_lane(so, g, so + MDL.o_tmpv);                /* lane d: (M (warm - qacc_smooth))_d */
  real cw = efc_eval(so, g, false, so + MDL.o_ejar) + (dl ? 0.5f * tmpv[d] * mw : 0.0f), cs0 = efc_eval(so, g, false, so + MDL.o_eJv);
  gsum2(g, cw, cs0);
  const bool use_warm = cw < cs0;
  if (dl) qacc[d] = use_warm ? warm[d] : qas[d];
  if (!use_warm) for (int r = g.lane; r < nefc; r += RSB_LANES) jar[r] = jas[r];
  gsync(g);
  /* As in mj_solNewton the residuals jar = J qacc - aref and Ma = M (qacc - qacc_smooth) are carried along the iterations and updated with
     the step (jar += alpha Jv, Ma += alpha Mv) instead of being recomputed from qacc; lane d keeps Ma_d and Mv_d in registers. */
  real ma = use_warm ? mw : 0.0f;
  const real scale = 1.0f / (MDL.meaninertia * (real)(nv > 1 ? nv : 1));
  PROF(10);                                            /* warm-start selection */
  /* `active` is uniform within a group; every branch that encloses a shuffle tests a warp vote, so the groups of a warp stay converged.
     A finished group keeps executing the body (recomputing identical residuals/forces for its unchanged qacc) until its neighbour is done. */
  int iter = 0; bool active = true, last = false;                  /* last: the previous update improved the cost by less than the tolerance */
#pragma unroll 1
  for (int it = 0; it <= MDL.solver_iters; it++) {
    /* forces and Hessian weights at the current residuals; gradient = M (qacc - qacc_smooth) - J^T f (lane per dof) */
    efc_eval(so, g, true, so + MDL.o_ejar);
    gsync(g);
    const real f = mulJT_lane(so, g, nefc), a = ma - f;
    if (dl) { grad[d] = a; qfc[d] = f; }
    const real gn = gsum(g, a * a); PROF(11);            /* residual, forces, gradient */
#ifdef RSB_EMU_TRACE
    if (g.lane == 0) printf("  it %d scaled|grad| %.3e\n", iter, scale * sqrtf(gn));
#endif
    if (active && (last || it == MDL.solver_iters || scale * sqrtf(gn) < MDL.solver_tol)) active = false;      /* mj_solNewton: stop on small gradient OR small improvement (or the iteration limit) */
    if (!sany(active)) break;
    newton_hessian(so, g, nefc); PROF(12); PROF_COUNT(12);      /* (profile build: iteration count in the high bits) */
    if (dl) search[d] = -grad[d];
    gsync(g);
    chol_factor_solve(so + MDL.o_L, nv, 0, so + MDL.o_search, g); PROF(13);
    /* directional quantities */
    real gq1 = 0, gq2 = 0, sg = 0;
    const real mv = mulM_lane(so, g, so + MDL.o_search);
    if (dl) { gq2 = search[d] * mv; gq1 = search[d] * ma; sg = search[d] * grad[d]; }
    efc_mulJ(so, g, nefc, so + MDL.o_search, so + MDL.o_eJv, 0);
    gsum3(g, gq1, gq2, sg);                                     /* gq1 = s.(M a - M a_s): slope of the Gauss term at alpha = 0; sg: slope of the total cost */
    /* exact line search on the convex 1-D cost: safeguarded Newton on its derivative, first trial step 1 (newton_linesearch) */
    const real d1_0 = fabsf(sg), alpha = newton_linesearch(so, g, gq1, gq2, sg, scale, active);
    PROF(14);                                          /* directional quantities + line search */
    if (active && alpha == 0) active = false;
    if (active) { if (dl) qacc[d] += alpha * search[d]; ma += alpha * mv; iter++;
      for (int r = g.lane; r < nefc; r += RSB_LANES) jar[r] += alpha * jas[r];        /* jas = o_eJv holds J search here */
      last = scale * 0.5f * alpha * d1_0 < MDL.solver_tol; }       /* cost decrease of an exact line search on a (locally) quadratic cost: alpha |d1(0)| / 2 */
#ifdef RSB_EMU_TRACE
    if (g.lane == 0) printf("    scaled improvement %.3e (tolerance %.3e, scale %.3e)\n", scale * 0.5f * alpha * d1_0, MDL.solver_tol, scale);
#endif
    gsync(g);
    if (!sany(active)) break;
  }
  /* The loop evaluates forces at its top, so on every exit the forces correspond to the final qacc: an update is always followed by
     another pass of the top part (the loop runs to it == solver_iters, where a still-active group is stopped before its next update).
     Those forces come from the residuals carried along the iterations; the constraint force handed to the integrator is evaluated once
     more from residuals recomputed at the final qacc, so that the round-off of the incremental updates does not reach the state. */
  if (sany(iter > 0)) {
    efc_mulJ(so, g, nefc, so + MDL.o_qacc, so + MDL.o_ejar, 1);
    efc_eval(so, g, true, so + MDL.o_ejar);
    gsync(g);
    const real f = mulJT_lane(so, g, nefc); if (dl) qfc[d] = f;
  }
  if (g.lane == 0) { misc[MISC_ITER] = iter; misc[MISC_ITERSUM] += iter; }
  gsync(g);
}

/* ================================================================== A.3.8 semi-implicit Euler with implicit joint damping */
RSB_DN void st_euler(int so, Grp g) { real *s = RSB_SMEM + so;
  real *qpos = s + MDL.o_qpos, *qvel = s + MDL.o_qvel, *warm = s + MDL.o_warm, *L = s + MDL.o_L, *tmpv = s + MDL.o_tmpv;
  const real *M = s + MDL.o_M, *qacc = s + MDL.o_qacc, *smooth = s + MDL.o_smooth, *qfc = s + MDL.o_qfc; const real h = MDL.timestep;
  if (MDL.any_damping) {
    for (int e = g.lane; e < MDL.ntri; e += RSB_LANES) { const int ij = MDL.tri_ij[e], r = ij >> 8, c = ij & 255; L[e] = M[e] + ((r == c) ? h * MDL.dof_damping[r] : 0.0f); }
    for (int d = g.lane; d < MDL.nv; d += RSB_LANES) tmpv[d] = smooth[d] + qfc[d];
    gsync(g);
    chol_factor_solve(so + MDL.o_L, MDL.nv, 0, so + MDL.o_tmpv, g);
  } else { for (int d = g.lane; d < MDL.nv; d += RSB_LANES) tmpv[d] = qacc[d]; gsync(g); }
  for (int d = g.lane; d < MDL.nv; d += RSB_LANES) { qvel[d] += h * tmpv[d]; warm[d] = qacc[d]; }
  gsync(g);
  for (int j = g.lane; j < MDL.njnt; j += RSB_LANES) {
    int qa = MDL.jnt_qadr[j], d = MDL.jnt_dadr[j];
    if (MDL.jnt_type[j] == RSB_JNT_FREE) {
      for (int k = 0; k < 3; k++) qpos[qa + k] += h * qvel[d + k];
      real w[3] = {qvel[d + 3], qvel[d + 4], qvel[d + 5]}; real wn = sqrtf(dot3(w, w)), ang = wn * h;
      if (ang > 0) { real sn, c; rsb_sincos(0.5f * ang, &sn, &c); real f = sn / wn; real dq[4] = {c, f * w[0], f * w[1], f * w[2]}, qn[4];
        quatmul(qn, qpos + qa + 3, dq); quatnorm(qn); qpos[qa + 3] = qn[0]; qpos[qa + 4] = qn[1]; qpos[qa + 5] = qn[2]; qpos[qa + 6] = qn[3]; }
    } else qpos[qa] += h * qvel[d];
  }
  gsync(g);
}

/* ================================================================== one physics substep, the control step, reward, observation */
/* Stage sequence of one physics substep.  RSB_CTA_SYNC() between stages keeps all warps of the CTA in the same stage
   (they then share instruction-cache lines: the whole step is far larger than the I-cache); it carries no data dependency. */
#define STAGE_SYNC(k) do { PROF(k); RSB_CTA_SYNC(k); PROF(15); } while (0)
RSB_D void substep(int so, Grp g, bool policy_step, long long &pt_) {
  /* order matters for the shared-memory overlays (rsb_devmodel.h): everything that reads the kinematics/dynamics temporaries runs
     before the constraint rows are built, because the Jacobian overlays them */
  st_kinematics(so, g); STAGE_SYNC(0); st_inertia(so, g); st_crb(so, g); STAGE_SYNC(1); st_collision(so, g); STAGE_SYNC(2);
  st_bias(so, g); STAGE_SYNC(3);
  if (policy_step) ctrl_set_goal(so, g);
  ctrl_run(so, g); STAGE_SYNC(4);
  st_actuation(so, g); STAGE_SYNC(5); st_constraint(so, g); STAGE_SYNC(6); st_smooth_acc(so, g); PROF(9); st_solve(so, g); STAGE_SYNC(7); st_euler(so, g); STAGE_SYNC(8);
}

RSB_D bool geom_in(const int *set, int n, int gm) { for (int i = 0; i < n; i++) if (set[i] == gm) return true; return false; }
RSB_D bool check_grasp_range(int so, int ri, int lo, int hi) { const real *s = RSB_SMEM + so;       /* both fingers of robot ri touch a geom lo..hi of the object */
  const DevRobot &rb = MDL.robot[ri]; const real *con = s + MDL.o_con; int ncon = ((const int *)(s + MDL.o_misc))[MISC_NCON]; bool tl = false, tr = false;
  for (int c = 0; c < ncon; c++) {
    const int *ci = (const int *)(con + c * RSB_CONW); int g1 = CON_G1_OF(ci), g2 = CON_G2_OF(ci); const bool o1 = g1 >= lo && g1 <= hi, o2 = g2 >= lo && g2 <= hi;
    if ((geom_in(rb.lfg, rb.nlfg, g1) && o2) || (geom_in(rb.lfg, rb.nlfg, g2) && o1)) tl = true;
    if ((geom_in(rb.rfg, rb.nrfg, g1) && o2) || (geom_in(rb.rfg, rb.nrfg, g2) && o1)) tr = true;
  }
  return tl && tr;
}
RSB_D bool check_grasp(int so, int ri, int obj_geom) { return check_grasp_range(so, ri, obj_geom, obj_geom); }

/* TwoArmPegInHole._compute_orientation (see the oracle's peg_hole_orientation): out = {t, d, cos} */
RSB_DN void peg_hole_orientation(int so, real *out) { const real *s = RSB_SMEM + so;
  const real *hp = s + MDL.o_xpos + 3 * MDL.obj_body[0], *Rh = s + MDL.o_xmat + 9 * MDL.obj_body[0], *pp = s + MDL.o_xpos + 3 * MDL.obj_body[1], *Rp = s + MDL.o_xmat + 9 * MDL.obj_body[1];
  real v[3] = {Rp[2], Rp[5], Rp[8]}, n[3] = {Rh[2], Rh[5], Rh[8]}, cp[3], pc[3], x[3];
  for (int k = 0; k < 3; k++) { cp[k] = hp[k] + MDL.task_par[0] * Rh[3 * k] - pp[k]; pc[k] = -cp[k]; }
  cross3(x, v, pc);
  const real vv = dot3(v, v);
  out[0] = dot3(cp, v) / vv; out[1] = sqrtf(dot3(x, x) / vv); out[2] = fabsf(dot3(n, v)) / sqrtf(dot3(n, n) * vv);
}

/* A.6 staged task rewards (evaluated by every lane identically; cheap) */
RSB_DN real task_reward(int so) { const real *s = RSB_SMEM + so;
  const real *xpos = s + MDL.o_xpos, *sxpos = s + MDL.o_sxpos, *qpos = s + MDL.o_qpos; real r = 0;
  const real *eef = sxpos + 3 * MDL.robot[0].eef_site;
  if (MDL.task_id == RSB_TASK_LIFT) {
    const real *cube = xpos + 3 * MDL.obj_body[0];
    if (cube[2] > MDL.table_height + 0.04f) r = 2.25f;
    else if (MDL.reward_shaping) {
      real d[3] = {eef[0] - cube[0], eef[1] - cube[1], eef[2] - cube[2]}; r += 1 - tanhf(10.0f * sqrtf(dot3(d, d)));
      if (check_grasp(so, 0, MDL.obj_geom[0])) r += 0.25f;
    }
    return r * MDL.reward_scale / 2.25f;
  }
  if (MDL.task_id == RSB_TASK_STACK) {
    const real *A = xpos + 3 * MDL.obj_body[0], *B = xpos + 3 * MDL.obj_body[1];
    real d[3] = {eef[0] - A[0], eef[1] - A[1], eef[2] - A[2]}; real dist = sqrtf(dot3(d, d));
    bool grasp = check_grasp(so, 0, MDL.obj_geom[0]);
    real r_reach = (1 - tanhf(10.0f * dist)) * 0.25f + (grasp ? 0.25f : 0.0f);
    bool lifted = A[2] > MDL.table_height + 0.04f; real r_lift = lifted ? 1.0f : 0.0f;
    if (lifted) { real hd = sqrtf((A[0] - B[0]) * (A[0] - B[0]) + (A[1] - B[1]) * (A[1] - B[1])); r_lift += 0.5f * (1 - tanhf(hd)); }
    bool touch = false; const real *con = s + MDL.o_con; int ncon = ((const int *)(s + MDL.o_misc))[MISC_NCON];
    for (int c = 0; c < ncon; c++) { const int *ci = (const int *)(con + c * RSB_CONW); int g1 = CON_G1_OF(ci), g2 = CON_G2_OF(ci);
      if ((g1 == MDL.obj_geom[0] && g2 == MDL.obj_geom[1]) || (g2 == MDL.obj_geom[0] && g1 == MDL.obj_geom[1])) touch = true; }
    real r_stack = (!grasp && r_lift > 0 && touch) ? 2.0f : 0.0f;
    if (MDL.reward_shaping) r = fmaxf(r_reach, fmaxf(r_lift, r_stack)); else r = r_stack > 0 ? 2.0f : 0.0f;
    return r * MDL.reward_scale / 2.0f;
  }
  if (MDL.task_id == RSB_TASK_DOOR) {
    real hinge = qpos[MDL.obj_qadr[0]], handle = qpos[MDL.obj_qadr[1]];
    if (hinge > 0.3f) r = 1.0f;
    else if (MDL.reward_shaping) {
      const real *hs = sxpos + 3 * MDL.obj_site[0]; real d[3] = {eef[0] - hs[0], eef[1] - hs[1], eef[2] - hs[2]};
      r += 0.25f * (1 - tanhf(10.0f * sqrtf(dot3(d, d))));
      r += fminf(0.25f * fabsf(handle / (0.5f * RSB_PI)), 0.25f);
    }
    return r * MDL.reward_scale;
  }
  if (MDL.task_id == RSB_TASK_TWOARMLIFT) {
    const real *pot = xpos + 3 * MDL.obj_body[0]; const real *Rp = s + MDL.o_xmat + 9 * MDL.obj_body[0];
    real gate = Rp[8] >= 0.8660254037844387f ? 1.0f : 0.0f, elev = pot[2] - MDL.obj_half[0][2] - MDL.table_height;
    if (elev > 0.10f) r = 3.0f * gate;
    else if (MDL.reward_shaping) {
      r += 10.0f * gate * clampf(elev - 0.05f, 0.0f, 0.15f);
      for (int ri = 0; ri < 2; ri++) {
        const real *ee = sxpos + 3 * MDL.robot[ri].eef_site, *hs = sxpos + 3 * MDL.obj_site[ri]; real d[3] = {ee[0] - hs[0], ee[1] - hs[1], ee[2] - hs[2]};
        if (check_grasp(so, ri, MDL.obj_geom[ri])) r += 0.25f; else r += 0.5f * (1 - tanhf(10.0f * sqrtf(dot3(d, d))));
      }
    }
    return r * MDL.reward_scale / 3.0f;
  }
  if (MDL.task_id == RSB_TASK_PICKPLACE) {          /* single-object PickPlace: success 1, else max(reach, grasp, lift, hover) -- see the oracle's task_reward */
    const real *obj = xpos + 3 * MDL.obj_body[0], *tp = MDL.task_par;
    real d[3] = {eef[0] - obj[0], eef[1] - obj[1], eef[2] - obj[2]}; const real reach = 1 - tanhf(10.0f * sqrtf(dot3(d, d)));
    const bool above = fabsf(obj[0] - tp[0]) < 0.25f * tp[3] && fabsf(obj[1] - tp[1]) < 0.25f * tp[4];
    const bool in_bin = above && obj[2] > tp[2] && obj[2] < tp[2] + 0.1f;
    if (in_bin && reach < 0.6f) r = 1.0f;
    else if (MDL.reward_shaping) {
      const real r_grasp = check_grasp(so, 0, MDL.obj_geom[0]) ? 0.35f : 0.0f; real r_lift = 0;
      if (r_grasp > 0) r_lift = 0.35f + (1 - tanhf(15.0f * fmaxf(tp[2] + tp[5] - obj[2], 0.0f))) * 0.15f;
      const real hx = obj[0] - tp[0], hy = obj[1] - tp[1];
      const real r_hover = (above ? 0.5f : r_lift) + (1 - tanhf(10.0f * sqrtf(hx * hx + hy * hy))) * 0.2f;
      r = fmaxf(fmaxf(0.1f * reach, r_grasp), fmaxf(r_lift, r_hover));
    }
    return r * MDL.reward_scale;
  }
  if (MDL.task_id == RSB_TASK_NUTASSEMBLY) {        /* single-object NutAssembly: success 1, else max(reach to the handle, grasp, lift, hover over the peg) -- see the oracle */
    const real *nut = xpos + 3 * MDL.obj_body[0], *handle = s + MDL.o_gxpos + 3 * MDL.obj_geom[1], *tp = MDL.task_par;
    real d[3] = {eef[0] - nut[0], eef[1] - nut[1], eef[2] - nut[2]};
    const bool on_peg = fabsf(nut[0] - tp[0]) < 0.03f && fabsf(nut[1] - tp[1]) < 0.03f && nut[2] < tp[2] + 0.05f;
    if (on_peg && 1 - tanhf(10.0f * sqrtf(dot3(d, d))) < 0.6f) r = 1.0f;
    else if (MDL.reward_shaping) {
      real dh[3] = {eef[0] - handle[0], eef[1] - handle[1], eef[2] - handle[2]};
      const real r_reach = 0.1f * (1 - tanhf(10.0f * sqrtf(dot3(dh, dh)))), r_grasp = check_grasp_range(so, 0, MDL.obj_geom[0], MDL.obj_geom[1]) ? 0.35f : 0.0f; real r_lift = 0;
      if (r_grasp > 0) r_lift = 0.35f + (1 - tanhf(15.0f * fmaxf(tp[3] - nut[2], 0.0f))) * 0.15f;
      const real hx = nut[0] - tp[0], hy = nut[1] - tp[1];
      const real r_hover = r_lift + (1 - tanhf(10.0f * sqrtf(hx * hx + hy * hy))) * 0.2f;
      r = fmaxf(fmaxf(r_reach, r_grasp), fmaxf(r_lift, r_hover));
    }
    return r * MDL.reward_scale;
  }
  if (MDL.task_id == RSB_TASK_HANDOFF) {            /* staged handoff reward (stage values as logged by the committed runs) -- see the oracle's task_reward */
    const real *handle = s + MDL.o_gxpos + 3 * MDL.obj_geom[0], *e1 = sxpos + 3 * MDL.robot[1].eef_site;
    const bool g0 = check_grasp_range(so, 0, MDL.obj_geom[0], MDL.obj_geom[1]), g1 = check_grasp(so, 1, MDL.obj_geom[0]);
    const bool lifted = handle[2] - MDL.task_par[1] - MDL.table_height > MDL.task_par[0];
    if (MDL.reward_shaping) {
      if (lifted) {
        if (g1) r = g0 ? 1.5f : 2.0f;
        else { real d[3] = {handle[0] - e1[0], handle[1] - e1[1], handle[2] - e1[2]}; r = 1.0f + 0.25f * (1 - tanhf(sqrtf(dot3(d, d)))); }
      } else if (g0) r = 0.5f;
      else { real d[3] = {handle[0] - eef[0], handle[1] - eef[1], handle[2] - eef[2]}; r = 0.25f * (1 - tanhf(sqrtf(dot3(d, d)))); }
    } else r = (lifted && g1 && !g0) ? 2.0f : 0.0f;
    return r * MDL.reward_scale * 0.5f;
  }
  if (MDL.task_id == RSB_TASK_PEGINHOLE) {          /* success 1 + (reach, d, t, cos shaping terms), / 5 -- see the oracle's task_reward */
    real o3[3]; peg_hole_orientation(so, o3);
    if (o3[1] < 0.06f && o3[0] >= -0.12f && o3[0] <= 0.14f && o3[2] > 0.95f) r = 1.0f;
    if (MDL.reward_shaping) {
      const real *hp = xpos + 3 * MDL.obj_body[0], *pp = xpos + 3 * MDL.obj_body[1]; real dv[3] = {pp[0] - hp[0], pp[1] - hp[1], pp[2] - hp[2]};
      r += (1 - tanhf(sqrtf(dot3(dv, dv)))) + (1 - tanhf(o3[1])) + (1 - tanhf(fabsf(o3[0]))) + o3[2];
    } else r *= 5.0f;
    return r * MDL.reward_scale / 5.0f;
  }
  return 0;
}

/* PickPlace object-state (element i of 14): obj pos, quat (xyzw), then the object's pose in the gripper frame: R_eef^T (p_obj - p_eef), conj(q_eef) q_obj with w >= 0.
   Not inlined: its quaternion temporaries stay out of the step kernel's own stack frame. */
RSB_DN real obs_pickplace(int so, int i) { const real *s = RSB_SMEM + so;
  const real *xpos = s + MDL.o_xpos, *xquat = s + MDL.o_xquat, *eef = s + MDL.o_sxpos + 3 * MDL.robot[0].eef_site;
  const real *obj = xpos + 3 * MDL.obj_body[0], *qo = xquat + 4 * MDL.obj_body[0], *qe = xquat + 4 * MDL.robot[0].eef_body;
  if (i < 3) return obj[i];
  if (i < 7) { int k = i - 3; return qo[k == 3 ? 0 : k + 1]; }
  if (i < 10) { real Re[9], d[3] = {obj[0] - eef[0], obj[1] - eef[1], obj[2] - eef[2]}, rel[3]; quat2mat(Re, qe); matTvec3(rel, Re, d); return rel[i - 7]; }
  real qc[4] = {qe[0], -qe[1], -qe[2], -qe[3]}, qr[4]; quatmul(qr, qc, qo); quatnorm(qr);
  const real sg = qr[0] < 0 ? -1.0f : 1.0f; int k = i - 10; return sg * qr[k == 3 ? 0 : k + 1];
}

/* TwoArmPegInHole object-state (element i of 17): hole_pos, hole_quat, peg - hole, peg_quat, cos, t, d.  Not inlined, like obs_pickplace. */
RSB_DN real obs_peginhole(int so, int i) { const real *s = RSB_SMEM + so;
  const real *xpos = s + MDL.o_xpos, *xquat = s + MDL.o_xquat, *hp = xpos + 3 * MDL.obj_body[0], *pp = xpos + 3 * MDL.obj_body[1];
  if (i < 3) return hp[i];
  if (i < 7) { int k = i - 3; return xquat[4 * MDL.obj_body[0] + (k == 3 ? 0 : k + 1)]; }
  if (i < 10) return pp[i - 7] - hp[i - 7];
  if (i < 14) { int k = i - 10; return xquat[4 * MDL.obj_body[1] + (k == 3 ? 0 : k + 1)]; }
  real o3[3]; peg_hole_orientation(so, o3); return i == 14 ? o3[2] : (i == 15 ? o3[0] : o3[1]);
}

/* TwoArmHandoff object-state (element i of 22): hammer_pos, hammer_quat, handle_xpos, eef0, eef1, handle - eef0, handle - eef1.  Not inlined, like obs_pickplace. */
RSB_DN real obs_handoff(int so, int i) { const real *s = RSB_SMEM + so;
  const real *hp = s + MDL.o_xpos + 3 * MDL.obj_body[0], *handle = s + MDL.o_gxpos + 3 * MDL.obj_geom[0];
  const real *e0 = s + MDL.o_sxpos + 3 * MDL.robot[0].eef_site, *e1 = s + MDL.o_sxpos + 3 * MDL.robot[1].eef_site;
  if (i < 3) return hp[i];
  if (i < 7) { int k = i - 3; return s[MDL.o_xquat + 4 * MDL.obj_body[0] + (k == 3 ? 0 : k + 1)]; }
  if (i < 10) return handle[i - 7];
  if (i < 13) return e0[i - 10];
  if (i < 16) return e1[i - 13];
  if (i < 19) return handle[i - 16] - e0[i - 16];
  return handle[i - 19] - e1[i - 19];
}

/* A.1.4 observation vector, robosuite v1.0 order: per robot [sin q, cos q, qd, eef_pos, eef_quat(xyzw), grip q, grip qd], then object-state.
   Lane-parallel: lane i produces element i (and i+LANES, ...) and writes it straight to `obs` (global memory, coalesced). */
RSB_D real obs_element(int so, int i) { const real *s = RSB_SMEM + so;
  const real *qpos = s + MDL.o_qpos, *qvel = s + MDL.o_qvel, *xpos = s + MDL.o_xpos, *xquat = s + MDL.o_xquat, *sxpos = s + MDL.o_sxpos;
  for (int ri = 0; ri < MDL.nrobot; ri++) {
    const DevRobot &rb = MDL.robot[ri]; int n = 28 + 2 * rb.grip_ndof;
    if (i < n) {
      if (i < 7) return sinf(qpos[rb.arm_qadr[i]]);
      if (i < 14) return cosf(qpos[rb.arm_qadr[i - 7]]);
      if (i < 21) return qvel[rb.arm_dadr[i - 14]];
      if (i < 24) return sxpos[3 * rb.eef_site + i - 21];
      if (i < 28) { int k = i - 24; return xquat[4 * rb.eef_body + (k == 3 ? 0 : k + 1)]; }
      int k = i - 28; return k < rb.grip_ndof ? qpos[rb.grip_qadr[k]] : qvel[rb.grip_dadr[k - rb.grip_ndof]];
    }
    i -= n;
  }
  const real *eef = sxpos + 3 * MDL.robot[0].eef_site;
  if (MDL.task_id == RSB_TASK_LIFT) {
    const real *cube = xpos + 3 * MDL.obj_body[0];
    if (i < 3) return cube[i];
    if (i < 7) { int k = i - 3; return xquat[4 * MDL.obj_body[0] + (k == 3 ? 0 : k + 1)]; }
    return eef[i - 7] - cube[i - 7];
  }
  if (MDL.task_id == RSB_TASK_STACK) {
    const real *A = xpos + 3 * MDL.obj_body[0], *B = xpos + 3 * MDL.obj_body[1];
    if (i < 3) return A[i];
    if (i < 7) { int k = i - 3; return xquat[4 * MDL.obj_body[0] + (k == 3 ? 0 : k + 1)]; }
    if (i < 10) return B[i - 7];
    if (i < 14) { int k = i - 10; return xquat[4 * MDL.obj_body[1] + (k == 3 ? 0 : k + 1)]; }
    if (i < 17) return eef[i - 14] - A[i - 14];
    if (i < 20) return eef[i - 17] - B[i - 17];
    return A[i - 20] - B[i - 20];
  }
  if (MDL.task_id == RSB_TASK_DOOR) {
    const real *door = xpos + 3 * MDL.obj_body[0], *hs = sxpos + 3 * MDL.obj_site[0];
    if (i < 3) return door[i];
    if (i < 6) return hs[i - 3];
    if (i < 9) return door[i - 6] - eef[i - 6];
    if (i < 12) return hs[i - 9] - eef[i - 9];
    return qpos[MDL.obj_qadr[i - 12]];
  }
  if (MDL.task_id == RSB_TASK_TWOARMLIFT) {
    const real *pot = xpos + 3 * MDL.obj_body[0], *e0 = sxpos + 3 * MDL.robot[0].eef_site, *e1 = sxpos + 3 * MDL.robot[1].eef_site;
    const real *h0 = sxpos + 3 * MDL.obj_site[0], *h1 = sxpos + 3 * MDL.obj_site[1];
    if (i < 3) return pot[i];
    if (i < 7) { int k = i - 3; return xquat[4 * MDL.obj_body[0] + (k == 3 ? 0 : k + 1)]; }
    if (i < 10) return e0[i - 7];
    if (i < 13) return e1[i - 10];
    if (i < 16) return h0[i - 13];
    if (i < 19) return h1[i - 16];
    if (i < 22) return h0[i - 19] - e0[i - 19];
    return h1[i - 22] - e1[i - 22];
  }
  if (MDL.task_id == RSB_TASK_PICKPLACE || MDL.task_id == RSB_TASK_NUTASSEMBLY) return obs_pickplace(so, i);
  if (MDL.task_id == RSB_TASK_PEGINHOLE) return obs_peginhole(so, i);
  if (MDL.task_id == RSB_TASK_HANDOFF) return obs_handoff(so, i);
  return 0;
}

/* persistent state <-> shared memory (state record layout: DevModel::st_*) */
RSB_D void load_state(int so, const real *st, Grp g) { real *s = RSB_SMEM + so;
  for (int i = g.lane; i < MDL.nq; i += RSB_LANES) s[MDL.o_qpos + i] = st[MDL.st_qpos + i];
  for (int i = g.lane; i < MDL.nv; i += RSB_LANES) { s[MDL.o_qvel + i] = st[MDL.st_qvel + i]; s[MDL.o_warm + i] = st[MDL.st_warm + i]; }
  for (int ri = 0; ri < MDL.nrobot; ri++) for (int i = g.lane; i < MDL.cs_words; i += RSB_LANES) s[MDL.o_cs + ri * MDL.cs_words + i] = st[MDL.st_cs + ri * RSB_CS_WORDS + i];
  for (int i = g.lane; i < MDL.nu; i += RSB_LANES) s[MDL.o_ctrl + i] = 0;
  if (g.lane < 7) s[MDL.o_bpose + g.lane] = st[MDL.st_bpose + g.lane];
  gsync(g);
}
RSB_D void store_state(int so, real *st, Grp g) { const real *s = RSB_SMEM + so;
  for (int i = g.lane; i < MDL.nq; i += RSB_LANES) st[MDL.st_qpos + i] = s[MDL.o_qpos + i];
  for (int i = g.lane; i < MDL.nv; i += RSB_LANES) { st[MDL.st_qvel + i] = s[MDL.o_qvel + i]; st[MDL.st_warm + i] = s[MDL.o_warm + i]; }
  for (int ri = 0; ri < MDL.nrobot; ri++) for (int i = g.lane; i < MDL.cs_words; i += RSB_LANES) st[MDL.st_cs + ri * RSB_CS_WORDS + i] = s[MDL.o_cs + ri * MDL.cs_words + i];
  if (g.lane < 7) st[MDL.st_bpose + g.lane] = s[MDL.o_bpose + g.lane];
}

/* One control step of one env (robosuite MujocoEnv.step): 25 x (forward, controller, mj_step), then reward + observation.
   `st` = this env's state record, `action` = [act_dim], `obs` = [obs_dim] output row; `obs2` (may be null) = a second destination of the
   same observation row -- in ring mode `obs` is next_obs[slot] and `obs2` is observations[slot + N], the row the next control step's policy
   forward reads (rlkit's EnvReplayBuffer keeps both arrays).  Returns via *reward, *done; `iters` (may be null) receives the env's Newton
   iterations of this control step.  Stepping a finished episode leaves the state untouched, reports done = 2 and bumps counter 2 (the
   host raises ValueError). */
RSB_D void env_step(int so, Grp g, real *st, const real *action, real *obs, real *obs2, real *reward, unsigned char *done, unsigned int *iters, bool commit) { real *s = RSB_SMEM + so;
  int t = f2i(st[MDL.st_time]);
  const bool finished = (t >= MDL.horizon) && !MDL.ignore_done;
  load_state(so, st, g);
  for (int i = g.lane; i < MDL.act_dim; i += RSB_LANES) s[MDL.o_act + i] = action[i];
  if (g.lane == 0) { int *misc = (int *)(s + MDL.o_misc); misc[MISC_OVF] = 0; misc[MISC_ITERSUM] = 0; }
  gsync(g);
  PROF_DECL;
  for (int k = 0; k < MDL.substeps; k++) substep(so, g, k == 0, pt_);
  st_kinematics(so, g); RSB_CTA_SYNC(0); st_collision(so, g);   /* observations / reward read the post-step kinematics and contacts */
  if (!commit) return;                               /* padding warp of the last CTA: ran only to reach the barriers */
  if (finished) { if (g.lane == 0) { *done = 2; *reward = 0; count_event(2, 1); } return; }      /* state untouched; host raises ValueError */
  real r = task_reward(so);
  for (int i = g.lane; i < MDL.obs_dim; i += RSB_LANES) { const real v = obs_element(so, i); obs[i] = v; if (obs2) obs2[i] = v; }
  store_state(so, st, g);
  if (g.lane == 0) { t++; st[MDL.st_time] = i2f(t); *reward = r; *done = ((t >= MDL.horizon) && !MDL.ignore_done) ? 1 : 0;
    const int *misc = (const int *)(s + MDL.o_misc); const int ovf = misc[MISC_OVF];
    if (ovf & 1) count_event(0, 1);
    if (ovf & 2) count_event(1, 1);
    if (iters) *iters = (unsigned int)misc[MISC_ITERSUM]; }
}

/* robosuite MujocoEnv.reset (hard_reset = False): sim.reset, noisy arm init, object placement, new controller, forward, obs.
   Randomness: Philox keyed (seed, env_id), stream 0, counter = episode*8 + block -- identical to the oracle's orc_reset. */
RSB_D void env_reset(int so, Grp g, real *st, uint64_t seed, uint64_t env_id, real *obs, bool commit) { real *s = RSB_SMEM + so;
  int episode = f2i(st[MDL.st_episode]);
  real *qpos = s + MDL.o_qpos;
  for (int i = g.lane; i < MDL.nq; i += RSB_LANES) qpos[i] = MDL.qpos0[i];
  for (int i = g.lane; i < MDL.nv; i += RSB_LANES) { s[MDL.o_qvel + i] = 0; s[MDL.o_warm + i] = 0; }
  if (g.lane < 7) { int ob = MDL.override_body; s[MDL.o_bpose + g.lane] = ob < 0 ? (g.lane == 3 ? 1.0f : 0.0f) : (g.lane < 3 ? MDL.body_pos[3 * ob + g.lane] : MDL.body_quat[4 * ob + g.lane - 3]); }
  gsync(g);
  if (g.lane < MDL.nrobot) {
    const DevRobot &rb = MDL.robot[g.lane]; real z[8]; uint32_t r[4];
    for (int blk = 0; blk < 2; blk++) {
      rsb_philox(seed, env_id, 0, (uint32_t)(episode * 8 + (g.lane * 2 + blk)), r);
      box_muller(r[0], r[1], &z[4 * blk], &z[4 * blk + 1]); box_muller(r[2], r[3], &z[4 * blk + 2], &z[4 * blk + 3]);
    }
    for (int k = 0; k < RSB_ARM_DOF; k++) qpos[rb.arm_qadr[k]] = rb.init_qpos[k] + MDL.init_noise * z[k];
    for (int k = 0; k < rb.grip_ndof; k++) qpos[rb.grip_qadr[k]] = rb.grip_init[k];
  }
  if (g.lane == RSB_LANES - 1) {                      /* object placement: uniform xy + yaw, rejection on overlap (serial over objects) */
    for (int o = 0; o < RSB_MAX_OBJ; o++) {
      if ((MDL.obj_qadr[o] < 0 && MDL.place_body[o] < 0) || MDL.place_z[o] <= 0) continue;
      int qa = MDL.obj_qadr[o]; real x = 0, y = 0, yaw = 0; uint32_t r[4];
      for (int attempt = 0; attempt < 16; attempt++) {
        rsb_philox(seed, env_id, 0, (uint32_t)(episode * 8 + 4 + o) + 0x10000u * (uint32_t)attempt, r);
        double u0 = ((double)r[0] + 0.5) * (1.0 / 4294967296.0), u1 = ((double)r[1] + 0.5) * (1.0 / 4294967296.0), u2 = ((double)r[2] + 0.5) * (1.0 / 4294967296.0);
        x = (real)(MDL.place_x[o][0] + (MDL.place_x[o][1] - MDL.place_x[o][0]) * u0);
        y = (real)(MDL.place_y[o][0] + (MDL.place_y[o][1] - MDL.place_y[o][0]) * u1);
        yaw = (real)(MDL.place_yaw[o][0] + (MDL.place_yaw[o][1] - MDL.place_yaw[o][0]) * u2);
        bool ok = true;
        for (int p = 0; p < o; p++) if (MDL.obj_qadr[p] >= 0 && MDL.place_body[o] < 0 && MDL.place_z[p] > 0) {
          real dx = x + MDL.place_ref[0] - qpos[MDL.obj_qadr[p]], dy = y + MDL.place_ref[1] - qpos[MDL.obj_qadr[p] + 1];
          real rr = sqrtf(MDL.obj_half[o][0] * MDL.obj_half[o][0] + MDL.obj_half[o][1] * MDL.obj_half[o][1]) + sqrtf(MDL.obj_half[p][0] * MDL.obj_half[p][0] + MDL.obj_half[p][1] * MDL.obj_half[p][1]);
          if (dx * dx + dy * dy < rr * rr) ok = false;
        }
        if (ok) break;
      }
      real sn, c; rsb_sincos(0.5f * yaw, &sn, &c);
      real *dst = MDL.place_body[o] >= 0 ? s + MDL.o_bpose : qpos + qa;
      dst[0] = MDL.place_ref[0] + x; dst[1] = MDL.place_ref[1] + y; dst[2] = MDL.place_z[o];
      dst[3] = c; dst[4] = 0; dst[5] = 0; dst[6] = sn;
      if (MDL.task_id == RSB_TASK_HANDOFF) { dst[4] = (r[3] & 1u) ? -sn : sn; dst[6] = 0; }        /* laid down by a quarter turn about the world x axis, head towards robot 0 or robot 1 */
    }
  }
  gsync(g);
  st_kinematics(so, g);
  ctrl_reset(so, g);
  if (!commit) return;                                /* masked-off or padding group: ran only to keep the warp converged */
  for (int i = g.lane; i < MDL.obs_dim; i += RSB_LANES) obs[i] = obs_element(so, i);
  store_state(so, st, g);
  if (g.lane == 0) { st[MDL.st_time] = i2f(0); st[MDL.st_episode] = i2f(episode + 1); }
}

/* synthetic action stream shared with the oracle: a = tanh(N(0,1)) keyed (seed, env, step, dim), stream 1 */
RSB_D void random_action_block(uint64_t seed, uint64_t env_id, uint64_t step, int blk, int n, real *action) {
  uint32_t r[4]; rsb_philox(seed, env_id, 1, (uint32_t)(step * 4 + (uint64_t)blk), r);
  real z[4]; box_muller(r[0], r[1], &z[0], &z[1]); box_muller(r[2], r[3], &z[2], &z[3]);
  for (int k = 0; k < 4; k++) if (4 * blk + k < n) action[4 * blk + k] = tanhf(z[k]);
}

/* debug record (floats) for parity tests; layout documented in rsb_devmodel.h / tests/emu */
RSB_D void dump_debug(int so, Grp g, real *out) { const real *s = RSB_SMEM + so;
  const int *misc = (const int *)(s + MDL.o_misc); int nv = MDL.nv, nc = MDL.ncon_max, ne = MDL.nefc_max;
  if (g.lane == 0) { out[0] = (real)misc[MISC_NCON]; out[1] = (real)misc[MISC_NEFC]; out[2] = (real)misc[MISC_ITER]; }
  real *o = out + 8;
  for (int i = g.lane; i < nv * nv; i += RSB_LANES) o[i] = msym(s + MDL.o_M, i / nv, i % nv);
  o += nv * nv;
  const int vecs[8] = {MDL.o_bias, MDL.o_passive, MDL.o_actuator, MDL.o_qacc_smooth, MDL.o_qacc, MDL.o_qfc, MDL.o_smooth, MDL.o_warm};
  for (int k = 0; k < 8; k++) for (int i = g.lane; i < nv; i += RSB_LANES) o[k * nv + i] = s[vecs[k] + i];
  o += 8 * nv;
  for (int i = g.lane; i < 14; i += RSB_LANES) o[i] = i < 7 * MDL.nrobot ? s[MDL.o_tau + i] : 0.0f;
  o += 14;
  for (int i = g.lane; i < nc * 16; i += RSB_LANES) { int c = i / 16, k = i % 16; const real *cr = s + MDL.o_con + c * RSB_CONW; const int *ci = (const int *)cr;
    real v = 0; if (c < misc[MISC_NCON]) { real fr[9] = {cr[3], cr[4], cr[5], 0, 0, 0, 0, 0, 0}; make_frame(fr);
      if (k < 3) v = cr[k]; else if (k < 12) v = fr[k - 3]; else if (k == 12) v = cr[CON_DIST]; else if (k == 13) v = (real)CON_G1_OF(ci); else if (k == 14) v = (real)CON_G2_OF(ci); else v = cr[CON_MU]; } o[i] = v; }
  o += nc * 16;
  const int ev[5] = {MDL.o_earef, MDL.o_eD, MDL.o_eforce, MDL.o_earef, MDL.o_ejar};    /* slots 1 and 3 used to be R and pos: now D and aref again (temporaries are overlaid) */
  for (int k = 0; k < 5; k++) for (int i = g.lane; i < ne; i += RSB_LANES) o[k * ne + i] = i < misc[MISC_NEFC] ? s[ev[k] + i] : 0;
  for (int i = g.lane; i < ne; i += RSB_LANES) o[5 * ne + i] = i < misc[MISC_NEFC] ? (real)ET_TYPE(((const int *)(s + MDL.o_etype))[i]) : -1;
  o += 6 * ne;
  for (int i = g.lane; i < ne * nv; i += RSB_LANES) { int r = i / nv, d = i % nv; o[i] = r < misc[MISC_NEFC] ? s[MDL.o_J + r * MDL.ldj + d] : 0; }
}

#endif /* RSB_DEV_H */
