/*
 * rsb_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C, double-precision, single-env restatement of the hot path the reference reaches
 * through `env.step` (reference call sites: util/rlkit_custom.py:438 `env.step(a)`,
 * util/rlkit_utils.py:49-59 `suite.make` + `GymWrapper`): robosuite's 25-substep control step
 * (controller + MuJoCo mj_step), the task reward and the observation vector.
 *
 * PARITY UNPINNED: the algorithm lives in un-vendored third-party packages that are absent
 * from /root/reference and from this image -- robosuite>=1.0.1 (requirements.txt:1),
 * mujoco-py>=2.0.2.9 -> MuJoCo 2.0 (requirements.txt:4, Dockerfile:47).  This file restates
 * their published algorithms as summarised in SURVEY.md Appendix A (each function cites the
 * paragraph it follows); it is validated by physics invariants (tests/test_oracle_*.py), not
 * by outputs of the real MuJoCo.  The one tie to the real simulator is statistical: the
 * reference's committed 2020 policies (trained against robosuite + MuJoCo) are rolled out in
 * the CUDA env that this oracle checks, and the Panda families reach 0.78-1.02 of their logged
 * returns with equal best episodes (profiles/r2_policy_transfer_all.txt; the
 * OSC orientation rule, the JOINT_VELOCITY law and several asset choices were selected by
 * that transfer, COMPAT.md).  Numerically the parity stays unpinned.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  It deliberately uses different formulations from the CUDA kernels
 * (world-origin spatial algebra, dense matrices, Jacobi-eigen pseudo-inverse) so that parity
 * between the two is a real check.
 */
#define _USE_MATH_DEFINES
#include <math.h>
#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include "../include/rsb_model.h"

typedef double real;

#define MAXQ 48
#define MAXV 40
#define MAXB 64
#define MAXJ 48
#define MAXG 128
#define MAXS 16
#define MAXU 24
#define MAXCON 160
#define MAXEFC 640
#define MINVAL 1e-15

enum { EFC_FRICTION = 0, EFC_LIMIT = 1, EFC_CONTACT_NORMAL = 2, EFC_CONTACT_FRICTION = 3 };

typedef struct {
  int geom1, geom2, pair, dim, efc_address;
  real pos[3], frame[9], dist, includemargin, friction[5], solref[2], solimp[5], mu;
} orc_contact;

typedef struct {
  real goal_pos[3], goal_ori[9], initial_joint[7], grip_cur[2];
  real goal_vel[7], summed_err[7], last_err[7], derr_buf[5][7];
  int derr_n, derr_ptr, saturated;
  real torques[7];
  real osc_cond;                           /* condition number of J M^-1 J^T (6x6) at the last OSC evaluation: how much the torque law amplifies input round-off */
  real torques_raw[7];                     /* the control law's output before the actuator torque limits clip it */
} orc_ctrl;

typedef struct orc_env {
  rsb_model m;
  rsb_task t;
  int ncon_max, nefc_max;
  /* state */
  real qpos[MAXQ], qvel[MAXV], qacc_warmstart[MAXV], ctrl[MAXU];
  real bpose[7]; int override_body;          /* per-env pose of the task's placed FIXED body (rsb_task.place_body), -1 if none */
  orc_ctrl rc[RSB_MAX_ROBOTS];
  int timestep, done;
  uint64_t episode;
  /* position-dependent */
  real xpos[MAXB][3], xquat[MAXB][4], xmat[MAXB][9], xipos[MAXB][3], ximat[MAXB][9];
  real xanchor[MAXJ][3], xaxis[MAXJ][3];
  real geom_xpos[MAXG][3], geom_xmat[MAXG][9], site_xpos[MAXS][3], site_xmat[MAXS][9];
  real cdof[MAXV][6];            /* [ang; lin at world origin] */
  real M[MAXV * MAXV];
  /* velocity-dependent */
  real qfrc_bias[MAXV], qfrc_passive[MAXV], qfrc_actuator[MAXV], qfrc_smooth[MAXV];
  real qacc_smooth[MAXV], qacc[MAXV], qfrc_constraint[MAXV];
  /* contacts + constraints */
  int ncon, nefc, solver_iter;
  orc_contact con[MAXCON];
  real *efc_J;                   /* [MAXEFC][MAXV] */
  real efc_pos[MAXEFC], efc_margin[MAXEFC], efc_aref[MAXEFC], efc_R[MAXEFC], efc_D[MAXEFC];
  real efc_floss[MAXEFC], efc_force[MAXEFC], efc_vel[MAXEFC];
  int efc_type[MAXEFC], efc_id[MAXEFC];
} orc_env;

/* ------------------------------------------------------------------ small vector helpers */
static void v3set(real *o, real a, real b, real c) { o[0] = a; o[1] = b; o[2] = c; }
static void v3copy(real *o, const real *a) { o[0] = a[0]; o[1] = a[1]; o[2] = a[2]; }
static real v3dot(const real *a, const real *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static void v3cross(real *o, const real *a, const real *b) {
  real x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  o[0] = x; o[1] = y; o[2] = z;
}
static real v3norm(const real *a) { return sqrt(v3dot(a, a)); }
static void v3addscl(real *o, const real *a, const real *b, real s) { for (int k = 0; k < 3; k++) o[k] = a[k] + s * b[k]; }
static void v3sub(real *o, const real *a, const real *b) { for (int k = 0; k < 3; k++) o[k] = a[k] - b[k]; }
static real v3normalize(real *a) { real n = v3norm(a); if (n < MINVAL) { a[0] = 1; a[1] = 0; a[2] = 0; } else { a[0] /= n; a[1] /= n; a[2] /= n; } return n; }
/* row-major 3x3 */
static void m3mulv(real *o, const real *R, const real *v) {
  real x = R[0] * v[0] + R[1] * v[1] + R[2] * v[2], y = R[3] * v[0] + R[4] * v[1] + R[5] * v[2], z = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
  o[0] = x; o[1] = y; o[2] = z;
}
static void m3Tmulv(real *o, const real *R, const real *v) {
  real x = R[0] * v[0] + R[3] * v[1] + R[6] * v[2], y = R[1] * v[0] + R[4] * v[1] + R[7] * v[2], z = R[2] * v[0] + R[5] * v[1] + R[8] * v[2];
  o[0] = x; o[1] = y; o[2] = z;
}
static void m3mul(real *o, const real *A, const real *B) {
  real t[9];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) t[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
  memcpy(o, t, sizeof t);
}
static void qmul(real *o, const real *a, const real *b) {
  real w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  real x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  real y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  real z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  o[0] = w; o[1] = x; o[2] = y; o[3] = z;
}
static void qnormalize(real *q) {
  real n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < MINVAL) { q[0] = 1; q[1] = q[2] = q[3] = 0; } else { for (int k = 0; k < 4; k++) q[k] /= n; }
}
static void q2mat(real *R, const real *q) {
  real w = q[0], x = q[1], y = q[2], z = q[3];
  R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - w * z); R[2] = 2 * (x * z + w * y);
  R[3] = 2 * (x * y + w * z); R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - w * x);
  R[6] = 2 * (x * z - w * y); R[7] = 2 * (y * z + w * x); R[8] = 1 - 2 * (x * x + y * y);
}
static void qrot(real *o, const real *q, const real *v) { real R[9]; q2mat(R, q); m3mulv(o, R, v); }
static void axisangle2q(real *q, const real *axis, real ang) {
  real s = sin(0.5 * ang); q[0] = cos(0.5 * ang); q[1] = s * axis[0]; q[2] = s * axis[1]; q[3] = s * axis[2];
}

/* ------------------------------------------------------------------ model deep copy */
static void *dup(const void *p, size_t n) { void *o = malloc(n ? n : 1); if (p && n) memcpy(o, p, n); return o; }
#define DUPD(f, n) e->m.f = (const double *)dup(src->f, sizeof(double) * (size_t)(n))
#define DUPI(f, n) e->m.f = (const int *)dup(src->f, sizeof(int) * (size_t)(n))

orc_env *orc_create(const rsb_model *src, const rsb_task *task, int ncon_max) {
  if (src->nq > MAXQ || src->nv > MAXV || src->nbody > MAXB || src->njnt > MAXJ || src->ngeom > MAXG ||
      src->nsite > MAXS || src->nu > MAXU) return NULL;
  orc_env *e = (orc_env *)calloc(1, sizeof(orc_env));
  e->m = *src; e->t = *task;
  e->ncon_max = ncon_max > 0 && ncon_max < MAXCON ? ncon_max : MAXCON;
  e->nefc_max = MAXEFC;
  int nb = src->nbody, nj = src->njnt, nv = src->nv, ng = src->ngeom, ns = src->nsite, np = src->npair, nu = src->nu;
  DUPI(body_parentid, nb); DUPI(body_rootid, nb); DUPI(body_jntadr, nb); DUPI(body_jntnum, nb); DUPI(body_dofadr, nb); DUPI(body_dofnum, nb);
  DUPD(body_pos, 3 * nb); DUPD(body_quat, 4 * nb); DUPD(body_ipos, 3 * nb); DUPD(body_iquat, 4 * nb); DUPD(body_mass, nb); DUPD(body_inertia, 3 * nb);
  DUPD(body_invweight0, 2 * nb);
  DUPI(jnt_type, nj); DUPI(jnt_qposadr, nj); DUPI(jnt_dofadr, nj); DUPI(jnt_bodyid, nj); DUPI(jnt_limited, nj);
  DUPD(jnt_pos, 3 * nj); DUPD(jnt_axis, 3 * nj); DUPD(jnt_range, 2 * nj); DUPD(jnt_stiffness, nj); DUPD(jnt_margin, nj);
  DUPD(jnt_solref, 2 * nj); DUPD(jnt_solimp, 5 * nj);
  DUPI(dof_bodyid, nv); DUPI(dof_jntid, nv); DUPI(dof_parentid, nv); DUPI(dof_Madr, nv);
  DUPD(dof_armature, nv); DUPD(dof_damping, nv); DUPD(dof_frictionloss, nv); DUPD(dof_invweight0, nv);
  DUPD(dof_solref, 2 * nv); DUPD(dof_solimp, 5 * nv); DUPD(qpos0, src->nq); DUPD(qpos_spring, src->nq);
  DUPI(geom_type, ng); DUPI(geom_bodyid, ng); DUPD(geom_size, 3 * ng); DUPD(geom_pos, 3 * ng); DUPD(geom_quat, 4 * ng); DUPD(geom_rbound, ng);
  DUPI(site_bodyid, ns); DUPD(site_pos, 3 * ns); DUPD(site_quat, 4 * ns);
  DUPI(pair_geom1, np); DUPI(pair_geom2, np); DUPI(pair_condim, np);
  DUPD(pair_friction, 5 * np); DUPD(pair_solref, 2 * np); DUPD(pair_solimp, 5 * np); DUPD(pair_margin, np); DUPD(pair_gap, np);
  DUPI(act_dofid, nu); DUPI(act_ctrllimited, nu); DUPI(act_forcelimited, nu);
  DUPD(act_gain, nu); DUPD(act_bias, 3 * nu); DUPD(act_ctrlrange, 2 * nu); DUPD(act_forcerange, 2 * nu); DUPD(act_gear, nu);
  e->efc_J = (real *)calloc((size_t)MAXEFC * MAXV, sizeof(real));
  e->override_body = -1;
  for (int o = 0; o < RSB_MAX_OBJ; o++) if (task->place_body[o] >= 0) { int b = task->place_body[o]; e->override_body = b;
    for (int k = 0; k < 3; k++) e->bpose[k] = src->body_pos[3 * b + k]; for (int k = 0; k < 4; k++) e->bpose[3 + k] = src->body_quat[4 * b + k]; }
  for (int i = 0; i < src->nq; i++) e->qpos[i] = src->qpos0[i];
  return e;
}
/* same row cap as the CUDA library's nefc_max (contacts that do not fit are skipped, scalar rows are clipped) */
void orc_set_nefc_max(orc_env *e, int n) { e->nefc_max = n > 0 && n < MAXEFC ? n : MAXEFC; }
void orc_destroy(orc_env *e) { if (e) { free(e->efc_J); free(e); } /* model arrays leak by design: test helper */ }

/* ------------------------------------------------------------------ A.3.1 kinematics */
static void kinematics(orc_env *e) {
  const rsb_model *m = &e->m;
  v3set(e->xpos[0], 0, 0, 0); e->xquat[0][0] = 1; e->xquat[0][1] = e->xquat[0][2] = e->xquat[0][3] = 0;
  q2mat(e->xmat[0], e->xquat[0]); v3set(e->xipos[0], 0, 0, 0); q2mat(e->ximat[0], e->xquat[0]);
  for (int b = 1; b < m->nbody; b++) {
    int p = m->body_parentid[b], jn = m->body_jntnum[b], ja = m->body_jntadr[b];
    real pos[3], quat[4];
    if (jn == 1 && m->jnt_type[ja] == RSB_JNT_FREE) {
      int a = m->jnt_qposadr[ja];
      qnormalize(e->qpos + a + 3);                           /* mj_kinematics normalises in place */
      v3copy(pos, e->qpos + a); memcpy(quat, e->qpos + a + 3, sizeof quat);
      v3copy(e->xanchor[ja], pos); v3set(e->xaxis[ja], 0, 0, 1);
    } else {
      real t[3]; const real *bp = b == e->override_body ? e->bpose : m->body_pos + 3 * b, *bq = b == e->override_body ? e->bpose + 3 : m->body_quat + 4 * b;
      m3mulv(t, e->xmat[p], bp); v3addscl(pos, e->xpos[p], t, 1);
      qmul(quat, e->xquat[p], bq);
      for (int k = 0; k < jn; k++) {
        int j = ja + k; real dq = e->qpos[m->jnt_qposadr[j]] - m->qpos0[m->jnt_qposadr[j]];
        real anchor[3], axis[3];
        qrot(t, quat, m->jnt_pos + 3 * j); v3addscl(anchor, pos, t, 1);
        qrot(axis, quat, m->jnt_axis + 3 * j);
        if (m->jnt_type[j] == RSB_JNT_SLIDE) v3addscl(pos, pos, axis, dq);
        else {
          real ql[4], qn[4]; axisangle2q(ql, m->jnt_axis + 3 * j, dq); qmul(qn, quat, ql); memcpy(quat, qn, sizeof qn);
          qrot(t, quat, m->jnt_pos + 3 * j); v3sub(pos, anchor, t);
        }
        v3copy(e->xanchor[j], anchor); v3copy(e->xaxis[j], axis);
      }
    }
    qnormalize(quat);
    v3copy(e->xpos[b], pos); memcpy(e->xquat[b], quat, sizeof quat); q2mat(e->xmat[b], quat);
    real t[3], Ri[9];
    m3mulv(t, e->xmat[b], m->body_ipos + 3 * b); v3addscl(e->xipos[b], pos, t, 1);
    q2mat(Ri, m->body_iquat + 4 * b); m3mul(e->ximat[b], e->xmat[b], Ri);
  }
  for (int g = 0; g < m->ngeom; g++) {
    int b = m->geom_bodyid[g]; real t[3], R[9];
    m3mulv(t, e->xmat[b], m->geom_pos + 3 * g); v3addscl(e->geom_xpos[g], e->xpos[b], t, 1);
    q2mat(R, m->geom_quat + 4 * g); m3mul(e->geom_xmat[g], e->xmat[b], R);
  }
  for (int s = 0; s < m->nsite; s++) {
    int b = m->site_bodyid[s]; real t[3], R[9];
    m3mulv(t, e->xmat[b], m->site_pos + 3 * s); v3addscl(e->site_xpos[s], e->xpos[b], t, 1);
    q2mat(R, m->site_quat + 4 * s); m3mul(e->site_xmat[s], e->xmat[b], R);
  }
}

/* A.3.2: motion subspace of every dof as a spatial vector [w; v_O] about the WORLD origin */
static void make_cdof(orc_env *e) {
  const rsb_model *m = &e->m;
  for (int j = 0; j < m->njnt; j++) {
    int d = m->jnt_dofadr[j], b = m->jnt_bodyid[j];
    if (m->jnt_type[j] == RSB_JNT_FREE) {
      for (int k = 0; k < 3; k++) { memset(e->cdof[d + k], 0, sizeof e->cdof[0]); e->cdof[d + k][3 + k] = 1; }
      for (int k = 0; k < 3; k++) {
        real ax[3] = { e->xmat[b][k], e->xmat[b][3 + k], e->xmat[b][6 + k] };   /* body-local axis k */
        v3copy(e->cdof[d + 3 + k], ax); v3cross(e->cdof[d + 3 + k] + 3, e->xpos[b], ax);
      }
    } else if (m->jnt_type[j] == RSB_JNT_SLIDE) {
      v3set(e->cdof[d], 0, 0, 0); v3copy(e->cdof[d] + 3, e->xaxis[j]);
    } else {
      v3copy(e->cdof[d], e->xaxis[j]); v3cross(e->cdof[d] + 3, e->xanchor[j], e->xaxis[j]);
    }
  }
}

/* spatial inertia about the world origin: mass, first moment, second moment (sym 3x3) */
typedef struct { real m, h[3], I[9]; } sinertia;
static void body_sinertia(const orc_env *e, int b, sinertia *s) {
  const rsb_model *m = &e->m; real ms = m->body_mass[b]; const real *c = e->xipos[b], *R = e->ximat[b];
  s->m = ms; for (int k = 0; k < 3; k++) s->h[k] = ms * c[k];
  real cc = v3dot(c, c);
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
    real v = 0; for (int k = 0; k < 3; k++) v += R[3 * i + k] * m->body_inertia[3 * b + k] * R[3 * j + k];
    s->I[3 * i + j] = v + ms * ((i == j ? cc : 0) - c[i] * c[j]);
  }
}
/* f = I * s : s = [w; v], f = [n; p] */
static void sinertia_mul(real *f, const sinertia *I, const real *s) {
  real t[3]; v3cross(t, s, I->h);                       /* w x (m c) */
  for (int k = 0; k < 3; k++) f[3 + k] = I->m * s[3 + k] + t[k];
  m3mulv(f, I->I, s); v3cross(t, I->h, s + 3); for (int k = 0; k < 3; k++) f[k] += t[k];
}
static real sdot(const real *s, const real *f) { return v3dot(s, f) + v3dot(s + 3, f + 3); }

/* A.3.3 composite rigid body + armature -> dense M */
static void mass_matrix(orc_env *e) {
  const rsb_model *m = &e->m; int nv = m->nv;
  static sinertia crb[MAXB];
  for (int b = 0; b < m->nbody; b++) body_sinertia(e, b, &crb[b]);
  for (int b = m->nbody - 1; b > 0; b--) {
    int p = m->body_parentid[b];
    crb[p].m += crb[b].m; for (int k = 0; k < 3; k++) crb[p].h[k] += crb[b].h[k]; for (int k = 0; k < 9; k++) crb[p].I[k] += crb[b].I[k];
  }
  memset(e->M, 0, sizeof(real) * MAXV * MAXV);
  for (int i = 0; i < nv; i++) {
    real f[6]; sinertia_mul(f, &crb[m->dof_bodyid[i]], e->cdof[i]);
    for (int j = i; j >= 0; j = m->dof_parentid[j]) { real v = sdot(e->cdof[j], f); e->M[i * nv + j] = v; e->M[j * nv + i] = v; }
    e->M[i * nv + i] += m->dof_armature[i];
  }
}

/* spatial cross products */
static void crossm(real *o, const real *v, const real *s) {            /* motion: v x s */
  real a[3], b[3], c[3]; v3cross(a, v, s); v3cross(b, v, s + 3); v3cross(c, v + 3, s);
  for (int k = 0; k < 3; k++) { o[k] = a[k]; o[3 + k] = b[k] + c[k]; }
}
static void crossf(real *o, const real *v, const real *f) {            /* force: v x* f */
  real a[3], b[3], c[3]; v3cross(a, v, f); v3cross(b, v + 3, f + 3); v3cross(c, v, f + 3);
  for (int k = 0; k < 3; k++) { o[k] = a[k] + b[k]; o[3 + k] = c[k]; }
}

/* A.3.7 rne(acc=0) with gravity -> qfrc_bias; also passive forces */
static void bias_forces(orc_env *e) {
  const rsb_model *m = &e->m; int nv = m->nv;
  static real V[MAXB][6], A[MAXB][6], F[MAXB][6];
  memset(V[0], 0, sizeof V[0]); memset(A[0], 0, sizeof A[0]);
  for (int k = 0; k < 3; k++) A[0][3 + k] = -m->gravity[k];
  for (int b = 1; b < m->nbody; b++) {
    int p = m->body_parentid[b];
    memcpy(V[b], V[p], sizeof V[b]); memcpy(A[b], A[p], sizeof A[b]);
    int d0 = m->body_dofadr[b], dn = m->body_dofnum[b];
    for (int k = 0; k < dn; k++) {
      int d = d0 + k; real sd[6];
      int jt = m->jnt_type[m->dof_jntid[d]];
      int is_free_trans = (jt == RSB_JNT_FREE && k < 3);
      if (jt == RSB_JNT_FREE && k == 3) {            /* the three rotational axes are body-fixed: use the same V for all */
        real Vt[6]; memcpy(Vt, V[b], sizeof Vt);
        for (int r = 0; r < 3; r++) { crossm(sd, Vt, e->cdof[d + r]); for (int c = 0; c < 6; c++) A[b][c] += sd[c] * e->qvel[d + r]; }
        for (int r = 0; r < 3; r++) for (int c = 0; c < 6; c++) V[b][c] += e->cdof[d + r][c] * e->qvel[d + r];
        break;
      }
      if (!is_free_trans) { crossm(sd, V[b], e->cdof[d]); for (int c = 0; c < 6; c++) A[b][c] += sd[c] * e->qvel[d]; }
      for (int c = 0; c < 6; c++) V[b][c] += e->cdof[d][c] * e->qvel[d];
    }
  }
  for (int b = 1; b < m->nbody; b++) {
    sinertia I; body_sinertia(e, b, &I);
    real IA[6], IV[6], t[6]; sinertia_mul(IA, &I, A[b]); sinertia_mul(IV, &I, V[b]); crossf(t, V[b], IV);
    for (int c = 0; c < 6; c++) F[b][c] = IA[c] + t[c];
  }
  memset(F[0], 0, sizeof F[0]);
  for (int b = m->nbody - 1; b > 0; b--) { int p = m->body_parentid[b]; for (int c = 0; c < 6; c++) F[p][c] += F[b][c]; }
  for (int d = 0; d < nv; d++) e->qfrc_bias[d] = sdot(e->cdof[d], F[m->dof_bodyid[d]]);
  /* passive: joint springs and dampers (no fluid: density = viscosity = 0) */
  for (int d = 0; d < nv; d++) e->qfrc_passive[d] = -m->dof_damping[d] * e->qvel[d];
  for (int j = 0; j < m->njnt; j++) if (m->jnt_type[j] != RSB_JNT_FREE && m->jnt_stiffness[j] != 0)
    e->qfrc_passive[m->jnt_dofadr[j]] -= m->jnt_stiffness[j] * (e->qpos[m->jnt_qposadr[j]] - m->qpos_spring[m->jnt_qposadr[j]]);
}

/* Jacobian of a world point attached to `body`: jacp, jacr are [3][nv] row-major */
static void jac_point(const orc_env *e, int body, const real *p, real *jacp, real *jacr) {
  const rsb_model *m = &e->m; int nv = m->nv;
  memset(jacp, 0, sizeof(real) * 3 * nv); if (jacr) memset(jacr, 0, sizeof(real) * 3 * nv);
  while (body > 0 && m->body_dofnum[body] == 0) body = m->body_parentid[body];
  if (body <= 0) return;
  for (int d = m->body_dofadr[body] + m->body_dofnum[body] - 1; d >= 0; d = m->dof_parentid[d]) {
    real t[3]; v3cross(t, e->cdof[d], p);
    for (int k = 0; k < 3; k++) { jacp[k * nv + d] = e->cdof[d][3 + k] + t[k]; if (jacr) jacr[k * nv + d] = e->cdof[d][k]; }
  }
}

/* ------------------------------------------------------------------ A.3.5 collision */
static void make_frame(real *frame) {           /* frame[0..2] = normal given; fill the tangents (mju_makeFrame) */
  real *x = frame, *y = frame + 3, *z = frame + 6;
  if (x[1] < 0.5 && x[1] > -0.5) v3set(y, 0, 1, 0); else v3set(y, 0, 0, 1);
  real d = v3dot(x, y); v3addscl(y, y, x, -d); v3normalize(y); v3cross(z, x, y);
}

typedef struct { real pos[3], normal[3], dist; } rawcon;

static int plane_box(const real *ppos, const real *pmat, const real *bpos, const real *bmat, const real *size, real margin, rawcon *out) {
  real n[3] = { pmat[2], pmat[5], pmat[8] }; int cnt = 0;
  real dif[3]; v3sub(dif, bpos, ppos); real d0 = v3dot(dif, n);
  for (int i = 0; i < 8 && cnt < 4; i++) {
    real loc[3] = { (i & 1 ? size[0] : -size[0]), (i & 2 ? size[1] : -size[1]), (i & 4 ? size[2] : -size[2]) }, w[3];
    m3mulv(w, bmat, loc);
    real ld = d0 + v3dot(w, n);
    if (ld > margin) continue;
    rawcon *c = &out[cnt++]; c->dist = ld; v3copy(c->normal, n);
    for (int k = 0; k < 3; k++) c->pos[k] = bpos[k] + w[k] - n[k] * ld * 0.5;
  }
  return cnt;
}

static int plane_sphere(const real *ppos, const real *pmat, const real *spos, real r, real margin, rawcon *out) {
  real n[3] = { pmat[2], pmat[5], pmat[8] }, dif[3]; v3sub(dif, spos, ppos);
  real d = v3dot(dif, n) - r; if (d > margin) return 0;
  out->dist = d; v3copy(out->normal, n); for (int k = 0; k < 3; k++) out->pos[k] = spos[k] - n[k] * (r + 0.5 * d);
  return 1;
}

static int plane_capsule(const real *ppos, const real *pmat, const real *cpos, const real *cmat, const real *size, real margin, rawcon *out) {
  real ax[3] = { cmat[2], cmat[5], cmat[8] }; int cnt = 0;
  for (int s = -1; s <= 1; s += 2) { real p[3]; v3addscl(p, cpos, ax, s * size[1]); cnt += plane_sphere(ppos, pmat, p, size[0], margin, out + cnt); }
  return cnt;
}

static int sphere_sphere(const real *p1, real r1, const real *p2, real r2, real margin, rawcon *out) {
  real d[3]; v3sub(d, p2, p1); real len = v3norm(d), dist = len - r1 - r2;
  if (dist > margin) return 0;
  if (len < MINVAL) v3set(d, 1, 0, 0); else { d[0] /= len; d[1] /= len; d[2] /= len; }
  out->dist = dist; v3copy(out->normal, d); for (int k = 0; k < 3; k++) out->pos[k] = p1[k] + d[k] * (r1 + 0.5 * dist);
  return 1;
}

static int capsule_capsule(const real *p1, const real *m1, const real *s1, const real *p2, const real *m2, const real *s2, real margin, rawcon *out) {
  /* closest points between the two axis segments, then sphere-sphere */
  real a1[3] = { m1[2], m1[5], m1[8] }, a2[3] = { m2[2], m2[5], m2[8] }, d[3]; v3sub(d, p1, p2);
  real b = v3dot(a1, a2), c1 = v3dot(a1, d), c2 = v3dot(a2, d), den = 1 - b * b, t1, t2;
  if (den < 1e-12) { t1 = 0; t2 = c2; }
  else { t1 = (b * c2 - c1) / den; t2 = (c2 - b * c1) / den; }
  if (t1 > s1[1]) t1 = s1[1]; if (t1 < -s1[1]) t1 = -s1[1];
  t2 = c2 + b * t1; if (t2 > s2[1]) t2 = s2[1]; if (t2 < -s2[1]) t2 = -s2[1];
  t1 = -c1 + b * t2; if (t1 > s1[1]) t1 = s1[1]; if (t1 < -s1[1]) t1 = -s1[1];
  real q1[3], q2[3]; v3addscl(q1, p1, a1, t1); v3addscl(q2, p2, a2, t2);
  return sphere_sphere(q1, s1[0], q2, s2[0], margin, out);
}

static int sphere_box(const real *sp, real r, const real *bp, const real *bm, const real *size, real margin, rawcon *out) {
  real d[3], loc[3], cl[3]; v3sub(d, sp, bp); m3Tmulv(loc, bm, d); int inside = 1;
  for (int k = 0; k < 3; k++) { cl[k] = loc[k]; if (cl[k] > size[k]) { cl[k] = size[k]; inside = 0; } if (cl[k] < -size[k]) { cl[k] = -size[k]; inside = 0; } }
  real nl[3], dist;
  if (!inside) { v3sub(nl, loc, cl); real len = v3normalize(nl); dist = len - r; }
  else {        /* centre inside the box: push out through the nearest face */
    int bk = 0; real best = 1e30;
    for (int k = 0; k < 3; k++) { real f = size[k] - fabs(loc[k]); if (f < best) { best = f; bk = k; } }
    v3set(nl, 0, 0, 0); nl[bk] = loc[bk] >= 0 ? 1 : -1; cl[bk] = nl[bk] * size[bk]; dist = -best - r;
  }
  if (dist > margin) return 0;
  real nw[3], cw[3]; m3mulv(nw, bm, nl); m3mulv(cw, bm, cl);
  /* normal must point from geom1 (sphere) to geom2 (box) */
  out->dist = dist; for (int k = 0; k < 3; k++) { out->normal[k] = -nw[k]; out->pos[k] = bp[k] + cw[k] + nw[k] * 0.5 * dist; }
  return 1;
}

static int capsule_box(const real *cp, const real *cm, const real *cs, const real *bp, const real *bm, const real *size, real margin, rawcon *out) {
  /* two end spheres + the mid sphere: a primitive-only approximation of mjc_CapsuleBox (documented deviation) */
  real ax[3] = { cm[2], cm[5], cm[8] }; int cnt = 0;
  for (int s = -1; s <= 1; s++) { real p[3]; v3addscl(p, cp, ax, s * cs[1]); cnt += sphere_box(p, cs[0], bp, bm, size, margin, out + cnt); }
  return cnt;
}

/* box-box: separating-axis test + reference-face clipping (own algorithm; same as csrc/ narrow phase by
   construction of the spec in DESIGN.md "box-box", not MuJoCo's mjc_BoxBox -- documented deviation). */
static int box_box(const real *pa, const real *Ra, const real *ha, const real *pb, const real *Rb, const real *hb, real margin, rawcon *out) {
  real R[9], Q[9], t[3], d[3];
  /* R = Ra^T Rb */
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { real v = 0; for (int k = 0; k < 3; k++) v += Ra[3 * k + i] * Rb[3 * k + j]; R[3 * i + j] = v; Q[3 * i + j] = fabs(v) + 1e-9; }
  v3sub(d, pb, pa); m3Tmulv(t, Ra, d);
  real best = -1e30; int code = -1; real bsign = 1;
  /* face axes of A */
  for (int i = 0; i < 3; i++) {
    real s = fabs(t[i]) - (ha[i] + hb[0] * Q[3 * i] + hb[1] * Q[3 * i + 1] + hb[2] * Q[3 * i + 2]);
    if (s > margin) return 0;
    if (s > best + 1e-6) { best = s; code = i; bsign = t[i] >= 0 ? 1 : -1; }
  }
  /* face axes of B */
  for (int j = 0; j < 3; j++) {
    real tb = t[0] * R[j] + t[1] * R[3 + j] + t[2] * R[6 + j];
    real s = fabs(tb) - (hb[j] + ha[0] * Q[j] + ha[1] * Q[3 + j] + ha[2] * Q[6 + j]);
    if (s > margin) return 0;
    if (s > best + 1e-6) { best = s; code = 3 + j; bsign = tb >= 0 ? 1 : -1; }
  }
  /* edge x edge axes */
  real en[3] = { 0, 0, 0 };
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
    int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
    /* L = a_i x b_j in A's frame */
    real L[3] = { 0, 0, 0 }; L[i1] = -R[3 * i2 + j]; L[i2] = R[3 * i1 + j];
    real len = sqrt(L[i1] * L[i1] + L[i2] * L[i2]);
    if (len < 1e-6) continue;
    real tl = (t[i1] * L[i1] + t[i2] * L[i2]) / len;
    real ra = (ha[i1] * Q[3 * i2 + j] + ha[i2] * Q[3 * i1 + j]) / len;
    real rb = (hb[j1] * Q[3 * i + j2] + hb[j2] * Q[3 * i + j1]) / len;
    real s = fabs(tl) - (ra + rb);
    if (s > margin) return 0;
    if (s > best + 1e-4) { best = s; code = 6 + 3 * i + j; bsign = tl >= 0 ? 1 : -1; for (int k = 0; k < 3; k++) en[k] = L[k] / len; }
  }
  if (code < 0) return 0;
  if (code >= 6) {
    int i = (code - 6) / 3, j = (code - 6) % 3;
    real nA[3] = { en[0] * bsign, en[1] * bsign, en[2] * bsign }, n[3]; m3mulv(n, Ra, nA);   /* world, A -> B */
    /* supporting edge centres */
    real ca[3], cb[3]; v3copy(ca, pa); v3copy(cb, pb);
    for (int k = 0; k < 3; k++) if (k != i) { real sg = nA[k] > 0 ? 1 : -1; for (int c = 0; c < 3; c++) ca[c] += sg * ha[k] * Ra[3 * c + k]; }
    real nB[3]; m3Tmulv(nB, Rb, n);
    for (int k = 0; k < 3; k++) if (k != j) { real sg = nB[k] > 0 ? -1 : 1; for (int c = 0; c < 3; c++) cb[c] += sg * hb[k] * Rb[3 * c + k]; }
    real ua[3] = { Ra[i], Ra[3 + i], Ra[6 + i] }, ub[3] = { Rb[j], Rb[3 + j], Rb[6 + j] }, w[3]; v3sub(w, ca, cb);
    real bb = v3dot(ua, ub), dd = v3dot(ua, w), ee = v3dot(ub, w), den = 1 - bb * bb;
    real sa = den > 1e-12 ? (bb * ee - dd) / den : 0, sb = den > 1e-12 ? (ee - bb * dd) / den : 0;
    if (sa > ha[i]) sa = ha[i]; if (sa < -ha[i]) sa = -ha[i]; if (sb > hb[j]) sb = hb[j]; if (sb < -hb[j]) sb = -hb[j];
    real qa[3], qb[3]; v3addscl(qa, ca, ua, sa); v3addscl(qb, cb, ub, sb);
    out->dist = best; v3copy(out->normal, n); for (int k = 0; k < 3; k++) out->pos[k] = 0.5 * (qa[k] + qb[k]);
    return 1;
  }
  /* face contact: reference box owns the axis; clip the incident face of the other box */
  const real *pr, *Rr, *hr, *pi, *Ri, *hi; int ax; real nsign;
  if (code < 3) { pr = pa; Rr = Ra; hr = ha; pi = pb; Ri = Rb; hi = hb; ax = code; nsign = bsign; }
  else { pr = pb; Rr = Rb; hr = hb; pi = pa; Ri = Ra; hi = ha; ax = code - 3; nsign = -bsign; }
  /* outward reference-face normal in world */
  real nr[3] = { Rr[ax] * nsign, Rr[3 + ax] * nsign, Rr[6 + ax] * nsign };
  /* incident face: most anti-parallel face of the incident box */
  real nl[3]; m3Tmulv(nl, Ri, nr); int ia = 0; real am = fabs(nl[0]);
  for (int k = 1; k < 3; k++) if (fabs(nl[k]) > am) { am = fabs(nl[k]); ia = k; }
  real isg = nl[ia] > 0 ? -1 : 1; int u = (ia + 1) % 3, v = (ia + 2) % 3;
  real poly[16][3], tmp[16][3]; int np = 4;
  for (int c = 0; c < 4; c++) {
    real su = (c == 0 || c == 3) ? -1 : 1, sv = (c < 2) ? -1 : 1;
    for (int k = 0; k < 3; k++) poly[c][k] = pi[k] + isg * hi[ia] * Ri[3 * k + ia] + su * hi[u] * Ri[3 * k + u] + sv * hi[v] * Ri[3 * k + v];
  }
  /* to reference-local coordinates */
  for (int c = 0; c < 4; c++) { real w[3]; v3sub(w, poly[c], pr); m3Tmulv(poly[c], Rr, w); }
  int ru = (ax + 1) % 3, rv = (ax + 2) % 3;
  for (int side = 0; side < 4; side++) {
    int k = side < 2 ? ru : rv; real sg = (side & 1) ? -1 : 1, lim = hr[k];
    int nn = 0;
    for (int c = 0; c < np; c++) {
      const real *P = poly[c], *Qp = poly[(c + 1) % np];
      real dp = sg * P[k] - lim, dq = sg * Qp[k] - lim;
      if (dp <= 0) { v3copy(tmp[nn], P); nn++; }
      if ((dp < 0 && dq > 0) || (dp > 0 && dq < 0)) { real f = dp / (dp - dq); for (int x = 0; x < 3; x++) tmp[nn][x] = P[x] + f * (Qp[x] - P[x]); nn++; }
    }
    np = nn; memcpy(poly, tmp, sizeof(real) * 3 * (size_t)np);
    if (np == 0) return 0;
  }
  int cnt = 0;
  real nout[3]; for (int k = 0; k < 3; k++) nout[k] = (code < 3) ? nr[k] : -nr[k];    /* A -> B */
  for (int c = 0; c < np && cnt < 8; c++) {
    real depth = nsign * poly[c][ax] - hr[ax];          /* signed distance of the incident vertex above the reference face */
    if (depth > margin) continue;
    real pl[3]; v3copy(pl, poly[c]); pl[ax] -= nsign * 0.5 * depth;
    rawcon *o = &out[cnt++]; real w[3]; m3mulv(w, Rr, pl); v3addscl(o->pos, pr, w, 1); o->dist = depth; v3copy(o->normal, nout);
  }
  return cnt;
}

static void collision(orc_env *e) {
  const rsb_model *m = &e->m; e->ncon = 0;
  for (int p = 0; p < m->npair; p++) {
    int g1 = m->pair_geom1[p], g2 = m->pair_geom2[p], t1 = m->geom_type[g1], t2 = m->geom_type[g2];
    real margin = m->pair_margin[p]; const real *p1 = e->geom_xpos[g1], *p2 = e->geom_xpos[g2], *R1 = e->geom_xmat[g1], *R2 = e->geom_xmat[g2];
    const real *s1 = m->geom_size + 3 * g1, *s2 = m->geom_size + 3 * g2;
    /* mid phase: bounding spheres */
    if (t1 == RSB_GEOM_PLANE) {
      real n[3] = { R1[2], R1[5], R1[8] }, d[3]; v3sub(d, p2, p1);
      if (v3dot(d, n) > m->geom_rbound[g2] + margin) continue;
    } else {
      real d[3]; v3sub(d, p2, p1); real bound = m->geom_rbound[g1] + m->geom_rbound[g2] + margin;
      if (v3dot(d, d) > bound * bound) continue;
    }
    rawcon rc[16]; int n = 0;
    if (t1 == RSB_GEOM_PLANE && t2 == RSB_GEOM_BOX) n = plane_box(p1, R1, p2, R2, s2, margin, rc);
    else if (t1 == RSB_GEOM_PLANE && t2 == RSB_GEOM_SPHERE) n = plane_sphere(p1, R1, p2, s2[0], margin, rc);
    else if (t1 == RSB_GEOM_PLANE && t2 == RSB_GEOM_CAPSULE) n = plane_capsule(p1, R1, p2, R2, s2, margin, rc);
    else if (t1 == RSB_GEOM_SPHERE && t2 == RSB_GEOM_SPHERE) n = sphere_sphere(p1, s1[0], p2, s2[0], margin, rc);
    else if (t1 == RSB_GEOM_SPHERE && t2 == RSB_GEOM_BOX) n = sphere_box(p1, s1[0], p2, R2, s2, margin, rc);
    else if (t1 == RSB_GEOM_CAPSULE && t2 == RSB_GEOM_CAPSULE) n = capsule_capsule(p1, R1, s1, p2, R2, s2, margin, rc);
    else if (t1 == RSB_GEOM_CAPSULE && t2 == RSB_GEOM_BOX) n = capsule_box(p1, R1, s1, p2, R2, s2, margin, rc);
    else if (t1 == RSB_GEOM_BOX && t2 == RSB_GEOM_BOX) n = box_box(p1, R1, s1, p2, R2, s2, margin, rc);
    for (int k = 0; k < n && e->ncon < e->ncon_max; k++) {
      real inc = margin - m->pair_gap[p];
      if (!(rc[k].dist < inc)) continue;
      orc_contact *c = &e->con[e->ncon++];
      c->geom1 = g1; c->geom2 = g2; c->pair = p; c->dim = m->pair_condim[p]; c->dist = rc[k].dist; c->includemargin = inc;
      v3copy(c->pos, rc[k].pos); v3copy(c->frame, rc[k].normal); make_frame(c->frame);
      memcpy(c->friction, m->pair_friction + 5 * p, sizeof c->friction);
      memcpy(c->solref, m->pair_solref + 2 * p, sizeof c->solref); memcpy(c->solimp, m->pair_solimp + 5 * p, sizeof c->solimp);
      c->mu = 0; c->efc_address = -1;
    }
  }
}

/* ------------------------------------------------------------------ A.3.6 constraint rows */
static real impedance(const real *si, real pos, real margin) {
  real d0 = si[0], d1 = si[1], w = si[2], mid = si[3], pw = si[4];
  if (d0 == d1 || w <= MINVAL) return 0.5 * (d0 + d1);
  real x = fabs(pos - margin) / w, y;
  if (x >= 1) return d1;
  if (x <= 0) return d0;
  if (pw == 1) y = x;
  else if (x <= mid) y = pow(x, pw) / pow(mid, pw - 1);
  else y = 1 - pow(1 - x, pw) / pow(1 - mid, pw - 1);
  return d0 + y * (d1 - d0);
}

static void kb(const orc_env *e, const real *solref, const real *solimp, real *K, real *B) {
  real dmax = solimp[1];
  if (solref[0] > 0) {
    real tc = solref[0], dr = solref[1]; if (tc < 2 * e->m.timestep) tc = 2 * e->m.timestep;       /* refsafe */
    real k = dmax * dmax * tc * tc * dr * dr; *K = 1 / (k > MINVAL ? k : MINVAL);
    real b = dmax * tc; *B = 2 / (b > MINVAL ? b : MINVAL);
  } else { *K = -solref[0] / (dmax * dmax > MINVAL ? dmax * dmax : MINVAL); *B = -solref[1] / (dmax > MINVAL ? dmax : MINVAL); }
}

static int add_row(orc_env *e, int type, int id) {
  int r = e->nefc++; memset(e->efc_J + (size_t)r * MAXV, 0, sizeof(real) * MAXV);
  e->efc_type[r] = type; e->efc_id[r] = id; e->efc_floss[r] = 0; e->efc_pos[r] = 0; e->efc_margin[r] = 0;
  return r;
}

static void make_constraint(orc_env *e) {
  const rsb_model *m = &e->m; int nv = m->nv; e->nefc = 0;
  /* dof friction loss */
  for (int d = 0; d < nv; d++) if (m->dof_frictionloss[d] > 0) {
    int r = add_row(e, EFC_FRICTION, d); e->efc_J[(size_t)r * MAXV + d] = 1; e->efc_floss[r] = m->dof_frictionloss[d];
  }
  /* joint limits (hinge / slide) */
  for (int j = 0; j < m->njnt; j++) if (m->jnt_limited[j] && m->jnt_type[j] != RSB_JNT_FREE) {
    real q = e->qpos[m->jnt_qposadr[j]];
    for (int side = -1; side <= 1; side += 2) {
      real dist = side < 0 ? q - m->jnt_range[2 * j] : m->jnt_range[2 * j + 1] - q;
      if (dist < m->jnt_margin[j]) {
        int r = add_row(e, EFC_LIMIT, j); e->efc_J[(size_t)r * MAXV + m->jnt_dofadr[j]] = -side;
        e->efc_pos[r] = dist; e->efc_margin[r] = m->jnt_margin[j];
      }
    }
  }
  /* contacts (elliptic cones: dim rows each; pyramidal handled as frictionless+note) */
  static real jp1[3 * MAXV], jr1[3 * MAXV], jp2[3 * MAXV], jr2[3 * MAXV];
  for (int c = 0; c < e->ncon; c++) {
    orc_contact *k = &e->con[c]; if (e->nefc + k->dim > e->nefc_max) { k->efc_address = -1; continue; }
    int b1 = m->geom_bodyid[k->geom1], b2 = m->geom_bodyid[k->geom2];
    jac_point(e, b1, k->pos, jp1, jr1); jac_point(e, b2, k->pos, jp2, jr2);
    k->efc_address = e->nefc;
    for (int r = 0; r < k->dim; r++) {
      int row = add_row(e, r == 0 ? EFC_CONTACT_NORMAL : EFC_CONTACT_FRICTION, c); real *J = e->efc_J + (size_t)row * MAXV;
      const real *ax = k->frame + 3 * (r < 3 ? r : r - 3);
      const real *A = r < 3 ? jp1 : jr1, *B = r < 3 ? jp2 : jr2;
      for (int d = 0; d < nv; d++) J[d] = ax[0] * (B[d] - A[d]) + ax[1] * (B[nv + d] - A[nv + d]) + ax[2] * (B[2 * nv + d] - A[2 * nv + d]);
      if (r == 0) { e->efc_pos[row] = k->dist; e->efc_margin[row] = k->includemargin; }
    }
  }
  /* impedance, regularisation, reference acceleration (mj_makeImpedance) */
  for (int r = 0; r < e->nefc; r++) {
    const real *solref, *solimp; real diag; int type = e->efc_type[r], id = e->efc_id[r];
    if (type == EFC_FRICTION) { solref = m->dof_solref + 2 * id; solimp = m->dof_solimp + 5 * id; diag = m->dof_invweight0[id]; }
    else if (type == EFC_LIMIT) { solref = m->jnt_solref + 2 * id; solimp = m->jnt_solimp + 5 * id; diag = m->dof_invweight0[m->jnt_dofadr[id]]; }
    else {
      orc_contact *k = &e->con[id]; solref = k->solref; solimp = k->solimp; int rr = r - k->efc_address;
      int b1 = m->geom_bodyid[k->geom1], b2 = m->geom_bodyid[k->geom2], o = rr < 3 ? 0 : 1;
      diag = m->body_invweight0[2 * b1 + o] + m->body_invweight0[2 * b2 + o];
    }
    real K, B; kb(e, solref, solimp, &K, &B);
    if (type == EFC_FRICTION || type == EFC_CONTACT_FRICTION) K = 0;
    real imp = impedance(solimp, e->efc_pos[r], e->efc_margin[r]);
    real R = (1 - imp) / imp * diag; if (R < MINVAL) R = MINVAL; e->efc_R[r] = R;
    real vel = 0; const real *J = e->efc_J + (size_t)r * MAXV; for (int d = 0; d < nv; d++) vel += J[d] * e->qvel[d];
    e->efc_vel[r] = vel;
    e->efc_aref[r] = -B * vel - K * imp * (e->efc_pos[r] - e->efc_margin[r]);
  }
  /* elliptic friction rows: R from the normal row and impratio; regularised cone slope mu */
  for (int c = 0; c < e->ncon; c++) {
    orc_contact *k = &e->con[c]; int i = k->efc_address; if (i < 0 || k->dim < 2) { if (i >= 0) k->mu = k->friction[0]; continue; }
    real ir = m->impratio > MINVAL ? m->impratio : MINVAL;
    e->efc_R[i + 1] = e->efc_R[i] / ir;
    k->mu = k->friction[0] * sqrt(e->efc_R[i + 1] / e->efc_R[i]);
    for (int j = 1; j < k->dim - 1; j++) e->efc_R[i + 1 + j] = e->efc_R[i + 1] * k->friction[0] * k->friction[0] / (k->friction[j] * k->friction[j]);
  }
  for (int r = 0; r < e->nefc; r++) e->efc_D[r] = 1 / e->efc_R[r];
}

/* ------------------------------------------------------------------ dense Cholesky helpers */
static int chol(real *A, int n, int ld) {           /* in place lower; returns 0 ok */
  for (int j = 0; j < n; j++) {
    real s = A[j * ld + j]; for (int k = 0; k < j; k++) s -= A[j * ld + k] * A[j * ld + k];
    if (s < MINVAL) return 1; s = sqrt(s); A[j * ld + j] = s;
    for (int i = j + 1; i < n; i++) { real t = A[i * ld + j]; for (int k = 0; k < j; k++) t -= A[i * ld + k] * A[j * ld + k]; A[i * ld + j] = t / s; }
  }
  return 0;
}
static void chol_solve(const real *L, int n, int ld, real *x) {
  for (int i = 0; i < n; i++) { real s = x[i]; for (int k = 0; k < i; k++) s -= L[i * ld + k] * x[k]; x[i] = s / L[i * ld + i]; }
  for (int i = n - 1; i >= 0; i--) { real s = x[i]; for (int k = i + 1; k < n; k++) s -= L[k * ld + i] * x[k]; x[i] = s / L[i * ld + i]; }
}

/* ------------------------------------------------------------------ A.3.7 constraint cost and Newton solver */
typedef struct { real cost, d1, d2; } lsval;

/* Evaluate constraint cost at jar (+ alpha*Jv when Jv != NULL).  Writes forces when force != NULL, adds the constraint
   Hessian J^T W J into H when H != NULL, and accumulates directional derivatives into ls when ls != NULL. */
static real constraint_eval(const orc_env *e, const real *jar0, const real *Jv, real alpha, real *force, real *H, lsval *ls) {
  const rsb_model *m = &e->m; int nv = m->nv; real cost = 0;
  for (int r = 0; r < e->nefc; r++) {
    int type = e->efc_type[r]; real D = e->efc_D[r], R = e->efc_R[r];
    real x = jar0[r] + (Jv ? alpha * Jv[r] : 0), dx = Jv ? Jv[r] : 0; const real *J = e->efc_J + (size_t)r * MAXV;
    if (type == EFC_FRICTION) {
      real fl = e->efc_floss[r], rf = R * fl; real w = 0;
      if (x <= -rf) { cost += -0.5 * rf * fl - fl * x; if (force) force[r] = fl; if (ls) ls->d1 += -fl * dx; }
      else if (x >= rf) { cost += -0.5 * rf * fl + fl * x; if (force) force[r] = -fl; if (ls) ls->d1 += fl * dx; }
      else { cost += 0.5 * D * x * x; if (force) force[r] = -D * x; if (ls) { ls->d1 += D * x * dx; ls->d2 += D * dx * dx; } w = D; }
      if (H && w > 0) for (int a = 0; a < nv; a++) if (J[a] != 0) for (int b = 0; b < nv; b++) H[a * nv + b] += w * J[a] * J[b];
    } else if (type == EFC_LIMIT || (type == EFC_CONTACT_NORMAL && e->con[e->efc_id[r]].dim == 1)) {
      if (x < 0) {
        cost += 0.5 * D * x * x; if (force) force[r] = -D * x; if (ls) { ls->d1 += D * x * dx; ls->d2 += D * dx * dx; }
        if (H) for (int a = 0; a < nv; a++) if (J[a] != 0) for (int b = 0; b < nv; b++) H[a * nv + b] += D * J[a] * J[b];
      } else if (force) force[r] = 0;
    } else if (type == EFC_CONTACT_NORMAL) {
      const orc_contact *k = &e->con[e->efc_id[r]]; int dim = k->dim; real mu = k->mu;
      real sc[6], U[6], dU[6], xs[6]; sc[0] = mu; for (int j = 1; j < dim; j++) sc[j] = k->friction[j - 1];
      for (int j = 0; j < dim; j++) { xs[j] = jar0[r + j] + (Jv ? alpha * Jv[r + j] : 0); U[j] = xs[j] * sc[j]; dU[j] = (Jv ? Jv[r + j] : 0) * sc[j]; }
      real N = U[0], T = 0; for (int j = 1; j < dim; j++) T += U[j] * U[j]; T = sqrt(T);
      if (N >= mu * T || (T <= 0 && N >= 0)) { if (force) for (int j = 0; j < dim; j++) force[r + j] = 0; }
      else if (mu * N + T <= 0 || (T <= 0 && N < 0)) {
        for (int j = 0; j < dim; j++) {
          real Dj = e->efc_D[r + j], dxj = Jv ? Jv[r + j] : 0; cost += 0.5 * Dj * xs[j] * xs[j]; if (force) force[r + j] = -Dj * xs[j];
          if (ls) { ls->d1 += Dj * xs[j] * dxj; ls->d2 += Dj * dxj * dxj; }
          if (H) { const real *Jj = J + (size_t)j * MAXV; for (int a = 0; a < nv; a++) if (Jj[a] != 0) for (int b = 0; b < nv; b++) H[a * nv + b] += Dj * Jj[a] * Jj[b]; }
        }
      } else {
        real Dm = D / (mu * mu * (1 + mu * mu)), NmT = N - mu * T;
        cost += 0.5 * Dm * NmT * NmT;
        /* gradient and Hessian in U space */
        real g[6], HU[36]; g[0] = Dm * NmT; for (int j = 1; j < dim; j++) g[j] = -Dm * mu * NmT * U[j] / T;
        if (force) for (int j = 0; j < dim; j++) force[r + j] = -g[j] * sc[j];
        if (ls || H) {
          for (int a = 0; a < dim; a++) for (int b = 0; b < dim; b++) {
            real h;
            if (a == 0 && b == 0) h = Dm;
            else if (a == 0 || b == 0) h = -Dm * mu * U[a + b] / T;
            else h = Dm * mu * mu * U[a] * U[b] / (T * T) - Dm * mu * NmT * ((a == b ? 1.0 : 0.0) / T - U[a] * U[b] / (T * T * T));
            HU[a * 6 + b] = h;
          }
          if (ls) {
            for (int a = 0; a < dim; a++) { ls->d1 += g[a] * dU[a]; for (int b = 0; b < dim; b++) ls->d2 += dU[a] * HU[a * 6 + b] * dU[b]; }
          }
          if (H) for (int a = 0; a < dim; a++) for (int b = 0; b < dim; b++) {
            real w = HU[a * 6 + b] * sc[a] * sc[b]; const real *Ja = J + (size_t)a * MAXV, *Jb = J + (size_t)b * MAXV;
            for (int p = 0; p < nv; p++) if (Ja[p] != 0) for (int q = 0; q < nv; q++) H[p * nv + q] += w * Ja[p] * Jb[q];
          }
        }
      }
      r += dim - 1;
    }
  }
  return cost;
}

static void mulJ(const orc_env *e, const real *v, real *out) {
  int nv = e->m.nv; for (int r = 0; r < e->nefc; r++) { const real *J = e->efc_J + (size_t)r * MAXV; real s = 0; for (int d = 0; d < nv; d++) s += J[d] * v[d]; out[r] = s; }
}
static void mulJT(const orc_env *e, const real *f, real *out) {
  int nv = e->m.nv; for (int d = 0; d < nv; d++) out[d] = 0;
  for (int r = 0; r < e->nefc; r++) { const real *J = e->efc_J + (size_t)r * MAXV; for (int d = 0; d < nv; d++) out[d] += J[d] * f[r]; }
}
static void mulM(const orc_env *e, const real *v, real *out) {
  int nv = e->m.nv; for (int i = 0; i < nv; i++) { real s = 0; for (int j = 0; j < nv; j++) s += e->M[i * nv + j] * v[j]; out[i] = s; }
}

static real total_cost(const orc_env *e, const real *qacc, real *jar_out) {
  int nv = e->m.nv; static real jar[MAXEFC]; real d[MAXV], Md[MAXV];
  mulJ(e, qacc, jar); for (int r = 0; r < e->nefc; r++) jar[r] -= e->efc_aref[r];
  for (int i = 0; i < nv; i++) d[i] = qacc[i] - e->qacc_smooth[i];
  mulM(e, d, Md); real g = 0; for (int i = 0; i < nv; i++) g += 0.5 * d[i] * Md[i];
  if (jar_out) memcpy(jar_out, jar, sizeof(real) * (size_t)e->nefc);
  return g + constraint_eval(e, jar, NULL, 0, NULL, NULL, NULL);
}

static void solve_constraints(orc_env *e) {
  const rsb_model *m = &e->m; int nv = m->nv, nefc = e->nefc;
  e->solver_iter = 0;
  if (nefc == 0) { memcpy(e->qacc, e->qacc_smooth, sizeof(real) * (size_t)nv); memset(e->qfrc_constraint, 0, sizeof(real) * (size_t)nv); return; }
  static real jar[MAXEFC], Jv[MAXEFC], H[MAXV * MAXV];
  real Ma[MAXV], grad[MAXV], search[MAXV], Mv[MAXV], qfc[MAXV];
  /* warm start: pick the cheaper of qacc_warmstart and qacc_smooth (mj_fwdConstraint) */
  real cw = total_cost(e, e->qacc_warmstart, NULL), cs = total_cost(e, e->qacc_smooth, NULL);
  memcpy(e->qacc, cw < cs ? e->qacc_warmstart : e->qacc_smooth, sizeof(real) * (size_t)nv);
  real scale = 1.0 / (m->meaninertia * (nv > 1 ? nv : 1));
  real tol = 1e-12;                                /* tighter than opt.tolerance: the oracle reports the minimiser */
  real cost = total_cost(e, e->qacc, jar);
  for (int iter = 0; iter < 200; iter++) {
    /* gradient and Hessian */
    constraint_eval(e, jar, NULL, 0, e->efc_force, NULL, NULL);
    mulJT(e, e->efc_force, qfc); mulM(e, e->qacc, Ma);
    real gn = 0;
    for (int i = 0; i < nv; i++) { real qs = 0; for (int j = 0; j < nv; j++) qs += e->M[i * nv + j] * e->qacc_smooth[j]; grad[i] = Ma[i] - qs - qfc[i]; gn += grad[i] * grad[i]; }
    if (scale * sqrt(gn) < tol) break;
    memcpy(H, e->M, sizeof(real) * (size_t)nv * nv);   /* M is stored with leading dim nv */
    constraint_eval(e, jar, NULL, 0, NULL, H, NULL);
    if (chol(H, nv, nv)) break;
    memcpy(search, grad, sizeof(real) * (size_t)nv); chol_solve(H, nv, nv, search); for (int i = 0; i < nv; i++) search[i] = -search[i];
    mulM(e, search, Mv); mulJ(e, search, Jv);
    /* exact line search on the convex 1-D cost: safeguarded Newton on the derivative */
    real gq1 = 0, gq2 = 0; for (int i = 0; i < nv; i++) { gq1 += search[i] * grad[i] + 0; gq2 += search[i] * Mv[i]; }
    /* derivative of the Gauss part at alpha: s.(Ma - M a_s) + alpha s.Ms ; constraint part added below (grad already has -J^T f) */
    real gaussd1 = 0; for (int i = 0; i < nv; i++) { real qs = 0; for (int j = 0; j < nv; j++) qs += e->M[i * nv + j] * e->qacc_smooth[j]; gaussd1 += search[i] * (Ma[i] - qs); }
    real lo = 0, hi = -1, alpha = 0, d1 = 0, d2 = 0;
    for (int it = 0; it < 100; it++) {
      lsval v = { 0, 0, 0 }; constraint_eval(e, jar, Jv, alpha, NULL, NULL, &v);
      d1 = gaussd1 + alpha * gq2 + v.d1; d2 = gq2 + v.d2;
      if (fabs(d1) < 1e-14 * (1 + fabs(gq1))) break;
      if (d1 < 0) lo = alpha; else hi = alpha;
      real an = d2 > MINVAL ? alpha - d1 / d2 : alpha;
      if (hi >= 0 && (an <= lo || an >= hi)) an = 0.5 * (lo + hi);
      else if (hi < 0 && an <= lo) an = lo > 0 ? 2 * lo : 1;
      if (an == alpha) break;
      alpha = an;
    }
    (void)gq1;
    if (alpha == 0) break;
    for (int i = 0; i < nv; i++) e->qacc[i] += alpha * search[i];
    for (int r = 0; r < nefc; r++) jar[r] += alpha * Jv[r];
    real newcost = total_cost(e, e->qacc, jar);
    e->solver_iter = iter + 1;
    real improvement = scale * (cost - newcost); cost = newcost;
    if (improvement < 1e-15) break;
  }
  constraint_eval(e, jar, NULL, 0, e->efc_force, NULL, NULL);
  mulJT(e, e->efc_force, e->qfrc_constraint);
}

/* ------------------------------------------------------------------ actuation, acceleration, Euler */
static void actuation(orc_env *e) {
  const rsb_model *m = &e->m; for (int d = 0; d < m->nv; d++) e->qfrc_actuator[d] = 0;
  for (int a = 0; a < m->nu; a++) {
    real c = e->ctrl[a]; int d = m->act_dofid[a];
    if (m->act_ctrllimited[a]) { if (c < m->act_ctrlrange[2 * a]) c = m->act_ctrlrange[2 * a]; if (c > m->act_ctrlrange[2 * a + 1]) c = m->act_ctrlrange[2 * a + 1]; }
    int qa = m->jnt_qposadr[m->dof_jntid[d]]; real len = e->qpos[qa] * m->act_gear[a], vel = e->qvel[d] * m->act_gear[a];
    real f = m->act_gain[a] * c + m->act_bias[3 * a] + m->act_bias[3 * a + 1] * len + m->act_bias[3 * a + 2] * vel;
    if (m->act_forcelimited[a]) { if (f < m->act_forcerange[2 * a]) f = m->act_forcerange[2 * a]; if (f > m->act_forcerange[2 * a + 1]) f = m->act_forcerange[2 * a + 1]; }
    e->qfrc_actuator[d] += m->act_gear[a] * f;
  }
}

static void acceleration(orc_env *e) {
  int nv = e->m.nv; static real L[MAXV * MAXV];
  for (int d = 0; d < nv; d++) e->qfrc_smooth[d] = e->qfrc_passive[d] - e->qfrc_bias[d] + e->qfrc_actuator[d];
  memcpy(L, e->M, sizeof(real) * (size_t)nv * nv); chol(L, nv, nv);
  memcpy(e->qacc_smooth, e->qfrc_smooth, sizeof(real) * (size_t)nv); chol_solve(L, nv, nv, e->qacc_smooth);
}

static void euler(orc_env *e) {
  const rsb_model *m = &e->m; int nv = m->nv; real h = m->timestep; real a[MAXV]; static real L[MAXV * MAXV];
  int damp = 0; for (int d = 0; d < nv; d++) if (m->dof_damping[d] > 0) damp = 1;
  if (damp) {
    memcpy(L, e->M, sizeof(real) * (size_t)nv * nv); for (int d = 0; d < nv; d++) L[d * nv + d] += h * m->dof_damping[d];
    chol(L, nv, nv); for (int d = 0; d < nv; d++) a[d] = e->qfrc_smooth[d] + e->qfrc_constraint[d]; chol_solve(L, nv, nv, a);
  } else memcpy(a, e->qacc, sizeof(real) * (size_t)nv);
  for (int d = 0; d < nv; d++) e->qvel[d] += h * a[d];
  for (int j = 0; j < m->njnt; j++) {
    int qa = m->jnt_qposadr[j], d = m->jnt_dofadr[j];
    if (m->jnt_type[j] == RSB_JNT_FREE) {
      for (int k = 0; k < 3; k++) e->qpos[qa + k] += h * e->qvel[d + k];
      real w[3] = { e->qvel[d + 3], e->qvel[d + 4], e->qvel[d + 5] }, ang = v3norm(w) * h;
      if (ang > 0) { real ax[3] = { w[0], w[1], w[2] }; v3normalize(ax); real dq[4], qn[4]; axisangle2q(dq, ax, ang); qmul(qn, e->qpos + qa + 3, dq); qnormalize(qn); memcpy(e->qpos + qa + 3, qn, sizeof qn); }
    } else e->qpos[qa] += h * e->qvel[d];
  }
  memcpy(e->qacc_warmstart, e->qacc, sizeof(real) * (size_t)nv);
}

/* ------------------------------------------------------------------ A.2 controllers */
/* symmetric Jacobi eigen-decomposition; A (n x n, ld 6) destroyed, V columns = eigenvectors */
static void jacobi_eig(real *A, real *V, int n) {
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) V[i * 6 + j] = i == j;
  for (int sweep = 0; sweep < 60; sweep++) {
    real off = 0; for (int i = 0; i < n; i++) for (int j = i + 1; j < n; j++) off += A[i * 6 + j] * A[i * 6 + j];
    if (off < 1e-300) break;
    for (int p = 0; p < n; p++) for (int q = p + 1; q < n; q++) {
      if (fabs(A[p * 6 + q]) < 1e-300) continue;
      real th = (A[q * 6 + q] - A[p * 6 + p]) / (2 * A[p * 6 + q]), t = (th >= 0 ? 1 : -1) / (fabs(th) + sqrt(th * th + 1)), c = 1 / sqrt(t * t + 1), s = t * c;
      for (int k = 0; k < n; k++) { real akp = A[k * 6 + p], akq = A[k * 6 + q]; A[k * 6 + p] = c * akp - s * akq; A[k * 6 + q] = s * akp + c * akq; }
      for (int k = 0; k < n; k++) { real apk = A[p * 6 + k], aqk = A[q * 6 + k]; A[p * 6 + k] = c * apk - s * aqk; A[q * 6 + k] = s * apk + c * aqk; }
      for (int k = 0; k < n; k++) { real vkp = V[k * 6 + p], vkq = V[k * 6 + q]; V[k * 6 + p] = c * vkp - s * vkq; V[k * 6 + q] = s * vkp + c * vkq; }
    }
  }
}
/* numpy.linalg.pinv of a symmetric PSD matrix (rcond 1e-15), ld 6 */
static void sym_pinv(const real *Ain, real *out, int n) {
  real A[36], V[36]; memcpy(A, Ain, sizeof A); jacobi_eig(A, V, n);
  real lmax = 0; for (int i = 0; i < n; i++) if (fabs(A[i * 6 + i]) > lmax) lmax = fabs(A[i * 6 + i]);
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) {
    real s = 0; for (int k = 0; k < n; k++) { real l = A[k * 6 + k]; if (fabs(l) > 1e-15 * lmax) s += V[i * 6 + k] * V[j * 6 + k] / l; }
    out[i * 6 + j] = s;
  }
}

static void arm_quantities(const orc_env *e, const rsb_robot *rb, real *J /*6x7*/, real *Marm /*7x7*/, real *Minv /*7x7*/, real *v6) {
  const rsb_model *m = &e->m; int nv = m->nv; static real jp[3 * MAXV], jr[3 * MAXV];
  int sb = m->site_bodyid[rb->eef_site];
  jac_point(e, sb, e->site_xpos[rb->eef_site], jp, jr);
  for (int r = 0; r < 3; r++) for (int c = 0; c < 7; c++) { J[r * 7 + c] = jp[r * nv + rb->arm_dofadr[c]]; J[(3 + r) * 7 + c] = jr[r * nv + rb->arm_dofadr[c]]; }
  /* site velocity = full Jacobian times full qvel (mujoco-py site_xvelp / site_xvelr) */
  for (int r = 0; r < 3; r++) { real a = 0, b = 0; for (int d = 0; d < nv; d++) { a += jp[r * nv + d] * e->qvel[d]; b += jr[r * nv + d] * e->qvel[d]; } v6[r] = a; v6[3 + r] = b; }
  for (int r = 0; r < 7; r++) for (int c = 0; c < 7; c++) Marm[r * 7 + c] = e->M[rb->arm_dofadr[r] * nv + rb->arm_dofadr[c]];
  real L[49]; memcpy(L, Marm, sizeof L); chol(L, 7, 7);
  for (int c = 0; c < 7; c++) { real x[7] = { 0 }; x[c] = 1; chol_solve(L, 7, 7, x); for (int r = 0; r < 7; r++) Minv[r * 7 + c] = x[r]; }
}

static real scale_action(const rsb_robot *rb, int k, real a) {
  real lo = rb->input_min[k], hi = rb->input_max[k]; if (a < lo) a = lo; if (a > hi) a = hi;
  real sc = fabs(rb->output_max[k] - rb->output_min[k]) / fabs(hi - lo);
  return (a - 0.5 * (hi + lo)) * sc + 0.5 * (rb->output_max[k] + rb->output_min[k]);
}

static void controller_reset(orc_env *e, int ri) {
  const rsb_robot *rb = &e->t.robot[ri]; orc_ctrl *c = &e->rc[ri]; memset(c, 0, sizeof *c);
  for (int k = 0; k < 7; k++) c->initial_joint[k] = e->qpos[rb->arm_qposadr[k]];
  v3copy(c->goal_pos, e->site_xpos[rb->eef_site]); memcpy(c->goal_ori, e->site_xmat[rb->eef_site], sizeof c->goal_ori);
}

static void controller_set_goal(orc_env *e, int ri, const real *action) {
  const rsb_robot *rb = &e->t.robot[ri]; orc_ctrl *c = &e->rc[ri];
  if (rb->ctrl_type == RSB_CTRL_OSC_POSE || rb->ctrl_type == RSB_CTRL_OSC_POSITION) {
    real d[6] = { 0 }; for (int k = 0; k < rb->control_dim; k++) d[k] = scale_action(rb, k, action[k]);
    if (rb->ctrl_type == RSB_CTRL_OSC_POSE && (d[3] != 0 || d[4] != 0 || d[5] != 0)) {
      real Rm[9];
      if (rb->ori_delta_mode == RSB_ORI_DELTA_AXIS_ANGLE) {          /* robosuite >= 1.1: quat2mat(axisangle2quat(d)) */
        real ang = v3norm(d + 3), ax[3] = { d[3], d[4], d[5] }, q[4];
        if (ang < 1e-15) { q[0] = 1; q[1] = q[2] = q[3] = 0; } else { v3normalize(ax); axisangle2q(q, ax, ang); }
        q2mat(Rm, q);
      } else {
        /* euler2mat(d)^T, euler2mat as in mujoco-py / robosuite transform_utils.  That function's closed form equals E = X(d0) Y(d1) Z(d2) (elementary
           rotations about the fixed axes); composed here from the three transposed elementary rotations instead of the closed form the device code
           uses: E^T = Z(d2)^T Y(d1)^T X(d0)^T. */
        real cx = cos(d[3]), sx = sin(d[3]), cy = cos(d[4]), sy = sin(d[4]), cz = cos(d[5]), sz = sin(d[5]);
        real Xt[9] = { 1, 0, 0, 0, cx, sx, 0, -sx, cx }, Yt[9] = { cy, 0, -sy, 0, 1, 0, sy, 0, cy }, Zt[9] = { cz, sz, 0, -sz, cz, 0, 0, 0, 1 }, ZY[9];
        m3mul(ZY, Zt, Yt); m3mul(Rm, ZY, Xt);
      }
      m3mul(c->goal_ori, Rm, e->site_xmat[rb->eef_site]);
    }
    for (int k = 0; k < 3; k++) c->goal_pos[k] = e->site_xpos[rb->eef_site][k] + d[k];
  } else if (rb->ctrl_type == RSB_CTRL_JOINT_VELOCITY) {
    for (int k = 0; k < 7; k++) {
      real v = scale_action(rb, k, action[k]);
      if (rb->has_velocity_limits) { if (v < rb->velocity_limit_lo[k]) v = rb->velocity_limit_lo[k]; if (v > rb->velocity_limit_hi[k]) v = rb->velocity_limit_hi[k]; }
      c->goal_vel[k] = v;
    }
  } else if (rb->ctrl_type == RSB_CTRL_JOINT_POSITION) {      /* robosuite JointPositionController.set_goal: goal_qpos = joint_pos + scaled delta (kept in goal_vel) */
    for (int k = 0; k < 7; k++) c->goal_vel[k] = e->qpos[rb->arm_qposadr[k]] + scale_action(rb, k, action[k]);
  } else { for (int k = 0; k < 7; k++) c->goal_vel[k] = scale_action(rb, k, action[k]); }   /* JOINT_TORQUE: goal torque */
}

static void controller_run(orc_env *e, int ri) {
  const rsb_robot *rb = &e->t.robot[ri]; orc_ctrl *c = &e->rc[ri]; real tau[7];
  if (rb->ctrl_type == RSB_CTRL_OSC_POSE || rb->ctrl_type == RSB_CTRL_OSC_POSITION) {
    real J[42], Ma[49], Mi[49], v6[6]; arm_quantities(e, rb, J, Ma, Mi, v6);
    const real *ep = e->site_xpos[rb->eef_site], *Rc = e->site_xmat[rb->eef_site], *Rd = c->goal_ori;
    real F[6]; real eo[3] = { 0, 0, 0 };
    for (int k = 0; k < 3; k++) { real a[3] = { Rc[k], Rc[3 + k], Rc[6 + k] }, b[3] = { Rd[k], Rd[3 + k], Rd[6 + k] }, x[3]; v3cross(x, a, b); for (int q = 0; q < 3; q++) eo[q] += 0.5 * x[q]; }
    for (int k = 0; k < 3; k++) { F[k] = rb->kp[k] * (c->goal_pos[k] - ep[k]) - rb->kd[k] * v6[k]; F[3 + k] = rb->kp[3 + k] * eo[k] - rb->kd[3 + k] * v6[3 + k]; }
    /* JMi = J M^-1 (6x7); Lfi = J M^-1 J^T */
    real JMi[42], Lfi[36], Lf[36], Lp[36], Lo[36], T3[36];
    for (int r = 0; r < 6; r++) for (int cc = 0; cc < 7; cc++) { real s = 0; for (int k = 0; k < 7; k++) s += J[r * 7 + k] * Mi[k * 7 + cc]; JMi[r * 7 + cc] = s; }
    for (int r = 0; r < 6; r++) for (int cc = 0; cc < 6; cc++) { real s = 0; for (int k = 0; k < 7; k++) s += JMi[r * 7 + k] * J[cc * 7 + k]; Lfi[r * 6 + cc] = s; }
    sym_pinv(Lfi, Lf, 6);
    { real E[36], V[36], lo = 1e300, hi = 0; memcpy(E, Lfi, sizeof E); jacobi_eig(E, V, 6);
      for (int k = 0; k < 6; k++) { real l = fabs(E[k * 6 + k]); if (l < lo) lo = l; if (l > hi) hi = l; }
      c->osc_cond = hi / (lo > 1e-300 ? lo : 1e-300); }
    for (int r = 0; r < 3; r++) for (int cc = 0; cc < 3; cc++) T3[r * 6 + cc] = Lfi[r * 6 + cc];
    sym_pinv(T3, Lp, 3);
    for (int r = 0; r < 3; r++) for (int cc = 0; cc < 3; cc++) T3[r * 6 + cc] = Lfi[(3 + r) * 6 + 3 + cc];
    sym_pinv(T3, Lo, 3);
    real w[6];
    if (rb->uncouple_pos_ori) for (int r = 0; r < 3; r++) { real a = 0, b = 0; for (int k = 0; k < 3; k++) { a += Lp[r * 6 + k] * F[k]; b += Lo[r * 6 + k] * F[3 + k]; } w[r] = a; w[3 + r] = b; }
    else for (int r = 0; r < 6; r++) { real s = 0; for (int k = 0; k < 6; k++) s += Lf[r * 6 + k] * F[k]; w[r] = s; }
    if (rb->ctrl_type == RSB_CTRL_OSC_POSITION) {
      /* position-only OSC: orientation held at the goal captured at reset (robosuite use_ori=False keeps a fixed ori goal) */
    }
    for (int k = 0; k < 7; k++) { real s = 0; for (int r = 0; r < 6; r++) s += J[r * 7 + k] * w[r]; tau[k] = s + e->qfrc_bias[rb->arm_dofadr[k]]; }
    /* nullspace: N = I - Jbar J, Jbar = M^-1 J^T Lf ;  tau += N^T M (kn (q0-q) - 2 sqrt(kn) qd) */
    real Jbar[42], pose[7], pt[7], kn = rb->nullspace_kp, kv = 2 * sqrt(kn);
    for (int r = 0; r < 7; r++) for (int cc = 0; cc < 6; cc++) { real s = 0; for (int k = 0; k < 6; k++) s += JMi[k * 7 + r] * Lf[k * 6 + cc]; Jbar[r * 6 + cc] = s; }
    for (int k = 0; k < 7; k++) pose[k] = kn * (c->initial_joint[k] - e->qpos[rb->arm_qposadr[k]]) - kv * e->qvel[rb->arm_dofadr[k]];
    for (int r = 0; r < 7; r++) { real s = 0; for (int k = 0; k < 7; k++) s += Ma[r * 7 + k] * pose[k]; pt[r] = s; }
    for (int cc = 0; cc < 7; cc++) {        /* (N^T pt)[cc] = pt[cc] - sum_r (Jbar J)[r][cc] pt[r] */
      real s = pt[cc]; for (int r = 0; r < 7; r++) { real jj = 0; for (int k = 0; k < 6; k++) jj += Jbar[r * 6 + k] * J[k * 7 + cc]; s -= jj * pt[r]; }
      tau[cc] += s;
    }
  } else if (rb->ctrl_type == RSB_CTRL_JOINT_VELOCITY) {
    real err[7], raw[7], avg[7] = { 0 };
    for (int k = 0; k < 7; k++) { err[k] = c->goal_vel[k] - e->qvel[rb->arm_dofadr[k]]; c->derr_buf[c->derr_ptr][k] = err[k] - c->last_err[k]; c->last_err[k] = err[k]; }
    c->derr_ptr = (c->derr_ptr + 1) % 5; if (c->derr_n < 5) c->derr_n++;
    /* robosuite RingBuffer.average divides by the buffer LENGTH once full, by the fill count before */
    for (int i = 0; i < 5; i++) for (int k = 0; k < 7; k++) avg[k] += c->derr_buf[i][k] / c->derr_n;
    if (!c->saturated) for (int k = 0; k < 7; k++) c->summed_err[k] += err[k];
    int sat = 0;
    for (int k = 0; k < 7; k++) {
      raw[k] = rb->kp[k] * err[k] + rb->ki[k] * c->summed_err[k] + rb->kd[k] * avg[k] + e->qfrc_bias[rb->arm_dofadr[k]];
      tau[k] = raw[k]; if (tau[k] < rb->torque_limit_lo[k]) tau[k] = rb->torque_limit_lo[k]; if (tau[k] > rb->torque_limit_hi[k]) tau[k] = rb->torque_limit_hi[k];
      if (tau[k] != raw[k]) sat = 1;
    }
    c->saturated = sat;
  } else if (rb->ctrl_type == RSB_CTRL_JOINT_POSITION) {
    /* robosuite v1.0 JointPositionController.run_controller: desired = kp (goal_qpos - q) + kd (-qd); torques = mass_matrix . desired + torque_compensation,
       mass_matrix = the arm's block of the joint-space inertia */
    real des[7];
    for (int k = 0; k < 7; k++) des[k] = rb->kp[k] * (c->goal_vel[k] - e->qpos[rb->arm_qposadr[k]]) - rb->kd[k] * e->qvel[rb->arm_dofadr[k]];
    for (int k = 0; k < 7; k++) { real t = e->qfrc_bias[rb->arm_dofadr[k]]; for (int j = 0; j < 7; j++) t += e->M[rb->arm_dofadr[k] * e->m.nv + rb->arm_dofadr[j]] * des[j]; tau[k] = t; }
  } else {
    for (int k = 0; k < 7; k++) tau[k] = c->goal_vel[k] + e->qfrc_bias[rb->arm_dofadr[k]];
  }
  for (int k = 0; k < 7; k++) {
    real t = tau[k]; if (t < rb->torque_limit_lo[k]) t = rb->torque_limit_lo[k]; if (t > rb->torque_limit_hi[k]) t = rb->torque_limit_hi[k];
    c->torques_raw[k] = tau[k]; c->torques[k] = t; e->ctrl[rb->arm_act[k]] = t;
  }
}

static void gripper_action(orc_env *e, int ri, real g) {
  const rsb_robot *rb = &e->t.robot[ri]; orc_ctrl *c = &e->rc[ri]; const rsb_model *m = &e->m;
  real sg = g > 0 ? 1 : (g < 0 ? -1 : 0);
  for (int k = 0; k < rb->grip_ndof; k++) {
    real v = c->grip_cur[k] + rb->grip_sign[k] * rb->grip_speed * sg; if (v > 1) v = 1; if (v < -1) v = -1; c->grip_cur[k] = v;
    int a = rb->grip_act[k]; real lo = m->act_ctrlrange[2 * a], hi = m->act_ctrlrange[2 * a + 1];
    e->ctrl[a] = 0.5 * (hi + lo) + 0.5 * (hi - lo) * v;
  }
}

/* ------------------------------------------------------------------ position / velocity stages */
static void fwd_position(orc_env *e) { kinematics(e); make_cdof(e); mass_matrix(e); collision(e); }

void orc_forward(orc_env *e) {           /* everything that depends on (qpos, qvel) only */
  fwd_position(e); bias_forces(e); make_constraint(e);
}

/* one physics substep: robosuite `_pre_action` (controller) + mj_step (A.1.3, A.3) */
static void substep(orc_env *e, const real *action, int policy_step) {
  orc_forward(e);
  int off = 0;
  for (int ri = 0; ri < e->t.nrobot; ri++) {
    const rsb_robot *rb = &e->t.robot[ri];
    if (policy_step) controller_set_goal(e, ri, action + off);
    controller_run(e, ri);
    if (rb->grip_action_dim > 0) gripper_action(e, ri, action[off + rb->control_dim]);
    off += rb->control_dim + rb->grip_action_dim;
  }
  actuation(e); acceleration(e); solve_constraints(e); euler(e);
}

/* ------------------------------------------------------------------ A.6 rewards, A.1.4 observations */
static int geom_in(const int *set, int n, int g) { for (int i = 0; i < n; i++) if (set[i] == g) return 1; return 0; }
/* both fingers of robot ri touch a geom of the object (geom ids lo..hi) */
static int check_grasp_range(const orc_env *e, int ri, int lo, int hi) {
  const rsb_robot *rb = &e->t.robot[ri]; int tl = 0, tr = 0;
  for (int c = 0; c < e->ncon; c++) {
    int g1 = e->con[c].geom1, g2 = e->con[c].geom2, o1 = g1 >= lo && g1 <= hi, o2 = g2 >= lo && g2 <= hi;
    if ((geom_in(rb->left_finger_geoms, rb->n_left_finger_geoms, g1) && o2) || (geom_in(rb->left_finger_geoms, rb->n_left_finger_geoms, g2) && o1)) tl = 1;
    if ((geom_in(rb->right_finger_geoms, rb->n_right_finger_geoms, g1) && o2) || (geom_in(rb->right_finger_geoms, rb->n_right_finger_geoms, g2) && o1)) tr = 1;
  }
  return tl && tr;
}
static int check_grasp(const orc_env *e, int ri, int obj_geom) { return check_grasp_range(e, ri, obj_geom, obj_geom); }

/* TwoArmPegInHole._compute_orientation: v = peg axis (the peg body's z), c = hole centre (plate origin + offset along the plate's x); t = (c - p) . v, the
   signed distance of the centre along the axis from the peg's origin; d = |v x (p - c)|, the distance of the centre from the axis; cosn = |n . v| with n the plate's z */
static void peg_hole_orientation(const orc_env *e, real *t_out, real *d_out, real *cos_out) {
  const rsb_task *t = &e->t; const real *hp = e->xpos[t->obj_body[0]], *Rh = e->xmat[t->obj_body[0]], *pp = e->xpos[t->obj_body[1]], *Rp = e->xmat[t->obj_body[1]];
  real v[3] = {Rp[2], Rp[5], Rp[8]}, n[3] = {Rh[2], Rh[5], Rh[8]}, c[3], pc[3], cp[3], x[3];
  for (int k = 0; k < 3; k++) { c[k] = hp[k] + t->task_par[0] * Rh[3 * k]; cp[k] = c[k] - pp[k]; pc[k] = pp[k] - c[k]; }
  v3cross(x, v, pc);
  *t_out = v3dot(cp, v) / v3dot(v, v); *d_out = v3norm(x) / v3norm(v); *cos_out = fabs(v3dot(n, v) / v3norm(n) / v3norm(v));
}

static real task_reward(const orc_env *e) {
  const rsb_task *t = &e->t; real r = 0;
  if (t->task_id == RSB_TASK_LIFT) {
    const real *cube = e->xpos[t->obj_body[0]], *eef = e->site_xpos[t->robot[0].eef_site];
    if (cube[2] > t->table_height + 0.04) r = 2.25;
    else if (t->reward_shaping) {
      real d[3]; v3sub(d, eef, cube); r += 1 - tanh(10.0 * v3norm(d));
      if (check_grasp(e, 0, t->obj_geom[0])) r += 0.25;
    }
    return r * t->reward_scale / 2.25;
  }
  if (t->task_id == RSB_TASK_STACK) {
    const real *A = e->xpos[t->obj_body[0]], *B = e->xpos[t->obj_body[1]], *eef = e->site_xpos[t->robot[0].eef_site];
    real d[3]; v3sub(d, eef, A); real dist = v3norm(d);
    int grasp = check_grasp(e, 0, t->obj_geom[0]);
    real r_reach = (1 - tanh(10.0 * dist)) * 0.25 + (grasp ? 0.25 : 0);
    int lifted = A[2] > t->table_height + 0.04;
    real r_lift = lifted ? 1.0 : 0.0;
    if (lifted) { real h = sqrt((A[0] - B[0]) * (A[0] - B[0]) + (A[1] - B[1]) * (A[1] - B[1])); r_lift += 0.5 * (1 - tanh(h)); }
    int touch = 0; for (int c = 0; c < e->ncon; c++) { int g1 = e->con[c].geom1, g2 = e->con[c].geom2; if ((g1 == t->obj_geom[0] && g2 == t->obj_geom[1]) || (g2 == t->obj_geom[0] && g1 == t->obj_geom[1])) touch = 1; }
    real r_stack = (!grasp && r_lift > 0 && touch) ? 2.0 : 0.0;
    if (t->reward_shaping) { r = r_reach; if (r_lift > r) r = r_lift; if (r_stack > r) r = r_stack; }
    else r = r_stack > 0 ? 2.0 : 0.0;
    return r * t->reward_scale / 2.0;
  }
  if (t->task_id == RSB_TASK_DOOR) {
    real hinge = e->qpos[t->obj_qposadr[0]], handle = e->qpos[t->obj_qposadr[1]];
    if (hinge > 0.3) r = 1.0;
    else if (t->reward_shaping) {
      const real *eef = e->site_xpos[t->robot[0].eef_site], *hs = e->site_xpos[t->obj_site[0]]; real d[3]; v3sub(d, eef, hs);
      r += 0.25 * (1 - tanh(10.0 * v3norm(d)));
      real hr = 0.25 * fabs(handle / (0.5 * M_PI)); if (hr > 0.25) hr = 0.25; r += hr;
    }
    return r * t->reward_scale / 1.0;
  }
  if (t->task_id == RSB_TASK_TWOARMLIFT) {
    /* A.6: tilt gate cos(angle(z_pot, z)) >= cos 30 deg; success pot bottom > table + 0.10 -> 3 * gate; else lift shaping 10 * gate *
       clamp(elevation - 0.05, 0, 0.15) + per arm (0.25 if grasping its handle else 0.5 (1 - tanh(10 d))); scaled by 1/3 */
    const real *pot = e->xpos[t->obj_body[0]], *Rp = e->xmat[t->obj_body[0]];
    real cos_z = Rp[8], gate = cos_z >= cos(M_PI / 6.0) ? 1.0 : 0.0;
    real bottom = pot[2] - t->obj_half[0][2], elev = bottom - t->table_height;
    if (elev > 0.10) r = 3.0 * gate;
    else if (t->reward_shaping) {
      real lift = elev - 0.05; if (lift < 0) lift = 0; if (lift > 0.15) lift = 0.15; r += 10.0 * gate * lift;
      for (int ri = 0; ri < 2; ri++) {
        const real *eef = e->site_xpos[t->robot[ri].eef_site], *hs = e->site_xpos[t->obj_site[ri]]; real d[3]; v3sub(d, eef, hs);
        if (check_grasp(e, ri, t->obj_geom[ri])) r += 0.25; else r += 0.5 * (1 - tanh(10.0 * v3norm(d)));
      }
    }
    return r * t->reward_scale / 3.0;
  }
  if (t->task_id == RSB_TASK_PICKPLACE) {
    /* robosuite v1.0 PickPlace.reward / staged_rewards / not_in_bin in single-object mode (the other three objects sit cleared away at x = 10 and contribute nothing):
       success (object inside its bin-2 quadrant, 0 < z - bin2_z < 0.1, and the gripper away: 1 - tanh(10 d) < 0.6) -> 1; otherwise, when shaping, the maximum of
       reach 0.1 (1 - tanh(10 d)), grasp 0.35 (both fingers touch), lift 0.35 + 0.15 (1 - tanh(15 max(z_target - z, 0))) while grasped, and hover
       (0.5 when above the quadrant, else the lift reward) + 0.2 (1 - tanh(10 |xy - target|)).  The constants are pinned by the committed PickPlace runs' logs:
       reward plateaus at 0.35 and 0.5, maxima 0.63 and exactly 1.0.  Not divided by 4: single-object mode. */
    const real *obj = e->xpos[t->obj_body[0]], *eef = e->site_xpos[t->robot[0].eef_site], *tp = t->task_par;
    real d[3]; v3sub(d, eef, obj); real reach = 1 - tanh(10.0 * v3norm(d));
    int above = fabs(obj[0] - tp[0]) < tp[3] / 4 && fabs(obj[1] - tp[1]) < tp[4] / 4;
    int in_bin = above && obj[2] > tp[2] && obj[2] < tp[2] + 0.1;
    if (in_bin && reach < 0.6) r = 1.0;
    else if (t->reward_shaping) {
      real r_reach = 0.1 * reach, r_grasp = check_grasp(e, 0, t->obj_geom[0]) ? 0.35 : 0.0, r_lift = 0;
      if (r_grasp > 0) { real zd = tp[2] + tp[5] - obj[2]; if (zd < 0) zd = 0; r_lift = 0.35 + (1 - tanh(15.0 * zd)) * 0.15; }
      real hd = sqrt((obj[0] - tp[0]) * (obj[0] - tp[0]) + (obj[1] - tp[1]) * (obj[1] - tp[1]));
      real r_hover = (above ? 0.5 : r_lift) + (1 - tanh(10.0 * hd)) * 0.2;
      r = r_reach; if (r_grasp > r) r = r_grasp; if (r_lift > r) r = r_lift; if (r_hover > r) r = r_hover;
    }
    return r * t->reward_scale;
  }
  if (t->task_id == RSB_TASK_PEGINHOLE) {
    /* robosuite v1.0 TwoArmPegInHole.reward: success (d < 0.06, -0.12 <= t <= 0.14, cos > 0.95) 1, plus, when shaping, 1 - tanh(|peg - plate origin|),
       1 - tanh(d), 1 - tanh(|t|) and cos; sparse reward x 5; all scaled by reward_scale / 5.  The committed runs log a maximum of 0.98 = (1 + 0.9 + 3) / 5:
       the peg's origin at the hole's centre is 0.1 from the plate's origin, 1 - tanh(0.1) = 0.9 */
    real tt, d, cs; peg_hole_orientation(e, &tt, &d, &cs);
    if (d < 0.06 && tt >= -0.12 && tt <= 0.14 && cs > 0.95) r = 1.0;
    if (t->reward_shaping) {
      real dv[3]; v3sub(dv, e->xpos[t->obj_body[1]], e->xpos[t->obj_body[0]]);
      r += (1 - tanh(v3norm(dv))) + (1 - tanh(d)) + (1 - tanh(fabs(tt))) + cs;
    } else r *= 5.0;
    return r * t->reward_scale / 5.0;
  }
  if (t->task_id == RSB_TASK_NUTASSEMBLY) {
    /* robosuite v1.0 NutAssembly.reward / staged_rewards / on_peg in single-object mode: success (nut centre within 0.03 of its peg in x and y, below table + 0.05,
       gripper away: 1 - tanh(10 |eef - nut|) < 0.6) -> 1; otherwise, when shaping, the maximum of reach 0.1 (1 - tanh(10 |eef - handle geom|)), grasp 0.35 (both
       fingers touch any geom of the nut), lift 0.35 + 0.15 (1 - tanh(15 max(z_target - z, 0))) while grasped, hover = lift + 0.2 (1 - tanh(10 |xy - peg|)).
       Same plateaus in the committed NutAssemblyRound logs: 0.35, 0.5, up to 0.7, exactly 1.0. */
    const real *nut = e->xpos[t->obj_body[0]], *eef = e->site_xpos[t->robot[0].eef_site], *handle = e->geom_xpos[t->obj_geom[1]], *tp = t->task_par;
    real d[3]; v3sub(d, eef, nut);
    int on_peg = fabs(nut[0] - tp[0]) < 0.03 && fabs(nut[1] - tp[1]) < 0.03 && nut[2] < tp[2] + 0.05;
    if (on_peg && 1 - tanh(10.0 * v3norm(d)) < 0.6) r = 1.0;
    else if (t->reward_shaping) {
      v3sub(d, eef, handle);
      real r_reach = 0.1 * (1 - tanh(10.0 * v3norm(d))), r_grasp = check_grasp_range(e, 0, t->obj_geom[0], t->obj_geom[1]) ? 0.35 : 0.0, r_lift = 0;
      if (r_grasp > 0) { real zd = tp[3] - nut[2]; if (zd < 0) zd = 0; r_lift = 0.35 + (1 - tanh(15.0 * zd)) * 0.15; }
      real hd = sqrt((nut[0] - tp[0]) * (nut[0] - tp[0]) + (nut[1] - tp[1]) * (nut[1] - tp[1]));
      real r_hover = r_lift + (1 - tanh(10.0 * hd)) * 0.2;
      r = r_reach; if (r_grasp > r) r = r_grasp; if (r_lift > r) r = r_lift; if (r_hover > r) r = r_hover;
    }
    return r * t->reward_scale;
  }
  if (t->task_id == RSB_TASK_HANDOFF) {
    /* robosuite v1.0 TwoArmHandoff.reward with the stage values the committed runs log (x 1/2: reach <= 0.125, plateau at exactly 0.25, 0.5 .. 0.625, 1.0): hammer not
       lifted: arm 0 grasping any hammer geom -> 0.5, else 0.25 (1 - tanh |gripper0 - handle|); lifted (handle's underside > table + 0.1): arm 1 grasping the handle ->
       2.0 once arm 0 has let go (1.5 while both hold: this value is NOT pinned, no run logs it), else 1.0 + 0.25 (1 - tanh |gripper1 - handle|).  Sparse: 2.0 on
       success only.  All x reward_scale / 2. */
    const real *handle = e->geom_xpos[t->obj_geom[0]], *e0 = e->site_xpos[t->robot[0].eef_site], *e1 = e->site_xpos[t->robot[1].eef_site];
    int g0 = check_grasp_range(e, 0, t->obj_geom[0], t->obj_geom[1]), g1 = check_grasp(e, 1, t->obj_geom[0]);
    int lifted = handle[2] - t->task_par[1] - t->table_height > t->task_par[0];
    real d[3];
    if (t->reward_shaping) {
      if (lifted) {
        if (g1) r = g0 ? 1.5 : 2.0;
        else { v3sub(d, handle, e1); r = 1.0 + 0.25 * (1 - tanh(v3norm(d))); }
      } else if (g0) r = 0.5;
      else { v3sub(d, handle, e0); r = 0.25 * (1 - tanh(v3norm(d))); }
    } else r = (lifted && g1 && !g0) ? 2.0 : 0.0;
    return r * t->reward_scale / 2.0;
  }
  return 0;
}

static void put_quat_xyzw(real *o, const real *q) { o[0] = q[1]; o[1] = q[2]; o[2] = q[3]; o[3] = q[0]; }

static void observation(const orc_env *e, real *obs) {
  const rsb_task *t = &e->t; int n = 0;
  for (int ri = 0; ri < t->nrobot; ri++) {
    const rsb_robot *rb = &t->robot[ri];
    for (int k = 0; k < 7; k++) obs[n++] = sin(e->qpos[rb->arm_qposadr[k]]);
    for (int k = 0; k < 7; k++) obs[n++] = cos(e->qpos[rb->arm_qposadr[k]]);
    for (int k = 0; k < 7; k++) obs[n++] = e->qvel[rb->arm_dofadr[k]];
    for (int k = 0; k < 3; k++) obs[n++] = e->site_xpos[rb->eef_site][k];
    put_quat_xyzw(obs + n, e->xquat[rb->eef_body]); n += 4;
    for (int k = 0; k < rb->grip_ndof; k++) obs[n++] = e->qpos[rb->grip_qposadr[k]];
    for (int k = 0; k < rb->grip_ndof; k++) obs[n++] = e->qvel[rb->grip_dofadr[k]];
  }
  const real *eef = e->site_xpos[t->robot[0].eef_site];
  if (t->task_id == RSB_TASK_LIFT) {
    const real *cube = e->xpos[t->obj_body[0]];
    for (int k = 0; k < 3; k++) obs[n++] = cube[k];
    put_quat_xyzw(obs + n, e->xquat[t->obj_body[0]]); n += 4;
    for (int k = 0; k < 3; k++) obs[n++] = eef[k] - cube[k];
  } else if (t->task_id == RSB_TASK_STACK) {
    const real *A = e->xpos[t->obj_body[0]], *B = e->xpos[t->obj_body[1]];
    for (int k = 0; k < 3; k++) obs[n++] = A[k];
    put_quat_xyzw(obs + n, e->xquat[t->obj_body[0]]); n += 4;
    for (int k = 0; k < 3; k++) obs[n++] = B[k];
    put_quat_xyzw(obs + n, e->xquat[t->obj_body[1]]); n += 4;
    for (int k = 0; k < 3; k++) obs[n++] = eef[k] - A[k];
    for (int k = 0; k < 3; k++) obs[n++] = eef[k] - B[k];
    for (int k = 0; k < 3; k++) obs[n++] = A[k] - B[k];
  } else if (t->task_id == RSB_TASK_DOOR) {
    const real *door = e->xpos[t->obj_body[0]], *hs = e->site_xpos[t->obj_site[0]];
    for (int k = 0; k < 3; k++) obs[n++] = door[k];
    for (int k = 0; k < 3; k++) obs[n++] = hs[k];
    for (int k = 0; k < 3; k++) obs[n++] = door[k] - eef[k];
    for (int k = 0; k < 3; k++) obs[n++] = hs[k] - eef[k];
    obs[n++] = e->qpos[t->obj_qposadr[0]]; obs[n++] = e->qpos[t->obj_qposadr[1]];
  } else if (t->task_id == RSB_TASK_TWOARMLIFT) {
    const real *pot = e->xpos[t->obj_body[0]], *e0 = e->site_xpos[t->robot[0].eef_site], *e1 = e->site_xpos[t->robot[1].eef_site];
    const real *h0 = e->site_xpos[t->obj_site[0]], *h1 = e->site_xpos[t->obj_site[1]];
    for (int k = 0; k < 3; k++) obs[n++] = pot[k];
    put_quat_xyzw(obs + n, e->xquat[t->obj_body[0]]); n += 4;
    for (int k = 0; k < 3; k++) obs[n++] = e0[k];
    for (int k = 0; k < 3; k++) obs[n++] = e1[k];
    for (int k = 0; k < 3; k++) obs[n++] = h0[k];
    for (int k = 0; k < 3; k++) obs[n++] = h1[k];
    for (int k = 0; k < 3; k++) obs[n++] = h0[k] - e0[k];
    for (int k = 0; k < 3; k++) obs[n++] = h1[k] - e1[k];
  } else if (t->task_id == RSB_TASK_PICKPLACE || t->task_id == RSB_TASK_NUTASSEMBLY) {
    /* object-state of robosuite v1.0 PickPlace._get_observation, single-object mode: {obj}_pos, {obj}_quat (xyzw), then the object's pose in the gripper frame
       (pose_inv(eef pose) * object pose: {obj}_to_eef_pos = R_eef^T (p_obj - p_eef), {obj}_to_eef_quat = mat2quat(R_eef^T R_obj), xyzw with w >= 0) */
    const real *obj = e->xpos[t->obj_body[0]], *qe = e->xquat[t->robot[0].eef_body], *qo = e->xquat[t->obj_body[0]];
    for (int k = 0; k < 3; k++) obs[n++] = obj[k];
    put_quat_xyzw(obs + n, qo); n += 4;
    real d[3], Re[9], rel[3]; v3sub(d, obj, eef); q2mat(Re, qe); m3Tmulv(rel, Re, d);
    for (int k = 0; k < 3; k++) obs[n++] = rel[k];
    real qc[4] = {qe[0], -qe[1], -qe[2], -qe[3]}, qr[4]; qmul(qr, qc, qo); qnormalize(qr);
    if (qr[0] < 0) for (int k = 0; k < 4; k++) qr[k] = -qr[k];
    put_quat_xyzw(obs + n, qr); n += 4;
  } else if (t->task_id == RSB_TASK_PEGINHOLE) {
    /* hole_pos, hole_quat, cyl_to_hole = peg - hole, cyl_quat, angle (= cos), t, d */
    const real *hp = e->xpos[t->obj_body[0]], *pp = e->xpos[t->obj_body[1]];
    for (int k = 0; k < 3; k++) obs[n++] = hp[k];
    put_quat_xyzw(obs + n, e->xquat[t->obj_body[0]]); n += 4;
    for (int k = 0; k < 3; k++) obs[n++] = pp[k] - hp[k];
    put_quat_xyzw(obs + n, e->xquat[t->obj_body[1]]); n += 4;
    real tt, d, cs; peg_hole_orientation(e, &tt, &d, &cs);
    obs[n++] = cs; obs[n++] = tt; obs[n++] = d;
  } else if (t->task_id == RSB_TASK_HANDOFF) {
    /* hammer_pos, hammer_quat, handle_xpos, robot0 eef_xpos, robot1 eef_xpos, gripper0_to_handle, gripper1_to_handle (handle - eef) */
    const real *hp = e->xpos[t->obj_body[0]], *handle = e->geom_xpos[t->obj_geom[0]], *e0 = e->site_xpos[t->robot[0].eef_site], *e1 = e->site_xpos[t->robot[1].eef_site];
    for (int k = 0; k < 3; k++) obs[n++] = hp[k];
    put_quat_xyzw(obs + n, e->xquat[t->obj_body[0]]); n += 4;
    for (int k = 0; k < 3; k++) obs[n++] = handle[k];
    for (int k = 0; k < 3; k++) obs[n++] = e0[k];
    for (int k = 0; k < 3; k++) obs[n++] = e1[k];
    for (int k = 0; k < 3; k++) obs[n++] = handle[k] - e0[k];
    for (int k = 0; k < 3; k++) obs[n++] = handle[k] - e1[k];
  }
}

/* ------------------------------------------------------------------ Philox4x32-10 (Random123) and reset */
static void philox4x32(uint32_t c[4], const uint32_t key[2]) {
  uint32_t k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3; k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}
void orc_philox(uint64_t seed, uint64_t env_id, uint32_t stream, uint32_t index, uint32_t out[4]) {
  uint32_t c[4] = { (uint32_t)env_id, (uint32_t)(env_id >> 32), stream, index }, k[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
  philox4x32(c, k); memcpy(out, c, sizeof c);
}
static real u01(uint32_t x) { return ((real)x + 0.5) * (1.0 / 4294967296.0); }

/* stream ids: 0 = reset draws (index = episode*8 + block), 1 = synthetic actions (index = step*4 + block) */
void orc_reset(orc_env *e, uint64_t seed, uint64_t env_id, uint64_t episode) {
  const rsb_model *m = &e->m; const rsb_task *t = &e->t;
  for (int i = 0; i < m->nq; i++) e->qpos[i] = m->qpos0[i];
  for (int i = 0; i < m->nv; i++) { e->qvel[i] = 0; e->qacc_warmstart[i] = 0; }
  for (int i = 0; i < m->nu; i++) e->ctrl[i] = 0;
  e->episode = episode; e->timestep = 0; e->done = 0;
  uint32_t r[4];
  for (int ri = 0; ri < t->nrobot; ri++) {
    const rsb_robot *rb = &t->robot[ri]; real z[8];
    for (int blk = 0; blk < 2; blk++) {        /* 8 gaussians from 8 uniforms (Box-Muller) */
      orc_philox(seed, env_id, 0, (uint32_t)(episode * 8 + (uint64_t)(ri * 2 + blk)), r);
      for (int p = 0; p < 2; p++) { real u1 = u01(r[2 * p]), u2 = u01(r[2 * p + 1]), rad = sqrt(-2 * log(u1)); z[4 * blk + 2 * p] = rad * cos(2 * M_PI * u2); z[4 * blk + 2 * p + 1] = rad * sin(2 * M_PI * u2); }
    }
    for (int k = 0; k < 7; k++) e->qpos[rb->arm_qposadr[k]] = rb->init_qpos[k] + t->init_noise * z[k];
    for (int k = 0; k < rb->grip_ndof; k++) e->qpos[rb->grip_qposadr[k]] = rb->grip_init_qpos[k];
  }
  /* object placement: uniform xy + yaw per object (robosuite UniformRandomSampler; rejection for overlaps) */
  for (int o = 0; o < RSB_MAX_OBJ; o++) {
    if ((t->obj_qposadr[o] < 0 && t->place_body[o] < 0) || t->place_z[o] <= 0) continue;
    int qa = t->obj_qposadr[o]; real x = 0, y = 0, yaw = 0;
    for (int attempt = 0; attempt < 16; attempt++) {
      orc_philox(seed, env_id, 0, (uint32_t)(episode * 8 + 4 + (uint64_t)o) + 0x10000u * (uint32_t)attempt, r);
      x = t->place_x[o][0] + (t->place_x[o][1] - t->place_x[o][0]) * u01(r[0]);
      y = t->place_y[o][0] + (t->place_y[o][1] - t->place_y[o][0]) * u01(r[1]);
      yaw = t->place_yaw[o][0] + (t->place_yaw[o][1] - t->place_yaw[o][0]) * u01(r[2]);
      int ok = 1;
      for (int p = 0; p < o; p++) if (t->obj_qposadr[p] >= 0 && t->place_body[o] < 0 && t->place_z[p] > 0) {
        real dx = x + t->place_ref[0] - e->qpos[t->obj_qposadr[p]], dy = y + t->place_ref[1] - e->qpos[t->obj_qposadr[p] + 1];
        real rr = sqrt(t->obj_half[o][0] * t->obj_half[o][0] + t->obj_half[o][1] * t->obj_half[o][1]) + sqrt(t->obj_half[p][0] * t->obj_half[p][0] + t->obj_half[p][1] * t->obj_half[p][1]);
        if (dx * dx + dy * dy < rr * rr) ok = 0;
      }
      if (ok) break;
    }
    real *dst = t->place_body[o] >= 0 ? e->bpose : e->qpos + qa;
    dst[0] = t->place_ref[0] + x; dst[1] = t->place_ref[1] + y; dst[2] = t->place_z[o];
    dst[3] = cos(0.5 * yaw); dst[4] = 0; dst[5] = 0; dst[6] = sin(0.5 * yaw);
    if (t->task_id == RSB_TASK_HANDOFF) {         /* the hammer is laid down by a quarter turn about the world x axis, head towards robot 0 or robot 1 (fourth word of the draw) */
      real sg = (r[3] & 1u) ? -1.0 : 1.0; dst[4] = sg * dst[6]; dst[6] = 0;
    }
  }
  orc_forward(e);
  for (int ri = 0; ri < t->nrobot; ri++) controller_reset(e, ri);
}

/* synthetic action stream shared with the CUDA library: a = tanh(N(0,1)) keyed (seed, env, step, dim) */
void orc_random_action(const orc_env *e, uint64_t seed, uint64_t env_id, uint64_t step, real *action) {
  int n = e->t.act_dim; uint32_t r[4];
  for (int blk = 0; 4 * blk < n; blk++) {
    orc_philox(seed, env_id, 1, (uint32_t)(step * 4 + (uint64_t)blk), r);
    for (int p = 0; p < 2; p++) {
      real u1 = u01(r[2 * p]), u2 = u01(r[2 * p + 1]), rad = sqrt(-2 * log(u1));
      real z0 = rad * cos(2 * M_PI * u2), z1 = rad * sin(2 * M_PI * u2);
      if (4 * blk + 2 * p < n) action[4 * blk + 2 * p] = tanh(z0);
      if (4 * blk + 2 * p + 1 < n) action[4 * blk + 2 * p + 1] = tanh(z1);
    }
  }
}

/* ------------------------------------------------------------------ public step / state access */
int orc_step(orc_env *e, const real *action, real *obs, real *reward) {
  if (e->done) return -1;                              /* robosuite raises ValueError("executing action in terminated episode") */
  e->timestep++;
  for (int s = 0; s < e->t.substeps; s++) substep(e, action, s == 0);
  orc_forward(e);                                      /* observations / reward read the post-step kinematics and contacts */
  *reward = task_reward(e); observation(e, obs);
  e->done = (e->timestep >= e->t.horizon) && !e->t.ignore_done;
  return e->done;
}

void orc_substep(orc_env *e, const real *action, int policy_step) { substep(e, action, policy_step); }
void orc_observe(orc_env *e, real *obs, real *reward) { orc_forward(e); *reward = task_reward(e); observation(e, obs); }

/* controller state layout per robot (doubles): goal_pos3 goal_ori9 initial_joint7 grip_cur2 goal_vel7 summed_err7 last_err7
   derr_buf35 derr_n derr_ptr saturated  = 80 */
#define CS 80
void orc_get_state(const orc_env *e, real *qpos, real *qvel, real *warm, real *cs) {
  memcpy(qpos, e->qpos, sizeof(real) * (size_t)e->m.nq); memcpy(qvel, e->qvel, sizeof(real) * (size_t)e->m.nv);
  if (warm) memcpy(warm, e->qacc_warmstart, sizeof(real) * (size_t)e->m.nv);
  if (cs) for (int ri = 0; ri < e->t.nrobot; ri++) {
    const orc_ctrl *c = &e->rc[ri]; real *o = cs + CS * ri;
    memcpy(o, c->goal_pos, 3 * sizeof(real)); memcpy(o + 3, c->goal_ori, 9 * sizeof(real)); memcpy(o + 12, c->initial_joint, 7 * sizeof(real));
    memcpy(o + 19, c->grip_cur, 2 * sizeof(real)); memcpy(o + 21, c->goal_vel, 7 * sizeof(real)); memcpy(o + 28, c->summed_err, 7 * sizeof(real));
    memcpy(o + 35, c->last_err, 7 * sizeof(real)); memcpy(o + 42, c->derr_buf, 35 * sizeof(real)); o[77] = c->derr_n; o[78] = c->derr_ptr; o[79] = c->saturated;
  }
}
void orc_set_state(orc_env *e, const real *qpos, const real *qvel, const real *warm, const real *cs) {
  memcpy(e->qpos, qpos, sizeof(real) * (size_t)e->m.nq); memcpy(e->qvel, qvel, sizeof(real) * (size_t)e->m.nv);
  if (warm) memcpy(e->qacc_warmstart, warm, sizeof(real) * (size_t)e->m.nv); else memset(e->qacc_warmstart, 0, sizeof e->qacc_warmstart);
  if (cs) for (int ri = 0; ri < e->t.nrobot; ri++) {
    orc_ctrl *c = &e->rc[ri]; const real *o = cs + CS * ri;
    memcpy(c->goal_pos, o, 3 * sizeof(real)); memcpy(c->goal_ori, o + 3, 9 * sizeof(real)); memcpy(c->initial_joint, o + 12, 7 * sizeof(real));
    memcpy(c->grip_cur, o + 19, 2 * sizeof(real)); memcpy(c->goal_vel, o + 21, 7 * sizeof(real)); memcpy(c->summed_err, o + 28, 7 * sizeof(real));
    memcpy(c->last_err, o + 35, 7 * sizeof(real)); memcpy(c->derr_buf, o + 42, 35 * sizeof(real)); c->derr_n = (int)o[77]; c->derr_ptr = (int)o[78]; c->saturated = (int)o[79];
  }
  e->done = 0;
}
void orc_get_bpose(const orc_env *e, real *out) { memcpy(out, e->bpose, sizeof e->bpose); }
void orc_set_bpose(orc_env *e, const real *in) { memcpy(e->bpose, in, sizeof e->bpose); }
void orc_set_timestep(orc_env *e, int t) { e->timestep = t; e->done = 0; }

/* named getter for tests: returns element count written (doubles) or -1 */
int orc_get(orc_env *e, const char *name, double *out) {
  const rsb_model *m = &e->m; int nv = m->nv;
#define RET(ptr, n) do { memcpy(out, ptr, sizeof(double) * (size_t)(n)); return (int)(n); } while (0)
  if (!strcmp(name, "M")) RET(e->M, nv * nv);
  if (!strcmp(name, "qfrc_bias")) RET(e->qfrc_bias, nv);
  if (!strcmp(name, "qfrc_passive")) RET(e->qfrc_passive, nv);
  if (!strcmp(name, "qfrc_actuator")) RET(e->qfrc_actuator, nv);
  if (!strcmp(name, "qfrc_smooth")) RET(e->qfrc_smooth, nv);
  if (!strcmp(name, "qfrc_constraint")) RET(e->qfrc_constraint, nv);
  if (!strcmp(name, "qacc_smooth")) RET(e->qacc_smooth, nv);
  if (!strcmp(name, "qacc")) RET(e->qacc, nv);
  if (!strcmp(name, "ctrl")) RET(e->ctrl, m->nu);
  if (!strcmp(name, "xpos")) { for (int b = 0; b < m->nbody; b++) memcpy(out + 3 * b, e->xpos[b], 3 * sizeof(double)); return 3 * m->nbody; }
  if (!strcmp(name, "xquat")) { for (int b = 0; b < m->nbody; b++) memcpy(out + 4 * b, e->xquat[b], 4 * sizeof(double)); return 4 * m->nbody; }
  if (!strcmp(name, "xipos")) { for (int b = 0; b < m->nbody; b++) memcpy(out + 3 * b, e->xipos[b], 3 * sizeof(double)); return 3 * m->nbody; }
  if (!strcmp(name, "geom_xpos")) { for (int g = 0; g < m->ngeom; g++) memcpy(out + 3 * g, e->geom_xpos[g], 3 * sizeof(double)); return 3 * m->ngeom; }
  if (!strcmp(name, "site_xpos")) { for (int s = 0; s < m->nsite; s++) memcpy(out + 3 * s, e->site_xpos[s], 3 * sizeof(double)); return 3 * m->nsite; }
  if (!strcmp(name, "site_xmat")) { for (int s = 0; s < m->nsite; s++) memcpy(out + 9 * s, e->site_xmat[s], 9 * sizeof(double)); return 9 * m->nsite; }
  if (!strcmp(name, "efc_J")) { for (int r = 0; r < e->nefc; r++) memcpy(out + (size_t)r * nv, e->efc_J + (size_t)r * MAXV, sizeof(double) * (size_t)nv); return e->nefc * nv; }
  if (!strcmp(name, "efc_aref")) RET(e->efc_aref, e->nefc);
  if (!strcmp(name, "efc_R")) RET(e->efc_R, e->nefc);
  if (!strcmp(name, "efc_D")) RET(e->efc_D, e->nefc);
  if (!strcmp(name, "efc_force")) RET(e->efc_force, e->nefc);
  if (!strcmp(name, "efc_pos")) RET(e->efc_pos, e->nefc);
  if (!strcmp(name, "efc_type")) { for (int r = 0; r < e->nefc; r++) out[r] = e->efc_type[r]; return e->nefc; }
  if (!strcmp(name, "contact_geoms")) { for (int c = 0; c < e->ncon; c++) { out[2 * c] = e->con[c].geom1; out[2 * c + 1] = e->con[c].geom2; } return 2 * e->ncon; }
  if (!strcmp(name, "contact_pos")) { for (int c = 0; c < e->ncon; c++) memcpy(out + 3 * c, e->con[c].pos, 3 * sizeof(double)); return 3 * e->ncon; }
  if (!strcmp(name, "contact_frame")) { for (int c = 0; c < e->ncon; c++) memcpy(out + 9 * c, e->con[c].frame, 9 * sizeof(double)); return 9 * e->ncon; }
  if (!strcmp(name, "contact_dist")) { for (int c = 0; c < e->ncon; c++) out[c] = e->con[c].dist; return e->ncon; }
  if (!strcmp(name, "contact_mu")) { for (int c = 0; c < e->ncon; c++) out[c] = e->con[c].mu; return e->ncon; }
  if (!strcmp(name, "osc_cond")) { for (int ri = 0; ri < e->t.nrobot; ri++) out[ri] = e->rc[ri].osc_cond; return e->t.nrobot; }
  if (!strcmp(name, "torques_raw")) { for (int ri = 0; ri < e->t.nrobot; ri++) memcpy(out + 7 * ri, e->rc[ri].torques_raw, 7 * sizeof(double)); return 7 * e->t.nrobot; }
  if (!strcmp(name, "torques")) { for (int ri = 0; ri < e->t.nrobot; ri++) memcpy(out + 7 * ri, e->rc[ri].torques, 7 * sizeof(double)); return 7 * e->t.nrobot; }
  if (!strcmp(name, "counts")) { out[0] = e->ncon; out[1] = e->nefc; out[2] = e->solver_iter; out[3] = e->timestep; return 4; }
  return -1;
}

/* stage-level entry points for invariant tests */
void orc_fwd_actuation(orc_env *e) { actuation(e); acceleration(e); }
void orc_fwd_constraint(orc_env *e) { solve_constraints(e); }
double orc_cost(orc_env *e, const double *qacc) { return total_cost(e, qacc, NULL); }
void orc_set_ctrl(orc_env *e, const double *ctrl) { memcpy(e->ctrl, ctrl, sizeof(double) * (size_t)e->m.nu); }
void orc_euler(orc_env *e) { euler(e); }
int orc_box_box(const double *pa, const double *Ra, const double *ha, const double *pb, const double *Rb, const double *hb, double margin, double *out /* [8][7] pos,normal,dist */) {
  rawcon rc[16]; int n = box_box(pa, Ra, ha, pb, Rb, hb, margin, rc);
  for (int i = 0; i < n; i++) { memcpy(out + 7 * i, rc[i].pos, 3 * sizeof(double)); memcpy(out + 7 * i + 3, rc[i].normal, 3 * sizeof(double)); out[7 * i + 6] = rc[i].dist; }
  return n;
}
int orc_sizeof_model(void) { return (int)sizeof(rsb_model); }
int orc_sizeof_task(void) { return (int)sizeof(rsb_task); }
