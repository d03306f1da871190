/*
 * rsb_devmodel.h -- the fp32/int32 model the kernels read, and the host-side builder that flattens
 * include/rsb_model.h (`rsb_model`, `rsb_task`, doubles) into it.
 *
 * Stands where mujoco-py hands robosuite a compiled `MjSim` (reference call site
 * util/rlkit_utils.py:49-56 `suite.make`): compile once, upload once, every env of the batch shares it.
 * The struct is copied to __constant__ memory before a launch (rsb_cuda.cu bind_model): scalars and the per-env shared-memory
 * layout are read through the constant bank, arrays are device pointers into one arena in HBM (L1-resident: 98 % hit rate).
 */
#ifndef RSB_DEVMODEL_H
#define RSB_DEVMODEL_H

#include "../../include/rsb_model.h"

#define RSB_MAXDIM 4            /* contact condim supported: 1, 3, 4 */
#define RSB_CONW 10             /* words per contact record in shared memory */
#define RSB_CS_WORDS 80         /* controller state words per robot (same layout as the oracle's get/set_state) */

/* int arrays */
#define RSB_DM_INT_ARRAYS(X) \
  X(body_parent) X(body_root) X(body_jntadr) X(body_jntnum) X(body_dofadr) X(body_dofnum) X(body_lastdof) X(body_dofmask) \
  X(jnt_type) X(jnt_qadr) X(jnt_dadr) X(jnt_body) \
  X(dof_body) X(dof_jnt) X(dof_parent) X(dof_kind) X(dof_velstart) X(dof_velmask) X(dof_root) \
  X(mpair_i) X(mpair_j) X(tri_ij) \
  X(geom_type) X(geom_body) X(site_body) \
  X(pair_g1) X(pair_g2) X(pair_dim) X(pair_dm1) X(pair_dm2) \
  X(act_dof) X(act_climited) X(act_flimited) \
  X(fl_dof) X(lim_jnt)
/* float arrays */
#define RSB_DM_FLT_ARRAYS(X) \
  X(body_pos) X(body_quat) X(body_ipos) X(body_imat) X(body_mass) X(body_inertia) X(body_invw) \
  X(jnt_pos) X(jnt_axis) X(jnt_range) X(jnt_stiffness) X(jnt_margin) X(jnt_solref) X(jnt_solimp) \
  X(dof_armature) X(dof_damping) X(dof_floss) X(dof_invw) X(dof_solref) X(dof_solimp) \
  X(qpos0) X(qpos_spring) \
  X(geom_size) X(geom_pos) X(geom_mat) X(geom_rbound) X(site_pos) X(site_mat) \
  X(pair_friction) X(pair_solref) X(pair_solimp) X(pair_margin) X(pair_gap) X(pair_invw) X(pair_kb) X(dof_kb) X(jnt_kb) \
  X(act_gain) X(act_bias) X(act_crange) X(act_frange) X(act_gear)

enum { RSB_DOF_HINGE = 0, RSB_DOF_SLIDE = 1, RSB_DOF_FREE_T = 2, RSB_DOF_FREE_R = 3 };

typedef struct DevRobot {
  int arm_qadr[RSB_ARM_DOF], arm_dadr[RSB_ARM_DOF], arm_act[RSB_ARM_DOF];
  int grip_ndof, grip_qadr[2], grip_dadr[2], grip_act[2], grip_action_dim;
  float grip_sign[2], grip_speed, grip_init[2];
  int eef_site, eef_body;
  float init_qpos[RSB_ARM_DOF];
  int lfg[RSB_MAX_FINGER_GEOMS], nlfg, rfg[RSB_MAX_FINGER_GEOMS], nrfg;
  int ctrl_type, control_dim;
  float in_max[RSB_ARM_DOF], in_min[RSB_ARM_DOF], out_max[RSB_ARM_DOF], out_min[RSB_ARM_DOF];
  float kp[RSB_ARM_DOF], kd[RSB_ARM_DOF], ki[RSB_ARM_DOF], null_kp;
  int uncouple;
  int ori_mode;                /* RSB_ORI_DELTA_* */
  float tl_lo[RSB_ARM_DOF], tl_hi[RSB_ARM_DOF], vl_lo[RSB_ARM_DOF], vl_hi[RSB_ARM_DOF];
  int has_vl;
  int act_off;                 /* offset of this robot's slice in the action vector */
} DevRobot;

typedef struct DevModel {
  /* sizes */
  int nq, nv, nu, nbody, njnt, ngeom, nsite, npair, nmpair, nfl, nlimj, ncon_max, nefc_max;
  int ldm, ldj;                /* ldm: unused (M, H and the factor are PACKED lower triangles, entry (i,j), j<=i, at i(i+1)/2+j); ldj: row length of J (nv|1) */
  int frame_cache;             /* 1: contact tangents cached in the ew array during the constraint stage (3 ncon_max <= nefc_max) */
  int lockstep;                /* bit k: CTA barrier after stage k (all warps of a CTA fetch the same code together) */
  int cs_words;                /* controller state words per robot held in shared memory (21 for OSC laws, RSB_CS_WORDS otherwise) */
  int ntri, nvsh;              /* lower-triangle entry count nv(nv+1)/2 (table tri_ij = i<<8|j); log2 of the power of two >= nv */
  float timestep, gravity[3], impratio, meaninertia;
  int cone, any_damping, solver_iters, ls_iters;
  float solver_tol, ls_tol;     /* Newton: scaled gradient / improvement tolerance; line search: |slope| <= ls_tol * |slope at 0| */
  /* task */
  int task_id, nrobot, horizon, substeps, ignore_done, reward_shaping, obs_dim, act_dim;
  float reward_scale, init_noise, table_height;
  int obj_body[RSB_MAX_OBJ], obj_geom[RSB_MAX_OBJ], obj_site[RSB_MAX_OBJ], obj_qadr[RSB_MAX_OBJ], obj_dadr[RSB_MAX_OBJ];
  float obj_half[RSB_MAX_OBJ][3], place_x[RSB_MAX_OBJ][2], place_y[RSB_MAX_OBJ][2], place_yaw[RSB_MAX_OBJ][2];
  float place_z[RSB_MAX_OBJ], place_ref[3], task_par[RSB_TASK_NPAR];
  int place_body[RSB_MAX_OBJ], override_body;       /* fixed body whose pose is a per-env quantity (Door), -1 if none */
  DevRobot robot[RSB_MAX_ROBOTS];
  /* per-env persistent state record in HBM (words): qpos, qvel, warm, cs[nrobot*RSB_CS_WORDS], bpose[7], timestep, episode */
  int st_qpos, st_qvel, st_warm, st_cs, st_bpose, st_time, st_episode, st_words;
  /* per-env shared-memory layout (word offsets) */
  int o_qpos, o_qvel, o_warm, o_ctrl, o_cs, o_act, o_bpose;
  int o_xpos, o_xquat, o_xmat, o_xanchor, o_xaxis, o_jq;
  int o_cinert, o_crb, o_cdof, o_fi, o_M, o_L;
  int o_gxpos, o_gxmat, o_sxpos, o_sxmat;
  int o_cvel, o_cacc, o_cdofdot;
  int o_bias, o_passive, o_actuator, o_smooth, o_qacc_smooth, o_qacc, o_qfc, o_grad, o_search, o_Mv, o_tmpv;
  int o_con, o_J, o_epos, o_emargin, o_eR, o_eD, o_earef, o_efloss, o_ejar, o_eJv, o_eforce, o_ew, o_etype, o_eid, o_Hc;
  int o_cscr, o_tau;           /* controller scratch; last arm torques (kept for parity checks) */
  int o_misc;                  /* ncon, nefc, iters, ... (8 words) */
  int smem_words;
  unsigned int *counters;      /* per-batch event counters in HBM (may be null): [0] control steps of an env that dropped a contact beyond ncon_max,
                                  [1] ... that dropped a constraint row beyond nefc_max, [2] steps asked of a terminated episode */
#define X(n) const int *n;
  RSB_DM_INT_ARRAYS(X)
#undef X
#define X(n) const float *n;
  RSB_DM_FLT_ARRAYS(X)
#undef X
} DevModel;

/* debug record written by rsb_debug_substep / the emulator (floats):
   [0] ncon [1] nefc [2] solver iters, then at fixed offsets below */
#define RSB_DBG_M 8                                /* nv*nv dense */
#define RSB_DBG_WORDS(nv, ncon_max, nefc_max) (8 + (nv) * (nv) + 8 * (nv) + 14 + (ncon_max) * 16 + (nefc_max) * 6 + (nefc_max) * (nv))

#ifndef __CUDACC_RTC__
#ifdef __cplusplus
#include <vector>
#include <cstring>
#include <cmath>
#include <string>

/* Host-side arena: all arrays concatenated; pointers in DevModel are fixed up against a base address. */
struct RsbHostModel {
  DevModel dm;                              /* pointers hold BYTE OFFSETS into `arena` until fixup() */
  std::vector<unsigned char> arena;
  std::string error;
};

namespace rsbdm {
inline size_t push_i(RsbHostModel &h, const std::vector<int> &v) {
  size_t off = h.arena.size(); size_t n = v.size() ? v.size() : 1;
  h.arena.resize(off + ((n * 4 + 15) / 16) * 16, 0);
  if (v.size()) memcpy(h.arena.data() + off, v.data(), v.size() * 4);
  return off;
}
inline size_t push_f(RsbHostModel &h, const std::vector<float> &v) {
  size_t off = h.arena.size(); size_t n = v.size() ? v.size() : 1;
  h.arena.resize(off + ((n * 4 + 15) / 16) * 16, 0);
  if (v.size()) memcpy(h.arena.data() + off, v.data(), v.size() * 4);
  return off;
}
inline std::vector<int> vi(const int *p, int n) { return std::vector<int>(p, p + (n > 0 ? n : 0)); }
inline std::vector<float> vf(const double *p, int n) { std::vector<float> o((size_t)(n > 0 ? n : 0)); for (int i = 0; i < n; i++) o[(size_t)i] = (float)p[i]; return o; }
inline void q2m(const double *q, double *R) {
  double w = q[0], x = q[1], y = q[2], z = q[3];
  R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - w * z); R[2] = 2 * (x * z + w * y);
  R[3] = 2 * (x * y + w * z); R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - w * x);
  R[6] = 2 * (x * z - w * y); R[7] = 2 * (y * z + w * x); R[8] = 1 - 2 * (x * x + y * y);
}
}  // namespace rsbdm

/* Build the flattened model.  Returns false (and sets h.error) when the model uses a feature the kernels do not
   implement -- never silently ignored. */
inline bool rsb_build_host_model(const rsb_model *m, const rsb_task *t, int ncon_max, int nefc_max, RsbHostModel &h) {
  using namespace rsbdm;
  DevModel &d = h.dm; memset(&d, 0, sizeof d); h.arena.clear(); h.error.clear();
  if (m->nv > 32) { h.error = "nv > 32 unsupported"; return false; }
  if (m->cone != RSB_CONE_ELLIPTIC) { h.error = "only cone=elliptic is implemented"; return false; }
  for (int p = 0; p < m->npair; p++) {
    int dm_ = m->pair_condim[p];
    if (!(dm_ == 1 || dm_ == 3 || dm_ == 4)) { h.error = "condim must be 1, 3 or 4"; return false; }
    int t1 = m->geom_type[m->pair_geom1[p]], t2 = m->geom_type[m->pair_geom2[p]];
    bool ok = (t1 == RSB_GEOM_PLANE && (t2 == RSB_GEOM_BOX || t2 == RSB_GEOM_SPHERE || t2 == RSB_GEOM_CAPSULE)) ||
              (t1 == RSB_GEOM_SPHERE && (t2 == RSB_GEOM_SPHERE || t2 == RSB_GEOM_BOX)) ||
              (t1 == RSB_GEOM_CAPSULE && (t2 == RSB_GEOM_CAPSULE || t2 == RSB_GEOM_BOX)) || (t1 == RSB_GEOM_BOX && t2 == RSB_GEOM_BOX);
    if (!ok) { h.error = "unsupported geom type pair in candidate contact list"; return false; }
  }
  d.nq = m->nq; d.nv = m->nv; d.nu = m->nu; d.nbody = m->nbody; d.njnt = m->njnt; d.ngeom = m->ngeom; d.nsite = m->nsite; d.npair = m->npair;
  d.ncon_max = ncon_max; d.nefc_max = nefc_max; d.frame_cache = (3 * ncon_max <= nefc_max) ? 1 : 0; d.ldm = m->nv | 1; d.ldj = m->nv | 1;
  d.timestep = (float)m->timestep; for (int k = 0; k < 3; k++) d.gravity[k] = (float)m->gravity[k];
  d.impratio = (float)m->impratio; d.meaninertia = (float)m->meaninertia; d.cone = m->cone;
  /* <option iterations tolerance ls_iterations ls_tolerance> of the compiled model (MuJoCo's mjOption fields): never hard-coded here.  The fp32
     working budget (12 Newton iterations, improvement tolerance 1e-6, 24 line-search steps) is an explicit override applied by the HOST to the
     model it passes in (environments.py `solver=` argument, documented in DESIGN.md 4.1), not a property of this builder. */
  if (m->iterations < 1 || m->ls_iterations < 1 || !(m->tolerance >= 0) || !(m->ls_tolerance > 0)) { h.error = "model option: iterations / ls_iterations / tolerance / ls_tolerance out of range"; return false; }
  d.solver_iters = m->iterations; d.ls_iters = m->ls_iterations; d.solver_tol = (float)m->tolerance; d.ls_tol = (float)m->ls_tolerance; d.lockstep = 0x1ff;
  /* bodies */
  std::vector<int> lastdof((size_t)m->nbody, -1), dofmask((size_t)m->nbody, 0);
  for (int b = 1; b < m->nbody; b++) {
    int p = m->body_parentid[b];
    lastdof[(size_t)b] = m->body_dofnum[b] > 0 ? m->body_dofadr[b] + m->body_dofnum[b] - 1 : lastdof[(size_t)p];
    int mask = 0; for (int k = lastdof[(size_t)b]; k >= 0; k = m->dof_parentid[k]) mask |= 1 << k;
    dofmask[(size_t)b] = mask;
  }
  std::vector<float> imat((size_t)m->nbody * 9);
  for (int b = 0; b < m->nbody; b++) { double R[9]; q2m(m->body_iquat + 4 * b, R); for (int k = 0; k < 9; k++) imat[(size_t)b * 9 + k] = (float)R[k]; }
  /* dofs */
  std::vector<int> dkind((size_t)m->nv), dvs((size_t)m->nv), droot((size_t)m->nv), fl_dof, lim_jnt, mi, mj;
  for (int j = 0; j < m->njnt; j++) {
    int a = m->jnt_dofadr[j];
    if (m->jnt_type[j] == RSB_JNT_FREE) {
      for (int k = 0; k < 3; k++) { dkind[(size_t)(a + k)] = RSB_DOF_FREE_T; dvs[(size_t)(a + k)] = -2; }
      for (int k = 3; k < 6; k++) { dkind[(size_t)(a + k)] = RSB_DOF_FREE_R; dvs[(size_t)(a + k)] = a + 2; }
    } else { dkind[(size_t)a] = m->jnt_type[j] == RSB_JNT_SLIDE ? RSB_DOF_SLIDE : RSB_DOF_HINGE; dvs[(size_t)a] = m->dof_parentid[a]; }
    if (m->jnt_limited[j] && m->jnt_type[j] != RSB_JNT_FREE) lim_jnt.push_back(j);
  }
  std::vector<int> dvm((size_t)m->nv, 0);
  for (int i = 0; i < m->nv; i++) { for (int k = dvs[(size_t)i]; k >= 0; k = m->dof_parentid[k]) dvm[(size_t)i] |= 1 << k; }
  for (int i = 0; i < m->nv; i++) {
    droot[(size_t)i] = m->body_rootid[m->dof_bodyid[i]];
    if (m->dof_frictionloss[i] > 0) fl_dof.push_back(i);
    if (m->dof_damping[i] > 0) d.any_damping = 1;
    for (int k = i; k >= 0; k = m->dof_parentid[k]) { mi.push_back(i); mj.push_back(k); }
  }
  std::vector<int> tri; for (int i = 0; i < m->nv; i++) for (int j = 0; j <= i; j++) tri.push_back((i << 8) | j);
  d.ntri = (int)tri.size(); d.nvsh = 0; while ((1 << d.nvsh) < m->nv) d.nvsh++;
  d.nfl = (int)fl_dof.size(); d.nlimj = (int)lim_jnt.size(); d.nmpair = (int)mi.size();
  std::vector<float> gmat((size_t)m->ngeom * 9), smat((size_t)m->nsite * 9);
  for (int g = 0; g < m->ngeom; g++) { double R[9]; q2m(m->geom_quat + 4 * g, R); for (int k = 0; k < 9; k++) gmat[(size_t)g * 9 + k] = (float)R[k]; }
  for (int s = 0; s < m->nsite; s++) { double R[9]; q2m(m->site_quat + 4 * s, R); for (int k = 0; k < 9; k++) smat[(size_t)s * 9 + k] = (float)R[k]; }

#define SETI(name, vec) d.name = (const int *)push_i(h, vec)
#define SETF(name, vec) d.name = (const float *)push_f(h, vec)
  SETI(body_parent, vi(m->body_parentid, m->nbody)); SETI(body_root, vi(m->body_rootid, m->nbody));
  SETI(body_jntadr, vi(m->body_jntadr, m->nbody)); SETI(body_jntnum, vi(m->body_jntnum, m->nbody));
  SETI(body_dofadr, vi(m->body_dofadr, m->nbody)); SETI(body_dofnum, vi(m->body_dofnum, m->nbody));
  SETI(body_lastdof, lastdof); SETI(body_dofmask, dofmask);
  SETI(jnt_type, vi(m->jnt_type, m->njnt)); SETI(jnt_qadr, vi(m->jnt_qposadr, m->njnt)); SETI(jnt_dadr, vi(m->jnt_dofadr, m->njnt));
  SETI(jnt_body, vi(m->jnt_bodyid, m->njnt));
  SETI(dof_body, vi(m->dof_bodyid, m->nv)); SETI(dof_jnt, vi(m->dof_jntid, m->nv)); SETI(dof_parent, vi(m->dof_parentid, m->nv));
  SETI(dof_kind, dkind); SETI(dof_velstart, dvs); SETI(dof_velmask, dvm); SETI(dof_root, droot); SETI(mpair_i, mi); SETI(mpair_j, mj); SETI(tri_ij, tri);
  SETI(geom_type, vi(m->geom_type, m->ngeom)); SETI(geom_body, vi(m->geom_bodyid, m->ngeom)); SETI(site_body, vi(m->site_bodyid, m->nsite));
  SETI(pair_g1, vi(m->pair_geom1, m->npair)); SETI(pair_g2, vi(m->pair_geom2, m->npair)); SETI(pair_dim, vi(m->pair_condim, m->npair));
  /* per-pair / per-dof / per-joint constants the constraint stage would otherwise chase through three dependent loads or recompute with
     divisions every substep: the two bodies' ancestor-dof masks, the summed inverse weights (translation, rotation), and (K, B) of the
     reference acceleration (mj_makeImpedance: functions of solref, solimp[1] and the timestep only) */
  std::vector<int> pdm1((size_t)m->npair), pdm2((size_t)m->npair); std::vector<float> pinvw((size_t)m->npair * 2), pkb((size_t)m->npair * 2), dkb((size_t)m->nv * 2), jkb((size_t)m->njnt * 2);
  auto kb = [&](const double *solref, const double *solimp, float *out) {
    float dmax = (float)solimp[1], K, B, minval = 1e-15f;
    if ((float)solref[0] > 0) { float tc = fmaxf((float)solref[0], 2 * (float)m->timestep), dr = (float)solref[1]; float k = dmax * dmax * tc * tc * dr * dr; K = 1 / fmaxf(k, minval); B = 2 / fmaxf(dmax * tc, minval); }
    else { K = -(float)solref[0] / fmaxf(dmax * dmax, minval); B = -(float)solref[1] / fmaxf(dmax, minval); }
    out[0] = K; out[1] = B; };
  for (int p = 0; p < m->npair; p++) { int b1 = m->geom_bodyid[m->pair_geom1[p]], b2 = m->geom_bodyid[m->pair_geom2[p]];
    pdm1[(size_t)p] = dofmask[(size_t)b1]; pdm2[(size_t)p] = dofmask[(size_t)b2];
    for (int o = 0; o < 2; o++) pinvw[(size_t)p * 2 + o] = (float)m->body_invweight0[2 * b1 + o] + (float)m->body_invweight0[2 * b2 + o];
    kb(m->pair_solref + 2 * p, m->pair_solimp + 5 * p, &pkb[(size_t)p * 2]); }
  for (int i = 0; i < m->nv; i++) kb(m->dof_solref + 2 * i, m->dof_solimp + 5 * i, &dkb[(size_t)i * 2]);
  for (int j = 0; j < m->njnt; j++) kb(m->jnt_solref + 2 * j, m->jnt_solimp + 5 * j, &jkb[(size_t)j * 2]);
  SETI(pair_dm1, pdm1); SETI(pair_dm2, pdm2);
  SETI(act_dof, vi(m->act_dofid, m->nu)); SETI(act_climited, vi(m->act_ctrllimited, m->nu)); SETI(act_flimited, vi(m->act_forcelimited, m->nu));
  SETI(fl_dof, fl_dof); SETI(lim_jnt, lim_jnt);
  SETF(body_pos, vf(m->body_pos, 3 * m->nbody)); SETF(body_quat, vf(m->body_quat, 4 * m->nbody)); SETF(body_ipos, vf(m->body_ipos, 3 * m->nbody));
  SETF(body_imat, imat); SETF(body_mass, vf(m->body_mass, m->nbody)); SETF(body_inertia, vf(m->body_inertia, 3 * m->nbody));
  SETF(body_invw, vf(m->body_invweight0, 2 * m->nbody));
  SETF(jnt_pos, vf(m->jnt_pos, 3 * m->njnt)); SETF(jnt_axis, vf(m->jnt_axis, 3 * m->njnt)); SETF(jnt_range, vf(m->jnt_range, 2 * m->njnt));
  SETF(jnt_stiffness, vf(m->jnt_stiffness, m->njnt)); SETF(jnt_margin, vf(m->jnt_margin, m->njnt));
  SETF(jnt_solref, vf(m->jnt_solref, 2 * m->njnt)); SETF(jnt_solimp, vf(m->jnt_solimp, 5 * m->njnt));
  SETF(dof_armature, vf(m->dof_armature, m->nv)); SETF(dof_damping, vf(m->dof_damping, m->nv)); SETF(dof_floss, vf(m->dof_frictionloss, m->nv));
  SETF(dof_invw, vf(m->dof_invweight0, m->nv)); SETF(dof_solref, vf(m->dof_solref, 2 * m->nv)); SETF(dof_solimp, vf(m->dof_solimp, 5 * m->nv));
  SETF(qpos0, vf(m->qpos0, m->nq)); SETF(qpos_spring, vf(m->qpos_spring, m->nq));
  SETF(geom_size, vf(m->geom_size, 3 * m->ngeom)); SETF(geom_pos, vf(m->geom_pos, 3 * m->ngeom)); SETF(geom_mat, gmat);
  SETF(geom_rbound, vf(m->geom_rbound, m->ngeom)); SETF(site_pos, vf(m->site_pos, 3 * m->nsite)); SETF(site_mat, smat);
  SETF(pair_friction, vf(m->pair_friction, 5 * m->npair)); SETF(pair_solref, vf(m->pair_solref, 2 * m->npair));
  SETF(pair_solimp, vf(m->pair_solimp, 5 * m->npair)); SETF(pair_margin, vf(m->pair_margin, m->npair)); SETF(pair_gap, vf(m->pair_gap, m->npair));
  SETF(pair_invw, pinvw); SETF(pair_kb, pkb); SETF(dof_kb, dkb); SETF(jnt_kb, jkb);
  SETF(act_gain, vf(m->act_gain, m->nu)); SETF(act_bias, vf(m->act_bias, 3 * m->nu)); SETF(act_crange, vf(m->act_ctrlrange, 2 * m->nu));
  SETF(act_frange, vf(m->act_forcerange, 2 * m->nu)); SETF(act_gear, vf(m->act_gear, m->nu));
#undef SETI
#undef SETF
  /* task */
  d.task_id = t->task_id; d.nrobot = t->nrobot; d.horizon = t->horizon; d.substeps = t->substeps; d.ignore_done = t->ignore_done;
  d.reward_shaping = t->reward_shaping; d.obs_dim = t->obs_dim; d.act_dim = t->act_dim;
  d.override_body = -1;
  d.reward_scale = (float)t->reward_scale; d.init_noise = (float)t->init_noise; d.table_height = (float)t->table_height;
  for (int o = 0; o < RSB_MAX_OBJ; o++) {
    d.obj_body[o] = t->obj_body[o]; d.obj_geom[o] = t->obj_geom[o]; d.obj_site[o] = t->obj_site[o]; d.obj_qadr[o] = t->obj_qposadr[o]; d.obj_dadr[o] = t->obj_dofadr[o];
    for (int k = 0; k < 3; k++) d.obj_half[o][k] = (float)t->obj_half[o][k];
    for (int k = 0; k < 2; k++) { d.place_x[o][k] = (float)t->place_x[o][k]; d.place_y[o][k] = (float)t->place_y[o][k]; d.place_yaw[o][k] = (float)t->place_yaw[o][k]; }
    d.place_z[o] = (float)t->place_z[o]; d.place_body[o] = t->place_body[o]; if (t->place_body[o] >= 0) d.override_body = t->place_body[o];
  }
  for (int k = 0; k < 3; k++) d.place_ref[k] = (float)t->place_ref[k];
  for (int k = 0; k < RSB_TASK_NPAR; k++) d.task_par[k] = (float)t->task_par[k];
  int aoff = 0;
  for (int r = 0; r < t->nrobot; r++) {
    const rsb_robot *s = &t->robot[r]; DevRobot *q = &d.robot[r];
    for (int k = 0; k < RSB_ARM_DOF; k++) {
      q->arm_qadr[k] = s->arm_qposadr[k]; q->arm_dadr[k] = s->arm_dofadr[k]; q->arm_act[k] = s->arm_act[k]; q->init_qpos[k] = (float)s->init_qpos[k];
      q->in_max[k] = (float)s->input_max[k]; q->in_min[k] = (float)s->input_min[k]; q->out_max[k] = (float)s->output_max[k]; q->out_min[k] = (float)s->output_min[k];
      q->kp[k] = (float)s->kp[k]; q->kd[k] = (float)s->kd[k]; q->ki[k] = (float)s->ki[k];
      q->tl_lo[k] = (float)s->torque_limit_lo[k]; q->tl_hi[k] = (float)s->torque_limit_hi[k];
      q->vl_lo[k] = (float)s->velocity_limit_lo[k]; q->vl_hi[k] = (float)s->velocity_limit_hi[k];
    }
    q->grip_ndof = s->grip_ndof; q->grip_action_dim = s->grip_action_dim; q->grip_speed = (float)s->grip_speed;
    for (int k = 0; k < 2; k++) { q->grip_qadr[k] = s->grip_qposadr[k]; q->grip_dadr[k] = s->grip_dofadr[k]; q->grip_act[k] = s->grip_act[k]; q->grip_sign[k] = (float)s->grip_sign[k]; q->grip_init[k] = (float)s->grip_init_qpos[k]; }
    q->eef_site = s->eef_site; q->eef_body = s->eef_body;
    q->nlfg = s->n_left_finger_geoms; q->nrfg = s->n_right_finger_geoms;
    for (int k = 0; k < RSB_MAX_FINGER_GEOMS; k++) { q->lfg[k] = s->left_finger_geoms[k]; q->rfg[k] = s->right_finger_geoms[k]; }
    q->ctrl_type = s->ctrl_type; q->control_dim = s->control_dim; q->null_kp = (float)s->nullspace_kp; q->uncouple = s->uncouple_pos_ori; q->ori_mode = s->ori_delta_mode; q->has_vl = s->has_velocity_limits;
    q->act_off = aoff; aoff += s->control_dim + s->grip_action_dim;
    if (s->grip_ndof > 2) { h.error = "gripper with more than 2 dofs unsupported"; return false; }
  }
  /* persistent state record */
  int w = 0;
  d.st_qpos = w; w += d.nq; d.st_qvel = w; w += d.nv; d.st_warm = w; w += d.nv; d.st_cs = w; w += d.nrobot * RSB_CS_WORDS; d.st_bpose = w; w += 7;
  d.st_time = w++; d.st_episode = w++; d.st_words = w;
  /* shared-memory layout.  Arrays are grouped by LIFETIME inside one physics substep, and groups that are never alive together
     share words (stage order: kinematics, inertia, crb, collision, bias, controller, actuation | constraint rows, solve, Euler):
       P  persistent over the substep (state, M, force vectors, contact list, per-row D/aref/type)
       K  body/site poses and cdof: read up to the constraint stage -> overlaid by the factor workspace L and the gradient
       A  kinematics/dynamics/controller temporaries, all dead once the smooth forces are known
       J  the dense constraint Jacobian, written by the constraint stage -> overlays A
       B  solver temporaries (per-row jar/Jv/force/weight; the constraint stage's pos/margin/R temporaries overlay jar/Jv/force)
     The whole record is sized so that 28 Lift-Panda environments fit one SM (4096 envs = one wave on 148 SMs). */
  int o = 0; int nb = d.nbody, nv = d.nv, nj = d.njnt, ne = nefc_max, nc = ncon_max;
  d.cs_words = 21;                                   /* goal_pos3 goal_ori9 initial_joint7 grip_cur2: all an OSC law keeps */
  for (int r = 0; r < t->nrobot; r++) if (t->robot[r].ctrl_type != RSB_CTRL_OSC_POSE && t->robot[r].ctrl_type != RSB_CTRL_OSC_POSITION) d.cs_words = RSB_CS_WORDS;
#define L(name, n) d.name = o; o += (n)
  L(o_qpos, d.nq); L(o_qvel, nv); L(o_warm, nv); L(o_ctrl, d.nu > 0 ? d.nu : 1); L(o_cs, d.nrobot * d.cs_words); L(o_act, d.act_dim); L(o_bpose, 7);
  L(o_M, d.ntri);
  L(o_bias, nv); L(o_actuator, nv); L(o_smooth, nv); L(o_qacc_smooth, nv); L(o_qacc, nv); L(o_qfc, nv); L(o_tau, 7 * d.nrobot);
  L(o_con, nc * RSB_CONW); L(o_eD, ne); L(o_earef, ne); L(o_etype, ne); d.o_efloss = d.o_eid = 0; L(o_misc, 8);
  /* group K: kinematic results the constraint stage still reads (contact Jacobian); dead afterwards, so the factor workspace and the
     gradient (first written after the constraint stage) overlay them */
  const int k0 = o;
  L(o_xpos, 3 * nb); L(o_sxpos, 3 * d.nsite); L(o_sxmat, 9 * d.nsite); L(o_cdof, 6 * nv);
  const int end_k = o;
  o = k0; L(o_L, d.ntri); L(o_grad, nv);
  if (o < end_k) o = end_k;
  const int r0 = o;
  /* group A */
  L(o_xquat, 4 * nb); L(o_passive, nv); L(o_gxpos, 3 * d.ngeom + 1); L(o_gxmat, 9 * d.ngeom + 1); L(o_cvel, 6 * nb);
  { const int x1 = o; int kin = (7 * nb + 1) + 2 * (3 * nj + 1); int sz = kin > 208 ? kin : 208;          /* X1: controller scratch | per-body local transforms, joint anchors/axes */
    d.o_cscr = x1; d.o_jq = x1; d.o_xanchor = x1 + 7 * nb + 1; d.o_xaxis = d.o_xanchor + 3 * nj + 1; o = x1 + sz; }
  L(o_cinert, 10 * nb); L(o_crb, 10 * nb);
  { const int x3 = o; int sz = 6 * nb + 6 * nv; if (sz < 9 * nb) sz = 9 * nb;                                          /* X3: body rotation matrices (kinematics, inertia) | crb force temporaries | RNE temporaries */
    d.o_xmat = x3; d.o_cacc = x3; d.o_cdofdot = x3 + 6 * nb; d.o_fi = x3; o = x3 + sz; }
  const int end_a = o;
  /* group J over A */
  d.o_J = r0; const int end_j = r0 + ne * d.ldj;
  o = end_a > end_j ? end_a : end_j;
  /* group B */
  d.o_Hc = 0; L(o_ejar, ne); L(o_eJv, ne); L(o_eforce, ne); L(o_ew, ne);
  L(o_search, nv); L(o_tmpv, nv); d.o_Mv = d.o_tmpv;
  d.o_epos = d.o_ejar; d.o_emargin = d.o_eJv; d.o_eR = d.o_eforce;
#undef L
  d.smem_words = (o + 3) & ~3;
  return true;
}

/* turn the byte offsets stored in the pointer members into addresses relative to `base` */
inline void rsb_fixup_pointers(DevModel &d, const void *base) {
  const unsigned char *b = (const unsigned char *)base;
#define X(n) d.n = (const int *)(b + (size_t)d.n);
  RSB_DM_INT_ARRAYS(X)
#undef X
#define X(n) d.n = (const float *)(b + (size_t)d.n);
  RSB_DM_FLT_ARRAYS(X)
#undef X
}
#endif /* __cplusplus */
#endif
#endif /* RSB_DEVMODEL_H */
