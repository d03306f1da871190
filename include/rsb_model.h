/*
 * rsb_model.h -- flat, host-side description of one compiled scene ("model") and of the
 * robosuite task wrapped around it ("task").  This is the data format that crosses the
 * C-ABI once, at rsb_create(); the CUDA library converts it to fp32 and uploads it.
 *
 * It stands where the reference reaches `suite.make(...)` -> robosuite builds MJCF ->
 * mujoco-py compiles `MjSim` (reference call site: util/rlkit_utils.py:49-56).  Field names
 * follow MuJoCo's mjModel so that a maintainer holding a real
 * mjModel can fill the struct field by field.  All arrays are host pointers owned by the
 * caller and only read during rsb_create()/oracle calls.  Quaternions are (w,x,y,z).
 */
#ifndef RSB_MODEL_H
#define RSB_MODEL_H

#ifdef __cplusplus
extern "C" {
#endif

enum { RSB_JNT_FREE = 0, RSB_JNT_SLIDE = 2, RSB_JNT_HINGE = 3 };           /* mjtJoint values */
enum { RSB_GEOM_PLANE = 0, RSB_GEOM_SPHERE = 2, RSB_GEOM_CAPSULE = 3,
       RSB_GEOM_CYLINDER = 5, RSB_GEOM_BOX = 6 };                          /* mjtGeom values  */
enum { RSB_CONE_PYRAMIDAL = 0, RSB_CONE_ELLIPTIC = 1 };
enum { RSB_TASK_LIFT = 0, RSB_TASK_DOOR = 1, RSB_TASK_STACK = 2, RSB_TASK_TWOARMLIFT = 3,
       RSB_TASK_PICKPLACE = 4,     /* PickPlace in single-object mode (PickPlaceMilk / PickPlaceCan / ...): one object, bin 1 -> its quadrant of bin 2 */
       RSB_TASK_PEGINHOLE = 5,     /* TwoArmPegInHole: peg on robot 0's hand, plate with a hole on robot 1's hand */
       RSB_TASK_NUTASSEMBLY = 6,   /* NutAssembly in single-object mode (NutAssemblyRound / NutAssemblySquare): one nut, table -> its peg */
       RSB_TASK_HANDOFF = 7 };     /* TwoArmHandoff: a hammer on a narrow table beside robot 0, to be picked up and handed to robot 1 */
enum { RSB_CTRL_OSC_POSE = 0, RSB_CTRL_OSC_POSITION = 1, RSB_CTRL_JOINT_VELOCITY = 2,
       RSB_CTRL_JOINT_TORQUE = 3,
       RSB_CTRL_JOINT_POSITION = 4 };   /* goal = q + scaled action at the policy step; torque = M_arm (kp (goal - q) - kd qd) + bias (kp, kd: 7 used) */

/* OSC_POSE goal orientation from the scaled rotation action d (3-vector):
   EULER_T    : goal = euler2mat(d)^T R_ee, euler2mat as in mujoco-py / robosuite transform_utils (first order: rotation by -d).  This is the convention the
                reference's committed 2020 policies were trained under: the committed Lift-Panda-OSC_POSE-SEED17 policy scores 328 +- 8 (logged 364, same maximum
                487) with it and 21 with the axis-angle form (tools/diag_policy_transfer2.py).
   AXIS_ANGLE : goal = quat2mat(axisangle2quat(d)) R_ee (robosuite >= 1.1: rotation by +d). */
enum { RSB_ORI_DELTA_EULER_T = 0, RSB_ORI_DELTA_AXIS_ANGLE = 1 };

#define RSB_MAX_ROBOTS 2
#define RSB_ARM_DOF 7
#define RSB_MAX_FINGER_GEOMS 4
#define RSB_MAX_OBJ 4
#define RSB_TASK_NPAR 8

typedef struct rsb_model {
  /* sizes */
  int nq, nv, nu, nbody, njnt, ngeom, nsite, npair, nM;
  /* <option> */
  double timestep, gravity[3], impratio, tolerance, ls_tolerance, meaninertia;
  int cone, iterations, ls_iterations;
  /* bodies (tree order, body 0 = world) */
  const int *body_parentid, *body_rootid, *body_jntadr, *body_jntnum, *body_dofadr, *body_dofnum;
  const double *body_pos, *body_quat, *body_ipos, *body_iquat, *body_mass, *body_inertia;
  const double *body_invweight0;          /* [nbody][2] (translation, rotation) */
  /* joints */
  const int *jnt_type, *jnt_qposadr, *jnt_dofadr, *jnt_bodyid, *jnt_limited;
  const double *jnt_pos, *jnt_axis, *jnt_range, *jnt_stiffness, *jnt_margin;
  const double *jnt_solref, *jnt_solimp;  /* limit rows: [njnt][2], [njnt][5] */
  /* dofs */
  const int *dof_bodyid, *dof_jntid, *dof_parentid, *dof_Madr;
  const double *dof_armature, *dof_damping, *dof_frictionloss, *dof_invweight0;
  const double *dof_solref, *dof_solimp;  /* frictionloss rows */
  const double *qpos0, *qpos_spring;
  /* geoms */
  const int *geom_type, *geom_bodyid;
  const double *geom_size, *geom_pos, *geom_quat, *geom_rbound;
  /* sites */
  const int *site_bodyid;
  const double *site_pos, *site_quat;
  /* statically filtered candidate geom pairs, in MuJoCo contact order
     (body pair ascending, then geom ids), with the mixed contact parameters */
  const int *pair_geom1, *pair_geom2, *pair_condim;
  const double *pair_friction;            /* [npair][5] */
  const double *pair_solref, *pair_solimp;/* [npair][2], [npair][5] */
  const double *pair_margin, *pair_gap;
  /* actuators (joint transmission only) */
  const int *act_dofid, *act_ctrllimited, *act_forcelimited;
  const double *act_gain, *act_bias;      /* gainprm[0]; biasprm[0..2] -> [nu][3] */
  const double *act_ctrlrange, *act_forcerange, *act_gear;
} rsb_model;

/* One robot arm + gripper + its controller, as robosuite's SingleArm/Controller hold them
   (SURVEY.md A.2).  Index arrays are into the model's dof/qpos/actuator spaces. */
typedef struct rsb_robot {
  int arm_qposadr[RSB_ARM_DOF], arm_dofadr[RSB_ARM_DOF], arm_act[RSB_ARM_DOF];
  int grip_ndof;                           /* gripper joints/actuators (2 for Panda/Rethink) */
  int grip_qposadr[2], grip_dofadr[2], grip_act[2];
  int grip_action_dim;                     /* 1 */
  double grip_sign[2];                     /* Panda (-1,+1); Rethink (+1,-1) */
  double grip_speed;                       /* 0.01 */
  double grip_init_qpos[2];
  int eef_site, eef_body;                  /* grip_site, hand body (for eef_quat) */
  double init_qpos[RSB_ARM_DOF];
  int left_finger_geoms[RSB_MAX_FINGER_GEOMS], n_left_finger_geoms;
  int right_finger_geoms[RSB_MAX_FINGER_GEOMS], n_right_finger_geoms;
  /* controller */
  int ctrl_type, control_dim;              /* 6 / 3 / 7 */
  double input_max[RSB_ARM_DOF], input_min[RSB_ARM_DOF];
  double output_max[RSB_ARM_DOF], output_min[RSB_ARM_DOF];
  double kp[RSB_ARM_DOF], kd[RSB_ARM_DOF]; /* OSC: 6 used (kd = 2 sqrt(kp) damping); JV: PID kp, kd per joint */
  double ki[RSB_ARM_DOF];                  /* JV integral gain (0 -> pure P law of robosuite v1.0) */
  double nullspace_kp;                     /* 10 */
  int uncouple_pos_ori;
  int ori_delta_mode;                      /* OSC_POSE: how the scaled rotation action d sets the goal orientation (RSB_ORI_DELTA_*) */
  double torque_limit_lo[RSB_ARM_DOF], torque_limit_hi[RSB_ARM_DOF];
  double velocity_limit_lo[RSB_ARM_DOF], velocity_limit_hi[RSB_ARM_DOF];
  int has_velocity_limits;
} rsb_robot;

typedef struct rsb_task {
  int task_id, nrobot;
  rsb_robot robot[RSB_MAX_ROBOTS];
  int horizon, substeps, ignore_done, reward_shaping;
  double reward_scale;
  double init_noise;                       /* gaussian std on arm init qpos (0.02) */
  double table_height;                     /* table_offset z (0.8) */
  int obs_dim, act_dim;
  /* task objects: meaning depends on task_id
     LIFT:  obj_body[0]=cube, obj_geom[0]=cube geom, obj_qposadr[0]=free joint
     DOOR:  obj_body[0]=door, [1]=latch(handle), obj_site[0]=handle site,
            obj_qposadr[0]=hinge, [1]=latch hinge, obj_body[2]=door frame root (placed)
     STACK: cubeA=0, cubeB=1
     TWOARMLIFT: obj_body[0]=pot, obj_site[0..1]=handle sites, obj_geom[0..1]=handle geoms
     PICKPLACE: obj_body[0]=the object, obj_geom[0]=its collision geom, obj_qposadr[0]=free joint; task_par below
     PEGINHOLE: obj_body[0]=plate with the hole, obj_body[1]=peg; task_par below
     HANDOFF: obj_body[0]=hammer, obj_geom[0]..obj_geom[1] = its (contiguous) geoms, the FIRST being the handle, obj_qposadr[0]=free joint; the placement angle
              (place_yaw) turns the hammer about the world x axis, with a random sign (head towards robot 0 or towards robot 1), not about z
     NUTASSEMBLY: obj_body[0]=the nut, obj_geom[0]..obj_geom[1] = its (contiguous) collision geoms, the last one being the handle, obj_qposadr[0]=free joint */
  int obj_body[RSB_MAX_OBJ], obj_geom[RSB_MAX_OBJ], obj_site[RSB_MAX_OBJ];
  int obj_qposadr[RSB_MAX_OBJ], obj_dofadr[RSB_MAX_OBJ];
  double obj_half[RSB_MAX_OBJ][3];         /* box half sizes for placement z / overlap tests */
  /* placement sampler: uniform ranges (robosuite UniformRandomSampler) */
  double place_x[RSB_MAX_OBJ][2], place_y[RSB_MAX_OBJ][2], place_yaw[RSB_MAX_OBJ][2];
  double place_z[RSB_MAX_OBJ];             /* absolute z written to the free joint */
  double place_ref[3];                     /* reference point added to sampled xy */
  /* place_body[o] >= 0: object o is a FIXED body (no free joint; robosuite writes model.body_pos/body_quat at reset, e.g. the
     Door): the sampled pose (place_ref + xy, place_z, yaw about z) overrides that body's pos/quat for this env.  At most one. */
  int place_body[RSB_MAX_OBJ];
  /* task constants, meaning by task_id (0 where unused)
     PICKPLACE: [0..1] = xy of the object's target placement in bin 2 (centre of its quadrant), [2] = bin-2 z (bin2_pos[2]),
                [3..4] = bin_size xy (the quadrant check is |obj - target| < bin_size / 4 per axis), [5] = lift target height above bin-2 z (0.25)
     PEGINHOLE: [0] = distance of the hole's centre from the plate body's origin along the plate's x axis (0.1)
     HANDOFF: [0] = height above the table top (table_height) that counts as lifted (0.1), [1] = half thickness of the handle (its centre minus this is the hammer's height)
     NUTASSEMBLY: [0..1] = xy of the nut's peg, [2] = table top z (on-peg: |nut - peg| < 0.03 per axis and nut z < [2] + 0.05), [3] = lift target z */
  double task_par[RSB_TASK_NPAR];
} rsb_task;

#ifdef __cplusplus
}
#endif
#endif /* RSB_MODEL_H */
