/* odg_model.h — plain-C model descriptor shared by the CUDA library, the CPU oracle and Python.
 *
 * One free-floating trunk + `nleg` serial leg chains of `njl` hinge joints each (bodies welded to
 * the last link, e.g. the OpenDOG paws, are fused into it by the model compiler). This is what
 * `mujoco.MjModel.from_xml_path(".../walking_scene.xml")` provides to the reference hot path
 * (reference: Code/mujoco/environments/WalkEnvironment.py:34-39, sim2real/train.py:153-154);
 * field names follow mjModel where a 1:1 field exists. All reals are double here; the CUDA library
 * narrows to fp32 when it uploads the constants.
 *
 * DoF order (MuJoCo's): 3 trunk translations (world), 3 trunk rotations (trunk frame), then legs
 * in body order, joints root-to-tip.  qpos = (x y z qw qx qy qz, hinge angles ...).
 */
#ifndef ODG_MODEL_H
#define ODG_MODEL_H

#ifdef __cplusplus
extern "C" {
#endif

#define ODG_MAX_LEG 4
#define ODG_MAX_JL 3          /* hinge joints per leg */
#define ODG_MAX_NU 12
#define ODG_MAX_NQ (7 + ODG_MAX_LEG * ODG_MAX_JL)
#define ODG_MAX_NV (6 + ODG_MAX_LEG * ODG_MAX_JL)
#define ODG_MAX_GEOM 48
#define ODG_MAX_VERT 2048
#define ODG_MAX_CON_PER_GEOM 4

/* collision geoms against the floor plane: convex hull of a mesh (mjc_PlaneConvex), sphere (mjc_PlaneSphere), and the
 * primitives of unitree_go1/go1.xml:26-60 — capsule (mjc_PlaneCapsule), cylinder (mjc_PlaneCylinder), box (mjc_PlaneBox) */
enum { ODG_GEOM_HULL = 0, ODG_GEOM_SPHERE = 1, ODG_GEOM_CAPSULE = 2, ODG_GEOM_CYLINDER = 3, ODG_GEOM_BOX = 4 };

typedef struct OdgGeom {
  int leg;                 /* owning leg, -1 = trunk */
  int link;                /* joint index within the leg whose body carries the geom (-1 for trunk) */
  int type;                /* ODG_GEOM_* */
  int vert_start, vert_count; /* hull: slice of OdgModel.vert (link frame) */
  int mj_geom_id;          /* geom id in the MuJoCo model (floor plane is 0) */
  int mj_body_id;          /* body id in the MuJoCo model (paws: 4,7,10,13) */
  int condim;              /* contact dim after mixing with the floor: max(condim_geom, condim_floor); 1, 3 or 6 */
  double center[3];        /* geom position in the link frame (sphere centre; capsule / cylinder / box centre) */
  double radius;           /* sphere radius */
  double rot[9];           /* primitives: geom frame in the link frame, row-major (columns = geom axes) */
  double size[3];          /* capsule / cylinder: radius, half-length (along the geom z axis); box: half extents */
  double friction;         /* sliding friction after mixing with the floor (element-wise max, or the higher priority's) */
  double friction_torsion; /* torsional and rolling friction (condim 6: go1.xml:61-64), same mixing */
  double friction_roll;
  double margin;           /* includemargin = max(margin) - max(gap) */
  double solref[2];        /* contact solref after mixing */
  double solimp[5];        /* contact solimp after mixing */
  double invweight0;       /* body_invweight0[mj_body_id][0] (translational) */
} OdgGeom;

typedef struct OdgModel {
  int nleg, njl, nq, nv, nu, ngeom, nvert;
  int cone;                /* 1 = elliptic (the only cone the reference models use) */
  double timestep;
  double gravity[3];
  double impratio;
  /* default solver parameters used by friction-loss and joint-limit rows */
  double dof_solref[2], dof_solimp[5];
  double lim_solref[2], lim_solimp[5];

  /* trunk */
  double base_mass, base_ipos[3], base_inertia[9];   /* inertia about COM, trunk frame, row-major */
  double base_armature[6], base_frictionloss[6], base_damping[6], base_invweight0[6];

  /* leg links [leg][joint] */
  double body_pos[ODG_MAX_LEG][ODG_MAX_JL][3];       /* link frame origin in parent frame */
  double body_quat[ODG_MAX_LEG][ODG_MAX_JL][4];
  double jnt_pos[ODG_MAX_LEG][ODG_MAX_JL][3];        /* hinge anchor in link frame */
  double jnt_axis[ODG_MAX_LEG][ODG_MAX_JL][3];       /* hinge axis in link frame (unit) */
  double jnt_range[ODG_MAX_LEG][ODG_MAX_JL][2];
  int    jnt_limited[ODG_MAX_LEG][ODG_MAX_JL];
  double armature[ODG_MAX_LEG][ODG_MAX_JL];
  double frictionloss[ODG_MAX_LEG][ODG_MAX_JL];
  double damping[ODG_MAX_LEG][ODG_MAX_JL];
  double dof_invweight0[ODG_MAX_LEG][ODG_MAX_JL];
  double mass[ODG_MAX_LEG][ODG_MAX_JL];
  double ipos[ODG_MAX_LEG][ODG_MAX_JL][3];
  double inertia[ODG_MAX_LEG][ODG_MAX_JL][9];

  /* position actuators, ctrl order */
  int    act_leg[ODG_MAX_NU], act_joint[ODG_MAX_NU];
  double act_kp[ODG_MAX_NU], act_kv[ODG_MAX_NU];
  int    act_ctrllimited[ODG_MAX_NU], act_forcelimited[ODG_MAX_NU];
  double act_ctrlrange[ODG_MAX_NU][2], act_forcerange[ODG_MAX_NU][2];

  /* keyframe 0 ("home") */
  double key_qpos[ODG_MAX_NQ], key_ctrl[ODG_MAX_NU];

  /* collision geoms that can touch the floor plane z = 0, and their hull vertices */
  OdgGeom geom[ODG_MAX_GEOM];
  double vert[ODG_MAX_VERT][3];
  /* tilt of the three extra support directions used for multi-point plane/convex contacts */
  double multicontact_tilt;
} OdgModel;

#ifdef __cplusplus
}
#endif
#endif /* ODG_MODEL_H */
