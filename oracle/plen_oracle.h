/* ORACLE -- test infrastructure, NOT product code.
 *
 * float64 CPU restatement of the reference hot path:
 *   PlenWalkEnv.step / reset / compute_observation / compute_reward / compute_done
 *   (reference: plen_bullet/src/plen_bullet/plen_env.py:558-614, :638-692, :768-871, :873-1070, :1072-1093)
 * and of the PyBullet/Bullet3 btMultiBody tick those functions call (p.stepSimulation, plen_env.py:665-667).
 *
 * PARITY UNPINNED for the physics: Bullet3/PyBullet is a third-party dependency that is absent from
 * /root/reference and un-pinned (README.md:35 says only "Pybullet"; dating evidence points at
 * PyBullet 2.6.x-2.7.x, Q1 2020 -- SURVEY.md section 8c).  Its published algorithm (Featherstone ABA in
 * btMultiBody::computeAccelerationsArticulatedBodyAlgorithmMultiDof, unit-impulse responses in
 * calcAccelerationDeltasMultiDof, row setup in btMultiBodyConstraintSolver, projected Gauss-Seidel with
 * the PhysicsServerCommandProcessor defaults) is restated here from the published sources as recalled;
 * every such behaviour is a named field of plen_oracle_config so it can be re-calibrated.
 * The env logic (observation / reward / done / counters) IS pinned line by line to plen_env.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
 * build, link, load or call this code.
 */
#ifndef PLEN_ORACLE_H
#define PLEN_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAXL 32      /* links besides the base (plen.urdf: 32 joints) */
#define ORC_NDOF 24      /* 6 base + 18 revolute */
#define ORC_NJ 18
#define ORC_NFEET 2
#define ORC_NPTS 4       /* contact points per foot (persistent-manifold capacity) */
#define ORC_MAXBOX 32    /* box colliders besides the two foot hulls (plen.urdf: torso + 30 links) */
#define ORC_MAXXP (ORC_MAXBOX * 4)   /* contact points they can produce: <= 4 per box (btBoxBoxDetector's manifold) */
#define ORC_MAXHULL 256  /* convex-hull vertices of a foot mesh (plen.urdf feet: 209 each) */
#define ORC_MAXROWS (2 * ORC_NJ + ORC_NJ + ORC_NFEET * ORC_NPTS * 6 + 3 * ORC_MAXXP)

typedef struct {
    int n_links;
    int parent[ORC_MAXL];            /* -1 = base */
    int jtype[ORC_MAXL];             /* 0 fixed, 1 revolute */
    int dof[ORC_MAXL];               /* index into the 18 joint dofs, -1 if fixed */
    double axis[ORC_MAXL][3];        /* joint axis in the child (joint) frame */
    double R_pj[ORC_MAXL][9];        /* rotation parent link frame -> joint frame at q=0 (row major) */
    double p_pj[ORC_MAXL][3];        /* joint origin in parent link frame */
    double com[ORC_MAXL][3];         /* inertial origin in link frame */
    double mass[ORC_MAXL];
    double inertia[ORC_MAXL][3];     /* diagonal, inertial frame == link axes (all inertial rpy are 0) */
    double lower[ORC_MAXL], upper[ORC_MAXL];
    double base_mass, base_com[3], base_inertia[3];
    int foot_link[ORC_NFEET];                   /* [0]=right (11), [1]=left (19) */
    double foot_pts[ORC_NFEET][ORC_NPTS][3];    /* sole contact vertices, link frame */
    double foot_break[ORC_NFEET];               /* contact breaking threshold per foot */
    /* box colliders of every link except the feet (plen.urdf:504-1274), for ground contact of knees / hands / torso ... */
    int n_boxes;
    int box_link[ORC_MAXBOX];                   /* Bullet link index, -1 = base (torso) */
    double box_center[ORC_MAXBOX][3];           /* collision origin in the link frame */
    double box_rot[ORC_MAXBOX][9];              /* collision rpy as a rotation, link frame (row major) */
    double box_half[ORC_MAXBOX][3];             /* half extents */
    /* convex-hull vertices of the two foot meshes, link frame (manifold_mode 1 only) */
    int n_hull[ORC_NFEET];
    double foot_hull[ORC_NFEET][ORC_MAXHULL][3];
} plen_oracle_model;

typedef struct {
    /* --- cited from the reference --- */
    double dt;                 /* 1/240, plen_env.py:41 */
    int substeps;              /* 4, plen_env.py:40-42 */
    int reset_ticks;           /* 8, plen_env.py:569-570 */
    double gravity_z;          /* -9.81, plen_env.py:296 */
    double start_pos[3];       /* 0,0,0.158, plen_env.py:312 */
    double motor_max_force;    /* 0.15, plen_env.py:753 */
    int joint_act;             /* plen_env.py:34,652-654 */
    double linear_damping;     /* 0 / 0.1 (joint_act), plen_env.py:472-481 */
    double angular_damping;    /* 0, plen_env.py:480 */
    double mu_lateral;         /* foot 0.8 * plane 0.8, plen_env.py:309,444 */
    double mu_spinning;        /* foot 0.1 * plane 0.8, plen_env.py:445 */
    double mu_rolling;         /* foot 0.1|0.01 * plane 0.8, plen_env.py:439-442 */
    double restitution;        /* 0.5*0.5, plen_env.py:309,481 */
    double env_lo[ORC_NJ], env_hi[ORC_NJ];   /* env_ranges, plen_env.py:148-167 */
    int max_episode_steps;     /* 500, plen_env.py:15-19 */
    /* --- [RECALL] Bullet/PyBullet defaults --- */
    double motor_kp, motor_kd; /* 0.1, 1.0: POSITION_CONTROL defaults */
    int solver_iterations;     /* 50 */
    double residual_threshold; /* 1e-7 (max squared row velocity change) */
    double erp_contact;        /* erp2 = 0.08 */
    double erp_joint;          /* erp = 0.2 (joint limits) */
    double linear_slop;        /* 1e-5 */
    double warmstart_factor;   /* 0.1 */
    double restitution_vel_threshold; /* 0.2 */
    double hull_margin;        /* 0.001: convex hull inflation */
    double max_coord_velocity; /* 100 */
    int implicit_cone;         /* 1: implicit cone friction (multibody default) */
    /* --- ground contact of the 31 non-foot colliders (every link collider hits plane.urdf in Bullet) --- */
    int link_contacts;         /* 1: on (default); 0: only the soles collide (round-1 behaviour) */
    double mu_link;            /* lateral friction link 0.5 (URDF default, [RECALL]) * plane 0.8, plen_env.py:309 */
    double restitution_base;   /* torso 0 (not in the changeDynamics loop, plen_env.py:476-481) * plane 0.5 */
    int max_contact_points;    /* at most this many box contact points per tick, deepest first (-1: no cap = Bullet; the CUDA
                                  path keeps 4 and the parity tests give the oracle the same cap) */
    /* --- EXPERIMENT (profiles/r2_physics_pin.md section 5): how the sole contact points are generated ---
     * 0 (default, what the CUDA path does): the four extreme sole corners, each in the manifold while within the breaking
     *   threshold of the ground -- the limiting 4-point set of Bullet's manifold reduction, present from the first tick.
     * 1: btConvexPlaneCollisionAlgorithm + btPersistentManifold as recalled: ONE new point per tick (the hull's support
     *   vertex towards the ground, if within the breaking threshold), merged into a cache of <= 4 points per foot (replace the
     *   nearest cached point within the threshold, else add, else sortCachedPoints: keep the deepest, maximise the area), then
     *   refreshContactPoints (drop points that separated or drifted by more than the threshold); impulses travel with the
     *   cached points. */
    int manifold_mode;
    /* manifold_mode 1: hull vertices within this distance (m) of the lowest one count as equally low and the FIRST of them in
     * the vertex list is the support vertex.  A foot that is flat on the ground (the reset pose: every sole vertex at the same
     * height) makes the search an exact tie that rounding decides -- differently in float64 and float32, and the episode that
     * follows is sensitive to it (profiles/r2_physics_pin.md section 6).  1e-7 m: far above float32 rounding of a vertex height
     * (4e-9 m), resolved by a tilt of 2e-6 rad across a sole.  0 restores the plain first-minimum search. */
    double support_tie;
} plen_oracle_config;

typedef struct {
    /* physics */
    double pos[3], quat[4] /* xyzw */, omega[3] /* world */, vel[3] /* world, base origin */;
    double q[ORC_NJ], qd[ORC_NJ], target[ORC_NJ];
    double lam_n[ORC_NFEET][ORC_NPTS];   /* cached normal impulses (warm start) */
    int in_manifold[ORC_NFEET][ORC_NPTS];
    /* env bookkeeping, plen_env.py */
    int cnt, ds, hist_len, ep_t, dead;
    double last[6], sums[9], ep_ret;
    /* diagnostics of the last tick */
    int last_iterations, last_rows, last_box_points, last_boxes_touching;
    long long flops;                     /* instrumented FLOP counter (adds+muls), whole life */
    /* manifold_mode 1: the persistent manifold of either foot (lam_n / in_manifold above are its impulses / occupancy) */
    int man_n[ORC_NFEET];
    double man_local[ORC_NFEET][ORC_NPTS][3];    /* point on the (inflated) hull, foot link frame */
    double man_world[ORC_NFEET][ORC_NPTS][3];    /* point on the plane, world frame */
} plen_oracle_state;

void plen_oracle_default_config(plen_oracle_config *cfg, int joint_act);
void plen_oracle_init_state(const plen_oracle_config *cfg, plen_oracle_state *s);
/* one Bullet tick (collision -> ABA -> rows -> PGS -> integrate) */
void plen_oracle_tick(const plen_oracle_model *m, const plen_oracle_config *cfg, plen_oracle_state *s);
/* PlenWalkEnv.reset, plen_env.py:558-614 */
void plen_oracle_reset(const plen_oracle_model *m, const plen_oracle_config *cfg, plen_oracle_state *s, double *obs26);
/* PlenWalkEnv.step + TimeLimit, plen_env.py:638-692, :15-19.  done = dead || ep_t >= max; timeout = !dead && ep_t >= max */
void plen_oracle_step(const plen_oracle_model *m, const plen_oracle_config *cfg, plen_oracle_state *s,
                      const double *action18, double *obs26, double *reward, int *done, int *timeout);
/* observation only (no history side effects) + discrete predicates, for tests */
void plen_oracle_observe(const plen_oracle_model *m, const plen_oracle_config *cfg, const plen_oracle_state *s,
                         double *obs26, double *foot_euler6);
/* world pose of every link frame (and base): pos[(n+1)*3], rot[(n+1)*9]; index 0 = base */
void plen_oracle_fk(const plen_oracle_model *m, const plen_oracle_state *s, double *pos, double *rot);
/* joint-space mass matrix M[24*24] in coordinates [omega_w, v_w, qd] via 24 unit-impulse responses (returns M^-1) */
void plen_oracle_minv(const plen_oracle_model *m, const plen_oracle_state *s, double *Minv);
/* batch: n envs, OpenMP over envs; auto_reset!=0 resets finished envs after returning their terminal obs */
void plen_oracle_step_batch(const plen_oracle_model *m, const plen_oracle_config *cfg, plen_oracle_state *s, int n,
                            const double *actions, double *obs, double *reward, int *done, int *timeout,
                            int auto_reset, int n_threads);
void plen_oracle_reset_batch(const plen_oracle_model *m, const plen_oracle_config *cfg, plen_oracle_state *s, int n,
                             double *obs, int n_threads);
int plen_oracle_sizeof_state(void);
int plen_oracle_sizeof_model(void);
int plen_oracle_sizeof_config(void);

#ifdef __cplusplus
}
#endif
#endif
