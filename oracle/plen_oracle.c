/* ORACLE -- test infrastructure, NOT product code.  See plen_oracle.h for scope and the
 * "PARITY UNPINNED" statement.  Plain C99, float64, one robot per call (OpenMP only over envs).
 *
 * Threading: none inside; oracle/oracle.py runs one slice of envs per Python thread (ctypes drops the GIL).
 *
 * Layout of this file
 *   1. small linear algebra
 *   2. kinematics + articulated-body algorithm on the 33-link Bullet tree
 *        (restates btMultiBody::computeAccelerationsArticulatedBodyAlgorithmMultiDof and
 *         calcAccelerationDeltasMultiDof; Featherstone, "Rigid Body Dynamics Algorithms", ch. 7/9)
 *   3. constraint rows + projected Gauss-Seidel
 *        (restates btMultiBodyConstraintSolver::{convertMultiBodyContact, setupMultiBodyContactConstraint,
 *         solveSingleIteration, resolveSingleConstraintRowGeneric, resolveConeFrictionConstraintRows},
 *         btMultiBodyJointMotor / btMultiBodyJointLimitConstraint::createConstraintRows)
 *   4. tick = collision -> ABA -> rows -> PGS -> integrate  (btMultiBodyDynamicsWorld, p.stepSimulation)
 *   5. env logic: line-by-line restatement of plen_env.py (citations inline)
 */
#include "plen_oracle.h"

#include <math.h>
#include <string.h>

#define FL(s, n) ((s)->flops += (n))

/* ------------------------------------------------------------------ 1. linear algebra */
static void cross3(const double *a, const double *b, double *o) {
    double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
    o[0] = x; o[1] = y; o[2] = z;
}
static double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static void mat3_mul(const double *A, const double *B, double *C) { /* C = A B, row major */
    double t[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) t[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
    memcpy(C, t, sizeof t);
}
static void mat3_vec(const double *A, const double *v, double *o) {
    double x = A[0] * v[0] + A[1] * v[1] + A[2] * v[2], y = A[3] * v[0] + A[4] * v[1] + A[5] * v[2],
           z = A[6] * v[0] + A[7] * v[1] + A[8] * v[2];
    o[0] = x; o[1] = y; o[2] = z;
}
static void mat3T_vec(const double *A, const double *v, double *o) {
    double x = A[0] * v[0] + A[3] * v[1] + A[6] * v[2], y = A[1] * v[0] + A[4] * v[1] + A[7] * v[2],
           z = A[2] * v[0] + A[5] * v[1] + A[8] * v[2];
    o[0] = x; o[1] = y; o[2] = z;
}
static void mat3_T(const double *A, double *o) {
    double t[9] = {A[0], A[3], A[6], A[1], A[4], A[7], A[2], A[5], A[8]};
    memcpy(o, t, sizeof t);
}
static void quat_to_mat(const double *q, double *R) { /* q = xyzw, R maps body -> world */
    double x = q[0], y = q[1], z = q[2], w = q[3];
    R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - w * z); R[2] = 2 * (x * z + w * y);
    R[3] = 2 * (x * y + w * z); R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - w * x);
    R[6] = 2 * (x * z - w * y); R[7] = 2 * (y * z + w * x); R[8] = 1 - 2 * (x * x + y * y);
}
static void mat_to_quat(const double *R, double *q) { /* xyzw, Shepperd */
    double tr = R[0] + R[4] + R[8];
    if (tr > 0) {
        double s = sqrt(tr + 1.0) * 2;
        q[3] = 0.25 * s; q[0] = (R[7] - R[5]) / s; q[1] = (R[2] - R[6]) / s; q[2] = (R[3] - R[1]) / s;
    } else if (R[0] > R[4] && R[0] > R[8]) {
        double s = sqrt(1.0 + R[0] - R[4] - R[8]) * 2;
        q[3] = (R[7] - R[5]) / s; q[0] = 0.25 * s; q[1] = (R[1] + R[3]) / s; q[2] = (R[2] + R[6]) / s;
    } else if (R[4] > R[8]) {
        double s = sqrt(1.0 + R[4] - R[0] - R[8]) * 2;
        q[3] = (R[2] - R[6]) / s; q[0] = (R[1] + R[3]) / s; q[1] = 0.25 * s; q[2] = (R[5] + R[7]) / s;
    } else {
        double s = sqrt(1.0 + R[8] - R[0] - R[4]) * 2;
        q[3] = (R[3] - R[1]) / s; q[0] = (R[2] + R[6]) / s; q[1] = (R[5] + R[7]) / s; q[2] = 0.25 * s;
    }
}
/* pybullet getEulerFromQuaternion (plen_env.py:799-800, :1017, :1030) [RECALL: btQuaternion::getEulerZYX with
 * the +-0.99999 gimbal branches] */
static void quat_to_euler(const double *q, double *rpy) {
    double x = q[0], y = q[1], z = q[2], w = q[3];
    double sarg = -2.0 * (x * z - w * y);
    if (sarg <= -0.99999) {
        rpy[1] = -0.5 * M_PI; rpy[0] = 0; rpy[2] = 2 * atan2(x, -y);
    } else if (sarg >= 0.99999) {
        rpy[1] = 0.5 * M_PI; rpy[0] = 0; rpy[2] = 2 * atan2(-x, y);
    } else {
        rpy[0] = atan2(2 * (y * z + w * x), w * w - x * x - y * y + z * z);
        rpy[1] = asin(sarg);
        rpy[2] = atan2(2 * (x * y + w * z), w * w + x * x - y * y - z * z);
    }
}
static void rodrigues(const double *a, double q, double *R) {
    double c = cos(q), s = sin(q), t = 1 - c, x = a[0], y = a[1], z = a[2];
    R[0] = t * x * x + c; R[1] = t * x * y - s * z; R[2] = t * x * z + s * y;
    R[3] = t * x * y + s * z; R[4] = t * y * y + c; R[5] = t * y * z - s * x;
    R[6] = t * x * z - s * y; R[7] = t * y * z + s * x; R[8] = t * z * z + c;
}
/* solve A x = b in place for SPD 6x6 via Gauss-Jordan inverse (small, well conditioned) */
static void inv6(const double *A, double *Ainv) {
    double M[6][12];
    for (int i = 0; i < 6; i++)
        for (int j = 0; j < 6; j++) { M[i][j] = A[6 * i + j]; M[i][6 + j] = (i == j); }
    for (int c = 0; c < 6; c++) {
        int piv = c;
        for (int r = c + 1; r < 6; r++) if (fabs(M[r][c]) > fabs(M[piv][c])) piv = r;
        if (piv != c) for (int j = 0; j < 12; j++) { double t = M[c][j]; M[c][j] = M[piv][j]; M[piv][j] = t; }
        double d = 1.0 / M[c][c];
        for (int j = 0; j < 12; j++) M[c][j] *= d;
        for (int r = 0; r < 6; r++) if (r != c) {
            double f = M[r][c];
            if (f != 0) for (int j = 0; j < 12; j++) M[r][j] -= f * M[c][j];
        }
    }
    for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) Ainv[6 * i + j] = M[i][6 + j];
}

/* ------------------------------------------------------------------ 2. kinematics + ABA */
typedef struct {
    /* index 0 = base, link i -> i+1 */
    double Rw[ORC_MAXL + 1][9], pw[ORC_MAXL + 1][3];
    double E[ORC_MAXL][9];                 /* parent frame -> link frame rotation */
    double v[ORC_MAXL + 1][6], c[ORC_MAXL][6], a[ORC_MAXL + 1][6];
    double IA[ORC_MAXL + 1][36], pA[ORC_MAXL + 1][6];
    double U[ORC_MAXL][6], D[ORC_MAXL], u[ORC_MAXL];
    double IA0inv[36];
    double axis_w[ORC_MAXL][3];
} work_t;

/* motion vector parent -> child: w' = E w, v' = E (v - r x w) */
static void xm(const double *E, const double *r, const double *in, double *out) {
    double t[3], l[3];
    cross3(r, in, t);
    l[0] = in[3] - t[0]; l[1] = in[4] - t[1]; l[2] = in[5] - t[2];
    mat3_vec(E, in, out);
    mat3_vec(E, l, out + 3);
}
/* force vector child -> parent, ACCUMULATED: f_p += E^T f, n_p += E^T n + r x (E^T f) */
static void xf_add(const double *E, const double *r, const double *in, double *acc) {
    double n[3], f[3], t[3];
    mat3T_vec(E, in, n);
    mat3T_vec(E, in + 3, f);
    cross3(r, f, t);
    for (int k = 0; k < 3; k++) { acc[k] += n[k] + t[k]; acc[3 + k] += f[k]; }
}
/* 6x6 motion transform X = [[E,0],[-E rx, E]] */
static void build_X(const double *E, const double *r, double *X) {
    double rx[9] = {0, -r[2], r[1], r[2], 0, -r[0], -r[1], r[0], 0}, Erx[9];
    mat3_mul(E, rx, Erx);
    memset(X, 0, 36 * sizeof(double));
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            X[6 * i + j] = E[3 * i + j];
            X[6 * (i + 3) + (j + 3)] = E[3 * i + j];
            X[6 * (i + 3) + j] = -Erx[3 * i + j];
        }
}
/* Ip += X^T Ic X */
static void inertia_to_parent_add(const double *X, const double *Ic, double *Ip) {
    double T[36];
    for (int i = 0; i < 6; i++)
        for (int j = 0; j < 6; j++) {
            double s = 0;
            for (int k = 0; k < 6; k++) s += Ic[6 * i + k] * X[6 * k + j];
            T[6 * i + j] = s;
        }
    for (int i = 0; i < 6; i++)
        for (int j = 0; j < 6; j++) {
            double s = 0;
            for (int k = 0; k < 6; k++) s += X[6 * k + i] * T[6 * k + j];
            Ip[6 * i + j] += s;
        }
}
/* rigid-body spatial inertia about the link origin, [ang;lin] ordering */
static void rigid_inertia(double m, const double *c, const double *Ic, double *I) {
    double cx[9] = {0, -c[2], c[1], c[2], 0, -c[0], -c[1], c[0], 0};
    double cc[9];
    mat3_mul(cx, cx, cc); /* [c]x[c]x */
    memset(I, 0, 36 * sizeof(double));
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            I[6 * i + j] = (i == j ? Ic[i] : 0.0) - m * cc[3 * i + j];
            I[6 * i + (3 + j)] = m * cx[3 * i + j];
            I[6 * (3 + i) + j] = -m * cx[3 * i + j];
            I[6 * (3 + i) + (3 + j)] = (i == j) ? m : 0.0;
        }
}
static void mat6_vec(const double *A, const double *v, double *o) {
    double t[6];
    for (int i = 0; i < 6; i++) {
        double s = 0;
        for (int k = 0; k < 6; k++) s += A[6 * i + k] * v[k];
        t[i] = s;
    }
    memcpy(o, t, sizeof t);
}

/* forward kinematics of link frames: fills Rw, pw, E, axis_w */
static void kinematics(const plen_oracle_model *m, const plen_oracle_state *s, work_t *w) {
    quat_to_mat(s->quat, w->Rw[0]);
    memcpy(w->pw[0], s->pos, sizeof s->pos);
    for (int i = 0; i < m->n_links; i++) {
        int p = m->parent[i] + 1;
        double Rl[9], t[3];
        if (m->jtype[i]) {
            double Rq[9];
            rodrigues(m->axis[i], s->q[m->dof[i]], Rq);
            mat3_mul(m->R_pj[i], Rq, Rl);
        } else {
            memcpy(Rl, m->R_pj[i], sizeof Rl);
        }
        mat3_T(Rl, w->E[i]);
        mat3_mul(w->Rw[p], Rl, w->Rw[i + 1]);
        mat3_vec(w->Rw[p], m->p_pj[i], t);
        for (int k = 0; k < 3; k++) w->pw[i + 1][k] = w->pw[p][k] + t[k];
        mat3_vec(w->Rw[i + 1], m->axis[i], w->axis_w[i]);
    }
}

/* Articulated-body pass: link velocities, articulated inertias, bias forces; then accelerations.
 * out[24] = generalized acceleration in coordinates [omega_w(3), v_w(3), qdd(18)]. */
static void aba_forward_dynamics(const plen_oracle_model *m, const plen_oracle_config *cfg, plen_oracle_state *s,
                                 work_t *w, double *out) {
    const int n = m->n_links;
    double wb[3], vb[3];
    kinematics(m, s, w);
    /* base spatial velocity in base coordinates (Bullet: spatVel[0] = rot_from_parent[0] * base_omega/vel) */
    mat3T_vec(w->Rw[0], s->omega, wb);
    mat3T_vec(w->Rw[0], s->vel, vb);
    memcpy(w->v[0], wb, sizeof wb); memcpy(w->v[0] + 3, vb, sizeof vb);
    const double g[3] = {0, 0, cfg->gravity_z};

    for (int b = 0; b <= n; b++) { /* b: 0 base, i+1 link i */
        const double mass = b ? m->mass[b - 1] : m->base_mass;
        const double *com = b ? m->com[b - 1] : m->base_com;
        const double *Ic = b ? m->inertia[b - 1] : m->base_inertia;
        if (b) {
            int i = b - 1, p = m->parent[i] + 1;
            xm(w->E[i], m->p_pj[i], w->v[p], w->v[b]);
            memset(w->c[i], 0, sizeof w->c[i]);
            if (m->jtype[i]) {
                double qd = s->qd[m->dof[i]], sq[3] = {m->axis[i][0] * qd, m->axis[i][1] * qd, m->axis[i][2] * qd};
                /* c = v x (S qd), motion cross product with S = [axis;0] */
                cross3(w->v[b], sq, w->c[i]);
                cross3(w->v[b] + 3, sq, w->c[i] + 3);
                for (int k = 0; k < 3; k++) w->v[b][k] += sq[k];
                /* note: c uses the child's total velocity; (S qd) x (S qd) = 0 so parent or child velocity agree */
            }
            FL(s, 60);
        }
        rigid_inertia(mass, com, Ic, w->IA[b]);
        /* bias force p = v x* (I v) - f_ext */
        double Iv[6], fe[6], F[3], vc[3], t[3];
        mat6_vec(w->IA[b], w->v[b], Iv);
        cross3(w->v[b], Iv, w->pA[b]);           /* w x n */
        cross3(w->v[b] + 3, Iv + 3, t);          /* v x f */
        for (int k = 0; k < 3; k++) w->pA[b][k] += t[k];
        cross3(w->v[b], Iv + 3, w->pA[b] + 3);   /* w x f */
        /* external: gravity at the COM (btMultiBodyDynamicsWorld adds m*g per link) + Bullet's linear drag */
        mat3T_vec(w->Rw[b], g, F);
        for (int k = 0; k < 3; k++) F[k] *= mass;
        if (cfg->linear_damping != 0.0 || cfg->angular_damping != 0.0) {
            /* [RECALL] btMultiBody: force = -k*m*v_com*(1+|v_com|), torque = -k_a*I*w*(1+|w|), K1=K2=damping */
            cross3(w->v[b], com, t);
            for (int k = 0; k < 3; k++) vc[k] = w->v[b][3 + k] + t[k];
            double nv = sqrt(dot3(vc, vc));
            for (int k = 0; k < 3; k++) F[k] -= cfg->linear_damping * mass * vc[k] * (1.0 + nv);
        }
        cross3(com, F, fe);
        memcpy(fe + 3, F, sizeof F);
        if (cfg->angular_damping != 0.0) {
            double nw = sqrt(dot3(w->v[b], w->v[b]));
            for (int k = 0; k < 3; k++) fe[k] -= cfg->angular_damping * Ic[k] * w->v[b][k] * (1.0 + nw);
        }
        for (int k = 0; k < 6; k++) w->pA[b][k] -= fe[k];
        FL(s, 72 + 27 + 30);
    }
    /* inward pass */
    for (int i = n - 1; i >= 0; i--) {
        int b = i + 1, p = m->parent[i] + 1;
        double Ia[36], pa[6], X[36];
        memcpy(Ia, w->IA[b], sizeof Ia);
        memcpy(pa, w->pA[b], sizeof pa);
        if (m->jtype[i]) {
            const double *ax = m->axis[i];
            for (int r = 0; r < 6; r++) w->U[i][r] = w->IA[b][6 * r] * ax[0] + w->IA[b][6 * r + 1] * ax[1] + w->IA[b][6 * r + 2] * ax[2];
            w->D[i] = dot3(ax, w->U[i]);
            w->u[i] = 0.0 - dot3(ax, w->pA[b]); /* tau = 0: motors are constraint rows, not torques */
            double Dinv = 1.0 / w->D[i], Iac[6];
            for (int r = 0; r < 6; r++)
                for (int q = 0; q < 6; q++) Ia[6 * r + q] -= w->U[i][r] * w->U[i][q] * Dinv;
            mat6_vec(Ia, w->c[i], Iac);
            for (int r = 0; r < 6; r++) pa[r] += Iac[r] + w->U[i][r] * (w->u[i] * Dinv);
            FL(s, 30 + 5 + 5 + 108 + 72 + 18);
        }
        build_X(w->E[i], m->p_pj[i], X);
        inertia_to_parent_add(X, Ia, w->IA[p]);
        xf_add(w->E[i], m->p_pj[i], pa, w->pA[p]);
        FL(s, 45 + 864 + 45);
    }
    /* base */
    inv6(w->IA[0], w->IA0inv);
    FL(s, 400);
    {
        double t[6];
        mat6_vec(w->IA0inv, w->pA[0], t);
        for (int k = 0; k < 6; k++) w->a[0][k] = -t[k];
    }
    /* outward pass */
    for (int i = 0; i < n; i++) {
        int b = i + 1, p = m->parent[i] + 1;
        xm(w->E[i], m->p_pj[i], w->a[p], w->a[b]);
        for (int k = 0; k < 6; k++) w->a[b][k] += w->c[i][k];
        if (m->jtype[i]) {
            double Ua = 0;
            for (int k = 0; k < 6; k++) Ua += w->U[i][k] * w->a[b][k];
            double qdd = (w->u[i] - Ua) / w->D[i];
            for (int k = 0; k < 3; k++) w->a[b][k] += m->axis[i][k] * qdd;
            out[6 + m->dof[i]] = qdd;
            FL(s, 20);
        }
        FL(s, 45);
    }
    /* base acceleration to world generalized coordinates: d/dt of (omega_w, v_w) */
    {
        double t[3], al[3];
        cross3(wb, vb, t);
        for (int k = 0; k < 3; k++) al[k] = w->a[0][3 + k] + t[k];
        mat3_vec(w->Rw[0], w->a[0], out);
        mat3_vec(w->Rw[0], al, out + 3);
    }
}

/* delta = M^-1 tau for a generalized force tau[24] (coordinates [moment about base origin (world), force (world),
 * joint torques]); reuses IA/U/D of the last aba_forward_dynamics (Bullet: calcAccelerationDeltasMultiDof). */
static void impulse_response(const plen_oracle_model *m, plen_oracle_state *s, const work_t *w, const double *tau,
                             double *delta) {
    const int n = m->n_links;
    double p[ORC_MAXL + 1][6], uu[ORC_MAXL], a[ORC_MAXL + 1][6];
    memset(p, 0, sizeof p);
    for (int i = n - 1; i >= 0; i--) {
        int b = i + 1, par = m->parent[i] + 1;
        double pa[6];
        memcpy(pa, p[b], sizeof pa);
        if (m->jtype[i]) {
            uu[i] = tau[6 + m->dof[i]] - dot3(m->axis[i], p[b]);
            double f = uu[i] / w->D[i];
            for (int r = 0; r < 6; r++) pa[r] += w->U[i][r] * f;
            FL(s, 6 + 13);
        }
        xf_add(w->E[i], m->p_pj[i], pa, p[par]);
        FL(s, 45);
    }
    {
        double fb[6], t[6];
        mat3T_vec(w->Rw[0], tau, fb);
        mat3T_vec(w->Rw[0], tau + 3, fb + 3);
        for (int k = 0; k < 6; k++) fb[k] -= p[0][k];
        mat6_vec(w->IA0inv, fb, t);
        memcpy(a[0], t, sizeof t);
        FL(s, 30 + 6 + 66);
    }
    for (int i = 0; i < n; i++) {
        int b = i + 1, par = m->parent[i] + 1;
        xm(w->E[i], m->p_pj[i], a[par], a[b]);
        if (m->jtype[i]) {
            double Ua = 0;
            for (int k = 0; k < 6; k++) Ua += w->U[i][k] * a[b][k];
            double qdd = (uu[i] - Ua) / w->D[i];
            for (int k = 0; k < 3; k++) a[b][k] += m->axis[i][k] * qdd;
            delta[6 + m->dof[i]] = qdd;
            FL(s, 20);
        }
        FL(s, 39);
    }
    mat3_vec(w->Rw[0], a[0], delta);
    mat3_vec(w->Rw[0], a[0] + 3, delta + 3);
    FL(s, 30);
}

/* ------------------------------------------------------------------ 3. rows + PGS */
typedef struct {
    double J[ORC_NDOF], B[ORC_NDOF];
    double rhs, lo, hi, dinv, lam, mu;
    int normal_index; /* row index of the normal constraint this friction row belongs to, -1 otherwise */
} row_t;

static double dotn(const double *a, const double *b) {
    double s = 0;
    for (int k = 0; k < ORC_NDOF; k++) s += a[k] * b[k];
    return s;
}
/* J for a unit force along dir at world point pt on link `link` (linear row) or unit moment about dir (angular) */
static void point_jacobian(const plen_oracle_model *m, const work_t *w, int link, const double *pt, const double *dir,
                           int angular, double *J) {
    memset(J, 0, ORC_NDOF * sizeof(double));
    double r[3], t[3];
    if (angular) {
        memcpy(J, dir, 3 * sizeof(double));
    } else {
        for (int k = 0; k < 3; k++) r[k] = pt[k] - w->pw[0][k];
        cross3(r, dir, J);
        memcpy(J + 3, dir, 3 * sizeof(double));
    }
    for (int i = link; i >= 0; i = m->parent[i]) {
        if (!m->jtype[i]) continue;
        if (angular) {
            J[6 + m->dof[i]] = dot3(w->axis_w[i], dir);
        } else {
            for (int k = 0; k < 3; k++) r[k] = pt[k] - w->pw[i + 1][k];
            cross3(r, dir, t);
            J[6 + m->dof[i]] = dot3(w->axis_w[i], t);
        }
    }
}
static void finish_row(const plen_oracle_model *m, plen_oracle_state *s, const work_t *w, row_t *r) {
    impulse_response(m, s, w, r->J, r->B);
    double d = dotn(r->J, r->B);
    r->dinv = (d > 2.220446049250313e-16) ? 1.0 / d : 0.0; /* SIMD_EPSILON (double build) */
    r->lam = 0.0;
    r->normal_index = -1;
    r->mu = 0.0;
    FL(s, 2 * ORC_NDOF + 1);
}
/* btMultiBodyConstraintSolver::resolveSingleConstraintRowGeneric */
static double resolve_row(plen_oracle_state *s, row_t *r, double *dv) {
    double delta = r->rhs - dotn(r->J, dv) * r->dinv; /* cfm = 0 */
    double sum = r->lam + delta;
    if (sum < r->lo) { delta = r->lo - r->lam; r->lam = r->lo; }
    else if (sum > r->hi) { delta = r->hi - r->lam; r->lam = r->hi; }
    else r->lam = sum;
    for (int k = 0; k < ORC_NDOF; k++) dv[k] += r->B[k] * delta;
    FL(s, 4 * ORC_NDOF + 6);
    return r->dinv != 0.0 ? delta / r->dinv : 0.0;
}
/* btMultiBodyConstraintSolver::resolveConeFrictionConstraintRows */
static double resolve_cone(plen_oracle_state *s, row_t *a, row_t *b, double *dv) {
    double da = a->rhs - dotn(a->J, dv) * a->dinv, db = b->rhs - dotn(b->J, dv) * b->dinv;
    double sa = a->lam + da, sb = b->lam + db;
    if (sa < a->lo || sa > a->hi || sb < b->lo || sb > b->hi) {
        double ang = atan2(sa, sb);
        double ca = fabs(a->lo * sin(ang)), cb = fabs(b->lo * cos(ang));
        if (sa < -ca) { da = -ca - a->lam; a->lam = -ca; }
        else if (sa > ca) { da = ca - a->lam; a->lam = ca; }
        else a->lam = sa;
        if (sb < -cb) { db = -cb - b->lam; b->lam = -cb; }
        else if (sb > cb) { db = cb - b->lam; b->lam = cb; }
        else b->lam = sb;
    } else {
        a->lam = sa; b->lam = sb;
    }
    for (int k = 0; k < ORC_NDOF; k++) dv[k] += a->B[k] * da + b->B[k] * db;
    FL(s, 8 * ORC_NDOF + 20);
    return (a->dinv != 0.0 ? da / a->dinv : 0.0) + (b->dinv != 0.0 ? db / b->dinv : 0.0);
}
/* btPlaneSpace1 */
static void plane_space(const double *n, double *p, double *q) {
    if (fabs(n[2]) > 0.7071067811865475244008443621048490) {
        double a = n[1] * n[1] + n[2] * n[2], k = 1.0 / sqrt(a);
        p[0] = 0; p[1] = -n[2] * k; p[2] = n[1] * k;
        q[0] = a * k; q[1] = -n[0] * p[2]; q[2] = n[0] * p[1];
    } else {
        double a = n[0] * n[0] + n[1] * n[1], k = 1.0 / sqrt(a);
        p[0] = -n[1] * k; p[1] = n[0] * k; p[2] = 0;
        q[0] = -n[2] * p[1]; q[1] = n[2] * p[0]; q[2] = a * k;
    }
}
static double clampd(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }

/* ------------------------------------------------------------------ 4. tick */
/* EXPERIMENT (plen_oracle_config.manifold_mode == 1): the sole contact of a foot as btConvexPlaneCollisionAlgorithm +
 * btPersistentManifold produce it, restated from the published sources as recalled ([RECALL], like the rest of the physics):
 *   collideSingleContact : support vertex of the hull towards the ground (first maximum over the point list), moved out by the
 *                          collision margin; a contact if its distance is below the breaking threshold
 *   addContactPoint      : getCacheEntry (nearest cached point, in the foot's frame, within the threshold) -> replace it and keep
 *                          its impulse; else add; a full cache goes through sortCachedPoints (keep the deepest, drop the point
 *                          whose removal leaves the largest quadrilateral)
 *   refreshContactPoints : world positions and distance of every cached point from its two local copies; points that separated
 *                          by more than the threshold or drifted sideways by more than it are removed (last index swapped in)
 * Fills cp_pos / cp_dist of the cached points (slot k = cache index k) and s->in_manifold / s->lam_n accordingly. */
static void manifold_update(const plen_oracle_model *m, const plen_oracle_config *cfg, plen_oracle_state *s, const work_t *w,
                            double cp_pos[ORC_NFEET][ORC_NPTS][3], double cp_dist[ORC_NFEET][ORC_NPTS]) {
    for (int f = 0; f < ORC_NFEET; f++) {
        const int L = m->foot_link[f];
        const double *R = w->Rw[L + 1], *p = w->pw[L + 1];
        const double thr = m->foot_break[f];
        /* support vertex: lowest world z; vertices within support_tie of the lowest are ties -> first in the list.  (The height
         * is compared relative to the foot origin, as the device code does.) */
        int best = -1;
        double zmin = 1e300, zbest = 1e300;
        for (int i = 0; i < m->n_hull[f]; i++) {
            const double *v = m->foot_hull[f][i];
            const double z = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
            if (z < zmin) zmin = z;
        }
        for (int i = 0; i < m->n_hull[f]; i++) {
            const double *v = m->foot_hull[f][i];
            const double z = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
            if (z <= zmin + cfg->support_tie) { zbest = p[2] + z; best = i; break; }
        }
        if (best >= 0 && zbest - cfg->hull_margin < thr) {
            /* point on the inflated hull (world), its copy in the foot frame, and its projection on the plane */
            double t[3], pa[3], la[3], pb[3];
            mat3_vec(R, m->foot_hull[f][best], t);
            pa[0] = p[0] + t[0]; pa[1] = p[1] + t[1]; pa[2] = p[2] + t[2] - cfg->hull_margin;
            const double d[3] = {pa[0] - p[0], pa[1] - p[1], pa[2] - p[2]};
            for (int c = 0; c < 3; c++) la[c] = R[c] * d[0] + R[3 + c] * d[1] + R[6 + c] * d[2];      /* R^T d */
            pb[0] = pa[0]; pb[1] = pa[1]; pb[2] = 0.0;
            const double dist = pa[2];
            int n = s->man_n[f], slot = -1;
            double nearest = thr * thr;
            for (int k = 0; k < n; k++) {           /* getCacheEntry */
                double dd = 0;
                for (int c = 0; c < 3; c++) { const double e = s->man_local[f][k][c] - la[c]; dd += e * e; }
                if (dd < nearest) { nearest = dd; slot = k; }
            }
            if (slot < 0) {
                if (n < ORC_NPTS) { slot = n; s->man_n[f] = n + 1; s->lam_n[f][slot] = 0.0; }
                else {                                   /* sortCachedPoints */
                    int deepest = -1;
                    double maxpen = dist;
                    for (int k = 0; k < ORC_NPTS; k++) {
                        const double zk = p[2] + R[6] * s->man_local[f][k][0] + R[7] * s->man_local[f][k][1] + R[8] * s->man_local[f][k][2];
                        if (zk < maxpen) { deepest = k; maxpen = zk; }
                    }
                    static const int other[4][3] = {{1, 2, 3}, {0, 2, 3}, {0, 1, 3}, {0, 1, 2}};
                    double res[4] = {0, 0, 0, 0};
                    for (int k = 0; k < 4; k++) {
                        if (k == deepest) continue;
                        const double *q0 = s->man_local[f][other[k][0]], *q1 = s->man_local[f][other[k][1]], *q2 = s->man_local[f][other[k][2]];
                        /* as Bullet pairs them: a = new - first remaining, b = third remaining - second remaining */
                        const double a[3] = {la[0] - q0[0], la[1] - q0[1], la[2] - q0[2]};
                        const double b[3] = {q2[0] - q1[0], q2[1] - q1[1], q2[2] - q1[2]};
                        double cr[3];
                        cross3(a, b, cr);
                        res[k] = cr[0] * cr[0] + cr[1] * cr[1] + cr[2] * cr[2];
                    }
                    slot = 0;
                    for (int k = 1; k < 4; k++) if (res[k] > res[slot]) slot = k;
                    s->lam_n[f][slot] = 0.0;
                }
            }
            for (int c = 0; c < 3; c++) { s->man_local[f][slot][c] = la[c]; s->man_world[f][slot][c] = pb[c]; }
        }
        /* refreshContactPoints, last to first */
        for (int k = s->man_n[f] - 1; k >= 0; k--) {
            double t[3], pa[3];
            mat3_vec(R, s->man_local[f][k], t);
            for (int c = 0; c < 3; c++) pa[c] = p[c] + t[c];
            const double dist = pa[2] - s->man_world[f][k][2];
            const double dx = s->man_world[f][k][0] - pa[0], dy = s->man_world[f][k][1] - pa[1];
            if (dist > thr || dx * dx + dy * dy > thr * thr) {
                const int last = s->man_n[f] - 1;
                if (k != last) {
                    memcpy(s->man_local[f][k], s->man_local[f][last], sizeof s->man_local[f][k]);
                    memcpy(s->man_world[f][k], s->man_world[f][last], sizeof s->man_world[f][k]);
                    s->lam_n[f][k] = s->lam_n[f][last];
                }
                s->lam_n[f][last] = 0.0;
                s->man_n[f] = last;
            }
        }
        for (int k = 0; k < ORC_NPTS; k++) {
            s->in_manifold[f][k] = k < s->man_n[f];
            cp_dist[f][k] = 0.0;
            for (int c = 0; c < 3; c++) cp_pos[f][k][c] = 0.0;
            if (k >= s->man_n[f]) { s->lam_n[f][k] = 0.0; continue; }
            double t[3];
            mat3_vec(R, s->man_local[f][k], t);
            for (int c = 0; c < 3; c++) cp_pos[f][k][c] = p[c] + t[c];
            cp_dist[f][k] = cp_pos[f][k][2] - s->man_world[f][k][2];
        }
    }
}

void plen_oracle_tick(const plen_oracle_model *m, const plen_oracle_config *cfg, plen_oracle_state *s) {
    work_t w;
    double acc[ORC_NDOF], vstar[ORC_NDOF], dv[ORC_NDOF];
    const double dt = cfg->dt;
    row_t nc[3 * ORC_NJ];                         /* non-contact: limits then motors */
    row_t spin[ORC_NFEET * ORC_NPTS];             /* spinning (about normal) */
    row_t roll[2 * ORC_NFEET * ORC_NPTS];         /* rolling (about the two tangents) */
    static _Thread_local row_t nrm[ORC_NFEET * ORC_NPTS + ORC_MAXXP];        /* normals: soles first, then the link boxes */
    static _Thread_local row_t fric[2 * (ORC_NFEET * ORC_NPTS + ORC_MAXXP)]; /* lateral pairs in the same order */
    int n_nc = 0, n_nrm = 0, n_spin = 0, n_roll = 0, n_fric = 0;
    int nrm_foot[ORC_NFEET * ORC_NPTS], nrm_pt[ORC_NFEET * ORC_NPTS];

    /* (a) unconstrained velocities: v* = v + a dt (btMultiBodyDynamicsWorld::solveConstraints, first pass) */
    aba_forward_dynamics(m, cfg, s, &w, acc);
    for (int k = 0; k < 3; k++) { vstar[k] = s->omega[k]; vstar[3 + k] = s->vel[k]; }
    for (int k = 0; k < ORC_NJ; k++) vstar[6 + k] = s->qd[k];
    for (int k = 0; k < ORC_NDOF; k++) vstar[k] = clampd(vstar[k] + acc[k] * dt, -cfg->max_coord_velocity, cfg->max_coord_velocity);
    memset(dv, 0, sizeof dv);

    /* (b) collision detection at the start-of-tick poses: sole vertices vs the plane z = 0 */
    double cp_pos[ORC_NFEET][ORC_NPTS][3], cp_dist[ORC_NFEET][ORC_NPTS];
    if (cfg->manifold_mode == 1) manifold_update(m, cfg, s, &w, cp_pos, cp_dist);
    else
    for (int f = 0; f < ORC_NFEET; f++) {
        int L = m->foot_link[f];
        for (int k = 0; k < ORC_NPTS; k++) {
            double t[3];
            mat3_vec(w.Rw[L + 1], m->foot_pts[f][k], t);
            for (int c = 0; c < 3; c++) cp_pos[f][k][c] = w.pw[L + 1][c] + t[c];
            /* inflated hull vs plane: distance between surfaces = z_vertex - margin; point on A on the inflated surface */
            cp_dist[f][k] = cp_pos[f][k][2] - cfg->hull_margin;
            cp_pos[f][k][2] -= cfg->hull_margin;
            int in = cp_dist[f][k] <= m->foot_break[f];
            if (!in || !s->in_manifold[f][k]) s->lam_n[f][k] = 0.0; /* new or dropped point: empty cache */
            s->in_manifold[f][k] = in;
        }
    }

    /* (c) non-contact rows: joint limits (only when violated), then motors, in link order */
    for (int i = 0; i < m->n_links; i++) {
        if (!m->jtype[i]) continue;
        int d = m->dof[i];
        for (int side = 0; side < 2; side++) {
            double pen = side ? (m->upper[i] - s->q[d]) : (s->q[d] - m->lower[i]);
            if (pen > 0) continue;
            row_t *r = &nc[n_nc++];
            memset(r->J, 0, sizeof r->J);
            r->J[6 + d] = side ? -1.0 : 1.0;
            finish_row(m, s, &w, r);
            double rel = dotn(r->J, vstar);
            r->rhs = (-pen * cfg->erp_joint / dt - rel) * r->dinv;
            r->lo = 0; r->hi = 100.0; /* btMultiBodyConstraint default m_maxAppliedImpulse */
        }
    }
    for (int i = 0; i < m->n_links; i++) {
        if (!m->jtype[i]) continue;
        int d = m->dof[i];
        row_t *r = &nc[n_nc++];
        memset(r->J, 0, sizeof r->J);
        r->J[6 + d] = 1.0;
        finish_row(m, s, &w, r);
        /* btMultiBodyJointMotor::createConstraintRows, erp 1, desired velocity 0 */
        double cur = vstar[6 + d];
        double desired = cfg->motor_kp * (s->target[d] - s->q[d]) / dt + cur + cfg->motor_kd * (0.0 - cur);
        r->rhs = (desired - cur) * r->dinv;
        r->hi = cfg->motor_max_force * dt; r->lo = -r->hi;
    }

    /* (d) contact rows */
    const double nrmdir[3] = {0, 0, 1};
    double t1[3], t2[3];
    plane_space(nrmdir, t1, t2);
    for (int f = 0; f < ORC_NFEET; f++) {
        int L = m->foot_link[f];
        for (int k = 0; k < ORC_NPTS; k++) {
            if (!s->in_manifold[f][k]) continue;
            int ni = n_nrm++;
            row_t *r = &nrm[ni];
            nrm_foot[ni] = f; nrm_pt[ni] = k;
            point_jacobian(m, &w, L, cp_pos[f][k], nrmdir, 0, r->J);
            finish_row(m, s, &w, r);
            double rel = dotn(r->J, vstar);
            double rest = (fabs(rel) < cfg->restitution_vel_threshold) ? 0.0 : cfg->restitution * -rel;
            if (rest <= 0) rest = 0;
            double dist = cp_dist[f][k] + cfg->linear_slop;
            double velerr = rest - rel, poserr = 0;
            if (dist > 0) velerr -= dist / dt; else poserr = -dist * cfg->erp_contact / dt;
            r->rhs = (poserr + velerr) * r->dinv;
            r->lo = 0; r->hi = 1e10;
            r->lam = s->lam_n[f][k] * cfg->warmstart_factor;
            if (r->lam != 0.0) for (int c = 0; c < ORC_NDOF; c++) dv[c] += r->B[c] * r->lam;
            if (cfg->mu_spinning > 0) {
                row_t *q = &spin[n_spin++];
                point_jacobian(m, &w, L, cp_pos[f][k], nrmdir, 1, q->J);
                finish_row(m, s, &w, q);
                q->rhs = -dotn(q->J, vstar) * q->dinv;
                q->mu = cfg->mu_spinning; q->normal_index = ni; q->lo = -q->mu; q->hi = q->mu;
            }
            if (cfg->mu_rolling > 0) {
                for (int a = 0; a < 2; a++) {
                    row_t *q = &roll[n_roll++];
                    point_jacobian(m, &w, L, cp_pos[f][k], a ? t2 : t1, 1, q->J);
                    finish_row(m, s, &w, q);
                    q->rhs = -dotn(q->J, vstar) * q->dinv;
                    q->mu = cfg->mu_rolling; q->normal_index = ni; q->lo = -q->mu; q->hi = q->mu;
                }
            }
            for (int a = 0; a < 2; a++) {
                row_t *q = &fric[n_fric++];
                point_jacobian(m, &w, L, cp_pos[f][k], a ? t2 : t1, 0, q->J);
                finish_row(m, s, &w, q);
                q->rhs = -dotn(q->J, vstar) * q->dinv;
                q->mu = cfg->mu_lateral; q->normal_index = ni; q->lo = -q->mu; q->hi = q->mu;
            }
        }
    }

    /* (d2) ground contact of the link BOXES (everything but the two foot hulls), restating btBoxBoxDetector for a small
     * box on the big ground box of plane.urdf: the separating axis is the ground normal, the incident face is the face of
     * the link box whose outward normal points most downward, and its (up to four) vertices that PENETRATE (depth >= 0)
     * become contact points.  No manifold hysteresis and no warm start for these points (documented simplification: a
     * cached point only adds a speculative row while it hovers within ~1 mm above the ground).  Friction: lateral only
     * (links carry no spinning / rolling friction, plen_env.py:438-467 sets those on the feet alone). */
    const int n_foot_nrm = n_nrm;
    s->last_box_points = 0; s->last_boxes_touching = 0;
    if (cfg->link_contacts && m->n_boxes > 0) {
        double bp[ORC_MAXBOX][4][3], bdepth[ORC_MAXBOX][4], bmax[ORC_MAXBOX];
        int bn[ORC_MAXBOX], keep[ORC_MAXBOX], touching = 0;
        for (int b = 0; b < m->n_boxes; b++) {
            const int body = m->box_link[b] + 1;
            double Rb[9], cb[3], t[3];
            mat3_mul(w.Rw[body], m->box_rot[b], Rb);
            mat3_vec(w.Rw[body], m->box_center[b], t);
            for (int c = 0; c < 3; c++) cb[c] = w.pw[body][c] + t[c];
            int ks = 0;                                       /* box axis most aligned with the ground normal (first max) */
            for (int k = 1; k < 3; k++) if (fabs(Rb[6 + k]) > fabs(Rb[6 + ks])) ks = k;
            const int ki = (ks + 1) % 3, kj = (ks + 2) % 3;
            const double sgn = Rb[6 + ks] > 0 ? -1.0 : 1.0;   /* walk along the axis towards the ground */
            bn[b] = 0; bmax[b] = -1.0; keep[b] = 0;
            for (int v = 0; v < 4; v++) {
                const double si = (v & 1) ? 1.0 : -1.0, sj = (v & 2) ? 1.0 : -1.0;
                double pt[3];
                for (int c = 0; c < 3; c++)
                    pt[c] = cb[c] + sgn * m->box_half[b][ks] * Rb[3 * c + ks] + si * m->box_half[b][ki] * Rb[3 * c + ki] +
                            sj * m->box_half[b][kj] * Rb[3 * c + kj];
                const double depth = -pt[2];
                if (depth >= 0.0) {
                    memcpy(bp[b][bn[b]], pt, sizeof pt);
                    bdepth[b][bn[b]] = depth;
                    if (depth > bmax[b]) bmax[b] = depth;
                    bn[b]++;
                }
            }
            if (bn[b]) touching++;
        }
        s->last_boxes_touching = touching;
        /* the cap: keep the max_contact_points points of deepest penetration (ties: lower box index, then lower vertex
         * index); the kept points then enter the solver in (box, vertex) order */
        int total = 0, taken[ORC_MAXBOX][4];
        for (int b = 0; b < m->n_boxes; b++) { total += bn[b]; for (int v = 0; v < 4; v++) taken[b][v] = 0; }
        int quota = (cfg->max_contact_points < 0 || cfg->max_contact_points > total) ? total : cfg->max_contact_points;
        for (int q = 0; q < quota; q++) {
            int bb = -1, bv = -1;
            for (int b = 0; b < m->n_boxes; b++)
                for (int v = 0; v < bn[b]; v++)
                    if (!taken[b][v] && (bb < 0 || bdepth[b][v] > bdepth[bb][bv])) { bb = b; bv = v; }
            taken[bb][bv] = 1; keep[bb] = 1;
        }
        for (int b = 0; b < m->n_boxes; b++) {
            if (!keep[b]) continue;
            const int L = m->box_link[b];
            const double rest_coeff = (L < 0) ? cfg->restitution_base : cfg->restitution;
            for (int v = 0; v < bn[b]; v++) {
                if (!taken[b][v]) continue;
                int ni = n_nrm++;
                row_t *r = &nrm[ni];
                point_jacobian(m, &w, L, bp[b][v], nrmdir, 0, r->J);
                finish_row(m, s, &w, r);
                double rel = dotn(r->J, vstar);
                double rest = (fabs(rel) < cfg->restitution_vel_threshold) ? 0.0 : rest_coeff * -rel;
                if (rest <= 0) rest = 0;
                double dist = -bdepth[b][v] + cfg->linear_slop;
                double velerr = rest - rel, poserr = 0;
                if (dist > 0) velerr -= dist / dt; else poserr = -dist * cfg->erp_contact / dt;
                r->rhs = (poserr + velerr) * r->dinv;
                r->lo = 0; r->hi = 1e10;
                for (int a = 0; a < 2; a++) {
                    row_t *q = &fric[n_fric++];
                    point_jacobian(m, &w, L, bp[b][v], a ? t2 : t1, 0, q->J);
                    finish_row(m, s, &w, q);
                    q->rhs = -dotn(q->J, vstar) * q->dinv;
                    q->mu = cfg->mu_link; q->normal_index = ni; q->lo = -q->mu; q->hi = q->mu;
                }
                s->last_box_points++;
            }
        }
    }

    /* (e) projected Gauss-Seidel, btMultiBodyConstraintSolver::solveSingleIteration order */
    int it;
    for (it = 0; it < cfg->solver_iterations; it++) {
        double res = 0, x;
        for (int j = 0; j < n_nc; j++) {
            int idx = (it & 1) ? j : n_nc - 1 - j;
            x = resolve_row(s, &nc[idx], dv); if (x * x > res) res = x * x;
        }
        for (int j = 0; j < n_nrm; j++) { x = resolve_row(s, &nrm[j], dv); if (x * x > res) res = x * x; }
        for (int j = 0; j < n_spin; j++) {
            double tot = nrm[spin[j].normal_index].lam;
            if (tot > 0) { spin[j].lo = -spin[j].mu * tot; spin[j].hi = spin[j].mu * tot;
                           x = resolve_row(s, &spin[j], dv); if (x * x > res) res = x * x; }
        }
        for (int j = 0; j < n_roll; j++) {
            double tot = nrm[roll[j].normal_index].lam;
            if (tot > 0) { roll[j].lo = -roll[j].mu * tot; roll[j].hi = roll[j].mu * tot;
                           x = resolve_row(s, &roll[j], dv); if (x * x > res) res = x * x; }
        }
        if (cfg->implicit_cone) {
            for (int j = 0; j + 1 < n_fric; j += 2) {
                double tot = nrm[fric[j].normal_index].lam;
                fric[j].lo = -fric[j].mu * tot; fric[j].hi = fric[j].mu * tot;
                fric[j + 1].lo = -fric[j + 1].mu * tot; fric[j + 1].hi = fric[j + 1].mu * tot;
                x = resolve_cone(s, &fric[j], &fric[j + 1], dv); if (x * x > res) res = x * x;
            }
        } else {
            for (int j = 0; j < n_fric; j++) {
                double tot = nrm[fric[j].normal_index].lam;
                if (tot > 0) { fric[j].lo = -fric[j].mu * tot; fric[j].hi = fric[j].mu * tot;
                               x = resolve_row(s, &fric[j], dv); if (x * x > res) res = x * x; }
            }
        }
        if (res <= cfg->residual_threshold || it >= cfg->solver_iterations - 1) { it++; break; }
    }
    s->last_iterations = it;
    s->last_rows = n_nc + n_nrm + n_spin + n_roll + n_fric;
    for (int j = 0; j < n_foot_nrm; j++) s->lam_n[nrm_foot[j]][nrm_pt[j]] = nrm[j].lam;

    /* (f) apply delta v (clamped like applyDeltaVeeMultiDof) and integrate (stepPositionsMultiDof) */
    for (int k = 0; k < ORC_NDOF; k++) vstar[k] = clampd(vstar[k] + dv[k], -cfg->max_coord_velocity, cfg->max_coord_velocity);
    for (int k = 0; k < 3; k++) { s->omega[k] = vstar[k]; s->vel[k] = vstar[3 + k]; s->pos[k] += s->vel[k] * dt; }
    for (int k = 0; k < ORC_NJ; k++) { s->qd[k] = vstar[6 + k]; s->q[k] += s->qd[k] * dt; }
    {
        /* exponential-map quaternion update with the world angular velocity, then normalise */
        double fa = sqrt(dot3(s->omega, s->omega)), ax[3], sc;
        if (fa * dt > 0.7853981633974483) fa = 0.7853981633974483 / dt; /* ANGULAR_MOTION_THRESHOLD */
        if (fa < 0.001) sc = 0.5 * dt - dt * dt * dt * 0.020833333333 * fa * fa;
        else sc = sin(0.5 * fa * dt) / fa;
        for (int k = 0; k < 3; k++) ax[k] = s->omega[k] * sc;
        double dq[4] = {ax[0], ax[1], ax[2], cos(0.5 * fa * dt)}, *q = s->quat, r[4];
        r[0] = dq[3] * q[0] + dq[0] * q[3] + dq[1] * q[2] - dq[2] * q[1];
        r[1] = dq[3] * q[1] - dq[0] * q[2] + dq[1] * q[3] + dq[2] * q[0];
        r[2] = dq[3] * q[2] + dq[0] * q[1] - dq[1] * q[0] + dq[2] * q[3];
        r[3] = dq[3] * q[3] - dq[0] * q[0] - dq[1] * q[1] - dq[2] * q[2];
        double nn = 1.0 / sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2] + r[3] * r[3]);
        for (int k = 0; k < 4; k++) q[k] = r[k] * nn;
    }
    FL(s, 200);
}

/* ------------------------------------------------------------------ 5. env logic (plen_env.py) */
void plen_oracle_default_config(plen_oracle_config *c, int joint_act) {
    static const double lo[ORC_NJ] = {-1.57, -0.15, -0.95, -0.9, -0.95, -0.8, -1.57, -1.5, -0.75, -0.3, -1.2, -0.4,
                                      -1.57, -0.15, -0.2, -1.57, -0.15, -0.2};       /* plen_env.py:148-167 */
    static const double hi[ORC_NJ] = {1.57, 1.5, 0.75, 0.3, 1.2, 0.4, 1.57, 0.15, 0.95, 0.9, 0.95, 0.8,
                                      1.57, 1.57, 0.35, 1.57, 1.57, 0.35};
    memset(c, 0, sizeof *c);
    c->dt = 1.0 / 240.0; c->substeps = 4; c->reset_ticks = 8; c->gravity_z = -9.81;
    c->start_pos[0] = 0; c->start_pos[1] = 0; c->start_pos[2] = 0.158;
    c->motor_max_force = 0.15; c->joint_act = joint_act;
    c->linear_damping = joint_act ? 0.1 : 0.0; c->angular_damping = 0.0;
    c->mu_lateral = 0.8 * 0.8; c->mu_spinning = 0.1 * 0.8; c->mu_rolling = (joint_act ? 0.01 : 0.1) * 0.8;
    c->restitution = 0.5 * 0.5;
    memcpy(c->env_lo, lo, sizeof lo); memcpy(c->env_hi, hi, sizeof hi);
    c->max_episode_steps = 500;
    c->motor_kp = 0.1; c->motor_kd = 1.0; c->solver_iterations = 50; c->residual_threshold = 1e-7;
    c->erp_contact = 0.08; c->erp_joint = 0.2; c->linear_slop = 1e-5; c->warmstart_factor = 0.1;
    c->restitution_vel_threshold = 0.2; c->hull_margin = 0.001; c->max_coord_velocity = 100.0; c->implicit_cone = 1;
    c->link_contacts = 1; c->mu_link = 0.5 * 0.8; c->restitution_base = 0.0 * 0.5; c->max_contact_points = -1;
    c->manifold_mode = 0; c->support_tie = 1e-7;
}

void plen_oracle_init_state(const plen_oracle_config *cfg, plen_oracle_state *s) {
    memset(s, 0, sizeof *s);
    memcpy(s->pos, cfg->start_pos, sizeof s->pos);
    s->quat[3] = 1.0;
}

void plen_oracle_fk(const plen_oracle_model *m, const plen_oracle_state *s, double *pos, double *rot) {
    work_t w;
    kinematics(m, s, &w);
    for (int b = 0; b <= m->n_links; b++) { memcpy(pos + 3 * b, w.pw[b], 3 * sizeof(double)); memcpy(rot + 9 * b, w.Rw[b], 9 * sizeof(double)); }
}

void plen_oracle_minv(const plen_oracle_model *m, const plen_oracle_state *s, double *Minv) {
    work_t w;
    plen_oracle_config cfg;
    plen_oracle_state t = *s;
    double acc[ORC_NDOF];
    plen_oracle_default_config(&cfg, 0);
    aba_forward_dynamics(m, &cfg, &t, &w, acc);
    for (int c = 0; c < ORC_NDOF; c++) {
        double e[ORC_NDOF] = {0}, col[ORC_NDOF];
        e[c] = 1.0;
        impulse_response(m, &t, &w, e, col);
        for (int r = 0; r < ORC_NDOF; r++) Minv[ORC_NDOF * r + c] = col[r];
    }
}

/* compute_observation (plen_env.py:768-822) without the history side effects */
void plen_oracle_observe(const plen_oracle_model *m, const plen_oracle_config *cfg, const plen_oracle_state *s,
                         double *obs, double *foot_euler6) {
    (void)cfg;
    double rpy[3];
    for (int k = 0; k < ORC_NJ; k++) obs[k] = s->q[k];            /* :807-814 */
    quat_to_euler(s->quat, rpy);                                  /* :799-803 */
    int R = 0, L = 0;
    for (int k = 0; k < ORC_NPTS; k++) { R |= s->in_manifold[0][k]; L |= s->in_manifold[1][k]; } /* :774-790 */
    obs[18] = s->pos[2]; obs[19] = s->vel[0]; obs[20] = rpy[0]; obs[21] = rpy[1]; obs[22] = rpy[2];
    obs[23] = s->pos[1]; obs[24] = R; obs[25] = L;                /* :816-822 */
    if (foot_euler6) {
        work_t w;
        kinematics(m, s, &w);
        for (int f = 0; f < ORC_NFEET; f++) {
            double q[4];
            mat_to_quat(w.Rw[m->foot_link[f] + 1], q);             /* getLinkState(...)[1], :1016, :1029 */
            quat_to_euler(q, foot_euler6 + 3 * f);
        }
    }
}

static const int PAIR_L[3] = {2, 3, 4};   /* "lhip/lknee/lankle" = JointStates[2],[3],[4]  (plen_env.py:853-866) */
static const int PAIR_R[3] = {8, 9, 10};  /* "rhip/rknee/rankle" = JointStates[8],[9],[10] */

static void clear_gait(plen_oracle_state *s) { /* plen_env.py:914-923 / :581-590 */
    s->hist_len = 0; memset(s->sums, 0, sizeof s->sums); s->cnt = 0; s->ds = 0;
}

void plen_oracle_reset(const plen_oracle_model *m, const plen_oracle_config *cfg, plen_oracle_state *s, double *obs) {
    int dead = s->dead;
    /* resetBasePositionAndOrientation + resetJointState(j, 0), plen_env.py:561-565 */
    memcpy(s->pos, cfg->start_pos, sizeof s->pos);
    s->quat[0] = s->quat[1] = s->quat[2] = 0; s->quat[3] = 1;
    memset(s->omega, 0, sizeof s->omega); memset(s->vel, 0, sizeof s->vel);
    memset(s->q, 0, sizeof s->q); memset(s->qd, 0, sizeof s->qd);
    /* the teleport invalidates every cached manifold point (positions are stored per body) */
    memset(s->lam_n, 0, sizeof s->lam_n); memset(s->in_manifold, 0, sizeof s->in_manifold);
    memset(s->man_n, 0, sizeof s->man_n);
    memset(s->target, 0, sizeof s->target);                    /* move_joints(zeros), raw radians, :568 */
    for (int i = 0; i < cfg->reset_ticks; i++) plen_oracle_tick(m, cfg, s);   /* :569-570 */
    if (obs) plen_oracle_observe(m, cfg, s, obs, 0);           /* :574 (history side effects cleared below) */
    s->ep_ret = 0; s->ep_t = 0;                                /* :578-579 */
    clear_gait(s);                                             /* :581-590 */
    s->dead = dead;                                            /* `dead` is not touched by reset */
}

void plen_oracle_step(const plen_oracle_model *m, const plen_oracle_config *cfg, plen_oracle_state *s,
                      const double *action, double *obs, double *reward_out, int *done_out, int *timeout_out) {
    /* agent_to_env, plen_env.py:694-714 (bypassed when joint_act, :652-654) */
    for (int i = 0; i < ORC_NJ; i++) {
        if (cfg->joint_act) { s->target[i] = action[i]; continue; }
        double lo = cfg->env_lo[i], hi = cfg->env_hi[i];
        double mm = (hi - lo) / (1.0 - (-1.0));
        double b = hi - (mm * 1.0);
        double y = mm * action[i] + b;
        if (y >= hi) y = hi - 0.001; else if (y <= lo) y = lo + 0.001;
        s->target[i] = y;
    }
    for (int i = 0; i < cfg->substeps; i++) plen_oracle_tick(m, cfg, s);      /* :663-667 */

    /* ---- compute_observation :768-871 */
    double fe[6];
    plen_oracle_observe(m, cfg, s, obs, fe);
    const int R = (int)obs[24], L = (int)obs[25];
    const double z = obs[18], vx = obs[19], roll = obs[20], pitch = obs[21], yaw = obs[22], y = obs[23];
    double cur[6], diff[6];
    for (int k = 0; k < 3; k++) { cur[2 * k] = s->q[PAIR_L[k]]; cur[2 * k + 1] = s->q[PAIR_R[k]]; }
    int first_pass;
    if (s->hist_len > 0) { for (int k = 0; k < 6; k++) diff[k] = s->last[k] - cur[k]; first_pass = 0; }   /* :825-841 */
    else { for (int k = 0; k < 6; k++) diff[k] = 0; first_pass = 1; }                                      /* :842-849 */
    for (int k = 0; k < 6; k++) s->last[k] = cur[k];                                                        /* :853-866 */
    for (int k = 0; k < 3; k++) {
        s->sums[3 * k] += cur[2 * k] * cur[2 * k + 1];
        s->sums[3 * k + 1] += cur[2 * k] * cur[2 * k];
        s->sums[3 * k + 2] += cur[2 * k + 1] * cur[2 * k + 1];
    }
    s->hist_len += 1;

    /* ---- compute_done :1072-1093 (one-sided) */
    int dead = (roll > fabs(M_PI / 3.)) || (pitch > fabs(M_PI / 3.)) || (z < 0.08) || (y > 1);
    s->dead = dead;

    /* ---- compute_reward :873-1070 */
    double r = 0.0;                                                   /* alive_reward = 0, :65, :880 */
    if (vx < 0) r -= exp(vx * 3.0); else r += (vx * 3.0) * (vx * 3.0);  /* :885-889 (np.sign(vx) < 0) */
    { double h = fabs(0.160178937611 - z) * 40.0; r -= h * h; }        /* :894-895 */
    r -= fabs(y) * fabs(y) * 1.0;                                      /* :901 */
    r -= fabs(roll) * fabs(roll) * 1.0;                                /* :903 */
    r -= fabs(pitch) * fabs(pitch) * 0.5;                              /* :905 */
    r -= fabs(yaw) * fabs(yaw) * 1.0;                                  /* :907 */
    double jrew = 0, jpen = 0;
    if (s->cnt >= 80 && R == 1) {                                      /* :913-923 */
        clear_gait(s);
    } else if (s->cnt >= 1.5 * 80) {                                   /* :924-925 */
        r -= 2;
    } else if (s->cnt > 0) {                                           /* :926-968 */
        for (int k = 0; k < 3; k++)                                    /* dot / (norm * norm), NaN if a norm is 0 */
            jrew += s->sums[3 * k] / (sqrt(s->sums[3 * k + 1]) * sqrt(s->sums[3 * k + 2]));
        jrew *= 1.0 / 3.0;                                             /* :948 */
        if (!first_pass) {                                             /* :952-968 */
            for (int k = 0; k < 6; k++) jpen -= 1.0 / exp(fabs(diff[k]));
            jpen *= 0.5 * (1.0 / 3.0);
        }
    }
    r += jrew; r += jpen;                                              /* :972-973 */
    if (L == 1) {                                                      /* :978-982 (python3 true division) */
        double a = (s->cnt * 10.0 / 80.0) - 0.5 * 10;
        r += 0.5 * (1 - tanh(a * a));
    }
    if (s->cnt < 80 / 2.0) {                                           /* :988-994 */
        if (R == 1 && L == 0) r += 0.1; else if (R == 0) r -= 0.1;
    } else if (s->cnt < 80) {                                          /* :995-1001 */
        if (L == 1 && R == 0) r += 0.1; else if (L == 0) r -= 0.1;
    }
    if (R == 1 && L == 1) {                                            /* :1004-1007 */
        s->ds += 1;
        if (s->ds >= 16) r -= 2;
    }
    if (L == 1 && fabs(fe[3]) <= 0.1 && fabs(fe[4]) <= 0.1) r += 0.1;  /* :1014-1023 */
    if (R == 1 && fabs(fe[0]) <= 0.1 && fabs(fe[1]) <= 0.1) r += 0.1;  /* :1027-1036 */
    if (s->dead) { r -= 100.0; s->dead = 0; }                          /* :1057-1059 */

    /* ---- bookkeeping :674-678 */
    s->ep_ret += r; s->ep_t += 1; s->cnt += 1;
    *reward_out = r;
    /* TimeLimit wrapper from the registration, :15-19 */
    int timeout = (!dead) && (s->ep_t >= cfg->max_episode_steps);
    *done_out = dead || timeout;
    if (timeout_out) *timeout_out = timeout;
}

void plen_oracle_step_batch(const plen_oracle_model *m, const plen_oracle_config *cfg, plen_oracle_state *s, int n,
                            const double *actions, double *obs, double *reward, int *done, int *timeout,
                            int auto_reset, int n_threads) {
    (void)n_threads; /* threading is done by the caller: one ctypes call per slice, GIL released */
    for (int e = 0; e < n; e++) {
        int to = 0;
        plen_oracle_step(m, cfg, &s[e], actions + ORC_NJ * e, obs + 26 * e, &reward[e], &done[e], &to);
        if (timeout) timeout[e] = to;
        if (auto_reset && done[e]) plen_oracle_reset(m, cfg, &s[e], 0);
    }
}

void plen_oracle_reset_batch(const plen_oracle_model *m, const plen_oracle_config *cfg, plen_oracle_state *s, int n,
                             double *obs, int n_threads) {
    (void)n_threads; /* threading is done by the caller: one ctypes call per slice, GIL released */
    for (int e = 0; e < n; e++) plen_oracle_reset(m, cfg, &s[e], obs ? obs + 26 * e : 0);
}

int plen_oracle_sizeof_state(void) { return (int)sizeof(plen_oracle_state); }
int plen_oracle_sizeof_model(void) { return (int)sizeof(plen_oracle_model); }
int plen_oracle_sizeof_config(void) { return (int)sizeof(plen_oracle_config); }
