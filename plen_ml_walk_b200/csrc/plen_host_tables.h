// plen_host_tables.h -- host-side packing of plen_model / plen_config (C ABI structs) into the lane-major tables the
// kernels read.  Plain C++; shared by the CUDA library (plen_b200.cu) and the warp-emulation test harness.
#pragma once

#include <string.h>

#include "plen_env.cuh"

namespace plen {

inline void build_table(const plen_model *m, const plen_config *c, float *tab /* [T_ROWS*32] */) {
    memset(tab, 0, sizeof(float) * T_ROWS * 32);
    for (int l = 0; l < 32; l++) {
        const bool joint = l >= 6 && l < 24;
        for (int k = 0; k < 9; k++) tab[(T_RPJ + k) * 32 + l] = joint ? m->R_pj[l][k] : ((k % 4 == 0) ? 1.0f : 0.0f);
        for (int k = 0; k < 3; k++) {
            tab[(T_PPJ + k) * 32 + l] = joint ? m->p_pj[l][k] : 0.0f;
            tab[(T_AXIS + k) * 32 + l] = joint ? m->axis[l][k] : 0.0f;
            tab[(T_COM + k) * 32 + l] = (joint || l == 0) ? m->com[l][k] : 0.0f;
        }
        tab[T_MASS * 32 + l] = (joint || l == 0) ? m->mass[l] : 0.0f;
        for (int k = 0; k < 6; k++) tab[(T_INERTIA + k) * 32 + l] = (joint || l == 0) ? m->inertia[l][k] : 0.0f;
        tab[T_LOWER * 32 + l] = joint ? m->lower[l] : 0.0f;
        tab[T_UPPER * 32 + l] = joint ? m->upper[l] : 0.0f;
        int cs = l, ce = l;
        if (joint) {
            cs = m->chain_start[l];
            ce = l;
            while (ce + 1 < 24 && m->chain_start[ce + 1] == cs) ce++;
        }
        tab[T_CS * 32 + l] = (float)cs;
        tab[T_CE * 32 + l] = (float)ce;
        tab[T_ENVLO * 32 + l] = joint ? (float)c->env_lo[l - 6] : 0.0f;
        tab[T_ENVHI * 32 + l] = joint ? (float)c->env_hi[l - 6] : 0.0f;
        // box collider `l` (lane l tests box l against the ground; lanes >= n_boxes test nothing)
        const bool has = l < m->n_boxes && l < PLEN_MAX_BOXES;
        tab[T_BOX_LANE * 32 + l] = has ? (float)m->box_lane[l] : 0.0f;
        for (int k = 0; k < 3; k++) {
            tab[(T_BOX_C + k) * 32 + l] = has ? m->box_center[l][k] : 0.0f;
            tab[(T_BOX_H + k) * 32 + l] = has ? m->box_half[l][k] : 0.0f;
        }
        for (int k = 0; k < 9; k++) tab[(T_BOX_R + k) * 32 + l] = has ? m->box_rot[l][k] : ((k % 4 == 0) ? 1.0f : 0.0f);
        tab[T_BOX_REST * 32 + l] = has ? m->box_rest[l] : 0.0f;
    }
}

inline void build_devconfig(const plen_model *m, const plen_config *c, DevConfig *d, EnvRanges *r) {
    memset(d, 0, sizeof *d);
    d->dt = c->dt; d->inv_dt = 1.0f / c->dt; d->gravity_z = c->gravity_z;
    d->motor_imp = c->motor_max_force * c->dt;
    d->kp_over_dt = c->motor_kp / c->dt; d->one_minus_kd = 1.0f - c->motor_kd;
    d->linear_damping = c->linear_damping;
    d->mu_lateral = c->mu_lateral; d->mu_spinning = c->mu_spinning; d->mu_rolling = c->mu_rolling;
    d->restitution = c->restitution; d->rest_thresh = c->restitution_vel_threshold;
    d->erp_contact_over_dt = c->erp_contact / c->dt; d->erp_joint_over_dt = c->erp_joint / c->dt;
    d->linear_slop = c->linear_slop; d->warm = c->warmstart_factor; d->hull_margin = c->hull_margin;
    d->vmax = c->max_coord_velocity; d->residual_threshold = c->residual_threshold;
    d->link_contacts = (c->link_contacts && m->n_boxes > 0) ? 1 : 0; d->mu_link = c->mu_link;
    d->n_boxes = m->n_boxes < PLEN_MAX_BOXES ? m->n_boxes : PLEN_MAX_BOXES;
    // persistent sole manifold: needs the hull vertex lists; d->hull is set by the owner of the table memory
    d->sole_manifold = (c->sole_manifold && m->n_hull[0] > 0 && m->n_hull[1] > 0) ? 1 : 0;
    for (int f = 0; f < 2; f++) d->n_hull[f] = m->n_hull[f] < PLEN_MAX_HULL ? m->n_hull[f] : PLEN_MAX_HULL;
    d->hull = nullptr; d->support_tie = c->support_tie;
    for (int f = 0; f < 2; f++) {
        d->foot_break[f] = m->foot_break[f];
        d->foot_lane[f] = m->foot_lane[f];
        for (int p = 0; p < 4; p++)
            for (int k = 0; k < 3; k++) d->foot_pts[f][p][k] = m->foot_pts[f][p][k];
    }
    for (int k = 0; k < 3; k++) d->start_pos[k] = c->start_pos[k];
    d->substeps = c->substeps; d->reset_ticks = c->reset_ticks; d->iterations = c->solver_iterations;
    d->joint_act = c->joint_act; d->max_episode_steps = c->max_episode_steps; d->auto_reset = c->auto_reset;
    for (int k = 0; k < 18; k++) { r->lo[k] = c->env_lo[k]; r->hi[k] = c->env_hi[k]; }
}

// record of a robot teleported to the start pose with zero joints (plen_env.py:561-565), before the settle ticks
inline void init_record(const plen_config *c, float *rec /* [96] */) {
    memset(rec, 0, sizeof(float) * PLEN_STATE_WORDS);
    for (int k = 0; k < 3; k++) rec[W_POS + k] = c->start_pos[k];
    rec[W_QUAT + 3] = 1.0f;
}

inline int default_config(plen_config *c, int joint_act) {
    static const double lo[18] = {-1.57, -0.15, -0.95, -0.9, -0.95, -0.8, -1.57, -1.5, -0.75, -0.3, -1.2, -0.4,
                                  -1.57, -0.15, -0.2, -1.57, -0.15, -0.2};   // plen_env.py:148-167
    static const double hi[18] = {1.57, 1.5, 0.75, 0.3, 1.2, 0.4, 1.57, 0.15, 0.95, 0.9, 0.95, 0.8,
                                  1.57, 1.57, 0.35, 1.57, 1.57, 0.35};
    memset(c, 0, sizeof *c);
    c->dt = 1.0f / 240.0f; c->substeps = 4; c->reset_ticks = 8; c->gravity_z = -9.81f;
    c->start_pos[0] = 0.0f; c->start_pos[1] = 0.0f; c->start_pos[2] = 0.158f;
    c->motor_max_force = 0.15f; c->joint_act = joint_act ? 1 : 0;
    c->linear_damping = joint_act ? 0.1f : 0.0f;
    c->mu_lateral = 0.8f * 0.8f; c->mu_spinning = 0.1f * 0.8f; c->mu_rolling = (joint_act ? 0.01f : 0.1f) * 0.8f;
    c->restitution = 0.25f;
    for (int k = 0; k < 18; k++) { c->env_lo[k] = lo[k]; c->env_hi[k] = hi[k]; }
    c->max_episode_steps = 500;
    c->motor_kp = 0.1f; c->motor_kd = 1.0f; c->solver_iterations = 50; c->residual_threshold = 1e-7f;
    c->erp_contact = 0.08f; c->erp_joint = 0.2f; c->linear_slop = 1e-5f; c->warmstart_factor = 0.1f;
    c->restitution_vel_threshold = 0.2f; c->hull_margin = 0.001f; c->max_coord_velocity = 100.0f;
    c->auto_reset = 1;
    c->link_contacts = 1; c->mu_link = 0.5f * 0.8f;
    c->sole_manifold = 0; c->support_tie = 1.0e-7f;
    return 0;
}

}  // namespace plen
