// plen_solve.cuh -- second half of a physics tick (k_solve): projected Gauss-Seidel over the rows set up by
// tick_dynamics, then delta-v and semi-implicit integration.  sm_100a device code (also compiled by tests/emu).
//
// Mapping: EIGHT lanes per robot, four robots per warp.  The solver state of a robot is the 30-vector
//   x = [18 joint velocity changes; right-foot twist change (6); left-foot twist change (6)]
// spread over the 8 lanes g of its group as four registers s[slot] = x[g + 8 slot]:
//   s[0] = joint g      s[1] = joint 8+g      s[2] = joint 16+g (g < 2) | right-foot component g-2 (g >= 2)
//   s[3] = left-foot component g-2 (g >= 2)
// Contact points are COMPACTED per foot (slot q = 4 f + k = k-th active point of foot f), so a warp only loops to the
// largest active-point count among its four robots; k_solve additionally sorts robots by contact load inside 64-robot
// tiles so that the four robots of a warp have similar counts.
// foot twist components: 0..2 angular (wx wy wz), 3..5 linear (vx vy vz) about the base origin, world axes.
// A row update with impulse change delta is   s += (float4 of a column of G, one LDS.128) * delta   on every lane;
// the owner lane of a row (the lane whose register holds the row's "own" velocity component) computes the candidate.
// Row order, clamps, friction cone, warm start and early exit follow Bullet's btMultiBodyConstraintSolver as restated
// in oracle/plen_oracle.c (the parity checker): per iteration [limits, servos] (alternating direction), contact
// normals, spinning rows, rolling rows, lateral pairs with the implicit cone; exit when the largest squared row
// velocity change of the iteration is <= residual_threshold.  A robot that has converged is frozen (its row scalars are
// zeroed, so every later update is exactly zero) while the other robots of the warp keep iterating.
#pragma once

#include "plen_device.cuh"

namespace plen {

#ifndef PLEN_HOST_EMU
struct __align__(16) vec4 { float x, y, z, w; };
#else
struct vec4 { float x, y, z, w; };
#endif

PLEN_DEV float sel3(const float *a, int k) { return (k == 0) ? a[0] : ((k == 1) ? a[1] : a[2]); }

// solver lane that holds twist component k of either foot
#define PLEN_LN(f, k) (2 + (k))
// G column of twist component k of foot f
#define PLEN_COL(f, k) (18 + 6 * (f) + (k))

// s01 += (x0, x1) * d  as one packed FFMA2 (sm_100a fma.rn.f32x2 with a broadcast scalar multiplier)
PLEN_DEV void fma2(float &a0, float &a1, float x0, float x1, float d) {
#ifndef PLEN_HOST_EMU
    asm("{ .reg .b64 ra, rb, rc; mov.b64 ra, {%0, %1}; mov.b64 rb, {%2, %3}; mov.b64 rc, {%4, %4};\n\t"
        "fma.rn.f32x2 ra, rb, rc, ra; mov.b64 {%0, %1}, ra; }"
        : "+f"(a0), "+f"(a1) : "f"(x0), "f"(x1), "f"(d));
#else
    a0 = fmaf(x0, d, a0); a1 = fmaf(x1, d, a1);
#endif
}

struct SolveState {
    float s[4];
    float res;
};

PLEN_DEV void apply_col(SolveState &S, const vec4 *Gs4, int g, int c, float db) {
    const vec4 v = Gs4[c * 8 + g];
    fma2(S.s[0], S.s[1], v.x, v.y, db);
    fma2(S.s[2], S.s[3], v.z, v.w, db);
}

// One robot = lanes (lane & 24) .. +7 of the warp.  srec: this robot's solve record (global); Gs: this robot's 960-word
// shared staging area; state: this robot's 96-word state record (global), updated in place.
PLEN_DEV void solve_tick(const DevConfig &cfg, const float *__restrict__ srec, float *Gs, float *__restrict__ state,
                         int lane, bool valid) {
    const int g = lane & 7, gb = lane & 24;
#define GSH(v, l) shfl((v), gb | (l))
    vec4 *Gs4 = reinterpret_cast<vec4 *>(Gs);
    {
        const vec4 *src4 = reinterpret_cast<const vec4 *>(srec + SR_G);
        const vec4 z4 = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll 6
        for (int k = g; k < 240; k += 8) Gs4[k] = valid ? src4[k] : z4;
    }
    warp_sync();

    // ---- per-lane row scalars
    float m_rhs[3] = {0, 0, 0}, m_dinv[3] = {0, 0, 0}, m_d[3] = {0, 0, 0}, m_lam[3] = {0, 0, 0};
    float l_dir[3] = {0, 0, 0}, l_rhs[3] = {0, 0, 0}, l_lam[3] = {0, 0, 0};
    unsigned man = 0;
    if (valid) {
        vec4 t;
        t = *reinterpret_cast<const vec4 *>(srec + SR_MRHS + 4 * g); m_rhs[0] = t.x; m_rhs[1] = t.y; m_rhs[2] = t.z;
        t = *reinterpret_cast<const vec4 *>(srec + SR_MDINV + 4 * g); m_dinv[0] = t.x; m_dinv[1] = t.y; m_dinv[2] = t.z;
        t = *reinterpret_cast<const vec4 *>(srec + SR_LDIR + 4 * g); l_dir[0] = t.x; l_dir[1] = t.y; l_dir[2] = t.z;
        t = *reinterpret_cast<const vec4 *>(srec + SR_LRHS + 4 * g); l_rhs[0] = t.x; l_rhs[1] = t.y; l_rhs[2] = t.z;
        m_d[0] = Gs[g * 32 + 4 * g];                 // G[j][j], j = g
        m_d[1] = Gs[(8 + g) * 32 + 4 * g + 1];       // j = 8 + g
        m_d[2] = (g < 2) ? Gs[(16 + g) * 32 + 4 * g + 2] : 0.0f;
        man = (unsigned)srec[SR_BASE + 13];
    }
    // active-point counts per foot: this robot's and the largest among the four robots of the warp
    const int n0 = popc_(man & 15u), n1 = popc_((man >> 4) & 15u);
    const int nmax0 = (int)redux_max((unsigned)n0), nmax1 = (int)redux_max((unsigned)n1);
    const bool man_any = (nmax0 | nmax1) != 0;
    const unsigned lim_any = redux_or(((l_dir[0] != 0.0f) ? (1u << g) : 0u) | ((l_dir[1] != 0.0f) ? (256u << g) : 0u) |
                                      ((l_dir[2] != 0.0f) ? (65536u << g) : 0u));
    float c_rhs[8], c_dinv[8], c_d[8], c_lam[8], lamN[8], px[8], py[8], pz[8];
#pragma unroll
    for (int p = 0; p < 8; p++) { c_rhs[p] = c_dinv[p] = c_d[p] = c_lam[p] = lamN[p] = px[p] = py[p] = pz[p] = 0.0f; }
    if (man_any && valid) {
#pragma unroll
        for (int p = 0; p < 8; p++) {
            c_rhs[p] = srec[SR_CRHS + 8 * p + g];
            c_dinv[p] = srec[SR_CDINV + 8 * p + g];
            c_d[p] = srec[SR_CD + 8 * p + g];
            const vec4 t = *reinterpret_cast<const vec4 *>(srec + SR_PT + 4 * p);
            px[p] = t.x; py[p] = t.y; pz[p] = t.z;
            lamN[p] = srec[SR_LAMC + p] * cfg.warm;      // zero for unused slots
        }
    }

    SolveState S;
    S.s[0] = S.s[1] = S.s[2] = S.s[3] = 0.0f;
    S.res = 0.0f;

#define NORMAL_COLUMN(p, f, db)                                                   \
    {                                                                             \
        apply_col(S, Gs4, g, PLEN_COL(f, 0), py[p] * (db));                       \
        apply_col(S, Gs4, g, PLEN_COL(f, 1), -px[p] * (db));                      \
        apply_col(S, Gs4, g, PLEN_COL(f, 5), (db));                               \
    }

    // ---- warm start of the normal rows from the cached impulses
    if (man_any) {
#pragma unroll
        for (int p = 0; p < 8; p++)
            if ((p & 3) < ((p >> 2) ? nmax1 : nmax0)) NORMAL_COLUMN(p, (p >> 2), lamN[p]);
    }

#define SERVO_ROW(slot, l)                                                                              \
    {                                                                                                   \
        const float x_ = fmaf(-S.s[slot], m_dinv[slot], m_rhs[slot]);                                   \
        const float dl_ = fminf(fmaxf(x_, -cfg.motor_imp - m_lam[slot]), cfg.motor_imp - m_lam[slot]); \
        const float db_ = GSH(dl_, l);                                                                  \
        if (g == (l)) { m_lam[slot] += dl_; S.res = fmaxf(S.res, fabsf(dl_ * m_d[slot])); }             \
        apply_col(S, Gs4, g, 8 * (slot) + (l), db_);                                                    \
    }

    // spinning / rolling row of point p: own component k of foot f, friction coefficient mu
#define TORSION_ROW(p, f, k, mu)                                                                \
    {                                                                                           \
        const float x_ = fmaf(-S.s[2 + (f)], c_dinv[p], c_rhs[p]);                              \
        const float lim_ = (mu) * lamN[p];                                                      \
        float dl_ = fminf(fmaxf(x_, -lim_ - c_lam[p]), lim_ - c_lam[p]);                        \
        dl_ = (lamN[p] > 0.0f) ? dl_ : 0.0f;   /* row skipped while the normal impulse is not positive */ \
        const float db_ = GSH(dl_, PLEN_LN(f, k));                                              \
        if (g == PLEN_LN(f, k)) { c_lam[p] += dl_; S.res = fmaxf(S.res, fabsf(dl_ * c_d[p])); } \
        apply_col(S, Gs4, g, PLEN_COL(f, k), db_);                                              \
    }

    bool alive = valid;
    int my_iters = 0;
    for (int it = 0; it < cfg.iterations; it++) {
        S.res = 0.0f;
        const bool fwd = (it & 1) != 0;
        // ---- non-contact rows: list = [limits in joint order, servos in joint order]; odd iterations forward, even reversed
        for (int half = 0; half < 2; half++) {
            const bool do_limits = fwd == (half == 0);
            if (do_limits) {
                unsigned msk = lim_any;     // rare: a joint beyond its +-1.7 rad URDF limit on some robot of the warp
                while (msk) {
                    const int j = fwd ? lowest_bit(msk) : highest_bit(msk);
                    msk &= ~(1u << j);
                    const int slot = j >> 3, l = j & 7;
                    const float ldir = sel3(l_dir, slot), llam = sel3(l_lam, slot);
                    const float x_ = sel3(l_rhs, slot) - (ldir * sel3(S.s, slot)) * sel3(m_dinv, slot);
                    const float nl = clampf(llam + x_, 0.0f, 100.0f);
                    const float dl_ = nl - llam;
                    const float db_ = GSH(dl_ * ldir, l);
                    if (g == l) {
                        l_lam[0] = (slot == 0) ? nl : l_lam[0];
                        l_lam[1] = (slot == 1) ? nl : l_lam[1];
                        l_lam[2] = (slot == 2) ? nl : l_lam[2];
                        S.res = fmaxf(S.res, fabsf(dl_ * sel3(m_d, slot)));
                    }
                    apply_col(S, Gs4, g, j, db_);
                }
            } else if (fwd) {
#pragma unroll
                for (int l = 0; l < 8; l++) SERVO_ROW(0, l);
#pragma unroll
                for (int l = 0; l < 8; l++) SERVO_ROW(1, l);
#pragma unroll
                for (int l = 0; l < 2; l++) SERVO_ROW(2, l);
            } else {
#pragma unroll
                for (int l = 1; l >= 0; l--) SERVO_ROW(2, l);
#pragma unroll
                for (int l = 7; l >= 0; l--) SERVO_ROW(1, l);
#pragma unroll
                for (int l = 7; l >= 0; l--) SERVO_ROW(0, l);
            }
        }
        if (man_any) {
            // ---- contact normals (lower bound 0; the 1e10 upper bound of the reference never binds)
#pragma unroll
            for (int p = 0; p < 8; p++) {
                if ((p & 3) >= ((p >> 2) ? nmax1 : nmax0)) continue;
                const int f = p >> 2;
                const float wx = GSH(S.s[2 + f], PLEN_LN(f, 0)), wy = GSH(S.s[2 + f], PLEN_LN(f, 1));
                const float r_ = fmaf(wx, py[p], fmaf(-wy, px[p], S.s[2 + f]));
                const float x_ = fmaf(-r_, c_dinv[p], c_rhs[p]);
                const float dl_ = fmaxf(x_, -lamN[p]);
                const float db_ = GSH(dl_, PLEN_LN(f, 5));
                lamN[p] += db_;
                if (g == PLEN_LN(f, 5)) S.res = fmaxf(S.res, fabsf(dl_ * c_d[p]));
                NORMAL_COLUMN(p, f, db_);
            }
            // ---- all spinning rows, then the rolling rows point by point (t1, t2)
#pragma unroll
            for (int p = 0; p < 8; p++) {
                if ((p & 3) >= ((p >> 2) ? nmax1 : nmax0)) continue;
                TORSION_ROW(p, (p >> 2), 2, cfg.mu_spinning);
            }
#pragma unroll
            for (int p = 0; p < 8; p++) {
                if ((p & 3) >= ((p >> 2) ? nmax1 : nmax0)) continue;
                TORSION_ROW(p, (p >> 2), 1, cfg.mu_rolling);
                TORSION_ROW(p, (p >> 2), 0, cfg.mu_rolling);
            }
            // ---- lateral pairs with the implicit friction cone (resolveConeFrictionConstraintRows)
#pragma unroll
            for (int p = 0; p < 8; p++) {
                if ((p & 3) >= ((p >> 2) ? nmax1 : nmax0)) continue;
                const int f = p >> 2, LA = PLEN_LN(f, 4), LB = PLEN_LN(f, 3);
                const float wx = GSH(S.s[2 + f], PLEN_LN(f, 0)), wy = GSH(S.s[2 + f], PLEN_LN(f, 1)),
                            wz = GSH(S.s[2 + f], PLEN_LN(f, 2));
                const float rA = fmaf(wz, px[p], fmaf(-wx, pz[p], S.s[2 + f]));     // lane LA: own = vy
                const float rB = fmaf(wy, pz[p], fmaf(-wz, py[p], S.s[2 + f]));     // lane LB: own = vx
                const float r_ = (g == LA) ? rA : rB;
                const float sum_ = c_lam[p] + fmaf(-r_, c_dinv[p], c_rhs[p]);
                const float sumA = GSH(sum_, LA), sumB = GSH(sum_, LB);
                const float lim = cfg.mu_lateral * lamN[p];
                float nA = sumA, nB = sumB;
                if (fabsf(sumA) > lim || fabsf(sumB) > lim) {
                    const float inv = rsqrtf(sumA * sumA + sumB * sumB);
                    const float cA = fabsf(lim * sumA * inv), cB = fabsf(lim * sumB * inv);
                    nA = clampf(sumA, -cA, cA);
                    nB = clampf(sumB, -cB, cB);
                }
                const float nmine = (g == LA) ? nA : nB;
                const float dmine = nmine - c_lam[p];
                const float dA = GSH(dmine, LA), dB = GSH(dmine, LB);
                const float v_ = dmine * c_d[p];
                const float vB = GSH(v_, LB);
                if (g == LA || g == LB) c_lam[p] = nmine;
                if (g == LA) S.res = fmaxf(S.res, fabsf(v_ + vB));     // pair residual = dA/dinvA + dB/dinvB
                apply_col(S, Gs4, g, PLEN_COL(f, 0), -pz[p] * dA);
                apply_col(S, Gs4, g, PLEN_COL(f, 1), pz[p] * dB);
                apply_col(S, Gs4, g, PLEN_COL(f, 2), px[p] * dA - py[p] * dB);
                apply_col(S, Gs4, g, PLEN_COL(f, 3), dB);
                apply_col(S, Gs4, g, PLEN_COL(f, 4), dA);
            }
        }
        // ---- residual of this iteration, per robot
        float res = S.res;
        res = fmaxf(res, shfl_xor(res, 1));
        res = fmaxf(res, shfl_xor(res, 2));
        res = fmaxf(res, shfl_xor(res, 4));
        if (alive) {
            my_iters = it + 1;
            if (res * res <= cfg.residual_threshold) {
                alive = false;      // freeze: every later row update of this robot is exactly zero
#pragma unroll
                for (int k = 0; k < 3; k++) { m_rhs[k] = 0.0f; m_dinv[k] = 0.0f; l_rhs[k] = 0.0f; }
#pragma unroll
                for (int p = 0; p < 8; p++) { c_rhs[p] = 0.0f; c_dinv[p] = 0.0f; }
            }
        }
        if (!ballot(alive)) break;
    }

    // ---- delta-v: joints = s[0..2]; base = B z with z = [servo + limit impulses; foot wrenches]
    float z[4];
    z[0] = m_lam[0] + l_dir[0] * l_lam[0];
    z[1] = m_lam[1] + l_dir[1] * l_lam[1];
    z[2] = (g < 2) ? m_lam[2] + l_dir[2] * l_lam[2] : 0.0f;
    z[3] = 0.0f;
    if (man_any) {
#pragma unroll
        for (int f = 0; f < 2; f++) {
            const int comp = g - 2;              // this lane's row type (negative: no contact rows)
            float w6[6] = {0, 0, 0, 0, 0, 0};
            float own = 0.0f;
#pragma unroll
            for (int pp = 0; pp < 4; pp++) {
                const int p = 4 * f + pp;
                const float lam = (comp == 5) ? lamN[p] : c_lam[p];
                own += lam;
                // angular parts of the 3-component rows: normal (py,-px,0), lateral t2 (0,pz,-py), flipped t1 (-pz,0,px)
                const float a0 = (comp == 5) ? py[p] : ((comp == 4) ? -pz[p] : 0.0f);
                const float a1 = (comp == 5) ? -px[p] : ((comp == 3) ? pz[p] : 0.0f);
                const float a2 = (comp == 3) ? -py[p] : ((comp == 4) ? px[p] : 0.0f);
                w6[0] = fmaf(a0, lam, w6[0]); w6[1] = fmaf(a1, lam, w6[1]); w6[2] = fmaf(a2, lam, w6[2]);
            }
#pragma unroll
            for (int k = 0; k < 6; k++) w6[k] += (comp == k) ? own : 0.0f;
#pragma unroll
            for (int m = 1; m < 8; m <<= 1)
#pragma unroll
                for (int k = 0; k < 6; k++) w6[k] += shfl_xor(w6[k], m);
            float mine = 0.0f;
#pragma unroll
            for (int k = 0; k < 6; k++) mine = (comp == k) ? w6[k] : mine;
            if (f == 0) z[2] = (g < 2) ? z[2] : mine; else z[3] = (g < 2) ? 0.0f : mine;
        }
    }
    float dvb[6];
#pragma unroll
    for (int k = 0; k < 6; k++) {
        float acc = 0.0f;
        if (valid) {
            const vec4 b = *reinterpret_cast<const vec4 *>(srec + SR_B + 32 * k + 4 * g);
            acc = b.x * z[0] + b.y * z[1] + b.z * z[2] + b.w * z[3];
        }
        dvb[k] = acc;
    }
#pragma unroll
    for (int m = 1; m < 8; m <<= 1)
#pragma unroll
        for (int k = 0; k < 6; k++) dvb[k] += shfl_xor(dvb[k], m);

    if (valid) {
        // ---- apply delta-v, integrate (semi-implicit; exponential-map quaternion update), write the state record back
        const vec4 vs = *reinterpret_cast<const vec4 *>(srec + SR_VSTAR + 4 * g);
        const vec4 q4 = *reinterpret_cast<const vec4 *>(srec + SR_Q + 4 * g);
        const float u0 = clampf(vs.x + S.s[0], -cfg.vmax, cfg.vmax), u1 = clampf(vs.y + S.s[1], -cfg.vmax, cfg.vmax),
                    u2 = clampf(vs.z + S.s[2], -cfg.vmax, cfg.vmax);
        state[W_U + 6 + g] = u0; state[W_Q + 6 + g] = q4.x + u0 * cfg.dt;
        state[W_U + 14 + g] = u1; state[W_Q + 14 + g] = q4.y + u1 * cfg.dt;
        if (g < 2) { state[W_U + 22 + g] = u2; state[W_Q + 22 + g] = q4.z + u2 * cfg.dt; }
        {   // cached normal impulse of ORIGINAL contact point g: slot = 4 f + (number of active points of the foot below it)
            const int f = g >> 2, q = 4 * f + popc_((man >> (4 * f)) & ((1u << (g & 3)) - 1u));
            float v = lamN[0];
#pragma unroll
            for (int k = 1; k < 8; k++) v = (q == k) ? lamN[k] : v;
            state[W_U + 24 + g] = ((man >> g) & 1u) ? v : 0.0f;
        }
        if (g == 0) {
            float ub[6];
#pragma unroll
            for (int k = 0; k < 6; k++) { ub[k] = clampf(srec[SR_BASE + k] + dvb[k], -cfg.vmax, cfg.vmax); state[W_U + k] = ub[k]; }

#pragma unroll
            for (int k = 0; k < 3; k++) state[W_POS + k] = srec[SR_BASE + 6 + k] + ub[3 + k] * cfg.dt;
            float fa = sqrtf(ub[0] * ub[0] + ub[1] * ub[1] + ub[2] * ub[2]);
            if (fa * cfg.dt > 0.78539816339f) fa = 0.78539816339f * cfg.inv_dt;
            float sc, cw;
            if (fa < 0.001f) {
                sc = 0.5f * cfg.dt - cfg.dt * cfg.dt * cfg.dt * 0.020833333333f * fa * fa;
                cw = cosf(0.5f * fa * cfg.dt);
            } else {
                float sn;
                sincos_(0.5f * fa * cfg.dt, &sn, &cw);
                sc = sn / fa;
            }
            const float dx = ub[0] * sc, dy = ub[1] * sc, dz = ub[2] * sc;
            const float qx = srec[SR_BASE + 9], qy = srec[SR_BASE + 10], qz = srec[SR_BASE + 11], qw = srec[SR_BASE + 12];
            const float rx = cw * qx + dx * qw + dy * qz - dz * qy;
            const float ry = cw * qy - dx * qz + dy * qw + dz * qx;
            const float rz = cw * qz + dx * qy - dy * qx + dz * qw;
            const float rw = cw * qw - dx * qx - dy * qy - dz * qz;
            const float nn = rsqrtf(rx * rx + ry * ry + rz * rz + rw * rw);
            state[W_QUAT + 0] = rx * nn; state[W_QUAT + 1] = ry * nn; state[W_QUAT + 2] = rz * nn; state[W_QUAT + 3] = rw * nn;
            state[W_MAN] = (float)man;
            state[W_ITERS] = (float)my_iters;
        }
    }
#undef GSH
#undef SERVO_ROW
#undef TORSION_ROW
#undef NORMAL_COLUMN
}

}  // namespace plen
