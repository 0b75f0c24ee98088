// plen_solve.cuh -- second half of a physics tick (k_solve): projected Gauss-Seidel over the rows set up by
// tick_dynamics, then delta-v and semi-implicit integration.  sm_100a device code (also compiled by tests/emu).
//
// Mapping: FOUR lanes per robot, eight robots per warp.  The solver state of a robot is the 32-slot vector
//   x = [18 joint velocity changes; right-foot twist change (6); 2 unused; left-foot twist change (6)]
// and lane g of the robot holds the eight consecutive entries s[k] = x[8 g + k]:
//   lane 0: joints 0..7    lane 1: joints 8..15    lane 2: joints 16, 17 + right-foot twist    lane 3: -, - + left-foot twist
// so either foot's twist (wx wy wz vx vy vz about the base origin, world axes) sits in slots 2..7 of ONE lane.  That lane
// owns every contact row of its foot: the row velocity is a local combination of its own registers (no gather), it
// computes the clamped candidate and broadcasts the impulse change with one SHFL.
// A row update with impulse change delta is   s += (8 entries of a column of G, two LDS.128) * delta   = four packed
// FFMA2 on every lane.  Servo rows are kept in velocity units (joint columns of G are pre-divided by G_jj in k_dyn):
// candidate = rhs - s, bounds (lo, hi) = -+ max impulse * G_jj minus the accumulated row value, and the row's velocity
// change -- the residual term -- is the broadcast value itself.
// Contact points are COMPACTED per foot (slot k = k-th active point), so a warp only loops to the largest active-point
// count among its eight robots; k_rank sorts robots by contact load so that the robots of a warp have similar counts.
// Row order, clamps, friction cone, warm start and early exit follow Bullet's btMultiBodyConstraintSolver as restated
// in oracle/plen_oracle.c (the parity checker): per iteration [limits, servos] (alternating direction), contact
// normals, spinning rows, rolling rows, lateral pairs with the implicit cone; exit when the largest squared row
// velocity change of the iteration is <= residual_threshold.  A robot that has converged is frozen (its bounds / row
// scalars are zeroed, so every later update is exactly zero) while the other robots of the warp keep iterating.
#pragma once

#include "plen_device.cuh"

namespace plen {


// Owner look-ahead (see solve_tick) per row family; each costs extra FFMA2 issue slots and removes one SHFL round trip
// from the dependency chain of the family.  Chosen by A/B timing on B200 (profiles/r1_v6_summary.md).
#ifndef PLEN_LA_SERVO
#define PLEN_LA_SERVO 1
#endif
#ifndef PLEN_REDUX_CONV
#define PLEN_REDUX_CONV 0      // convergence test: one partitioned redux.sync instead of two SHFL + FMNMX
#endif
#ifndef PLEN_LA_NORMAL
#define PLEN_LA_NORMAL 0
#endif
#ifndef PLEN_LA_TORSION
#define PLEN_LA_TORSION 0
#endif
#ifndef PLEN_LA_LATERAL
#define PLEN_LA_LATERAL 0
#endif
#ifndef PLEN_TORSION_BLOCK
#define PLEN_TORSION_BLOCK 1   // spinning / rolling rows of a foot as one private chain of the owner lane + one column update per functional
#endif

// Shared memory holds 24 of the 30 columns of G per robot: the 18 joint columns and the linear (vx vy vz) columns of
// either foot.  The six ANGULAR foot columns -- the ones the contact rows use most (normal 2 of 3, spinning, rolling,
// lateral 3 of 5 column updates) -- live in registers, which (a) takes 8 of the 11 LDS pairs out of every contact point
// and (b) shrinks the staging area to 24.1 KB per warp: 9 resident warps per SM instead of 7 (the kernel is bound by the
// latency of the Gauss-Seidel dependency chain, so resident warps are throughput).
// Words per robot: 24 columns x 32 words, plus one float4 so that the two robots of a quarter-warp hit disjoint banks
// with their LDS.128 (robot stride = 16 B mod 128 B)
#define PLEN_GS_COLS 24
#define PLEN_GS_WORDS (PLEN_GS_COLS * 32 + 4)
#define PLEN_SOLVE_ROBOTS 8      // robots per warp
// EXT == 2 only: the Jx / Bx vectors of the 3 x PLEN_MAX_BOX_POINTS box-contact rows of a robot (plen_device.cuh, XR_*),
// [row][Jx 32 | Bx 32] plus the same one-float4 skew between robots
#define PLEN_XS_WORDS (3 * PLEN_MAX_BOX_POINTS * 64 + 4)

// column of twist component k of foot f in the solve record (30 columns) / in the shared staging area (k >= 3 only)
#define PLEN_COL(f, k) (18 + 6 * (f) + (k))
#define PLEN_SCOL(f, k) (18 + 3 * (f) + (k) - 3)

// (a0, a1) += (x0, x1) * d  as one packed FFMA2 (sm_100a fma.rn.f32x2 with a broadcast scalar multiplier)
PLEN_DEV void fma2(float &a0, float &a1, float x0, float x1, float d) {
#ifndef PLEN_HOST_EMU
    asm("{ .reg .b64 ra, rb, rc; mov.b64 ra, {%0, %1}; mov.b64 rb, {%2, %3}; mov.b64 rc, {%4, %4};\n\t"
        "fma.rn.f32x2 ra, rb, rc, ra; mov.b64 {%0, %1}, ra; }"
        : "+f"(a0), "+f"(a1) : "f"(x0), "f"(x1), "f"(d));
#else
    a0 = fmaf(x0, d, a0); a1 = fmaf(x1, d, a1);
#endif
}

// Loop constants that come from the kernel parameters (friction coefficients, residual threshold, iteration cap) must sit
// in registers: ptxas otherwise re-reads each from the constant bank with an LDCU in front of every foot block and waits
// for it (~35 exposed cycles, eight times per iteration).  Neither an empty asm (only NVVM sees it) nor a SHFL of the
// value (folded: kernel parameters are warp-uniform) survives to SASS, so the constants take a round trip through shared
// memory with a volatile load, which cannot be rematerialised.
struct LoopConsts { float mu_spin, mu_roll, mu_lat, res_thr; int iterations, g; };
PLEN_DEV LoopConsts pin_loop_consts(const DevConfig &cfg, int lane) {
    LoopConsts c;
#ifndef PLEN_HOST_EMU
    __shared__ float pins[8];
    __shared__ int gpin[32];
    gpin[lane] = lane & 3;      // the lane's index inside its robot, too: ptxas otherwise re-derives it (S2R + LOP3) in every row
    if (lane == 0) {
        pins[0] = cfg.mu_spinning; pins[1] = cfg.mu_rolling; pins[2] = cfg.mu_lateral; pins[3] = cfg.residual_threshold;
        pins[4] = __int_as_float(cfg.iterations);
    }
    __syncwarp();
    volatile float *vp = pins;
    c.mu_spin = vp[0]; c.mu_roll = vp[1]; c.mu_lat = vp[2]; c.res_thr = vp[3]; c.iterations = __float_as_int(vp[4]);
    c.g = reinterpret_cast<volatile int *>(gpin)[lane];
#else
    (void)lane;
    c.mu_spin = cfg.mu_spinning; c.mu_roll = cfg.mu_rolling; c.mu_lat = cfg.mu_lateral; c.res_thr = cfg.residual_threshold;
    c.iterations = cfg.iterations;
    c.g = lane & 3;
#endif
    return c;
}
// 1/sqrt(x) as one MUFU.RSQ: rsqrtf() adds a subnormal-input rescue (FSETP + two predicated FMUL on the row chain)
PLEN_DEV float rsqrt_fast(float x) {
#ifndef PLEN_HOST_EMU
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return 1.0f / sqrtf(x);
#endif
}

// This lane's slice of G (8 rows of every column) is addressed through ONE 32-bit shared-window address plus immediate
// offsets: a generic pointer costs a 64-bit register pair, and under the register pressure of this kernel ptxas used to
// rematerialise it (S2R tid / S2R cta-window base / 10 integer instructions) in front of every contact row.
#ifndef PLEN_HOST_EMU
typedef unsigned GSlice;
PLEN_DEV GSlice g_slice(const float *Gs, int g) {
    unsigned a = (unsigned)__cvta_generic_to_shared(Gs) + 32u * (unsigned)g;
    asm volatile("" : "+r"(a) : : "memory");      // orders every later g_ld after the staging stores + warp_sync
    return a;
}
PLEN_DEV vec4 g_ld(GSlice a, int v4) {             // vec4 number v4 of the slice (column c -> 8 c, 8 c + 1)
    vec4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a + 16u * (unsigned)v4));
    return v;
}
#else
typedef const vec4 *GSlice;
PLEN_DEV GSlice g_slice(const float *Gs, int g) { return reinterpret_cast<const vec4 *>(Gs) + 2 * g; }
PLEN_DEV vec4 g_ld(GSlice a, int v4) { return a[v4]; }
#endif

PLEN_DEV void apply_col(float (&s)[8], GSlice Gl, int c, float db) {
    const vec4 a = g_ld(Gl, c * 8), b = g_ld(Gl, c * 8 + 1);
    fma2(s[0], s[1], a.x, a.y, db);
    fma2(s[2], s[3], a.z, a.w, db);
    fma2(s[4], s[5], b.x, b.y, db);
    fma2(s[6], s[7], b.z, b.w, db);
}

PLEN_DEV void apply_reg(float (&s)[8], const float (&a)[8], float db) {
    fma2(s[0], s[1], a[0], a[1], db);
    fma2(s[2], s[3], a[2], a[3], db);
    fma2(s[4], s[5], a[4], a[5], db);
    fma2(s[6], s[7], a[6], a[7], db);
}

// Owner look-ahead copies: entries 2..7 (a foot twist) of a register column / a staged column
PLEN_DEV void apply_reg6(float (&o)[8], const float (&a)[8], float d) {
    fma2(o[2], o[3], a[2], a[3], d);
    fma2(o[4], o[5], a[4], a[5], d);
    fma2(o[6], o[7], a[6], a[7], d);
}
PLEN_DEV void apply_vec(float (&o)[8], const vec4 &a, const vec4 &b, float d) {
    fma2(o[0], o[1], a.x, a.y, d);
    fma2(o[2], o[3], a.z, a.w, d);
    fma2(o[4], o[5], b.x, b.y, d);
    fma2(o[6], o[7], b.z, b.w, d);
}
PLEN_DEV void apply_vec6(float (&o)[8], const vec4 &a, const vec4 &b, float d) {
    fma2(o[2], o[3], a.z, a.w, d);
    fma2(o[4], o[5], b.x, b.y, d);
    fma2(o[6], o[7], b.z, b.w, d);
}

// vec4 number v4 of part (0: Jx, 1: Bx) of row `row` of an extension record; zeros for a robot without that row
PLEN_DEV vec4 x_vec(const float *srx, int row, int part, int v4, bool on) {
    const vec4 z4 = {0.0f, 0.0f, 0.0f, 0.0f};
    return on ? gld4(srx + XR_ROWS + row * XR_ROW_WORDS + 32 * part + 4 * v4) : z4;
}

PLEN_DEV void load8(const float *p, float (&o)[8]) {
    const vec4 a = gld4(p), b = gld4(p + 4);      // solve-record / state words: L2 only
    o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
}

// One robot = lanes (lane & 28) .. +3 of the warp.  srec: this robot's solve record (global); Gs: this robot's
// PLEN_GS_WORDS shared staging area; state: this robot's 96-word state record (global), updated in place.
// EXT != 0: the instance that also iterates the box-contact rows of the extension record srx (nx points of this robot, 0 for a
// robot without any).  Only warps that hold such a robot run it (k_rank puts them first in every tile), so the plain
// instance keeps its instruction stream.
template <int EXT>
PLEN_DEV void solve_tick(const DevConfig &cfg, const float *__restrict__ srec, float *Gs, float *__restrict__ state,
                         int lane, bool valid, const float *__restrict__ srx = nullptr, int nx = 0, float *Xs = nullptr) {
    const LoopConsts lc = pin_loop_consts(cfg, lane);
    const int g = lc.g;      // lane & 3, pinned in a register
#define GSH(v, l) shfl4((v), (l))
    {
        // stage the 24 columns: 48 16-byte pieces per lane, all in flight at once through cp.async (global -> shared without
        // a register round trip); the pieces are waited for after the register-resident parts of the record are requested
        vec4 *dst4 = reinterpret_cast<vec4 *>(Gs);
        const vec4 *src4 = reinterpret_cast<const vec4 *>(srec + SR_G);
        const vec4 z4 = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
        for (int k0 = 0; k0 < 8 * PLEN_GS_COLS; k0 += 4) {
            const int k = k0 + g;
            const int col = k0 >> 3, scol = (col < 18) ? col : ((col < 21) ? col + 3 : col + 6);      // skip the angular columns
#ifndef PLEN_HOST_EMU
            if (valid) {
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"((unsigned)__cvta_generic_to_shared(dst4 + k)),
                             "l"(src4 + 8 * scol + (k & 7)) : "memory");
            } else {
                dst4[k] = z4;
            }
#else
            dst4[k] = valid ? src4[8 * scol + (k & 7)] : z4;
#endif
        }
    }
    float A[2][3][8];      // this lane's 8 rows of the angular columns (wx wy wz) of either foot
#pragma unroll
    for (int f = 0; f < 2; f++)
#pragma unroll
        for (int c = 0; c < 3; c++) {
#pragma unroll
            for (int k = 0; k < 8; k++) A[f][c][k] = 0.0f;
            if (valid) load8(srec + SR_G + 32 * PLEN_COL(f, c) + 8 * g, A[f][c]);
        }

    // ---- servo rows of this lane (entries >= 18 carry zeros): velocity-unit rhs and bounds
    float m_rhs[8], m_lo[8], m_hi[8];
    unsigned limbits = 0;          // bit k: joint 8 g + k is beyond a URDF limit (its limit row exists this tick)
    unsigned man = 0;
    float fric_s = 1.0f;           // per-robot scale of the three friction coefficients (plen_set_env_scales)
#pragma unroll
    for (int k = 0; k < 8; k++) m_rhs[k] = m_lo[k] = m_hi[k] = 0.0f;
    if (valid) {
        float d[8], ld[8];
        fric_s = gld(srec + SR_BASE + 14);
        const float motor_s = gld(srec + SR_BASE + 15);
        load8(srec + SR_MRHS + 8 * g, m_rhs);
        load8(srec + SR_MD + 8 * g, d);
        load8(srec + SR_LDIR + 8 * g, ld);
#pragma unroll
        for (int k = 0; k < 8; k++) {
            m_hi[k] = cfg.motor_imp * motor_s * d[k];
            m_lo[k] = -m_hi[k];
            limbits |= (ld[k] != 0.0f) ? (1u << k) : 0u;
        }
        man = (unsigned)gld(srec + SR_BASE + 13);
    }
    // active-point counts per foot: this robot's and the largest among the robots of the warp
    const int n0 = popc_(man & 15u), n1 = popc_((man >> 4) & 15u);
    const int nmax0 = (int)redux_max((unsigned)n0), nmax1 = (int)redux_max((unsigned)n1);
    const bool man_any = (nmax0 | nmax1) != 0;
    const unsigned lim_any = redux_or(limbits << (8 * g));       // joint j <-> bit j (lanes 0..2 only carry joints)
    // limit-row impulses (velocity units; rare path only) are kept in shared memory, in the unused row 24 of column j
    // of this robot's G (k_dyn writes zeros there; lane 3 reads it into its unused slot s[0])
#define L_LAM(j) Gs[((j) < 18 ? (j) : 0) * 32 + 24]

    // ---- contact rows of this lane's foot (lane 2: right, lane 3: left; lanes 0, 1 hold zeros).  Twist component c of a
    //      row: 0 roll(wx) 1 roll(wy) 2 spin(wz) 3 lateral B (vx) 4 lateral A (vy) 5 normal (vz).  The spinning / rolling
    //      rows are functionals of ONE angular twist component, so their rhs / dinv / d are the same for every point
    //      of a foot (t_*[c], read from the foot's first slot); only their impulses are per point.
    float c_rhs[4][3], c_dinv[4][3], c_d[4][3];      // [point k][c - 3]
    float t_rhs[3], t_dinv[3], t_d[3];
    float c_lam[4][6];
    float px[8], py[8], pz[8];     // all eight compacted points (every lane scales the columns of either foot)
#pragma unroll
    for (int k = 0; k < 4; k++) {
#pragma unroll
        for (int c = 0; c < 3; c++) c_rhs[k][c] = c_dinv[k][c] = c_d[k][c] = 0.0f;
#pragma unroll
        for (int c = 0; c < 6; c++) c_lam[k][c] = 0.0f;
    }
#pragma unroll
    for (int c = 0; c < 3; c++) t_rhs[c] = t_dinv[c] = t_d[c] = 0.0f;
#pragma unroll
    for (int p = 0; p < 8; p++) px[p] = py[p] = pz[p] = 0.0f;
    if (man_any && valid) {
#pragma unroll
        for (int p = 0; p < 8; p++) {
            const vec4 t = gld4(srec + SR_PT + 4 * p);
            px[p] = t.x; py[p] = t.y; pz[p] = t.z;
        }
        if (g >= 2) {
            const int f = g - 2;
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const int o = 8 * (4 * f) + 2 + c;       // an empty foot has zeros here
                t_rhs[c] = gld(srec + SR_CRHS + o); t_dinv[c] = gld(srec + SR_CDINV + o); t_d[c] = gld(srec + SR_CD + o);
            }
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int o = 8 * (4 * f + k) + 2 + 3;
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    c_rhs[k][c] = gld(srec + SR_CRHS + o + c);
                    c_dinv[k][c] = gld(srec + SR_CDINV + o + c);
                    c_d[k][c] = gld(srec + SR_CD + o + c);
                }
                c_lam[k][5] = gld(srec + SR_LAMC + 4 * f + k) * cfg.warm;      // zero for unused slots
            }
        }
    }

    // ---- box-contact rows (EXT): lane q of the robot owns point q (its three impulses and row scalars live in that lane's
    //      registers); the 64-word Jx | Bx vectors of every row are staged in shared memory
    int nxmax = 0;
    float x_rhs[3] = {0.0f, 0.0f, 0.0f}, x_dinv[3] = {0.0f, 0.0f, 0.0f}, x_d[3] = {0.0f, 0.0f, 0.0f}, x_lam[3] = {0.0f, 0.0f, 0.0f};
    float resX = 0.0f;
    if (EXT) {
        nxmax = (int)redux_max((unsigned)nx);
        // the 64-word Jx | Bx vectors of every row are read from the extension record where they are used (ld.global.v4: a
        // robot's 3 KB stay in L1 / L2 for the 50 iterations): no shared memory, so this instance runs in the same launch
        // and with the same footprint as the plain one.  Robots without box contacts (nx = 0) get zero vectors.
        if (valid && g < nx) {
#pragma unroll
            for (int r = 0; r < 3; r++) {
                const float *row = srx + XR_ROWS + (3 * g + r) * XR_ROW_WORDS;
                x_rhs[r] = gld(row + XR_RHS); x_dinv[r] = gld(row + XR_DINV); x_d[r] = gld(row + XR_D);
            }
        }
        // EXT == 2 (k_solve_x, a launch of its own with room for it): the vectors are staged in shared memory instead,
        // [row][Jx 32 | Bx 32] per robot (PLEN_XS_WORDS); 1 % of the step at 131,072 robots against the global reads (r2_ab3)
        if (EXT == 2) {
            for (int row = 0; row < 3 * nxmax; row++) {
                vec4 *dst = reinterpret_cast<vec4 *>(Xs + row * 64);
#pragma unroll
                for (int v4 = 0; v4 < 2; v4++) {
                    dst[2 * g + v4] = x_vec(srx, row, 0, 2 * g + v4, valid && row < 3 * nx);
                    dst[8 + 2 * g + v4] = x_vec(srx, row, 1, 2 * g + v4, valid && row < 3 * nx);
                }
            }
            warp_sync();
        }
    }
#define X_VEC(row, part, half)                                                                              \
    ((EXT == 2) ? reinterpret_cast<const vec4 *>(Xs + (row) * 64 + 32 * (part))[2 * g + (half)]              \
                : x_vec(srx, (row), (part), 2 * g + (half), valid && (row) < 3 * nx))
#define X_DOT(row, out)                                                                                     \
    {                                                                                                       \
        const vec4 xja_ = X_VEC(row, 0, 0), xjb_ = X_VEC(row, 0, 1);                                          \
        float xr_ = xja_.x * s[0];                                                                            \
        xr_ = fmaf(xja_.y, s[1], xr_); xr_ = fmaf(xja_.z, s[2], xr_); xr_ = fmaf(xja_.w, s[3], xr_);                 \
        xr_ = fmaf(xjb_.x, s[4], xr_); xr_ = fmaf(xjb_.y, s[5], xr_); xr_ = fmaf(xjb_.z, s[6], xr_); xr_ = fmaf(xjb_.w, s[7], xr_); \
        xr_ += shfl_xor(xr_, 1);                                                                              \
        xr_ += shfl_xor(xr_, 2);                                                                              \
        out = xr_;                                                                                           \
    }

    float s[8];
#pragma unroll
    for (int k = 0; k < 8; k++) s[k] = 0.0f;
    float res = 0.0f, resF = 0.0f;
    // Owner look-ahead.  Consecutive rows of a block (8 joints of a lane / the points of one foot) are owned by the SAME
    // lane, and the next candidate only needs the owner's own entries of x.  The owner therefore keeps a private copy o of
    // its entries that it advances with its own impulse change dl_ BEFORE the broadcast, while s (every lane's slice of
    // x) takes the same update after the SHFL.  In the owner lane o and s receive identical operations in identical
    // order, so they agree bit for bit; in the other lanes o is scratch.  This takes the SHFL round trip (the longest
    // link of the Gauss-Seidel dependency chain) off the critical path except once per block, where o is refreshed from s.
    float o[8] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
    // servo blocks keep the private copy as the row ERROR u = x - rhs (so the candidate is clamp(-u): the subtraction of
    // the right-hand side leaves the row-to-row chain, FMNMX -> FMNMX -> FFMA2 instead of FADD -> FMNMX -> FMNMX -> FFMA2)
#define OWN_SYNC()  { _Pragma("unroll") for (int k_ = 0; k_ < 8; k_++) o[k_] = s[k_] - m_rhs[k_]; }
#define OWN_SYNC6() { _Pragma("unroll") for (int k_ = 2; k_ < 8; k_++) o[k_] = s[k_]; }

    // The active points of either foot occupy slots 0 .. n-1 (compacted), so the walk over a foot stops at the first empty
    // slot: min(n + 1, 4) warp-uniform tests per foot instead of 4 (k_rank groups robots of similar load into a warp).
    // Usage: FOR_ACTIVE_POINTS { body }.  p = slot (0..7), f = foot, k = point of the foot, all compile-time after unrolling.
#define FOR_ACTIVE_POINTS                                                          \
    _Pragma("unroll") for (int f = 0; f < 2; f++)                                  \
        _Pragma("unroll") for (int k = 0, p = 4 * f; k < 4; k++, p++)              \
            if (k >= (f ? nmax1 : nmax0)) break; else

#define NORMAL_COLUMN(p, f, db)                              \
    {                                                        \
        apply_reg(s, A[f][0], py[p] * (db));                 \
        apply_reg(s, A[f][1], -px[p] * (db));                \
        apply_col(s, Gl, PLEN_SCOL(f, 5), (db));             \
    }

    // ---- the staged columns must have landed before the first row update
#ifndef PLEN_HOST_EMU
    asm volatile("cp.async.wait_all;" ::: "memory");
#endif
    warp_sync();
    const GSlice Gl = g_slice(Gs, g);      // this lane's 8 rows of column c: vec4 8 c and 8 c + 1 of the slice

    // ---- warm start of the normal rows from the cached impulses
    if (man_any) {
        FOR_ACTIVE_POINTS {
            const float db_ = GSH(c_lam[k][5], 2 + f);
            NORMAL_COLUMN(p, f, db_);
        }
    }

    // servo row of joint 8 b + k (owner lane b)
#define SERVO_ROW(b, k)                                                \
    {                                                                  \
        const float dl_ = fminf(fmaxf(PLEN_LA_SERVO ? -o[k] : m_rhs[k] - s[k], m_lo[k]), m_hi[k]); \
        const vec4 ca_ = g_ld(Gl, (8 * (b) + (k)) * 8), cb_ = g_ld(Gl, (8 * (b) + (k)) * 8 + 1); \
        if (PLEN_LA_SERVO) apply_vec(o, ca_, cb_, dl_);                \
        const float db_ = GSH(dl_, b);                                 \
        if (g == (b)) { m_lo[k] -= dl_; m_hi[k] -= dl_; }              \
        res = fmaxf(res, fabsf(db_));                                  \
        apply_vec(s, ca_, cb_, db_);                                   \
    }

    // joint-limit row of joint 8 b + k (rare: the joint is beyond +-1.7 rad); lower bound 0, upper bound 100 (impulse units)
#define LIMIT_ROW(b, k)                                                                  \
    {                                                                                    \
        const float ldir_ = gld(srec + SR_LDIR + 8 * g + (k)), up_ = 100.0f * gld(srec + SR_MD + 8 * g + (k)); \
        const float x_ = gld(srec + SR_LRHS + 8 * g + (k)) - ldir_ * s[k];                     \
        const float ol_ = L_LAM(8 * g + (k));                                            \
        const float nl_ = alive ? clampf(ol_ + x_, 0.0f, up_) : ol_;                     \
        const float dl_ = (nl_ - ol_) * ldir_;                                           \
        const float db_ = GSH(dl_, b);                                                   \
        warp_sync();      /* lanes without a joint at slot k read word 0 above: keep that read ahead of lane 0's write */ \
        if (g == (b)) L_LAM(8 * g + (k)) = nl_;                                          \
        /* lane 3's slice of column j includes the L_LAM word (row 24, an unused slot of x): order its read after the  \
           owner's write -- the value is never used, but the access pair is a hazard for racecheck (rare path: free) */ \
        warp_sync();                                                                     \
        res = fmaxf(res, fabsf(db_));                                                    \
        apply_col(s, Gl, 8 * (b) + (k), db_);                                            \
        warp_sync();                                                                     \
    }

    // spinning / rolling row of point k of foot f: own twist component c, friction coefficient mu
#define TORSION_ROW(k, f, c, mu)                                                                          \
    {                                                                                                     \
        const float x_ = fmaf(-(PLEN_LA_TORSION ? o[2 + (c)] : s[2 + (c)]), t_dinv[c], t_rhs[c]);                                        \
        const float lim_ = (mu) * c_lam[k][5];                                                            \
        /* row skipped while the normal impulse is not positive: its bounds collapse to [0, 0] (the selects sit on the  \
           bounds, which do not depend on the row chain, instead of on the candidate) */                    \
        const bool on_ = c_lam[k][5] > 0.0f;                                                              \
        const float lo_ = on_ ? -lim_ - c_lam[k][c] : 0.0f, hi_ = on_ ? lim_ - c_lam[k][c] : 0.0f;        \
        const float dl_ = fminf(fmaxf(x_, lo_), hi_);                                                     \
        if (PLEN_LA_TORSION) apply_reg6(o, A[f][c], dl_);                                                 \
        const float db_ = GSH(dl_, 2 + (f));                                                              \
        if (g == 2 + (f)) { c_lam[k][c] += dl_; resT[c] = fmaxf(resT[c], fabsf(dl_)); }                   \
        apply_reg(s, A[f][c], db_);                                                                       \
    }

    // one spinning / rolling row inside the owner's private chain (PLEN_TORSION_BLOCK): u = the owner's current value of the
    // row's twist component, sum += the impulse change; leaves the change in dl_own_ for the caller's updates of u
#define TORSION_OWN(k, f, c, mu, u, sum)                                                                  \
        const float x_##c##k = fmaf(-(u), t_dinv[c], t_rhs[c]);                                           \
        const float lim_##c##k = (mu) * c_lam[k][5];                                                      \
        const bool on_##c##k = c_lam[k][5] > 0.0f;                                                        \
        const float lo_##c##k = on_##c##k ? -lim_##c##k - c_lam[k][c] : 0.0f;                             \
        const float hi_##c##k = on_##c##k ? lim_##c##k - c_lam[k][c] : 0.0f;                              \
        dl_own_ = fminf(fmaxf(x_##c##k, lo_##c##k), hi_##c##k);                                           \
        (sum) += dl_own_;                                                                                 \
        if (g == 2 + (f)) { c_lam[k][c] += dl_own_; resT[c] = fmaxf(resT[c], fabsf(dl_own_)); }

    float mu_spin = lc.mu_spin * fric_s, mu_roll = lc.mu_roll * fric_s, mu_lat = lc.mu_lat * fric_s;      // opened at the freeze
    float mu_link = EXT ? cfg.mu_link * fric_s : 0.0f;
    const float res_thr = lc.res_thr;
    const int n_iterations = lc.iterations;
    bool alive = valid;
    int my_iters = 0;
    for (int it = 0; it < n_iterations; it++) {
        res = 0.0f; resF = 0.0f;
        float resT[3] = {0.0f, 0.0f, 0.0f};      // largest |impulse change| of the spinning / rolling rows per twist component
        const bool fwd = (it & 1) != 0;
        // ---- non-contact rows: list = [limits in joint order, servos in joint order]; odd iterations forward, even reversed
        for (int half = 0; half < 2; half++) {
            const bool do_limits = fwd == (half == 0);
            if (do_limits) {
                unsigned msk = lim_any;     // rare: a joint beyond its +-1.7 rad URDF limit on some robot of the warp
                while (msk) {
                    const int j = fwd ? lowest_bit(msk) : highest_bit(msk);
                    msk &= ~(1u << j);
                    const int b = j >> 3;
                    switch (j & 7) {
                        case 0: LIMIT_ROW(b, 0); break;
                        case 1: LIMIT_ROW(b, 1); break;
                        case 2: LIMIT_ROW(b, 2); break;
                        case 3: LIMIT_ROW(b, 3); break;
                        case 4: LIMIT_ROW(b, 4); break;
                        case 5: LIMIT_ROW(b, 5); break;
                        case 6: LIMIT_ROW(b, 6); break;
                        default: LIMIT_ROW(b, 7); break;
                    }
                }
            } else if (fwd) {
                if (PLEN_LA_SERVO) OWN_SYNC();
#pragma unroll
                for (int k = 0; k < 8; k++) SERVO_ROW(0, k);
                if (PLEN_LA_SERVO) OWN_SYNC();
#pragma unroll
                for (int k = 0; k < 8; k++) SERVO_ROW(1, k);
                if (PLEN_LA_SERVO) OWN_SYNC();
#pragma unroll
                for (int k = 0; k < 2; k++) SERVO_ROW(2, k);
            } else {
                if (PLEN_LA_SERVO) OWN_SYNC();
#pragma unroll
                for (int k = 1; k >= 0; k--) SERVO_ROW(2, k);
                if (PLEN_LA_SERVO) OWN_SYNC();
#pragma unroll
                for (int k = 7; k >= 0; k--) SERVO_ROW(1, k);
                if (PLEN_LA_SERVO) OWN_SYNC();
#pragma unroll
                for (int k = 7; k >= 0; k--) SERVO_ROW(0, k);
            }
        }
        if (man_any) {
            // ---- contact normals (lower bound 0; the 1e10 upper bound of the reference never binds)
            FOR_ACTIVE_POINTS {
                if (PLEN_LA_NORMAL && k == 0) OWN_SYNC6();
                const float(&on)[8] = PLEN_LA_NORMAL ? o : s;
                const float r_ = fmaf(on[2], py[p], fmaf(-on[3], px[p], on[7]));
                const float x_ = fmaf(-r_, c_dinv[k][2], c_rhs[k][2]);
                const float dl_ = fmaxf(x_, -c_lam[k][5]);
                const vec4 ca_ = g_ld(Gl, PLEN_SCOL(f, 5) * 8), cb_ = g_ld(Gl, PLEN_SCOL(f, 5) * 8 + 1);
                if (PLEN_LA_NORMAL) {
                    apply_reg6(o, A[f][0], py[p] * dl_);
                    apply_reg6(o, A[f][1], -px[p] * dl_);
                    apply_vec6(o, ca_, cb_, dl_);
                }
                const float db_ = GSH(dl_, 2 + f);
                if (g == 2 + f) { c_lam[k][5] += dl_; resF = fmaxf(resF, fabsf(dl_ * c_d[k][2])); }
                apply_reg(s, A[f][0], py[p] * db_);
                apply_reg(s, A[f][1], -px[p] * db_);
                apply_vec(s, ca_, cb_, db_);
            }
        }
        if (EXT && nxmax) {
            // ---- normals of the box contact points (after the soles', as in the oracle's row list)
            for (int q = 0; q < nxmax; q++) {
                float r_;
                X_DOT(3 * q, r_);
                const float x_ = fmaf(-r_, x_dinv[0], x_rhs[0]);
                const float dl_ = fmaxf(x_, -x_lam[0]);
                const float db_ = GSH(dl_, q);
                if (g == q) { x_lam[0] += dl_; resX = fmaxf(resX, fabsf(dl_ * x_d[0])); }
                apply_vec(s, X_VEC(3 * q, 1, 0), X_VEC(3 * q, 1, 1), db_);
            }
        }
        if (man_any) {
            // ---- all spinning rows, then the rolling rows point by point (t1, t2)
#if PLEN_TORSION_BLOCK
            // The spinning rows of a foot's points are consecutive in the row list and are the SAME functional (the foot's wz);
            // the rolling rows (t1, t2 point by point) are two functionals (wy, wx).  A block of them therefore interacts only
            // through one (two) entries of x and the 1 x 1 (2 x 2) diagonal block of G, both of which the owner lane holds: it
            // runs the block's Gauss-Seidel chain privately (TORSION_OWN: FFMA -> FMNMX -> FMNMX -> FFMA per row, no SHFL) and
            // broadcasts the SUM of the impulse changes per functional -- x += G[:,c] (sum of deltas) is the same update by
            // linearity (re-associated: equal to the row-by-row form up to rounding).  n rows cost one column application
            // instead of n, and the SHFL round trip leaves the chain except once per block.
#pragma unroll
            for (int f = 0; f < 2; f++) {
                if ((f ? nmax1 : nmax0) == 0) continue;
                float u2_ = s[4], sum2_ = 0.0f;
                const float a22_ = A[f][2][4];
                float dl_own_;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    if (k >= (f ? nmax1 : nmax0)) break;
                    TORSION_OWN(k, f, 2, mu_spin, u2_, sum2_);
                    u2_ = fmaf(a22_, dl_own_, u2_);
                }
                apply_reg(s, A[f][2], GSH(sum2_, 2 + f));
            }
#pragma unroll
            for (int f = 0; f < 2; f++) {
                if ((f ? nmax1 : nmax0) == 0) continue;
                float u1_ = s[3], u0_ = s[2], sum1_ = 0.0f, sum0_ = 0.0f;
                // response of (wx, wy) of this foot to unit impulses on wy (column 1) and wx (column 0)
                const float a11_ = A[f][1][3], a01_ = A[f][1][2], a10_ = A[f][0][3], a00_ = A[f][0][2];
                float dl_own_;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    if (k >= (f ? nmax1 : nmax0)) break;
                    TORSION_OWN(k, f, 1, mu_roll, u1_, sum1_);
                    u1_ = fmaf(a11_, dl_own_, u1_); u0_ = fmaf(a01_, dl_own_, u0_);
                    TORSION_OWN(k, f, 0, mu_roll, u0_, sum0_);
                    u0_ = fmaf(a00_, dl_own_, u0_); u1_ = fmaf(a10_, dl_own_, u1_);
                }
                const float d1_ = GSH(sum1_, 2 + f), d0_ = GSH(sum0_, 2 + f);
                apply_reg(s, A[f][1], d1_);
                apply_reg(s, A[f][0], d0_);
            }
#else
            FOR_ACTIVE_POINTS {
                if (PLEN_LA_TORSION && k == 0) OWN_SYNC6();
                TORSION_ROW(k, f, 2, mu_spin);
            }
            FOR_ACTIVE_POINTS {
                if (PLEN_LA_TORSION && k == 0) OWN_SYNC6();
                TORSION_ROW(k, f, 1, mu_roll);
                TORSION_ROW(k, f, 0, mu_roll);
            }
#endif
            // ---- lateral pairs with the implicit friction cone (resolveConeFrictionConstraintRows)
            FOR_ACTIVE_POINTS {
                if (PLEN_LA_LATERAL && k == 0) OWN_SYNC6();
                const float(&ol)[8] = PLEN_LA_LATERAL ? o : s;
                const float rA = fmaf(ol[4], px[p], fmaf(-ol[2], pz[p], ol[6]));     // row A: own = vy
                const float rB = fmaf(ol[3], pz[p], fmaf(-ol[4], py[p], ol[5]));     // row B: own = vx
                const float sumA = c_lam[k][4] + fmaf(-rA, c_dinv[k][1], c_rhs[k][1]);
                const float sumB = c_lam[k][3] + fmaf(-rB, c_dinv[k][0], c_rhs[k][0]);
                const float lim = mu_lat * c_lam[k][5];
                // branch free: the projection is always evaluated and selected by the cone test (inside the cone the
                // operands may be 0 * inf = NaN; fminf / fmaxf return the other operand and the select drops the result)
                const bool outside = fmaxf(fabsf(sumA), fabsf(sumB)) > lim;
                const float inv = rsqrt_fast(fmaf(sumA, sumA, sumB * sumB));
                const float cA = fabsf(lim * sumA * inv), cB = fabsf(lim * sumB * inv);
                const float nA = outside ? clampf(sumA, -cA, cA) : sumA;
                const float nB = outside ? clampf(sumB, -cB, cB) : sumB;
                const float dlA = nA - c_lam[k][4], dlB = nB - c_lam[k][3];
                const vec4 xa_ = g_ld(Gl, PLEN_SCOL(f, 3) * 8), xb_ = g_ld(Gl, PLEN_SCOL(f, 3) * 8 + 1);
                const vec4 ya_ = g_ld(Gl, PLEN_SCOL(f, 4) * 8), yb_ = g_ld(Gl, PLEN_SCOL(f, 4) * 8 + 1);
                if (PLEN_LA_LATERAL) {
                    apply_reg6(o, A[f][0], -pz[p] * dlA);
                    apply_reg6(o, A[f][1], pz[p] * dlB);
                    apply_reg6(o, A[f][2], px[p] * dlA - py[p] * dlB);
                    apply_vec6(o, xa_, xb_, dlB);
                    apply_vec6(o, ya_, yb_, dlA);
                }
                const float dA = GSH(dlA, 2 + f), dB = GSH(dlB, 2 + f);
                if (g == 2 + f) {
                    c_lam[k][4] = nA; c_lam[k][3] = nB;
                    resF = fmaxf(resF, fabsf(dlA * c_d[k][1] + dlB * c_d[k][0]));     // pair residual = dA/dinvA + dB/dinvB
                }
                apply_reg(s, A[f][0], -pz[p] * dA);
                apply_reg(s, A[f][1], pz[p] * dB);
                apply_reg(s, A[f][2], px[p] * dA - py[p] * dB);
                apply_vec(s, xa_, xb_, dB);
                apply_vec(s, ya_, yb_, dA);
            }
        }
        if (EXT && nxmax) {
            // ---- lateral pairs of the box contact points with the implicit cone (A = t1 row, B = t2 row)
            for (int q = 0; q < nxmax; q++) {
                float rA, rB;
                X_DOT(3 * q + 1, rA);
                X_DOT(3 * q + 2, rB);
                const float sumA = x_lam[1] + fmaf(-rA, x_dinv[1], x_rhs[1]);
                const float sumB = x_lam[2] + fmaf(-rB, x_dinv[2], x_rhs[2]);
                const float lim = mu_link * x_lam[0];
                const bool outside = fmaxf(fabsf(sumA), fabsf(sumB)) > lim;
                const float inv = rsqrt_fast(fmaf(sumA, sumA, sumB * sumB));
                const float cA = fabsf(lim * sumA * inv), cB = fabsf(lim * sumB * inv);
                const float nA = outside ? clampf(sumA, -cA, cA) : sumA;
                const float nB = outside ? clampf(sumB, -cB, cB) : sumB;
                const float dlA = nA - x_lam[1], dlB = nB - x_lam[2];
                const float dA = GSH(dlA, q), dB = GSH(dlB, q);
                if (g == q) {
                    x_lam[1] = nA; x_lam[2] = nB;
                    resX = fmaxf(resX, fabsf(dlA * x_d[1] + dlB * x_d[2]));
                }
                apply_vec(s, X_VEC(3 * q + 1, 1, 0), X_VEC(3 * q + 1, 1, 1), dA);
                apply_vec(s, X_VEC(3 * q + 2, 1, 0), X_VEC(3 * q + 2, 1, 1), dB);
            }
        }
        // ---- residual of this iteration, per robot: servo / limit part is already robot-uniform, the contact part
        //      lives in the two foot lanes
        // spinning / rolling rows of a foot share one d per twist component (t_d), so |dl * d| is scaled once here
        resF = fmaxf(resF, fmaxf(resT[0] * fabsf(t_d[0]), fmaxf(resT[1] * fabsf(t_d[1]), resT[2] * fabsf(t_d[2]))));
        float rf = (g >= 2) ? resF : 0.0f;
        if (EXT) { rf = fmaxf(rf, resX); resX = 0.0f; }
#if PLEN_REDUX_CONV && !defined(PLEN_HOST_EMU)
        // max over the robot's four lanes with ONE partitioned redux.sync (non-negative floats order like their bit patterns)
        rf = __uint_as_float(__reduce_max_sync(0xFu << (lane & 28), __float_as_uint(rf)));
#else
        rf = fmaxf(rf, shfl_xor(rf, 1));
        rf = fmaxf(rf, shfl_xor(rf, 2));
#endif
        const float rr = fmaxf(res, rf);
        if (alive) {
            my_iters = it + 1;
            if (rr * rr <= res_thr) {
                alive = false;      // freeze: every later row update of this robot is exactly zero
#pragma unroll
                for (int k = 0; k < 8; k++) {     // park the accumulated servo row value in m_rhs, close the bounds
                    m_rhs[k] = -0.5f * (m_lo[k] + m_hi[k]);
                    m_lo[k] = 0.0f; m_hi[k] = 0.0f;
                }
#pragma unroll
                for (int k = 0; k < 4; k++)
#pragma unroll
                    for (int c = 0; c < 3; c++) { c_rhs[k][c] = 0.0f; c_dinv[k][c] = 0.0f; }
#pragma unroll
                for (int c = 0; c < 3; c++) { t_rhs[c] = 0.0f; t_dinv[c] = 0.0f; }
                // ... and open the friction bounds: an impulse sitting ON its bound can exceed it by an ulp (lam + (lim - lam)
                // != lim in fp32), and the next clamp would then move it again.  With the rows' right-hand sides zeroed and
                // the bounds wide, every later candidate of this robot is exactly its current impulse (dl = 0), so a frozen
                // robot does not depend on how long the other robots of its warp keep iterating.
                mu_spin = mu_roll = mu_lat = 1.0e30f;
                if (EXT) {
                    mu_link = 1.0e30f;
#pragma unroll
                    for (int r = 0; r < 3; r++) { x_rhs[r] = 0.0f; x_dinv[r] = 0.0f; }
                }
            }
        }
        if (!ballot(alive)) break;
    }

    // ---- delta-v: joints = s; base = B z with z = [servo + limit row values (velocity units; B is pre-divided);
    //      foot wrenches about the base origin]
    float z[8];
#pragma unroll
    for (int k = 0; k < 8; k++) z[k] = alive ? -0.5f * (m_lo[k] + m_hi[k]) : m_rhs[k];
    if (lim_any && valid) {
#pragma unroll
        for (int k = 0; k < 8; k++) z[k] += (g < 3) ? gld(srec + SR_LDIR + 8 * g + k) * L_LAM(8 * g + k) : 0.0f;
    }
    if (man_any) {
        // wrench of this lane's foot (lanes 0, 1 hold zero rows): normal (py,-px,0 | vz), lateral A (-pz,0,px | vy),
        // lateral B (0,pz,-py | vx), roll / spin rows on their own component
        float w[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const float qx = (g == 3) ? px[4 + k] : px[k], qy = (g == 3) ? py[4 + k] : py[k], qz = (g == 3) ? pz[4 + k] : pz[k];
            const float lN = c_lam[k][5], lA = c_lam[k][4], lB = c_lam[k][3];
            w[0] += c_lam[k][0] + qy * lN - qz * lA;
            w[1] += c_lam[k][1] - qx * lN + qz * lB;
            w[2] += c_lam[k][2] + qx * lA - qy * lB;
            w[3] += lB; w[4] += lA; w[5] += lN;
        }
        if (g >= 2) {
#pragma unroll
            for (int c = 0; c < 6; c++) z[2 + c] = w[c];
        }
    }
    float dvb[6];
#pragma unroll
    for (int k = 0; k < 6; k++) {
        float acc = 0.0f;
        if (valid) {
            float b[8];
            load8(srec + SR_B + 32 * k + 8 * g, b);
#pragma unroll
            for (int m = 0; m < 8; m++) acc = fmaf(b[m], z[m], acc);
        }
        dvb[k] = acc;
    }
    if (EXT && nxmax) {
        // box rows act on the base through Bb = (M^-1 Jg^T) base entries; lane q adds its point's three rows
        if (valid && g < nx) {
#pragma unroll
            for (int r = 0; r < 3; r++) {
                const float *row = srx + XR_ROWS + (3 * g + r) * XR_ROW_WORDS + XR_BB;
#pragma unroll
                for (int k = 0; k < 6; k++) dvb[k] = fmaf(gld(row + k), x_lam[r], dvb[k]);
            }
        }
    }
#pragma unroll
    for (int m = 1; m < 4; m <<= 1)
#pragma unroll
        for (int k = 0; k < 6; k++) dvb[k] += shfl_xor(dvb[k], m);

    if (valid) {
        // ---- apply delta-v, integrate (semi-implicit; exponential-map quaternion update), write the state record back
        if (g < 3) {
            float vs[8], q8[8];
            load8(srec + SR_VSTAR + 8 * g, vs);
            load8(srec + SR_Q + 8 * g, q8);
#pragma unroll
            for (int k = 0; k < 8; k++) {
                if (g == 2 && k >= 2) continue;
                const float u = clampf(vs[k] + s[k], -cfg.vmax, cfg.vmax);
                state[W_U + 6 + 8 * g + k] = u;
                state[W_Q + 6 + 8 * g + k] = q8[k] + u * cfg.dt;
            }
        } else {
            state[W_MAN] = (float)man;
            state[W_ITERS] = (float)my_iters;
        }
        if (g >= 2) {   // cached normal impulses of this foot's ORIGINAL contact points: slot = number of active points below
            const int f = g - 2;
            const unsigned mf = (man >> (4 * f)) & 15u;
#pragma unroll
            for (int o = 0; o < 4; o++) {
                const int q = popc_(mf & ((1u << o) - 1u));
                float v = c_lam[0][5];
                v = (q == 1) ? c_lam[1][5] : v; v = (q == 2) ? c_lam[2][5] : v; v = (q == 3) ? c_lam[3][5] : v;
                state[W_U + 24 + 4 * f + o] = ((mf >> o) & 1u) ? v : 0.0f;
            }
        }
        if (g == 0) {
            float ub[6];
#pragma unroll
            for (int k = 0; k < 6; k++) { ub[k] = clampf(gld(srec + SR_BASE + k) + dvb[k], -cfg.vmax, cfg.vmax); state[W_U + k] = ub[k]; }

#pragma unroll
            for (int k = 0; k < 3; k++) state[W_POS + k] = gld(srec + SR_BASE + 6 + k) + ub[3 + k] * cfg.dt;
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
            const float qx = gld(srec + SR_BASE + 9), qy = gld(srec + SR_BASE + 10), qz = gld(srec + SR_BASE + 11), qw = gld(srec + SR_BASE + 12);
            const float rx = cw * qx + dx * qw + dy * qz - dz * qy;
            const float ry = cw * qy - dx * qz + dy * qw + dz * qx;
            const float rz = cw * qz + dx * qy - dy * qx + dz * qw;
            const float rw = cw * qw - dx * qx - dy * qy - dz * qz;
            const float nn = rsqrtf(rx * rx + ry * ry + rz * rz + rw * rw);
            state[W_QUAT + 0] = rx * nn; state[W_QUAT + 1] = ry * nn; state[W_QUAT + 2] = rz * nn; state[W_QUAT + 3] = rw * nn;
        }
    }
#undef X_VEC
#undef X_DOT
#undef GSH
#undef OWN_SYNC
#undef OWN_SYNC6
#undef SERVO_ROW
#undef LIMIT_ROW
#undef L_LAM
#undef TORSION_ROW
#undef TORSION_OWN
#undef NORMAL_COLUMN
}

}  // namespace plen
