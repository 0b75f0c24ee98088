// TEST-ONLY warp emulation: compiles the product's device source (plen_ml_walk_b200/csrc/plen_device.cuh,
// plen_env.cuh) for the host, with 32 threads + barriers standing in for one warp's shuffles.  Lets the CPU test
// suite run the real kernel code against the float64 oracle without a GPU.  Never linked into the product library.
#define PLEN_HOST_EMU 1
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/plen_b200.h"   // C ABI structs stay float32 in both builds

typedef float float32_t_;      // a real float even in the float64 build below
#ifdef PLEN_EMU_DOUBLE
// Algorithm check: the same device source evaluated in float64 (every `float` below this line becomes double and
// the f-suffixed libm calls are redirected), so that any disagreement with the oracle that survives is a logic
// difference, not fp32 rounding.
#define float double
#define sinf sin
#define cosf cos
#define sqrtf sqrt
#define expf exp
#define tanhf tanh
#define atan2f atan2
#define asinf asin
#define fabsf fabs
#define fminf fmin
#define fmaxf fmax
#define fmaf fma
#endif

#define PLEN_DEV static inline
#define PLEN_DEV_NOINLINE static


namespace plen {
static thread_local int t_lane = 0;
static float g_x[32];
static unsigned g_u[32];
static std::barrier<> *g_bar = nullptr;
static inline void bar() { g_bar->arrive_and_wait(); }
static inline int lane_id() { return t_lane; }
static inline float xchg(float v, int src) { g_x[t_lane] = v; bar(); float r = g_x[src & 31]; bar(); return r; }
static inline float shfl(float v, int src) { return xchg(v, src); }
static inline float shfl4(float v, int src) { return xchg(v, (t_lane & 28) | (src & 3)); }
static inline float shfl_up(float v, int d) { return xchg(v, t_lane - d >= 0 ? t_lane - d : t_lane); }
static inline float shfl_down(float v, int d) { return xchg(v, t_lane + d <= 31 ? t_lane + d : t_lane); }
static inline float shfl_xor(float v, int m) { return xchg(v, t_lane ^ m); }
static inline unsigned ballot(bool p) {
    g_u[t_lane] = p ? 1u : 0u; bar();
    unsigned r = 0; for (int i = 0; i < 32; i++) r |= g_u[i] << i;
    bar(); return r;
}
static inline unsigned redux_max(unsigned v) {
    g_u[t_lane] = v; bar();
    unsigned r = 0; for (int i = 0; i < 32; i++) r = g_u[i] > r ? g_u[i] : r;
    bar(); return r;
}
static inline unsigned redux_or(unsigned v) {
    g_u[t_lane] = v; bar();
    unsigned r = 0; for (int i = 0; i < 32; i++) r |= g_u[i];
    bar(); return r;
}
static inline void warp_sync() { bar(); }
static inline float f_as_u_max(float v) {   // max over lanes of a non-negative value
    g_x[t_lane] = v; bar();
    float r = 0; for (int i = 0; i < 32; i++) r = g_x[i] > r ? g_x[i] : r;
    bar(); return r;
}
static inline void sincos_(float x, float *s, float *c) { *s = sinf(x); *c = cosf(x); }
static inline float rcp_(float x) { return 1.0f / x; }
static inline unsigned f_bits(float v) { const float32_t_ f = (float32_t_)v; unsigned u; std::memcpy(&u, &f, 4); return u; }
static inline int lowest_bit(unsigned m) { return __builtin_ffs((int)m) - 1; }
static inline int highest_bit(unsigned m) { return 31 - __builtin_clz(m); }
static inline int popc_(unsigned m) { return __builtin_popcount(m); }
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
}  // namespace plen
using plen::rsqrtf;

#include "../../plen_ml_walk_b200/csrc/plen_host_tables.h"
#include "../../plen_ml_walk_b200/csrc/plen_solve.cuh"

using namespace plen;

namespace {
template <class F> void run_warp(F body) {
    std::barrier<> b(32);
    g_bar = &b;
    std::vector<std::thread> th;
    for (int l = 0; l < 32; l++) th.emplace_back([=] { t_lane = l; body(l); });
    for (auto &t : th) t.join();
    g_bar = nullptr;
}
}  // namespace

// persistent sole manifolds of the robots being ticked ([n][PLEN_MAN_WORDS], set by emu_set_manifold; nullptr = option off)
static float *g_man = nullptr;

extern "C" {

int emu_default_config(plen_config *c, int joint_act) { return default_config(c, joint_act); }
int emu_real_bytes() { return (int)sizeof(float); }
int emu_sizeof_config() { return (int)sizeof(plen_config); }
int emu_sizeof_model() { return (int)sizeof(plen_model); }
void emu_init_record(const plen_config *c, float *rec) { init_record(c, rec); }
void emu_set_manifold(float *man) { g_man = man; }

// (k_dyn, k_solve) x n_ticks over n robots, exactly as plen_b200.cu launches them: one emulated warp per robot for the
// dynamics, then one emulated warp per 8 robots for the solver (identity permutation: grouping never changes a result).  tgt [n][18] nullable (zero targets).
static void run_ticks(const DevConfig &dc, const float *tab, float *records, const float *tgt, int n, int n_ticks,
                      const DebugOut *dbg0, int lane) {
    static WarpScratch ws;
    static std::vector<float> srec, srx, Gs(PLEN_SOLVE_ROBOTS * PLEN_GS_WORDS);
    static std::vector<uint8_t> keys;
    if (lane == 0) { srec.assign((size_t)n * SR_WORDS, 0.0f); srx.assign((size_t)n * XR_WORDS, 0.0f); keys.assign(n, 0); }
    bar();
    for (int t = 0; t < n_ticks; t++) {
        for (int e = 0; e < n; e++) {
            LaneState L;
            load_record(records + 96 * e, ws, L, lane);
            if (lane >= 6 && lane < 24) L.tgt = tgt ? tgt[18 * e + lane - 6] : 0.0f;
            tick_dynamics(dc, tab, ws, L, lane, srec.data() + (size_t)e * SR_WORDS, keys.data() + e,
                          (e == 0 && t == n_ticks - 1) ? dbg0 : nullptr, srx.data() + (size_t)e * XR_WORDS,
                          g_man ? g_man + (size_t)e * PLEN_MAN_WORDS : nullptr);
        }
        for (int b = 0; b < n; b += PLEN_SOLVE_ROBOTS) {
            const int r = b + (lane >> 2);
            const bool valid = r < n;
            const int rr = valid ? r : 0;
            // groups are run through the instance k_rank would route them to: EXT iff one of the eight robots has box contacts
            bool anyx = false;
            for (int k = b; k < b + PLEN_SOLVE_ROBOTS && k < n; k++) anyx = anyx || keys[k] >= PLEN_KEY_EXT;
            if (anyx) {
                const int nx = (valid && keys[rr] >= PLEN_KEY_EXT) ? (int)srx[(size_t)rr * XR_WORDS + XR_NX] : 0;
                solve_tick<1>(dc, srec.data() + (size_t)rr * SR_WORDS, Gs.data() + (lane >> 2) * PLEN_GS_WORDS, records + 96 * rr, lane, valid,
                                 srx.data() + (size_t)rr * XR_WORDS, nx);
            } else {
                solve_tick<0>(dc, srec.data() + (size_t)rr * SR_WORDS, Gs.data() + (lane >> 2) * PLEN_GS_WORDS, records + 96 * rr, lane, valid);
            }
            bar();
        }
    }
}

// n_ticks physics ticks with raw targets; records [n][96], targets [n][18]; dbg_* nullable (first env, last tick)
void emu_tick(const plen_model *m, const plen_config *c, float *records, const float *targets, int n, int n_ticks,
              float *dbg_minv, float *dbg_pos, float *dbg_rot, int *iters_out) {
    std::vector<float> tab(T_ROWS * 32);
    DevConfig dc; EnvRanges er;
    build_table(m, c, tab.data());
    build_devconfig(m, c, &dc, &er);
    std::vector<float> hull(2 * PLEN_MAX_HULL * 3);      // (the model struct holds real floats even in the float64 build)
    for (size_t i = 0; i < hull.size(); i++) hull[i] = (&m->foot_hull[0][0][0])[i];
    dc.hull = hull.data();
    DebugOut dbg{dbg_minv, dbg_pos, dbg_rot};
    run_warp([&](int lane) { run_ticks(dc, tab.data(), records, targets, n, n_ticks, dbg_minv ? &dbg : nullptr, lane); });
    if (iters_out) for (int e = 0; e < n; e++) iters_out[e] = (int)records[96 * e + W_ITERS];
}

void emu_step(const plen_model *m, const plen_config *c, float *records, const float *actions, int n, float *obs,
              float *reward, uint8_t *done, uint8_t *timeout, float *terminal_obs, const float *snapshot) {
    std::vector<float> tab(T_ROWS * 32), tgt((size_t)n * 18);
    DevConfig dc; EnvRanges er;
    build_table(m, c, tab.data());
    build_devconfig(m, c, &dc, &er);
    std::vector<float> hull(2 * PLEN_MAX_HULL * 3);      // (the model struct holds real floats even in the float64 build)
    for (size_t i = 0; i < hull.size(); i++) hull[i] = (&m->foot_hull[0][0][0])[i];
    dc.hull = hull.data();
    for (int e = 0; e < n; e++)
        for (int j = 0; j < 18; j++) tgt[18 * e + j] = agent_target(dc, er, j, actions[18 * e + j]);
    static WarpScratch ws;
    run_warp([&](int lane) {
        run_ticks(dc, tab.data(), records, tgt.data(), n, dc.substeps, nullptr, lane);
        for (int e = 0; e < n; e++) {
            LaneState L;
            load_record(records + 96 * e, ws, L, lane);
            StepIO io{nullptr, obs + 26 * e, reward + e, done + e, timeout ? timeout + e : nullptr,
                      terminal_obs ? terminal_obs + 26 * e : nullptr, snapshot, nullptr};
            env_post(dc, tab.data(), ws, L, lane, io);
            store_record(records + 96 * e, ws, L, lane);
        }
    });
}

void emu_observe(const plen_model *m, const plen_config *c, const float *records, int n, float *obs) {
    (void)m; (void)c;
    static WarpScratch ws;
    run_warp([&](int lane) {
        for (int e = 0; e < n; e++) {
            LaneState L;
            load_record(records + 96 * e, ws, L, lane);
            observe(ws, L, lane);
            if (lane < 26) obs[26 * e + lane] = ws.obs[lane];
            warp_sync();
        }
    });
}
}
