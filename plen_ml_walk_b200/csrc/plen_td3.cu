// plen_td3.cu -- device-resident pieces of the reference's TD3 loop (plen_ros/src/plen_ros_helpers/td3.py), part of
// libplen_b200.so (C ABI in include/plen_b200.h).  sm_100a only, no CPU path.
//
//   replay ring   ReplayBuffer.add / .sample (td3.py:122-193): append until full, then overwrite from index 0 upwards;
//                 uniform sampling WITH replacement.  One transition = 72 floats [s 26 | a 18 | s' 26 | r | done].
//   actor forward Actor.forward (td3.py:45-57) for N observations at once: tanh(W3 relu(W2 relu(W1 s))) * max_action,
//                 plus the exploration noise + clip of plen_td3.py:101-104 when asked.  fp32 FMA (the checkpoints are
//                 fp32 and the parity bar is 1e-5 on actions): one CTA owns 32 observations, activations stay in shared
//                 memory across the three layers, W is streamed through shared memory in 32-wide K tiles.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>

#include "../../include/plen_b200.h"

#define TR_WORDS 72      // PLEN_OBS + PLEN_NJ + PLEN_OBS + 2
static_assert(TR_WORDS == PLEN_OBS + PLEN_NJ + PLEN_OBS + 2, "transition layout");

struct plen_replay {
    int device;
    long long capacity, size, ptr;     // host mirror of the ring (adds are stream-ordered; the ring arithmetic is host side)
    float *d_store;                    // [capacity][72]
    char err[256];
};

static char g_td3_err[256] = "";
extern "C" const char *plen_td3_last_error(void) { return g_td3_err; }
static int td3_fail(int code, const char *msg, const char *detail = "") {
    snprintf(g_td3_err, sizeof g_td3_err, "%s%s", msg, detail);
    return code;
}
extern "C" int plen_td3_set_error(int code, const char *msg, const char *detail) { return td3_fail(code, msg, detail); }
#define TCK(call)                                                                          \
    do {                                                                                   \
        cudaError_t e_ = (call);                                                           \
        if (e_ != cudaSuccess) return td3_fail(PLEN_E_CUDA, #call ": ", cudaGetErrorString(e_)); \
    } while (0)

// ---------------------------------------------------------------------------------------------- replay ring
// rows [row0, row0 + n) of the ring (mod capacity) <- n transitions starting at batch offset off
__global__ void k_replay_add(float *__restrict__ store, long long capacity, long long row0, int off, int n,
                             const float *__restrict__ s, const float *__restrict__ a, const float *__restrict__ s2,
                             const float *__restrict__ r, const uint8_t *__restrict__ done) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)n * TR_WORDS) return;
    const int e = (int)(i / TR_WORDS), w = (int)(i - (long long)e * TR_WORDS);
    const int src = off + e;
    float v;
    if (w < 26) v = s[(size_t)src * 26 + w];
    else if (w < 44) v = a[(size_t)src * 18 + (w - 26)];
    else if (w < 70) v = s2[(size_t)src * 26 + (w - 44)];
    else if (w == 70) v = r[src];
    else v = done[src] ? 1.0f : 0.0f;
    long long row = row0 + e;
    if (row >= capacity) row -= capacity;
    store[row * TR_WORDS + w] = v;
}

__device__ __forceinline__ uint32_t hash32(uint64_t x) {     // splitmix64 finaliser
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return (uint32_t)((x ^ (x >> 31)) >> 32);
}

// batch rows drawn uniformly with replacement (td3.py:175 np.random.randint(0, len(storage), size=batch_size));
// one warp per sampled transition
__global__ void k_replay_sample(const float *__restrict__ store, long long size, int batch, uint64_t seed,
                                float *__restrict__ s, float *__restrict__ a, float *__restrict__ s2,
                                float *__restrict__ r, float *__restrict__ not_done, int *__restrict__ idx_out) {
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (b >= batch) return;
    const uint32_t u = hash32(seed * 0x100000001B3ull + (uint64_t)b);
    const long long row = (long long)(((uint64_t)u * (uint64_t)size) >> 32);
    const float *t = store + row * TR_WORDS;
    for (int w = lane; w < TR_WORDS; w += 32) {
        const float v = t[w];
        if (w < 26) s[(size_t)b * 26 + w] = v;
        else if (w < 44) a[(size_t)b * 18 + (w - 26)] = v;
        else if (w < 70) s2[(size_t)b * 26 + (w - 44)] = v;
        else if (w == 70) r[b] = v;
        else not_done[b] = 1.0f - v;
    }
    if (idx_out && lane == 0) idx_out[b] = (int)row;
}

// ---------------------------------------------------------------------------------------------- actor forward
#define ACT_TM 32        // observations per CTA
#define ACT_TK 32        // K tile of the streamed weights
#define ACT_H 256        // hidden width (td3.py:37-41)
#define ACT_THREADS 256

// out[TM][N] = act(in[TM][K] W^T + b),  W row-major [N][K] (nn.Linear layout).  256 threads: warp wy owns rows 4 wy..4 wy+3,
// lane tx owns columns tx + 32 j.  in/out row stride = ACT_H + 1 (conflict-free column access is not needed: rows are
// read as warp broadcasts).
template <int K, int N, int ACT>   // ACT 0: relu, 1: identity
__device__ __forceinline__ void dense_layer(const float *__restrict__ W, const float *__restrict__ bias, const float *in,
                                            float *out, float *Ws) {
    constexpr int NJ = (N + 31) / 32;          // columns per thread
    const int tx = threadIdx.x & 31, wy = threadIdx.x >> 5;
    float acc[4][NJ];
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
        for (int j = 0; j < NJ; j++) acc[r][j] = 0.0f;
    for (int k0 = 0; k0 < K; k0 += ACT_TK) {
        __syncthreads();
        // stage W[:, k0 .. k0+TK) transposed: Ws[kk][n]  (stride N_pad = 32 NJ + 1)
        for (int e = threadIdx.x; e < 32 * NJ * ACT_TK; e += ACT_THREADS) {
            const int n = e / ACT_TK, kk = e - n * ACT_TK;
            Ws[kk * (32 * NJ + 1) + n] = (n < N && k0 + kk < K) ? W[(size_t)n * K + k0 + kk] : 0.0f;
        }
        __syncthreads();
#pragma unroll 8
        for (int kk = 0; kk < ACT_TK; kk++) {
            float a[4], w[NJ];
#pragma unroll
            for (int r = 0; r < 4; r++) a[r] = in[(4 * wy + r) * (ACT_H + 1) + k0 + kk];
#pragma unroll
            for (int j = 0; j < NJ; j++) w[j] = Ws[kk * (32 * NJ + 1) + tx + 32 * j];
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int j = 0; j < NJ; j++) acc[r][j] = fmaf(a[r], w[j], acc[r][j]);
        }
    }
#pragma unroll
    for (int j = 0; j < NJ; j++) {
        const int n = tx + 32 * j;
        const float b = (n < N) ? bias[n] : 0.0f;
#pragma unroll
        for (int r = 0; r < 4; r++) {
            float v = acc[r][j] + b;
            if (ACT == 0) v = fmaxf(v, 0.0f);
            if (n < N) out[(4 * wy + r) * (ACT_H + 1) + n] = v;
        }
    }
}

__global__ void __launch_bounds__(ACT_THREADS)
k_actor_forward(const float *__restrict__ w1, const float *__restrict__ b1, const float *__restrict__ w2,
                const float *__restrict__ b2, const float *__restrict__ w3, const float *__restrict__ b3,
                const float *__restrict__ obs, int n, float max_action, float noise_std, uint64_t seed,
                float *__restrict__ act) {
    extern __shared__ float sm[];
    float *A = sm;                                   // [TM][H+1]
    float *B = A + ACT_TM * (ACT_H + 1);             // [TM][H+1]
    float *Ws = B + ACT_TM * (ACT_H + 1);            // [TK][257]
    const int row0 = blockIdx.x * ACT_TM;
    for (int e = threadIdx.x; e < ACT_TM * 32; e += ACT_THREADS) {       // K of layer 1 padded 26 -> 32 with zeros
        const int r = e >> 5, k = e & 31;
        A[r * (ACT_H + 1) + k] = (row0 + r < n && k < PLEN_OBS) ? obs[(size_t)(row0 + r) * PLEN_OBS + k] : 0.0f;
    }
    dense_layer<PLEN_OBS, ACT_H, 0>(w1, b1, A, B, Ws);     // fc1 + relu   (td3.py:52)
    dense_layer<ACT_H, ACT_H, 0>(w2, b2, B, A, Ws);        // fc2 + relu   (td3.py:54)
    dense_layer<ACT_H, PLEN_NJ, 1>(w3, b3, A, B, Ws);      // fc3          (td3.py:56)
    __syncthreads();
    for (int e = threadIdx.x; e < ACT_TM * PLEN_NJ; e += ACT_THREADS) {
        const int r = e / PLEN_NJ, j = e - r * PLEN_NJ;
        if (row0 + r >= n) continue;
        float v = max_action * tanhf(B[r * (ACT_H + 1) + j]);            // td3.py:56
        if (noise_std > 0.0f) {
            // plen_td3.py:101-104: clip(action + N(0, max_action * expl_noise), -max_action, max_action); Box-Muller on a
            // counter-based hash of (seed, env, joint)
            const uint64_t c = seed * 0x100000001B3ull + (uint64_t)(row0 + r) * 32u + (uint64_t)j;
            const float u1 = (hash32(c) + 1.0f) * 2.3283064e-10f, u2 = hash32(c ^ 0xA5A5A5A5DEADBEEFull) * 2.3283064e-10f;
            v += noise_std * sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
            v = fminf(fmaxf(v, -max_action), max_action);
        }
        act[(size_t)(row0 + r) * PLEN_NJ + j] = v;
    }
}

static const size_t ACT_SMEM = sizeof(float) * (2 * ACT_TM * (ACT_H + 1) + ACT_TK * (ACT_H + 1));

// ---------------------------------------------------------------------------------------------- C ABI
extern "C" {

plen_replay *plen_replay_create(long long capacity, int device) {
    if (capacity <= 0) { td3_fail(PLEN_E_ARG, "plen_replay_create: capacity <= 0"); return nullptr; }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || device < 0 || device >= ndev) {
        td3_fail(PLEN_E_CUDA, "plen_replay_create: no usable CUDA device; this library has no CPU fallback");
        return nullptr;
    }
    plen_replay *rb = new (std::nothrow) plen_replay();
    if (!rb) return nullptr;
    memset(rb, 0, sizeof *rb);
    rb->device = device; rb->capacity = capacity;
    if (cudaSetDevice(device) != cudaSuccess ||
        cudaMalloc(&rb->d_store, sizeof(float) * TR_WORDS * (size_t)capacity) != cudaSuccess) {
        td3_fail(PLEN_E_CUDA, "plen_replay_create: cudaMalloc failed");
        delete rb;
        return nullptr;
    }
    return rb;
}

void plen_replay_destroy(plen_replay *rb) {
    if (!rb) return;
    cudaSetDevice(rb->device);
    cudaFree(rb->d_store);
    delete rb;
}

long long plen_replay_size(const plen_replay *rb) { return rb ? rb->size : 0; }
long long plen_replay_ptr(const plen_replay *rb) { return rb ? rb->ptr : 0; }
const float *plen_replay_storage(const plen_replay *rb) { return rb ? rb->d_store : nullptr; }

int plen_replay_add(plen_replay *rb, const float *state_dev, const float *action_dev, const float *next_state_dev,
                    const float *reward_dev, const uint8_t *done_dev, int n, void *stream) {
    if (!rb || n < 0) return td3_fail(PLEN_E_ARG, "plen_replay_add: bad arguments");
    if (n == 0) return PLEN_OK;        // an empty batch is a no-op (its buffers may be NULL)
    if (!state_dev || !action_dev || !next_state_dev || !reward_dev || !done_dev)
        return td3_fail(PLEN_E_ARG, "plen_replay_add: NULL buffer");
    TCK(cudaSetDevice(rb->device));
    cudaStream_t st = (cudaStream_t)stream;
    int off = 0;
    while (off < n) {
        // ReplayBuffer.add (td3.py:136-147), n tuples at a time: append while not full, else overwrite at ptr and advance
        long long row0, cnt;
        if (rb->size < rb->capacity) { row0 = rb->size; cnt = rb->capacity - rb->size; }
        else { row0 = rb->ptr; cnt = rb->capacity; }
        if (cnt > n - off) cnt = n - off;
        const long long words = cnt * TR_WORDS;
        k_replay_add<<<(unsigned)((words + 255) / 256), 256, 0, st>>>(rb->d_store, rb->capacity, row0, off, (int)cnt, state_dev,
                                                                      action_dev, next_state_dev, reward_dev, done_dev);
        if (rb->size < rb->capacity) rb->size += cnt; else rb->ptr = (rb->ptr + cnt) % rb->capacity;
        off += (int)cnt;
    }
    TCK(cudaGetLastError());
    return PLEN_OK;
}

int plen_replay_sample(plen_replay *rb, int batch, unsigned long long seed, float *state_dev, float *action_dev,
                       float *next_state_dev, float *reward_dev, float *not_done_dev, int *index_dev, void *stream) {
    if (!rb || batch <= 0 || !state_dev || !action_dev || !next_state_dev || !reward_dev || !not_done_dev)
        return td3_fail(PLEN_E_ARG, "plen_replay_sample: bad arguments");
    if (rb->size <= 0) return td3_fail(PLEN_E_STATE, "plen_replay_sample: the buffer is empty");
    TCK(cudaSetDevice(rb->device));
    k_replay_sample<<<(batch * 32 + 127) / 128, 128, 0, (cudaStream_t)stream>>>(rb->d_store, rb->size, batch, seed, state_dev,
                                                                              action_dev, next_state_dev, reward_dev,
                                                                              not_done_dev, index_dev);
    TCK(cudaGetLastError());
    return PLEN_OK;
}

int plen_actor_forward(int device, const float *w1, const float *b1, const float *w2, const float *b2, const float *w3,
                       const float *b3, const float *obs_dev, int n, float max_action, float noise_std,
                       unsigned long long seed, float *action_dev, void *stream) {
    if (!w1 || !b1 || !w2 || !b2 || !w3 || !b3 || !obs_dev || !action_dev || n <= 0)
        return td3_fail(PLEN_E_ARG, "plen_actor_forward: bad arguments");
    TCK(cudaSetDevice(device));
    static bool attr_set[64] = {false};
    if (device < 64 && !attr_set[device]) {
        TCK(cudaFuncSetAttribute(k_actor_forward, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ACT_SMEM));
        attr_set[device] = true;
    }
    k_actor_forward<<<(n + ACT_TM - 1) / ACT_TM, ACT_THREADS, ACT_SMEM, (cudaStream_t)stream>>>(
        w1, b1, w2, b2, w3, b3, obs_dev, n, max_action, noise_std, seed, action_dev);
    TCK(cudaGetLastError());
    return PLEN_OK;
}

}  // extern "C"
