// plen_td3_learn.cu -- TD3Agent.train (plen_ros/src/plen_ros_helpers/td3.py:259-356) as hand-written CUDA, part of
// libplen_b200.so (C ABI in include/plen_b200.h).  sm_100a only, no CPU path, no cuBLAS / autograd.
//
// One update = sample -> target Q (actor_target, critic_target, clipped noise) -> critic forward / MSE / backward -> Adam
// -> every policy_freq-th call: actor forward, Q1 forward, backward through Q1 into the actor, Adam, Polyak targets.
// The networks are the reference's (Actor 26-256-256-18, Critic 2 x 44-256-256-1, td3.py:19-117) with every parameter of
// a network in ONE flat float32 vector in state_dict order (fc1.weight, fc1.bias, ...), so an optimiser step or a target
// update is one launch and a data-parallel learner all-reduces one buffer.
//
// At the reference's batch of 100 every matrix product is tiny (100 x 256 x 256): the step is bound by launch and
// dependency latency, not by FLOPs, so the products run on the FP32 CUDA cores in one strided GEMM kernel with fused
// epilogues (bias / ReLU / tanh / ReLU-mask / tanh-gradient / bias-gradient) -- 18 kernels for a critic-only update, 36
// with the policy update, captured once as a CUDA graph by plen_td3_train and replayed with one launch per update on
// the caller's stream (no host sync, no allocation).  fp32 keeps the
// update within 1e-5 of the reference's torch arithmetic (tests/test_td3_gpu.py).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>

#include "../../include/plen_b200.h"

extern "C" int plen_td3_set_error(int code, const char *msg, const char *detail);      // plen_td3.cu

#define LCK(call)                                                                                        \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) return plen_td3_set_error(PLEN_E_CUDA, #call ": ", cudaGetErrorString(e_)); \
    } while (0)

namespace {

constexpr int S = PLEN_OBS, A = PLEN_NJ, SA = PLEN_OBS + PLEN_NJ, H = 256;
// flat layouts (floats), state_dict order
constexpr int AW1 = 0, AB1 = AW1 + H * S, AW2 = AB1 + H, AB2 = AW2 + H * H, AW3 = AB2 + H, AB3 = AW3 + A * H, ACTOR_N = AB3 + A;
constexpr int CW1 = 0, CB1 = CW1 + H * SA, CW2 = CB1 + H, CB2 = CW2 + H * H, CW3 = CB2 + H, CB3 = CW3 + H, TOWER_N = CB3 + 1;
constexpr int CRITIC_N = 2 * TOWER_N;
static_assert(ACTOR_N == PLEN_TD3_ACTOR_PARAMS && CRITIC_N == PLEN_TD3_CRITIC_PARAMS, "flat parameter layouts");

enum { EPI_NONE = 0, EPI_BIAS, EPI_BIAS_RELU, EPI_BIAS_TANH, EPI_BIAS_TANH_NOISE, EPI_RELUMASK, EPI_TANHGRAD };

// Per-call scalars of plen_td3_train live in device memory, so that the captured CUDA graph of an update never changes:
// the draw seed, the current size of the replay ring, and the bias-corrected Adam step sizes of the two optimisers.
struct CallParams {
    unsigned long long seed;
    long long size;
    float adam_c[2], adam_a[2];      // {lr / (1 - b1^t), 1 / sqrt(1 - b2^t)} of the critic / actor step
};

struct Gemm {
    // C[z][m][n] = epi( sum_k A(z, m, k) B(z, k, n) ),  A(z,m,k) = a[z az + m am + k ak],  B(z,k,n) = b[z bz + k bk + n bn]
    const float *a, *b;
    float *c;
    long long az, am, ak, bz, bk, bn, cz, ldc;
    int M, N, K, nz;
    int epi;
    const float *bias; long long bias_z;          // bias[z][n]
    const float *aux; long long aux_z, ld_aux;     // EPI_RELUMASK: activation; EPI_TANHGRAD: max_action*tanh; EPI_BIAS_TANH_NOISE: noise or NULL
    float p0, p1, p2;                              // max_action, policy_noise, noise_clip
    unsigned long long seed;
    float *rowsum; long long rowsum_z;             // nullable: rowsum[z][m] = sum_k A(z, m, k)   (bias gradient of a dW product)
    const CallParams *cp;                          // nullable: seed read from device memory (graph replay) instead of `seed`
    int splits, k_chunk;                           // split-K (dW products of large minibatches): blockIdx.z = z + nz * split; results are
                                                   // accumulated with atomicAdd into a zeroed C / rowsum
};

__device__ __forceinline__ uint32_t mix32(uint64_t x) {     // splitmix64 finaliser
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return (uint32_t)((x ^ (x >> 31)) >> 32);
}

constexpr int GT = 256;

// BM x BN output tile per CTA, 256 threads as 16 x 16, TM x TN outputs per thread (register tile), K in steps of BK with
// the next tiles prefetched into registers while the current ones are multiplied.  Strides are arbitrary (row- or
// column-major operands, strided sub-matrices such as the action columns of fc1); loads are coalesced along whichever
// operand stride is 1.  Two instances: 32 x 32 x 32 (2 x 2 per thread) for the reference's batch of 100, where the grid
// must be as wide as possible, and 64 x 64 x 16 (4 x 4 per thread) once the batch fills the device.
template <int BM, int BN, int BK, int TM, int TN>
__global__ void __launch_bounds__(GT) k_gemm(const Gemm g) {
    static_assert(BM == 16 * TM && BN == 16 * TN && BM * BK == 4 * GT && BN * BK == 4 * GT, "tile shape");
    __shared__ float As[BK][BM + 1], Bs[BK][BN + 1];
    const int z = blockIdx.z % g.nz, split = blockIdx.z / g.nz, m0 = blockIdx.y * BM, n0 = blockIdx.x * BN, tid = threadIdx.x;
    const int k_lo = split * g.k_chunk, k_hi = min(g.K, k_lo + g.k_chunk);
    const float *a = g.a + z * g.az, *b = g.b + z * g.bz;
    const bool a_kfast = g.ak == 1, b_nfast = g.bn == 1;
    float ra[4], rb[4];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int e = tid + GT * i;
            const int am = a_kfast ? (e / BK) : (e % BM), ak = a_kfast ? (e % BK) : (e / BM);
            const int bk = b_nfast ? (e / BN) : (e % BK), bn = b_nfast ? (e % BN) : (e / BK);
            ra[i] = (m0 + am < g.M && k0 + ak < k_hi) ? a[(m0 + am) * g.am + (k0 + ak) * g.ak] : 0.0f;
            rb[i] = (k0 + bk < k_hi && n0 + bn < g.N) ? b[(k0 + bk) * g.bk + (n0 + bn) * g.bn] : 0.0f;
        }
    };
    const int tx = tid & 15, ty = tid >> 4;
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; i++)
#pragma unroll
        for (int j = 0; j < TN; j++) acc[i][j] = 0.0f;
    float rs = 0.0f;                                   // row sum of A for row m0 + tid (threads < BM, first column tile only)
    const bool do_rs = g.rowsum != nullptr && blockIdx.x == 0;
    fetch(k_lo);
    for (int k0 = k_lo; k0 < k_hi; k0 += BK) {
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int e = tid + GT * i;
            const int am = a_kfast ? (e / BK) : (e % BM), ak = a_kfast ? (e % BK) : (e / BM);
            const int bk = b_nfast ? (e / BN) : (e % BK), bn = b_nfast ? (e % BN) : (e / BK);
            As[ak][am] = ra[i];
            Bs[bk][bn] = rb[i];
        }
        __syncthreads();
        if (k0 + BK < k_hi) fetch(k0 + BK);
#pragma unroll
        for (int k = 0; k < BK; k++) {
            float av[TM], bv[TN];
#pragma unroll
            for (int i = 0; i < TM; i++) av[i] = As[k][TM * ty + i];
#pragma unroll
            for (int j = 0; j < TN; j++) bv[j] = Bs[k][tx + 16 * j];
#pragma unroll
            for (int i = 0; i < TM; i++)
#pragma unroll
                for (int j = 0; j < TN; j++) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (do_rs && tid < BM) {
#pragma unroll
            for (int k = 0; k < BK; k++) rs += As[k][tid];
        }
    }
    if (do_rs && tid < BM && m0 + tid < g.M) {
        if (g.splits > 1) atomicAdd(&g.rowsum[z * g.rowsum_z + m0 + tid], rs);
        else g.rowsum[z * g.rowsum_z + m0 + tid] = rs;
    }
    float *c = g.c + z * g.cz;
#pragma unroll
    for (int i = 0; i < TM; i++)
#pragma unroll
        for (int j = 0; j < TN; j++) {
            const int m = m0 + TM * ty + i, n = n0 + tx + 16 * j;
            if (m >= g.M || n >= g.N) continue;
            float v = acc[i][j];
            switch (g.epi) {
                case EPI_BIAS: v += g.bias[z * g.bias_z + n]; break;
                case EPI_BIAS_RELU: v = fmaxf(v + g.bias[z * g.bias_z + n], 0.0f); break;
                case EPI_BIAS_TANH: v = g.p0 * tanhf(v + g.bias[z * g.bias_z + n]); break;
                case EPI_BIAS_TANH_NOISE: {
                    // td3.py:303-308: noise = clamp(randn * policy_noise, +-noise_clip); clamp(actor_target(s') + noise, +-max_action)
                    float nz;
                    if (g.aux) nz = g.aux[z * g.aux_z + m * g.ld_aux + n];
                    else {
                        const uint64_t sd = g.cp ? g.cp->seed : g.seed;
                        const uint64_t ctr = sd * 0x100000001B3ull + (uint64_t)m * 64u + (uint64_t)n;
                        const float u1 = (mix32(ctr) + 1.0f) * 2.3283064e-10f, u2 = mix32(ctr ^ 0x5DEECE66D1234567ull) * 2.3283064e-10f;
                        nz = sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
                    }
                    nz = fminf(fmaxf(nz * g.p1, -g.p2), g.p2);
                    v = g.p0 * tanhf(v + g.bias[z * g.bias_z + n]) + nz;
                    v = fminf(fmaxf(v, -g.p0), g.p0);
                    break;
                }
                case EPI_RELUMASK: v = (g.aux[z * g.aux_z + m * g.ld_aux + n] > 0.0f) ? v : 0.0f; break;
                case EPI_TANHGRAD: {
                    const float t = g.aux[z * g.aux_z + m * g.ld_aux + n] / g.p0;
                    v = v * g.p0 * (1.0f - t * t);
                    break;
                }
                default: break;
            }
            if (g.splits > 1) atomicAdd(&c[m * g.ldc + n], v);      // epilogue is EPI_NONE on split products
            else c[m * g.ldc + n] = v;
        }
}

}  // namespace
#include "plen_gemm_tc.cuh"      // k_gemm_tc: the same products on tcgen05 (TF32), see launch()
namespace {

// minibatch rows drawn uniformly with replacement (td3.py:175), written as the concatenated critic inputs:
// sa = [s | a], s2a = [s' | .] (the action columns are filled by the target actor), spi = [s | .] (filled by the actor)
__global__ void k_sample_sa(const float *__restrict__ store, long long size, int batch, uint64_t seed, float *__restrict__ sa,
                            float *__restrict__ s2a, float *__restrict__ spi, float *__restrict__ r, float *__restrict__ nd,
                            const CallParams *cp) {
    const int bi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (bi >= batch) return;
    if (cp) { size = cp->size; seed = cp->seed * 2654435761ull + 1ull; }      // plen_td3_train's draw, from device memory
    const uint32_t u = mix32(seed * 0x100000001B3ull + (uint64_t)bi);
    const long long row = (long long)(((uint64_t)u * (uint64_t)size) >> 32);
    const float *t = store + row * 72;
    for (int w = lane; w < 72; w += 32) {
        const float v = t[w];
        if (w < 44) { sa[(size_t)bi * SA + w] = v; if (w < 26) spi[(size_t)bi * SA + w] = v; }
        else if (w < 70) s2a[(size_t)bi * SA + (w - 44)] = v;
        else if (w == 70) r[bi] = v;
        else nd[bi] = 1.0f - v;
    }
}

__global__ void k_set_batch(int batch, const float *__restrict__ s, const float *__restrict__ a, const float *__restrict__ s2,
                            const float *__restrict__ r_in, const float *__restrict__ nd_in, float *__restrict__ sa,
                            float *__restrict__ s2a, float *__restrict__ spi, float *__restrict__ r, float *__restrict__ nd) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= batch * SA) return;
    const int bi = e / SA, w = e - bi * SA;
    if (w < S) { sa[e] = s[bi * S + w]; spi[e] = s[bi * S + w]; s2a[e] = s2[bi * S + w]; }
    else sa[e] = a[bi * A + (w - S)];
    if (w == 0) { r[bi] = r_in[bi]; nd[bi] = nd_in[bi]; }
}

// y = r + not_done * discount * min(Q1', Q2') (td3.py:311-313); critic_loss = mse(Q1, y) + mse(Q2, y) (:322-323);
// dq[z] = 2 (Q_z - y) / B.  One CTA, strided over the batch.
__global__ void k_td_loss(int batch, float discount, const float *__restrict__ qt, const float *__restrict__ q, const float *__restrict__ r,
                          const float *__restrict__ nd, float *__restrict__ dq, float *__restrict__ loss_out) {
    __shared__ float red[32];
    float acc = 0.0f;
    const float inv = 1.0f / (float)batch;
    for (int i = threadIdx.x; i < batch; i += blockDim.x) {
        const float y = r[i] + nd[i] * discount * fminf(qt[i], qt[batch + i]);
        const float e1 = q[i] - y, e2 = q[batch + i] - y;
        dq[i] = 2.0f * e1 * inv; dq[batch + i] = 2.0f * e2 * inv;
        acc += e1 * e1 + e2 * e2;
    }
    for (int m = 16; m > 0; m >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, m);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.0f;
        for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
        if (threadIdx.x == 0 && loss_out) *loss_out = v * inv;
    }
}

// actor_loss = -mean(Q1(s, actor(s))) (td3.py:342); dq = -1 / B
__global__ void k_actor_loss(int batch, const float *__restrict__ q, float *__restrict__ dq, float *__restrict__ loss_out) {
    __shared__ float red[32];
    float acc = 0.0f;
    const float inv = 1.0f / (float)batch;
    for (int i = threadIdx.x; i < batch; i += blockDim.x) { acc += q[i]; dq[i] = -inv; }
    for (int m = 16; m > 0; m >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, m);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.0f;
        for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
        if (threadIdx.x == 0 && loss_out) *loss_out = -v * inv;
    }
}

// torch.optim.Adam (td3.py:226-233: lr 3e-4, default betas / eps, no weight decay), same operation order as torch's
// single-tensor path: m.lerp_(g, 1 - b1); v = v b2 + (1 - b2) g g; p -= (lr / bc1) m / (sqrt(v) / sqrt(bc2) + eps)
__global__ void k_adam(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m, float *__restrict__ v, int n,
                       float one_minus_b1, float b2, float one_minus_b2, float step_size, float inv_bc2_sqrt, float eps,
                       const float *sc) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (sc) { step_size = sc[0]; inv_bc2_sqrt = sc[1]; }      // graph replay: the step-dependent scalars come from device memory
    const float gi = g[i];
    const float mi = m[i] + (gi - m[i]) * one_minus_b1;
    const float vi = v[i] * b2 + one_minus_b2 * gi * gi;
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) * inv_bc2_sqrt + eps;
    p[i] = p[i] - step_size * (mi / denom);
}

// target = tau * param + (1 - tau) * target (td3.py:352-360)
__global__ void k_soft_update(float *__restrict__ tgt, const float *__restrict__ src, int n, float tau) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) tgt[i] = tau * src[i] + (1.0f - tau) * tgt[i];
}

// minibatch size from which the dW products are split along K (= the minibatch) and accumulated atomically
constexpr int SPLIT_K_MIN_BATCH = 512;

// precision of the products of the entry point being enqueued: 0 = FP32 CUDA cores (parity path, default), 1 = 3xTF32 on the
// tensor cores for every product that fills a 128-row tile (plen_td3_set_precision, minibatches >= SPLIT_K_MIN_BATCH); set by
// the entry points from their context
thread_local int s_tc = 0;

int launch(Gemm g, cudaStream_t st) {
    g.splits = 1; g.k_chunk = g.K;
    if (g.epi == EPI_NONE && g.rowsum != nullptr && g.K >= SPLIT_K_MIN_BATCH) {      // dW = dY^T X of a large minibatch
        g.splits = g.K / 256 < 16 ? g.K / 256 : 16;
        g.k_chunk = ((g.K + g.splits - 1) / g.splits + 31) / 32 * 32;
        g.splits = (g.K + g.k_chunk - 1) / g.k_chunk;
    }
    if (s_tc && g.M >= 128 && g.N >= 16 && g.K >= 16) {
        // the three product kinds of the wide layers (the 18- and 1-column products stay on the FP32 path)
        dim3 grid((g.N + tcg::BN - 1) / tcg::BN, (g.M + tcg::BM - 1) / tcg::BM, g.nz * g.splits);
        if (g.ak == 1 && g.bk == 1 && g.epi == EPI_BIAS_RELU) { tcg::k_gemm_tc<true, true, EPI_BIAS_RELU><<<grid, tcg::NT, tcg::SMEM, st>>>(g); return 1; }
        if (g.ak == 1 && g.bk != 1 && g.epi == EPI_RELUMASK) { tcg::k_gemm_tc<true, false, EPI_RELUMASK><<<grid, tcg::NT, tcg::SMEM, st>>>(g); return 1; }
        if (g.ak != 1 && g.bk != 1 && g.epi == EPI_NONE && g.splits > 1) { tcg::k_gemm_tc<false, false, EPI_NONE><<<grid, tcg::NT, tcg::SMEM, st>>>(g); return 1; }
    }
    // wide tiles only when they still give every SM a CTA
    if ((long long)((g.N + 63) / 64) * ((g.M + 63) / 64) * g.nz * g.splits >= 148) {
        dim3 grid((g.N + 63) / 64, (g.M + 63) / 64, g.nz * g.splits);
        k_gemm<64, 64, 16, 4, 4><<<grid, GT, 0, st>>>(g);
    } else {
        dim3 grid((g.N + 31) / 32, (g.M + 31) / 32, g.nz * g.splits);
        k_gemm<32, 32, 32, 2, 2><<<grid, GT, 0, st>>>(g);
    }
    return 1;
}

// forward layer: C[z] = epi(X[z] W[z]^T + b[z]),  W in nn.Linear layout [N][K] with row stride ldw (a column sub-range of
// a wider weight is addressed by offsetting w)
Gemm fwd(const float *x, long long ldx, long long xz, const float *w, long long ldw, long long wz, const float *bias, float *c,
         long long ldc, long long cz, int M, int N, int K, int nz, int epi) {
    Gemm g;
    memset(&g, 0, sizeof g);
    g.a = x; g.am = ldx; g.ak = 1; g.az = xz;
    g.b = w; g.bk = 1; g.bn = ldw; g.bz = wz;
    g.c = c; g.ldc = ldc; g.cz = cz;
    g.M = M; g.N = N; g.K = K; g.nz = nz; g.epi = epi; g.bias = bias; g.bias_z = wz;
    return g;
}
// backward-data: dX[z] = epi(dY[z] W[z]),  dY [M][Nout] (ld ldy), W [Nout][Kin] (row stride ldw)
Gemm bwd_data(const float *dy, long long ldy, long long dyz, const float *w, long long ldw, long long wz, float *dx, long long ldx,
              long long dxz, int M, int Kin, int Nout, int nz, int epi, const float *aux, long long ld_aux, long long aux_z) {
    Gemm g;
    memset(&g, 0, sizeof g);
    g.a = dy; g.am = ldy; g.ak = 1; g.az = dyz;
    g.b = w; g.bk = ldw; g.bn = 1; g.bz = wz;
    g.c = dx; g.ldc = ldx; g.cz = dxz;
    g.M = M; g.N = Kin; g.K = Nout; g.nz = nz; g.epi = epi; g.aux = aux; g.ld_aux = ld_aux; g.aux_z = aux_z;
    return g;
}
// backward-weight: dW[z] = dY[z]^T X[z] ([Nout][Kin], row stride ldw), db[z] = column sums of dY[z]
Gemm bwd_weight(const float *dy, long long ldy, long long dyz, const float *x, long long ldx, long long xz, float *dw, long long ldw,
                float *db, long long gz, int Bt, int Nout, int Kin, int nz) {
    Gemm g;
    memset(&g, 0, sizeof g);
    g.a = dy; g.am = 1; g.ak = ldy; g.az = dyz;
    g.b = x; g.bk = ldx; g.bn = 1; g.bz = xz;
    g.c = dw; g.ldc = ldw; g.cz = gz;
    g.M = Nout; g.N = Kin; g.K = Bt; g.nz = nz; g.epi = EPI_NONE;
    g.rowsum = db; g.rowsum_z = gz;
    return g;
}

}  // namespace

struct plen_td3 {
    int device, max_batch, batch;
    int tc;                             // plen_td3_set_precision: 1 = TF32 tensor-core products
    float *ws;                          // one allocation, carved below
    float *sa, *s2a, *spi, *r, *nd;
    float *at_h1, *at_h2, *ct_h1, *ct_h2, *qt;
    float *c_h1, *c_h2, *q, *dq, *dh2, *dh1;
    float *a_h1, *a_h2, *p_h1, *p_h2, *qpi, *dqpi, *dp_h2, *dp_h1, *da3, *da_h2, *da_h1;
    long long launches;
    // plen_td3_train: per-call scalars in device memory + one captured CUDA graph per update kind (critic only / with policy)
    CallParams *d_cp;
    bool use_cp;                        // the piecewise entry points read seed / ring size / Adam scalars from d_cp
    cudaStream_t cap;                   // capture stream
    cudaStream_t side2;                 // third branch (plen_td3_train): the actor forward / the critic's target update next to the critic step
    bool serial;                        // env PLEN_TD3_SERIAL (read at create): every branch on the caller's stream -- the order the branches are tested against
    bool actor_fwd_done;                // the actor forward of this update was already enqueued (plen_td3_train)
    cudaStream_t side;                  // second branch of an update: independent products run next to the main chain (fork / join with ev)
    cudaEvent_t ev[8];
    cudaGraphExec_t exec[2];
    unsigned long long key[2];          // hash of everything a captured graph bakes in (pointers, batch, hyper-parameters)
    int nodes[2];
};

extern "C" {

int plen_td3_default_hyper(plen_td3_hyper *h) {
    if (!h) return plen_td3_set_error(PLEN_E_ARG, "plen_td3_default_hyper: NULL", "");
    // td3.py:211-219 (TD3Agent defaults), :226-233 (Adam lr 3e-4, torch default betas / eps)
    h->discount = 0.99f; h->tau = 0.005f; h->policy_noise = 0.2f; h->noise_clip = 0.5f; h->max_action = 1.0f;
    h->lr = 3e-4f; h->beta1 = 0.9f; h->beta2 = 0.999f; h->eps = 1e-8f; h->policy_freq = 2;
    return PLEN_OK;
}

plen_td3 *plen_td3_create(int max_batch, int device) {
    int ndev = 0;
    if (max_batch <= 0) { plen_td3_set_error(PLEN_E_ARG, "plen_td3_create: max_batch <= 0", ""); return nullptr; }
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
        plen_td3_set_error(PLEN_E_CUDA, "plen_td3_create: no usable CUDA device; this library has no CPU fallback", "");
        return nullptr;
    }
    plen_td3 *t = new (std::nothrow) plen_td3();
    if (!t) return nullptr;
    memset(t, 0, sizeof *t);
    t->device = device; t->max_batch = max_batch;
    t->serial = getenv("PLEN_TD3_SERIAL") != nullptr && atoi(getenv("PLEN_TD3_SERIAL")) != 0;
    const size_t Bm = (size_t)max_batch;
    const size_t words = 3 * Bm * SA + 2 * Bm + 2 * Bm * H + 4 * Bm * H + 2 * Bm + 4 * Bm * H + 4 * Bm + 4 * Bm * H +
                         4 * Bm * H + 2 * Bm + 2 * Bm * H + Bm * A + 2 * Bm * H;
    if (cudaSetDevice(device) != cudaSuccess || cudaMalloc(&t->ws, sizeof(float) * words) != cudaSuccess) {
        plen_td3_set_error(PLEN_E_CUDA, "plen_td3_create: cudaMalloc failed", "");
        delete t;
        return nullptr;
    }
    cudaMemset(t->ws, 0, sizeof(float) * words);
    float *p = t->ws;
    auto take = [&](size_t n) { float *q = p; p += n; return q; };
    t->sa = take(Bm * SA); t->s2a = take(Bm * SA); t->spi = take(Bm * SA); t->r = take(Bm); t->nd = take(Bm);
    t->at_h1 = take(Bm * H); t->at_h2 = take(Bm * H);
    t->ct_h1 = take(2 * Bm * H); t->ct_h2 = take(2 * Bm * H); t->qt = take(2 * Bm);
    t->c_h1 = take(2 * Bm * H); t->c_h2 = take(2 * Bm * H); t->q = take(2 * Bm); t->dq = take(2 * Bm);
    t->dh2 = take(2 * Bm * H); t->dh1 = take(2 * Bm * H);
    t->a_h1 = take(Bm * H); t->a_h2 = take(Bm * H); t->p_h1 = take(Bm * H); t->p_h2 = take(Bm * H);
    t->qpi = take(Bm); t->dqpi = take(Bm); t->dp_h2 = take(Bm * H); t->dp_h1 = take(Bm * H); t->da3 = take(Bm * A);
    t->da_h2 = take(Bm * H); t->da_h1 = take(Bm * H);
    bool ok = cudaStreamCreateWithFlags(&t->side, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&t->side2, cudaStreamNonBlocking) == cudaSuccess;
    for (int k = 0; k < 8 && ok; k++) ok = cudaEventCreateWithFlags(&t->ev[k], cudaEventDisableTiming) == cudaSuccess;
    if (!ok || cudaMalloc(&t->d_cp, sizeof(CallParams)) != cudaSuccess || cudaStreamCreateWithFlags(&t->cap, cudaStreamNonBlocking) != cudaSuccess) {
        plen_td3_set_error(PLEN_E_CUDA, "plen_td3_create: call-parameter buffer / capture stream", "");
        cudaFree(t->ws); delete t;
        return nullptr;
    }
    if ((size_t)(p - t->ws) != words) { plen_td3_set_error(PLEN_E_STATE, "plen_td3_create: workspace carve mismatch", ""); cudaFree(t->ws); delete t; return nullptr; }
    return t;
}

void plen_td3_destroy(plen_td3 *t) {
    if (!t) return;
    cudaSetDevice(t->device);
    for (int k = 0; k < 2; k++)
        if (t->exec[k]) cudaGraphExecDestroy(t->exec[k]);
    if (t->cap) cudaStreamDestroy(t->cap);
    if (t->side) cudaStreamDestroy(t->side);
    if (t->side2) cudaStreamDestroy(t->side2);
    for (int k = 0; k < 8; k++)
        if (t->ev[k]) cudaEventDestroy(t->ev[k]);
    cudaFree(t->d_cp);
    cudaFree(t->ws);
    delete t;
}

long long plen_td3_launches(const plen_td3 *t) { return t ? t->launches : 0; }

int plen_td3_set_precision(plen_td3 *t, int tf32) {
    if (!t || (tf32 != 0 && tf32 != 1)) return plen_td3_set_error(PLEN_E_ARG, "plen_td3_set_precision: bad arguments", "");
    if (tf32) {
        LCK(cudaSetDevice(t->device));
        LCK(cudaFuncSetAttribute(tcg::k_gemm_tc<true, true, EPI_BIAS_RELU>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tcg::SMEM));
        LCK(cudaFuncSetAttribute(tcg::k_gemm_tc<true, false, EPI_RELUMASK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tcg::SMEM));
        LCK(cudaFuncSetAttribute(tcg::k_gemm_tc<false, false, EPI_NONE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tcg::SMEM));
    }
    if (t->tc != tf32) {      // the captured update graphs bake the kernels in
        for (int k = 0; k < 2; k++)
            if (t->exec[k]) { cudaGraphExecDestroy(t->exec[k]); t->exec[k] = nullptr; }
    }
    t->tc = tf32;
    return PLEN_OK;
}

int plen_td3_tc_timed_out(void) {
    int v = 0;
    if (cudaMemcpyFromSymbol(&v, plen_tc::g_tc_timeout, sizeof v) != cudaSuccess) return -1;
    return v;
}

int plen_td3_sample(plen_td3 *t, plen_replay *rb, int batch, unsigned long long seed, void *stream) {
    if (!t || !rb || batch <= 0 || batch > t->max_batch) return plen_td3_set_error(PLEN_E_ARG, "plen_td3_sample: bad arguments", "");
    const long long size = plen_replay_size(rb);
    if (size <= 0) return plen_td3_set_error(PLEN_E_STATE, "plen_td3_sample: the replay buffer is empty", "");
    LCK(cudaSetDevice(t->device));
    k_sample_sa<<<(batch * 32 + 127) / 128, 128, 0, (cudaStream_t)stream>>>(plen_replay_storage(rb), size, batch, seed, t->sa, t->s2a,
                                                                          t->spi, t->r, t->nd, t->use_cp ? t->d_cp : nullptr);
    t->batch = batch; t->launches += 1;
    LCK(cudaGetLastError());
    return PLEN_OK;
}

int plen_td3_set_batch(plen_td3 *t, const float *state_dev, const float *action_dev, const float *next_state_dev,
                       const float *reward_dev, const float *not_done_dev, int batch, void *stream) {
    if (!t || !state_dev || !action_dev || !next_state_dev || !reward_dev || !not_done_dev || batch <= 0 || batch > t->max_batch)
        return plen_td3_set_error(PLEN_E_ARG, "plen_td3_set_batch: bad arguments", "");
    LCK(cudaSetDevice(t->device));
    k_set_batch<<<(batch * SA + 255) / 256, 256, 0, (cudaStream_t)stream>>>(batch, state_dev, action_dev, next_state_dev, reward_dev,
                                                                          not_done_dev, t->sa, t->s2a, t->spi, t->r, t->nd);
    t->batch = batch; t->launches += 1;
    LCK(cudaGetLastError());
    return PLEN_OK;
}

int plen_td3_critic_grads(plen_td3 *t, const plen_td3_params *P, const plen_td3_hyper *h, const float *noise_dev,
                          unsigned long long seed, float *critic_loss_dev, void *stream) {
    if (!t || !P || !h || !P->actor_target || !P->critic || !P->critic_target || !P->critic_grad)
        return plen_td3_set_error(PLEN_E_ARG, "plen_td3_critic_grads: bad arguments", "");
    if (t->batch <= 0) return plen_td3_set_error(PLEN_E_STATE, "plen_td3_critic_grads: no minibatch (call plen_td3_sample / plen_td3_set_batch)", "");
    LCK(cudaSetDevice(t->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int B = t->batch;
    const long long BH = (long long)B * H;
    int n = 0;
    s_tc = (t->tc && B >= SPLIT_K_MIN_BATCH) ? 1 : 0;      // (below that the gradient buffers are not zeroed and the update is launch bound)
    // Two branches (fork / join with events; captured as parallel graph branches by plen_td3_train): the update is a chain of
    // small launches whose fixed costs add up, so everything that does not depend on the chain runs next to it on t->side --
    // the current-Q forward next to the target chain, every dW = dY^T X next to the dX = dY W that continues the chain.
    cudaStream_t sd = t->serial ? st : t->side;
    auto fork = [&](int i) { if (sd != st) { cudaEventRecord(t->ev[i], st); cudaStreamWaitEvent(sd, t->ev[i], 0); } };      // side continues after st's work so far
    auto join = [&](int i) { if (sd != st) { cudaEventRecord(t->ev[i], sd); cudaStreamWaitEvent(st, t->ev[i], 0); } };      // st continues after side's work so far
    fork(0);
    // ---- side: current Q, td3.py:316 (and the zeroed gradient vector the split-K products accumulate into)
    const float *c = P->critic;
    float *gr = P->critic_grad;
    n += launch(fwd(t->sa, SA, 0, c + CW1, SA, TOWER_N, c + CB1, t->c_h1, H, BH, B, H, SA, 2, EPI_BIAS_RELU), sd);
    n += launch(fwd(t->c_h1, H, BH, c + CW2, H, TOWER_N, c + CB2, t->c_h2, H, BH, B, H, H, 2, EPI_BIAS_RELU), sd);
    n += launch(fwd(t->c_h2, H, BH, c + CW3, H, TOWER_N, c + CB3, t->q, 1, B, B, 1, H, 2, EPI_BIAS), sd);
    if (B >= SPLIT_K_MIN_BATCH) { LCK(cudaMemsetAsync(gr, 0, sizeof(float) * CRITIC_N, sd)); n += 1; }      // split-K accumulates
    // ---- main: target Q (no gradient), td3.py:303-313
    const float *at = P->actor_target;
    n += launch(fwd(t->s2a, SA, 0, at + AW1, S, 0, at + AB1, t->at_h1, H, 0, B, H, S, 1, EPI_BIAS_RELU), st);
    n += launch(fwd(t->at_h1, H, 0, at + AW2, H, 0, at + AB2, t->at_h2, H, 0, B, H, H, 1, EPI_BIAS_RELU), st);
    {
        Gemm g = fwd(t->at_h2, H, 0, at + AW3, H, 0, at + AB3, t->s2a + S, SA, 0, B, A, H, 1, EPI_BIAS_TANH_NOISE);
        g.p0 = h->max_action; g.p1 = h->policy_noise; g.p2 = h->noise_clip; g.seed = seed; g.aux = noise_dev; g.ld_aux = A; g.cp = t->use_cp ? t->d_cp : nullptr;
        n += launch(g, st);
    }
    const float *ct = P->critic_target;
    n += launch(fwd(t->s2a, SA, 0, ct + CW1, SA, TOWER_N, ct + CB1, t->ct_h1, H, BH, B, H, SA, 2, EPI_BIAS_RELU), st);
    n += launch(fwd(t->ct_h1, H, BH, ct + CW2, H, TOWER_N, ct + CB2, t->ct_h2, H, BH, B, H, H, 2, EPI_BIAS_RELU), st);
    n += launch(fwd(t->ct_h2, H, BH, ct + CW3, H, TOWER_N, ct + CB3, t->qt, 1, B, B, 1, H, 2, EPI_BIAS), st);
    join(1);
    k_td_loss<<<1, 256, 0, st>>>(B, h->discount, t->qt, t->q, t->r, t->nd, t->dq, critic_loss_dev);
    n += 1;
    // ---- backward through both towers, td3.py:333 (gradients land in P->critic_grad, flat layout of the critic):
    //      dX chain on the main branch, dW products on the side branch
    fork(2);
    n += launch(bwd_weight(t->dq, 1, B, t->c_h2, H, BH, gr + CW3, H, gr + CB3, TOWER_N, B, 1, H, 2), sd);
    n += launch(bwd_data(t->dq, 1, B, c + CW3, H, TOWER_N, t->dh2, H, BH, B, H, 1, 2, EPI_RELUMASK, t->c_h2, H, BH), st);
    fork(3);
    n += launch(bwd_weight(t->dh2, H, BH, t->c_h1, H, BH, gr + CW2, H, gr + CB2, TOWER_N, B, H, H, 2), sd);
    n += launch(bwd_data(t->dh2, H, BH, c + CW2, H, TOWER_N, t->dh1, H, BH, B, H, H, 2, EPI_RELUMASK, t->c_h1, H, BH), st);
    fork(4);
    n += launch(bwd_weight(t->dh1, H, BH, t->sa, SA, 0, gr + CW1, SA, gr + CB1, TOWER_N, B, H, SA, 2), sd);
    join(5);
    t->launches += n;
    LCK(cudaGetLastError());
    return PLEN_OK;
}

// actor(state) for the policy step: spi = [s | actor(s)] (3 launches).  Depends on the actor's parameters and the minibatch only.
static int actor_forward_launches(plen_td3 *t, const plen_td3_params *P, const plen_td3_hyper *h, cudaStream_t st) {
    const int B = t->batch;
    const float *ac = P->actor;
    int n = 0;
    s_tc = (t->tc && B >= SPLIT_K_MIN_BATCH) ? 1 : 0;
    n += launch(fwd(t->spi, SA, 0, ac + AW1, S, 0, ac + AB1, t->a_h1, H, 0, B, H, S, 1, EPI_BIAS_RELU), st);
    n += launch(fwd(t->a_h1, H, 0, ac + AW2, H, 0, ac + AB2, t->a_h2, H, 0, B, H, H, 1, EPI_BIAS_RELU), st);
    Gemm g = fwd(t->a_h2, H, 0, ac + AW3, H, 0, ac + AB3, t->spi + S, SA, 0, B, A, H, 1, EPI_BIAS_TANH);
    g.p0 = h->max_action;
    n += launch(g, st);
    return n;
}

int plen_td3_actor_grads(plen_td3 *t, const plen_td3_params *P, const plen_td3_hyper *h, float *actor_loss_dev, void *stream) {
    if (!t || !P || !h || !P->actor || !P->critic || !P->actor_grad)
        return plen_td3_set_error(PLEN_E_ARG, "plen_td3_actor_grads: bad arguments", "");
    if (t->batch <= 0) return plen_td3_set_error(PLEN_E_STATE, "plen_td3_actor_grads: no minibatch", "");
    LCK(cudaSetDevice(t->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int B = t->batch;
    int n = 0;
    s_tc = (t->tc && B >= SPLIT_K_MIN_BATCH) ? 1 : 0;      // (below that the gradient buffers are not zeroed and the update is launch bound)
    // ---- actor_loss = -critic.Q1(state, actor(state)).mean(), td3.py:342
    const float *ac = P->actor, *c = P->critic;
    if (!t->actor_fwd_done) n += actor_forward_launches(t, P, h, st);      // (plen_td3_train runs it next to the critic step)
    t->actor_fwd_done = false;
    n += launch(fwd(t->spi, SA, 0, c + CW1, SA, 0, c + CB1, t->p_h1, H, 0, B, H, SA, 1, EPI_BIAS_RELU), st);
    n += launch(fwd(t->p_h1, H, 0, c + CW2, H, 0, c + CB2, t->p_h2, H, 0, B, H, H, 1, EPI_BIAS_RELU), st);
    n += launch(fwd(t->p_h2, H, 0, c + CW3, H, 0, c + CB3, t->qpi, 1, 0, B, 1, H, 1, EPI_BIAS), st);
    k_actor_loss<<<1, 256, 0, st>>>(B, t->qpi, t->dqpi, actor_loss_dev);
    n += 1;
    // ---- backward: through Q1 down to its action inputs (no critic parameter gradients are needed: the critic
    //      optimiser zeroes them before its next step, td3.py:328), then through the actor
    n += launch(bwd_data(t->dqpi, 1, 0, c + CW3, H, 0, t->dp_h2, H, 0, B, H, 1, 1, EPI_RELUMASK, t->p_h2, H, 0), st);
    n += launch(bwd_data(t->dp_h2, H, 0, c + CW2, H, 0, t->dp_h1, H, 0, B, H, H, 1, EPI_RELUMASK, t->p_h1, H, 0), st);
    {
        Gemm g = bwd_data(t->dp_h1, H, 0, c + CW1 + S, SA, 0, t->da3, A, 0, B, A, H, 1, EPI_TANHGRAD, t->spi + S, SA, 0);
        g.p0 = h->max_action;
        n += launch(g, st);
    }
    // dX chain on the main branch, dW products on the side branch (see plen_td3_critic_grads)
    cudaStream_t sd = t->serial ? st : t->side;
    auto fork = [&](int i) { if (sd != st) { cudaEventRecord(t->ev[i], st); cudaStreamWaitEvent(sd, t->ev[i], 0); } };
    auto join = [&](int i) { if (sd != st) { cudaEventRecord(t->ev[i], sd); cudaStreamWaitEvent(st, t->ev[i], 0); } };
    float *gr = P->actor_grad;
    fork(0);
    if (B >= SPLIT_K_MIN_BATCH) { LCK(cudaMemsetAsync(gr, 0, sizeof(float) * ACTOR_N, sd)); n += 1; }
    n += launch(bwd_weight(t->da3, A, 0, t->a_h2, H, 0, gr + AW3, H, gr + AB3, 0, B, A, H, 1), sd);
    n += launch(bwd_data(t->da3, A, 0, ac + AW3, H, 0, t->da_h2, H, 0, B, H, A, 1, EPI_RELUMASK, t->a_h2, H, 0), st);
    fork(1);
    n += launch(bwd_weight(t->da_h2, H, 0, t->a_h1, H, 0, gr + AW2, H, gr + AB2, 0, B, H, H, 1), sd);
    n += launch(bwd_data(t->da_h2, H, 0, ac + AW2, H, 0, t->da_h1, H, 0, B, H, H, 1, EPI_RELUMASK, t->a_h1, H, 0), st);
    fork(2);
    n += launch(bwd_weight(t->da_h1, H, 0, t->spi, SA, 0, gr + AW1, S, gr + AB1, 0, B, H, S, 1), sd);
    join(3);
    t->launches += n;
    LCK(cudaGetLastError());
    return PLEN_OK;
}

// scalars as torch computes them (python doubles, torch/optim/adam.py _single_tensor_adam), then cast to float32
static void adam_scalars(const plen_td3_hyper *h, long long step, float out[2]) {
    const double bc1 = 1.0 - pow((double)h->beta1, (double)step), bc2 = 1.0 - pow((double)h->beta2, (double)step);
    out[0] = (float)((double)h->lr / bc1);
    out[1] = (float)(1.0 / sqrt(bc2));
}

// sc (nullable): device pointer to {step size, 1 / sqrt(bias correction 2)}; overrides the values computed from `step`
static int adam_launch(float *param_dev, const float *grad_dev, float *m_dev, float *v_dev, int n, long long step,
                       const plen_td3_hyper *h, int device, cudaStream_t st, const float *sc) {
    LCK(cudaSetDevice(device));
    float s2[2];
    adam_scalars(h, step, s2);
    k_adam<<<(n + 255) / 256, 256, 0, st>>>(param_dev, grad_dev, m_dev, v_dev, n, 1.0f - h->beta1, h->beta2, 1.0f - h->beta2, s2[0],
                                           s2[1], h->eps, sc);
    LCK(cudaGetLastError());
    return PLEN_OK;
}

int plen_td3_adam(float *param_dev, const float *grad_dev, float *m_dev, float *v_dev, int n, long long step,
                  const plen_td3_hyper *h, int device, void *stream) {
    if (!param_dev || !grad_dev || !m_dev || !v_dev || n <= 0 || step <= 0 || !h)
        return plen_td3_set_error(PLEN_E_ARG, "plen_td3_adam: bad arguments", "");
    return adam_launch(param_dev, grad_dev, m_dev, v_dev, n, step, h, device, (cudaStream_t)stream, nullptr);
}

int plen_td3_soft_update(float *target_dev, const float *source_dev, int n, float tau, int device, void *stream) {
    if (!target_dev || !source_dev || n <= 0) return plen_td3_set_error(PLEN_E_ARG, "plen_td3_soft_update: bad arguments", "");
    LCK(cudaSetDevice(device));
    k_soft_update<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(target_dev, source_dev, n, tau);
    LCK(cudaGetLastError());
    return PLEN_OK;
}

// the whole update enqueued on st, every per-call scalar read from t->d_cp (so the sequence can be captured once)
static int enqueue_train(plen_td3 *t, const plen_td3_params *P, const plen_td3_hyper *h, plen_replay *rb, int batch, bool policy,
                         float *losses_dev, cudaStream_t st) {
    int rc;
    t->use_cp = true;
    do {
        if (rb) { rc = plen_td3_sample(t, rb, batch, 0, st); if (rc) break; }
        cudaStream_t s2 = t->serial ? st : t->side2;
        if (policy) {       // third branch: the actor forward needs the minibatch and the actor only -- it runs next to the critic step
            if (s2 != st) { cudaEventRecord(t->ev[6], st); cudaStreamWaitEvent(s2, t->ev[6], 0); }
            t->launches += actor_forward_launches(t, P, h, s2);
            t->actor_fwd_done = true;
        }
        rc = plen_td3_critic_grads(t, P, h, nullptr, 0, losses_dev ? losses_dev + 1 : nullptr, st);
        if (rc) break;
        rc = adam_launch(P->critic, P->critic_grad, P->critic_m, P->critic_v, CRITIC_N, 1, h, t->device, st, t->d_cp->adam_c);
        if (rc) break;
        t->launches += 1;
        if (policy) {       // delayed policy update, td3.py:339-360
            // ... and so does the critic's Polyak update once its Adam step is in (the policy step reads the critic, not its target)
            if (s2 != st) {
                cudaEventRecord(t->ev[7], s2); cudaStreamWaitEvent(st, t->ev[7], 0);          // join: spi is complete
                cudaEventRecord(t->ev[6], st); cudaStreamWaitEvent(s2, t->ev[6], 0);          // fork after the critic's Adam step
            }
            rc = plen_td3_soft_update(P->critic_target, P->critic, CRITIC_N, h->tau, t->device, s2);
            if (rc) break;
            rc = plen_td3_actor_grads(t, P, h, losses_dev, st);
            if (rc) break;
            rc = adam_launch(P->actor, P->actor_grad, P->actor_m, P->actor_v, ACTOR_N, 1, h, t->device, st, t->d_cp->adam_a);
            if (rc) break;
            rc = plen_td3_soft_update(P->actor_target, P->actor, ACTOR_N, h->tau, t->device, st);
            if (rc) break;
            if (s2 != st) { cudaEventRecord(t->ev[7], s2); cudaStreamWaitEvent(st, t->ev[7], 0); }      // join the target update
            t->launches += 3;
        }
    } while (0);
    t->actor_fwd_done = false;
    t->use_cp = false;
    return rc;
}

static unsigned long long mix_key(unsigned long long k, unsigned long long v) {
    k ^= v + 0x9E3779B97F4A7C15ull + (k << 6) + (k >> 2);
    return k;
}

int plen_td3_train(plen_td3 *t, const plen_td3_params *P, const plen_td3_hyper *h, plen_replay *rb, int batch, long long total_it,
                   long long critic_step, long long actor_step, unsigned long long seed, float *losses_dev, void *stream) {
    if (!t || !P || !h || total_it <= 0 || critic_step <= 0 || h->policy_freq <= 0)
        return plen_td3_set_error(PLEN_E_ARG, "plen_td3_train: bad arguments", "");
    const bool policy = total_it % h->policy_freq == 0;
    if (policy && actor_step <= 0) return plen_td3_set_error(PLEN_E_ARG, "plen_td3_train: actor_step <= 0 on a policy update", "");
    if (rb && (batch <= 0 || batch > t->max_batch)) return plen_td3_set_error(PLEN_E_ARG, "plen_td3_train: bad batch", "");
    if (rb && plen_replay_size(rb) <= 0) return plen_td3_set_error(PLEN_E_STATE, "plen_td3_train: the replay buffer is empty", "");
    LCK(cudaSetDevice(t->device));
    cudaStream_t st = (cudaStream_t)stream;
    // per-call scalars -> device memory (a small pageable copy is staged before the call returns)
    CallParams cp;
    cp.seed = seed; cp.size = rb ? plen_replay_size(rb) : 0;
    adam_scalars(h, critic_step, cp.adam_c);
    adam_scalars(h, policy ? actor_step : 1, cp.adam_a);
    LCK(cudaMemcpyAsync(t->d_cp, &cp, sizeof cp, cudaMemcpyHostToDevice, st));
    // One CUDA graph per update kind: 18 / 36 kernel nodes replayed with one launch.  Everything the graph bakes in is
    // hashed; a change (new buffers, batch, hyper-parameters) re-captures.
    unsigned long long key = 0x1234;
    const void *ptrs[] = {P->actor, P->actor_target, P->critic, P->critic_target, P->actor_m, P->actor_v, P->critic_m, P->critic_v,
                          P->actor_grad, P->critic_grad, losses_dev, rb ? (const void *)plen_replay_storage(rb) : nullptr};
    for (const void *q : ptrs) key = mix_key(key, (unsigned long long)(uintptr_t)q);
    const float hf[] = {h->discount, h->tau, h->policy_noise, h->noise_clip, h->max_action, h->beta1, h->beta2, h->eps};
    for (float f : hf) { unsigned u; memcpy(&u, &f, 4); key = mix_key(key, u); }
    key = mix_key(key, (unsigned long long)(rb ? batch : t->batch));
    const int kind = policy ? 1 : 0;
    if (!t->exec[kind] || t->key[kind] != key) {
        if (t->exec[kind]) { cudaGraphExecDestroy(t->exec[kind]); t->exec[kind] = nullptr; }
        cudaGraph_t graph = nullptr;
        LCK(cudaStreamBeginCapture(t->cap, cudaStreamCaptureModeRelaxed));
        const long long l0 = t->launches;
        const int rc = enqueue_train(t, P, h, rb, batch, policy, losses_dev, t->cap);
        const cudaError_t ce = cudaStreamEndCapture(t->cap, &graph);
        if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
        if (ce != cudaSuccess) return plen_td3_set_error(PLEN_E_CUDA, "plen_td3_train: graph capture: ", cudaGetErrorString(ce));
        t->nodes[kind] = (int)(t->launches - l0);
        t->launches = l0;
        const cudaError_t ie = cudaGraphInstantiate(&t->exec[kind], graph, 0);
        cudaGraphDestroy(graph);
        if (ie != cudaSuccess) { t->exec[kind] = nullptr; return plen_td3_set_error(PLEN_E_CUDA, "plen_td3_train: graph instantiate: ", cudaGetErrorString(ie)); }
        t->key[kind] = key;
    }
    LCK(cudaGraphLaunch(t->exec[kind], st));
    t->launches += t->nodes[kind];
    return PLEN_OK;
}

}  // extern "C"
