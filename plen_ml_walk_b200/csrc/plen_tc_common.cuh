// plen_tc_common.cuh -- tcgen05 / TMEM / mbarrier helpers shared by the tensor-core kernels of libplen_b200.so
// (plen_actor_tc.cu: actor inference, kind::f16; plen_td3_learn.cu: learner GEMMs, kind::tf32).  sm_100a only.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace plen_tc {

// UMMA shared-memory descriptor (cute::UMMA::SmemDescriptor): start >> 4 [0,14), LBO >> 4 [16,30), SBO >> 4 [32,46),
// version 1 [46,48), layout type 0 = no swizzle [61,64)
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}

// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 (1 << 4), A = B = F16 (format 0 at [7,10) and [10,13)), both
// K-major, N >> 3 at [17,23), M >> 4 at [24,29)
__device__ __forceinline__ uint32_t instr_desc(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
// set when an mbarrier wait gave up (never expected; keeps a protocol bug from hanging the GPU).  One copy per translation
// unit (this build has no relocatable device code): plen_actor_tc_timed_out / plen_td3_tc_timed_out read theirs.
static __device__ int g_tc_timeout = 0;
// Returns false when the wait was abandoned: the caller then poisons its outputs with NaN (the env step's numeric guard and
// td3.actor_forward's check of plen_actor_tc_timed_out both catch that) instead of reading an unfinished accumulator.
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
    for (int tries = 0; tries < (1 << 22); tries++) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return true;
    }
    g_tc_timeout = 1;
    return false;
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }


// kind::tf32: A and B are fp32 words in shared memory read as TF32 (10-bit mantissa), 8 K entries (32 bytes) per instruction
__device__ __forceinline__ uint32_t instr_desc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
        :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}

}  // namespace plen_tc
