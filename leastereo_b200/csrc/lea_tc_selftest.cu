// First-contact self test of the tcgen05 building blocks, independent of TMA and of the conv pipeline:
// ONE tcgen05.mma (M=128, N=16, K=16, bf16 -> fp32) on operands written to shared memory by ordinary stores in the
// no-swizzle K-major core-matrix layout the conv kernel relies on, for several (start offset, LBO, SBO) choices -
// including a start address that is only 16-byte aligned and an 8-row-group stride of 160 B, which is what the
// "every tap is a start address into one staged halo slab" trick needs.  Integer-valued inputs make the expected
// result exact.  Prints one line per variant; returns 0 when all designed variants match.
#include "lea_common.h"
#include <cstdio>
#include <vector>

namespace {

struct StParams { uint32_t a_off, a_lbo, a_sbo, b_off, b_lbo, b_sbo; int swap; };

__device__ __forceinline__ uint32_t st_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t st_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__device__ __forceinline__ int a_val(int m, int k) { return ((m * 3 + k * 5) % 17) - 8; }
__device__ __forceinline__ int b_val(int n, int k) { return ((n * 7 + k * 3) % 13) - 6; }

__global__ void __launch_bounds__(128, 1) lea_tc_selftest_kernel(StParams p, float* out, int* flag) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t* A = smem;              // 16 KB region
    uint8_t* Bm = smem + 16384;     // 4 KB region
    for (int i = tid; i < (16384 + 4096) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x7fc07fc0u;  // NaN fill
    __syncthreads();
    // A[m][k]: row m at (m/8)*sbo + (m%8)*16, k half at (k/8)*lbo
    for (int e = tid; e < 128 * 16; e += 128) {
        const int m = e / 16, k = e % 16;
        const uint32_t off = p.a_off + (m / 8) * p.a_sbo + (m % 8) * 16 + (k / 8) * p.a_lbo + (k % 8) * 2;
        *reinterpret_cast<uint16_t*>(A + off) = lea_f32_to_bf16((float)a_val(m, k));
    }
    for (int e = tid; e < 16 * 16; e += 128) {
        const int n = e / 16, k = e % 16;
        const uint32_t off = p.b_off + (n / 8) * p.b_sbo + (n % 8) * 16 + (k / 8) * p.b_lbo + (k % 8) * 2;
        *reinterpret_cast<uint16_t*>(Bm + off) = lea_f32_to_bf16((float)b_val(n, k));
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(st_smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(st_smem_u32(&tmem_slot)), "r"(32) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy stores -> async proxy (UMMA)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t aa = st_smem_u32(A) + p.a_off, ba = st_smem_u32(Bm) + p.b_off;
        const uint64_t ad = p.swap ? st_desc(aa, p.a_sbo, p.a_lbo) : st_desc(aa, p.a_lbo, p.a_sbo);
        const uint64_t bd = p.swap ? st_desc(ba, p.b_sbo, p.b_lbo) : st_desc(ba, p.b_lbo, p.b_sbo);
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(0u) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                     ::"r"(st_smem_u32(&bar)) : "memory");
    }
    // bounded wait
    {
        uint32_t ok = 0;
        const unsigned long long t0 = clock64();
        while (!ok) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                         "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(st_smem_u32(&bar)), "r"(0u) : "memory");
            if (!ok && clock64() - t0 > 2000000000ull) { if (tid == 0) *flag = 1; break; }
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t r[16];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    const int m = warp * 32 + lane;
    for (int n = 0; n < 16; ++n) out[m * 16 + n] = __uint_as_float(r[n]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32) : "memory");
}

int host_a(int m, int k) { return ((m * 3 + k * 5) % 17) - 8; }
int host_b(int n, int k) { return ((n * 7 + k * 3) % 13) - 6; }

}  // namespace

extern "C" int lea_tc_selftest(int32_t verbose, void* stream) {
    float* d_out = nullptr; int* d_flag = nullptr;
    LEA_CHECK(cudaMalloc(&d_out, 128 * 16 * sizeof(float)) == cudaSuccess, "selftest: cudaMalloc failed");
    LEA_CHECK(cudaMalloc(&d_flag, sizeof(int)) == cudaSuccess, "selftest: cudaMalloc failed");
    cudaFuncSetAttribute(lea_tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 24 * 1024);
    struct Variant { const char* name; StParams p; bool designed; };
    const Variant variants[] = {
        // a_off a_lbo a_sbo  b_off b_lbo b_sbo swap
        {"dense rows, lbo=K-half stride, sbo=8-row stride", {0, 2048, 128, 0, 256, 128, 0}, true},
        {"halo pitch: sbo=160, start +16 B (tap shift kw=1)", {16, 2880, 160, 0, 256, 128, 0}, true},
        {"halo pitch: sbo=160, start +176 B (kh=1,kw=1)",     {176, 2880, 160, 0, 256, 128, 0}, true},
        {"B tile inside a wider image: b_lbo=1024 (64 rows)", {0, 2048, 128, 0, 1024, 128, 0}, true},
    };
    int failures = 0;
    std::vector<float> h(128 * 16);
    for (const Variant& v : variants) {
        cudaMemsetAsync(d_flag, 0, sizeof(int), (cudaStream_t)stream);
        cudaMemsetAsync(d_out, 0xff, 128 * 16 * sizeof(float), (cudaStream_t)stream);
        lea_tc_selftest_kernel<<<1, 128, 20480, (cudaStream_t)stream>>>(v.p, d_out, d_flag);
        cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream);
        if (e != cudaSuccess) {
            lea_set_error("selftest: kernel failed: %s", cudaGetErrorString(e));
            if (verbose) printf("[tc_selftest] %-55s CUDA ERROR %s\n", v.name, cudaGetErrorString(e));
            return 100;
        }
        int flag = 0;
        cudaMemcpy(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost);
        cudaMemcpy(h.data(), d_out, h.size() * sizeof(float), cudaMemcpyDeviceToHost);
        int bad = 0; double maxerr = 0;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < 16; ++n) {
                int ref = 0;
                for (int k = 0; k < 16; ++k) ref += host_a(m, k) * host_b(n, k);
                const double err = fabs((double)h[m * 16 + n] - ref);
                if (!(err <= 1e-3)) ++bad;
                if (err == err && err > maxerr) maxerr = err;
            }
        const bool ok = (bad == 0) && !flag;
        if (verbose)
            printf("[tc_selftest] %-55s %s (mismatches %d/2048, max err %.3g%s)%s\n", v.name, ok ? "MATCH" : "differ",
                   bad, maxerr, flag ? ", TIMEOUT" : "", v.designed ? "" : "  [control]");
        if (v.designed && !ok) ++failures;
    }
    fflush(stdout);
    cudaFree(d_out); cudaFree(d_flag);
    if (failures) lea_set_error("selftest: %d designed tcgen05 descriptor variants did not match", failures);
    return failures;
}
