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

// ---------------------------------------------------------------------------------------------------------
// Micro-benchmark of tcgen05.mma issue/throughput on this chip (diagnostic, not part of the public header):
// `iters` MMAs of shape M=128 x N x K=16 (bf16) from one elected thread, A from shared memory (SS) or from tensor
// memory (TS), accumulating into `nacc` different accumulators in rotation.  Operand contents are irrelevant.
// ---------------------------------------------------------------------------------------------------------
namespace {
struct MbParams { int n, nacc, a_in_tmem, iters, a_rot, sbo_a, b_rot, a_lbo, b_lbo, ld_col; };

__global__ void __launch_bounds__(128, 1) lea_tc_microbench_kernel(MbParams p, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 64 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(st_smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(st_smem_u32(&tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (warp == 0) {
        // the whole warp walks the loop so that descriptor arithmetic stays in uniform registers; one lane issues
        uint32_t elected = 0;
        asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, %1;\n\tselp.u32 %0, 1, 0, px;\n\t}"
                     : "=r"(elected) : "r"(0xffffffffu));
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t abase = st_smem_u32(smem) >> 4, bbase = (st_smem_u32(smem) + 32768) >> 4;
        const uint32_t a_tmem = tmem + 448;                      // 8 columns of packed bf16 A (TS mode)
        const uint32_t a_lo0 = abase | (((uint32_t)p.a_lbo >> 4) << 16), a_hi = ((uint32_t)p.sbo_a >> 4) | (1u << 14);
        const uint32_t b_lo0 = bbase | (((uint32_t)p.b_lbo >> 4) << 16), b_hi = 8u | (1u << 14);
        const uint32_t amask = (uint32_t)p.a_rot - 1u, dmask = (uint32_t)p.nacc - 1u, bmask = (uint32_t)p.b_rot - 1u;
        const long long t0 = clock64();
#pragma unroll 8
        for (int i = 0; i < p.iters; ++i) {
            const uint32_t d = tmem + ((uint32_t)i & dmask) * (uint32_t)p.n;
            const uint32_t a_lo = a_lo0 + ((uint32_t)i & amask);
            const uint32_t b_lo = b_lo0 + (((uint32_t)i >> 1) & bmask) * 24u;       // a different B tile every 2nd MMA
            if (p.a_in_tmem) {
                asm volatile(
                    "{\n\t.reg .pred p, q;\n\t.reg .b64 db;\n\tsetp.ne.b32 q, %0, 0;\n\tmov.b64 db, {%3, %4};\n\t"
                    "setp.ne.b32 p, %6, 0;\n\t"
                    "@q tcgen05.mma.cta_group::1.kind::f16 [%1], [%2], db, %5, p;\n\t}"
                    ::"r"(elected), "r"(d), "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(1u) : "memory");
            } else {
                asm volatile(
                    "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 q, %0, 0;\n\t"
                    "mov.b64 da, {%2, %3};\n\tmov.b64 db, {%4, %5};\n\tsetp.ne.b32 p, %7, 0;\n\t"
                    "@q tcgen05.mma.cta_group::1.kind::f16 [%1], da, db, %6, p;\n\t}"
                    ::"r"(elected), "r"(d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(1u) : "memory");
            }
        }
        const long long t_issue = clock64();
        asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %0, 0;\n\t"
                     "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%1];\n\t}"
                     ::"r"(elected), "r"(st_smem_u32(&bar)) : "memory");
        uint32_t ok = 0;
        while (!ok) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                         "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(st_smem_u32(&bar)), "r"(0u) : "memory");
            if (!ok && clock64() - t0 > 4000000000ll) break;
        }
        if (tid == 0) { out[2 * blockIdx.x] = clock64() - t0; out[2 * blockIdx.x + 1] = t_issue - t0; }
        if (tid == 0) *reinterpret_cast<volatile uint32_t*>(&tmem_slot) = 0xffffffffu;     // stop the loader warps
    } else if (p.ld_col >= 0) {
        // loader warps: back-to-back tcgen05.ld of 32 columns at ld_col while warp 0 issues MMAs into [0, nacc*n)
        uint32_t r[16], sink = 0;
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)p.ld_col;
        while (*reinterpret_cast<volatile uint32_t*>(&tmem_slot) != 0xffffffffu) {
#pragma unroll
            for (int rep = 0; rep < 2; ++rep) {
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
                    "tcgen05.wait::ld.sync.aligned;"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                      "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                    : "r"(taddr + (uint32_t)(rep * 16)) : "memory");
                sink ^= r[0] ^ r[15];
            }
        }
        if (sink == 0x12345678u) out[0] = 1;      // keep the loads alive
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}
}  // namespace

extern "C" int lea_tc_microbench(int32_t grid, void* stream) {
    long long* d_out = nullptr;
    if (grid < 1) grid = 1;
    LEA_CHECK(cudaMalloc(&d_out, 2 * grid * sizeof(long long)) == cudaSuccess, "microbench: cudaMalloc failed");
    cudaFuncSetAttribute(lea_tc_microbench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    std::vector<long long> h(2 * grid);
    const int iters = 2000;
    printf("[tc_microbench] grid=%d iters=%d   cycles per tcgen05.mma (M=128, K=16, bf16; floor = N/2)\n", grid, iters);
    printf("[tc_microbench] %5s %5s %6s %6s %6s | %9s %9s %9s\n", "N", "nacc", "A", "a_rot", "b_rot", "min", "avg", "issue");
    printf("[tc_microbench] (columns: N nacc A a_rot b_rot a_lbo b_lbo)\n");
    struct Case { int n, a_in_tmem, a_rot, b_rot, a_lbo, b_lbo; };
    const Case cases[] = {
        {96, 0, 1, 1, 2880, 1536}, {96, 0, 8, 8, 2880, 1536}, {96, 0, 8, 8, 5760, 1536}, {96, 0, 8, 8, 5760, 1600},
        {96, 0, 8, 8, 2880, 1600}, {96, 0, 8, 8, 5824, 1536}, {96, 0, 8, 8, 8640, 1536},
        {32, 0, 8, 8, 5760, 512}, {32, 0, 8, 8, 2880, 512}, {64, 0, 8, 8, 5760, 1024}, {64, 0, 8, 8, 2880, 1088},
        {192, 0, 8, 8, 5760, 3072}, {192, 0, 8, 8, 2880, 3136}, {144, 0, 8, 8, 5760, 2304}, {144, 0, 8, 8, 2880, 2368},
        {128, 0, 8, 8, 5760, 2048}, {128, 0, 8, 8, 2880, 2112}, {256, 0, 8, 8, 2880, 4160},
    };
    for (const Case& c : cases) {
        MbParams p{c.n, 1, c.a_in_tmem, iters, c.a_rot, 160, c.b_rot, c.a_lbo, c.b_lbo, -1};
        lea_tc_microbench_kernel<<<grid, 128, 64 * 1024, (cudaStream_t)stream>>>(p, d_out);
        cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream);
        if (e != cudaSuccess) { lea_set_error("microbench: %s", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h.data(), d_out, 2 * grid * sizeof(long long), cudaMemcpyDeviceToHost);
        long long mn = h[0]; double avg = 0, iss = 0;
        for (int k = 0; k < grid; ++k) { if (h[2 * k] < mn) mn = h[2 * k]; avg += (double)h[2 * k]; iss += (double)h[2 * k + 1]; }
        printf("[tc_microbench] %5d %5s a_rot %d b_rot %d a_lbo %5d b_lbo %5d | min %7.1f avg %7.1f issue %7.1f\n", c.n,
               c.a_in_tmem ? "tmem" : "smem", c.a_rot, c.b_rot, c.a_lbo, c.b_lbo, (double)mn / iters, avg / grid / iters,
               iss / grid / iters);
    }
    // MMAs accumulating into columns [0, 96) while three other warps read 32 TMEM columns at ld_col in a tight loop:
    // does a tcgen05.ld next to / in the same half as / in the other half of the accumulator slow the MMAs down?
    printf("[tc_microbench] N=96 MMAs into columns 0..95 with concurrent tcgen05.ld (3 warps x 32 columns) at column:\n");
    const int ld_cols[] = {-1, 0, 96, 128, 224, 256, 384, 480};
    for (int lc : ld_cols) {
        MbParams p{96, 1, 0, iters, 8, 160, 8, 2880, 1536, lc};
        lea_tc_microbench_kernel<<<grid, 128, 64 * 1024, (cudaStream_t)stream>>>(p, d_out);
        cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream);
        if (e != cudaSuccess) { lea_set_error("microbench: %s", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h.data(), d_out, 2 * grid * sizeof(long long), cudaMemcpyDeviceToHost);
        double avg = 0;
        for (int k = 0; k < grid; ++k) avg += (double)h[2 * k];
        printf("[tc_microbench]   ld_col %4d : %7.1f cycles per MMA\n", lc, avg / grid / iters);
    }
    fflush(stdout);
    cudaFree(d_out);
    return 0;
}
