// CPU-emulated build of the SIMT part of the C ABI (tests only).  g++ -std=c++20 -DLEA_CPU_EMU.
#include "../../leastereo_b200/csrc/lea_common.h"
#include <cstdarg>
#include <cstdio>

thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;
thread_local std::barrier<>* emu_barrier = nullptr;
unsigned char* emu_dyn_smem = nullptr;
float emu_shfl_scratch[1024];

void emu_launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
    const unsigned nthreads = block.x * block.y * block.z;
    std::vector<unsigned char> dyn(smem + 64);
    emu_dyn_smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(dyn.data()) + 63) & ~uintptr_t(63));
    std::barrier<> bar(nthreads);
    auto worker = [&](unsigned t) {
        emu_barrier = &bar;
        blockDim = block; gridDim = grid;
        threadIdx = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
        for (unsigned bz = 0; bz < grid.z; ++bz)
            for (unsigned by = 0; by < grid.y; ++by)
                for (unsigned bx = 0; bx < grid.x; ++bx) {
                    bar.arrive_and_wait();            // previous block fully retired (static __shared__ reuse)
                    blockIdx = dim3(bx, by, bz);
                    body();
                }
    };
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < nthreads; ++t) pool.emplace_back(worker, t);
    worker(0);
    for (auto& th : pool) th.join();
}

static thread_local char g_err[512];
void lea_set_error(const char* fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
}
extern "C" const char* lea_last_error(void) { return g_err; }
extern "C" int lea_abi_version(void) { return LEA_ABI_VERSION; }
extern "C" int lea_is_device_build(void) { return 0; }

#define LEA_POST_LAUNCH() 0
#include "../../leastereo_b200/csrc/lea_simt_kernels.cuh"
#include "../../leastereo_b200/csrc/lea_train_kernels.cuh"
#include "../../leastereo_b200/csrc/lea_io_kernels.cuh"
#include "../../leastereo_b200/csrc/lea_api_simt.inl"
#include "../../leastereo_b200/csrc/lea_api_train.inl"
#include "../../leastereo_b200/csrc/lea_api_io.inl"
