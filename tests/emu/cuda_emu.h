// Minimal CPU emulation of the CUDA execution model for the SIMT kernels - TEST INFRASTRUCTURE ONLY.
// One std::thread per CUDA thread of a block, blocks run one after another, __syncthreads() is a std::barrier.
// It exists so that tests/ can run the unmodified kernel source of leastereo_b200/csrc/lea_simt_kernels.cuh in a
// container with no GPU and compare it with the oracle.  The leastereo_b200 package never loads this.
#pragma once
#include <barrier>
#include <functional>
#include <thread>
#include <vector>
#include <algorithm>
#include <cstdlib>
#include <cmath>

struct dim3 { unsigned x, y, z; dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {} };
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
static inline float4 make_float4(float a, float b, float c, float d) { float4 r{a, b, c, d}; return r; }

extern thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;
extern thread_local std::barrier<>* emu_barrier;
extern unsigned char* emu_dyn_smem;

#define __global__
#define __device__
#define __host__
#define __shared__ static
#define __align__(n) alignas(n)
#define __restrict__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define LEA_HD inline
#define LEA_D inline
#define __ldg(p) (*(p))
#define __expf(x) expf(x)
using std::exp2f;
static inline void __syncthreads() { emu_barrier->arrive_and_wait(); }
#include <atomic>
static inline float atomicAdd(float* p, float v) { return std::atomic_ref<float>(*p).fetch_add(v, std::memory_order_relaxed); }
static inline double atomicAdd(double* p, double v) { return std::atomic_ref<double>(*p).fetch_add(v, std::memory_order_relaxed); }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) {
    return std::atomic_ref<unsigned long long>(*p).fetch_add(v, std::memory_order_relaxed);
}
// warp shuffle for the emulator: every thread of the block must call it (the kernels that use it have no early exit);
// values are exchanged through a per-block scratch array between two block-wide barriers
extern float emu_shfl_scratch[1024];
static inline float __shfl_xor_sync(unsigned, float v, int lane_mask) {
    const unsigned t = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
    emu_shfl_scratch[t] = v;
    emu_barrier->arrive_and_wait();
    const float r = emu_shfl_scratch[t ^ (unsigned)lane_mask];
    emu_barrier->arrive_and_wait();
    return r;
}
using std::min;
using std::max;

void emu_launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);
#define LEA_LAUNCH(kernel, grid, block, smem, stream, ...) \
    emu_launch(dim3(grid), dim3(block), (smem), [=]() { kernel(__VA_ARGS__); })
#define LEA_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(emu_dyn_smem)
