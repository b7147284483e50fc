// Shared host/device helpers: build-mode shim, bf16 plane splitting, planes-volume indexing.
//
// The SIMT kernels in lea_simt_kernels.cuh are written against this header so that the SAME source compiles
//   (a) with nvcc for sm_100a (the product), and
//   (b) with g++ under -DLEA_CPU_EMU (tests/emu: a thread-per-CUDA-thread emulator used ONLY by the no-GPU unit
//       tests to check index math before GPU time is spent; never loaded by the leastereo_b200 package).
#pragma once

#include <stdint.h>
#include <string.h>
#include <math.h>

#include "../../include/leastereo_b200.h"

#ifdef LEA_CPU_EMU
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#define LEA_LAUNCH(kernel, grid, block, smem, stream, ...) \
    kernel<<<(grid), (block), (smem), (cudaStream_t)(stream)>>>(__VA_ARGS__)
#define LEA_HD __host__ __device__ __forceinline__
#define LEA_D __device__ __forceinline__
#endif

// ---------------------------------------------------------------------------------------------------------
// bf16 <-> fp32 by bit manipulation (identical on host, device and emulator; round-to-nearest-even)
// ---------------------------------------------------------------------------------------------------------
LEA_HD uint32_t lea_f32_bits(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
LEA_HD float lea_bits_f32(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}
LEA_HD uint16_t lea_f32_to_bf16(float f) {
    uint32_t u = lea_f32_bits(f);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);   // NaN stays NaN
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}
LEA_HD float lea_bf16_to_f32(uint16_t h) { return lea_bits_f32(((uint32_t)h) << 16); }

// Split x into up to 3 bf16 planes with x ~= p0 + p1 + p2 (P=3 is exact for normal fp32 values).
LEA_HD void lea_split_planes(float x, int P, uint16_t* out /*[3]*/) {
    uint16_t h0 = lea_f32_to_bf16(x);
    out[0] = h0; out[1] = 0; out[2] = 0;
    if (P < 2) return;
    if ((h0 & 0x7f80u) == 0x7f80u) return;                     // inf / NaN: no residual planes
    float r1 = x - lea_bf16_to_f32(h0);
    uint16_t h1 = lea_f32_to_bf16(r1);
    out[1] = h1;
    if (P < 3) return;
    float r2 = r1 - lea_bf16_to_f32(h1);
    out[2] = lea_f32_to_bf16(r2);
}

// 8 bf16 values = one 16-byte group
struct __attribute__((aligned(16))) lea_u4 { uint32_t x, y, z, w; };

LEA_HD void lea_unpack8(const lea_u4& v, float* f /*[8]*/, bool accumulate) {
    const uint32_t q[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float lo = lea_bits_f32(q[i] << 16);
        float hi = lea_bits_f32(q[i] & 0xffff0000u);
        if (accumulate) { f[2 * i] += lo; f[2 * i + 1] += hi; }
        else            { f[2 * i] = lo;  f[2 * i + 1] = hi; }
    }
}

// ---------------------------------------------------------------------------------------------------------
// planes-volume indexing (see include/leastereo_b200.h)
// ---------------------------------------------------------------------------------------------------------
// index, in 16-byte groups, of (b, channel-block cb, plane p, d, h, w)
LEA_HD int64_t lea_vol_group(const lea_vol& v, int b, int cb, int p, int d, int h, int w) {
    return ((((((int64_t)b * (v.C >> 3) + cb) * v.P + p) * v.D + d) * v.H + h) * (int64_t)v.W + w);
}
LEA_HD int64_t lea_vol_plane_stride(const lea_vol& v) { return (int64_t)v.D * v.H * v.W; }   // in groups

// read 8 channels of one voxel as fp32 (sum of planes)
LEA_HD void lea_vol_load8(const lea_vol& v, int b, int cb, int d, int h, int w, float* f) {
    const lea_u4* base = (const lea_u4*)v.data;
    int64_t g = lea_vol_group(v, b, cb, 0, d, h, w);
    const int64_t ps = lea_vol_plane_stride(v);
    lea_unpack8(base[g], f, false);
    for (int p = 1; p < v.P; ++p) lea_unpack8(base[g + p * ps], f, true);
}
// write 8 fp32 values as P bf16 planes: plane p goes to g0[p * ps]
LEA_HD void lea_store8_at(lea_u4* g0, int64_t ps, int P, const float* f) {
    uint32_t q[3][4];
#if defined(__CUDA_ARCH__)
    // device: hardware round-to-nearest-even packing (cvt.rn.bf16x2.f32); same bits as the software path below for
    // finite inputs
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float x0 = f[2 * i], x1 = f[2 * i + 1];
#pragma unroll
        for (int p = 0; p < 3; ++p) {
            uint32_t hq = 0;
            if (p < P) {
                asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hq) : "f"(x1), "f"(x0));
                x0 -= __uint_as_float(hq << 16);
                x1 -= __uint_as_float(hq & 0xffff0000u);
            }
            q[p][i] = hq;
        }
    }
#else
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        uint16_t a[3], c[3];
        lea_split_planes(f[2 * i], P, a);
        lea_split_planes(f[2 * i + 1], P, c);
#pragma unroll
        for (int p = 0; p < 3; ++p) q[p][i] = (uint32_t)a[p] | ((uint32_t)c[p] << 16);
    }
#endif
#pragma unroll
    for (int p = 0; p < 3; ++p) {
        if (p < P) {
            lea_u4 o; o.x = q[p][0]; o.y = q[p][1]; o.z = q[p][2]; o.w = q[p][3];
            g0[p * ps] = o;
        }
    }
}
// write 8 channels of one voxel, splitting into the volume's planes
LEA_HD void lea_vol_store8(const lea_vol& v, int b, int cb, int d, int h, int w, const float* f) {
    lea_store8_at((lea_u4*)v.data + lea_vol_group(v, b, cb, 0, d, h, w), lea_vol_plane_stride(v), v.P, f);
}
// read 8 fp32 values stored as P planes at g0[p * ps]
LEA_HD void lea_load8_at(const lea_u4* g0, int64_t ps, int P, float* f) {
    lea_unpack8(g0[0], f, false);
#pragma unroll
    for (int p = 1; p < 3; ++p)
        if (p < P) lea_unpack8(g0[p * ps], f, true);
}

// Collapsed stem0 (see lea_stem0_assemble): output voxels (d, w in the 8-wide tile tw) whose 3x3x3 window lies entirely
// inside the un-masked part of the cost volume (w' >= d' for every tap), off the first/last depth slice and off the
// first/last column.  Tile-granular along w so that the two kernels that share the volume write disjoint regions.
LEA_HD bool lea_cv_interior(int d, int tw, int D, int W) {
    return d >= 1 && d <= D - 2 && 8 * tw >= d + 2 && 8 * tw + 7 <= W - 2;
}
// ... and the voxels whose whole window is masked (every tap has w' < d'): the conv output is exactly 0 there, so
// stem0's output is relu(bn(0)).  Also written by lea_stem0_assemble, also skipped by the tensor-core launch.
LEA_HD bool lea_cv_masked(int d, int tw) { return 8 * tw + 7 <= d - 3; }
LEA_HD bool lea_cv_collapsed(int d, int tw, int D, int W) { return lea_cv_interior(d, tw, D, W) || lea_cv_masked(d, tw); }

// error reporting shared by every API translation unit
void lea_set_error(const char* fmt, ...);
#define LEA_CHECK(cond, ...) do { if (!(cond)) { lea_set_error(__VA_ARGS__); return 1; } } while (0)
