// CUDA-core kernels of the LEAStereo hot path: cost-volume gather, planes pack/unpack, trilinear resample
// (align_corners=True), fp32-FMA ConvBR (k = 1, 3), fused disparity head, disparity regression.
// Reference arithmetic being replaced is cited per kernel.  Compiles with nvcc (sm_100a) and, for the no-GPU
// index-math tests, with g++ -DLEA_CPU_EMU (see lea_common.h).
#pragma once
#include "lea_common.h"

#ifndef LEA_CPU_EMU
#define LEA_DYN_SMEM(type, name) extern __shared__ __align__(16) unsigned char lea_smem_raw_[]; \
                                 type* name = reinterpret_cast<type*>(lea_smem_raw_)
#endif

// =========================================================================================================
// K1  cost volume, reference layout (retrain/LEAStereo.py:34-48), bit-exact copy:
//       cost[b, c,   d, h, w] = x[b, c, h, w]      if w >= d else 0
//       cost[b, C+c, d, h, w] = y[b, c, h, w - d]  if w >= d else 0
//     One thread produces 4 consecutive w of one (b, c2, d) plane and stores them as one 128-bit word.
//     HBM-bound on the write: 4*(2C*H*W + 2C*D3*H*W) algorithmic bytes per pair.
// =========================================================================================================
__global__ void __launch_bounds__(256)
lea_cost_volume_f32_kernel(const float* __restrict__ x, const float* __restrict__ y, float* __restrict__ cost,
                           int C, int H, int W, int D3, int hw4 /* ceil(H*W/4) when W%4==0 */, int chunks) {
    // grid: (chunks, 2C*D3, B); each block covers 1024 float4 of one plane
    const int plane = blockIdx.y;               // c2 * D3 + d
    const int c2 = plane / D3, d = plane - c2 * D3;
    const int b = blockIdx.z;
    const bool left = c2 < C;
    const float* __restrict__ src = (left ? x : y) + ((int64_t)b * C + (left ? c2 : c2 - C)) * H * W;
    float* __restrict__ dst = cost + (((int64_t)b * 2 * C + c2) * D3 + d) * (int64_t)H * W;
    const int shift = left ? 0 : d;
    const int w4n = W >> 2;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const int q = (blockIdx.x * 4 + it) * 256 + threadIdx.x;       // float4 index inside the plane
        if (q >= hw4) break;
        const int h = q / w4n;
        const int w = (q - h * w4n) << 2;
        const float* __restrict__ row = src + (int64_t)h * W;
        float4 v;
        v.x = (w + 0 >= d) ? __ldg(row + (w + 0 - shift)) : 0.0f;
        v.y = (w + 1 >= d) ? __ldg(row + (w + 1 - shift)) : 0.0f;
        v.z = (w + 2 >= d) ? __ldg(row + (w + 2 - shift)) : 0.0f;
        v.w = (w + 3 >= d) ? __ldg(row + (w + 3 - shift)) : 0.0f;
        *reinterpret_cast<float4*>(dst + (int64_t)h * W + w) = v;
    }
    (void)chunks;
}

// scalar variant for W % 4 != 0 (ragged widths); one thread per output element
__global__ void __launch_bounds__(256)
lea_cost_volume_f32_scalar_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                  float* __restrict__ cost, int C, int H, int W, int D3) {
    const int plane = blockIdx.y;
    const int c2 = plane / D3, d = plane - c2 * D3;
    const int b = blockIdx.z;
    const bool left = c2 < C;
    const float* __restrict__ src = (left ? x : y) + ((int64_t)b * C + (left ? c2 : c2 - C)) * H * W;
    float* __restrict__ dst = cost + (((int64_t)b * 2 * C + c2) * D3 + d) * (int64_t)H * W;
    const int shift = left ? 0 : d;
    const int q = blockIdx.x * 256 + threadIdx.x;
    if (q >= H * W) return;
    const int h = q / W, w = q - h * W;
    dst[q] = (w >= d) ? __ldg(src + (int64_t)h * W + (w - shift)) : 0.0f;
}

// K1b  same volume written in the planes layout stem0 consumes.  One thread = one (b, cb, d, h, w) 8-channel group.
__global__ void __launch_bounds__(256)
lea_cost_volume_planes_kernel(const float* __restrict__ x, const float* __restrict__ y, lea_vol vol, int C) {
    const int W = vol.W, H = vol.H;
    const int w = blockIdx.x * 256 + threadIdx.x;
    if (w >= W) return;
    const int h = blockIdx.y % H;
    const int d = blockIdx.y / H;
    const int cbn = vol.C >> 3;
    const int b = blockIdx.z / cbn, cb = blockIdx.z - b * cbn;
    const bool left = cb < (C >> 3);
    const int c0 = (left ? cb : cb - (C >> 3)) * 8;
    const float* __restrict__ src = (left ? x : y) + (((int64_t)b * C + c0) * H + h) * W;
    const int sw = left ? w : w - d;
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = (w >= d) ? __ldg(src + (int64_t)j * H * W + sw) : 0.0f;
    lea_vol_store8(vol, b, cb, d, h, w, f);
}

// =========================================================================================================
// planes pack / unpack: fp32 (B, c, D, H, W) <-> channel slice [c0, c0+c) of a planes volume
// =========================================================================================================
__global__ void __launch_bounds__(256)
lea_pack_planes_kernel(const float* __restrict__ src, lea_vol dst, int dst_c0, int c) {
    const int w = blockIdx.x * 256 + threadIdx.x;
    if (w >= dst.W) return;
    const int h = blockIdx.y % dst.H, d = blockIdx.y / dst.H;
    const int cbn = c >> 3;
    const int b = blockIdx.z / cbn, cb = blockIdx.z - b * cbn;
    const int64_t sp = (int64_t)dst.D * dst.H * dst.W;
    const float* __restrict__ s = src + ((int64_t)b * c + cb * 8) * sp + ((int64_t)d * dst.H + h) * dst.W + w;
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = __ldg(s + j * sp);
    lea_vol_store8(dst, b, (dst_c0 >> 3) + cb, d, h, w, f);
}

__global__ void __launch_bounds__(256)
lea_unpack_planes_kernel(lea_vol src, int src_c0, int c, float* __restrict__ dst) {
    const int w = blockIdx.x * 256 + threadIdx.x;
    if (w >= src.W) return;
    const int h = blockIdx.y % src.H, d = blockIdx.y / src.H;
    const int cbn = c >> 3;
    const int b = blockIdx.z / cbn, cb = blockIdx.z - b * cbn;
    const int64_t sp = (int64_t)src.D * src.H * src.W;
    float f[8];
    lea_vol_load8(src, b, (src_c0 >> 3) + cb, d, h, w, f);
    float* __restrict__ o = dst + ((int64_t)b * c + cb * 8) * sp + ((int64_t)d * src.H + h) * src.W + w;
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j * sp] = f[j];
}

// =========================================================================================================
// K5  trilinear resample, align_corners=True (retrain/skip_model_3d.py:44-51, :162-169).
//     PyTorch index rule restated: scale = (in-1)/(out-1) (0 if out == 1); src = scale*dst; i0 = min(floor(src), in-1);
//     l1 = clamp(src - i0, 0, 1); i1 = i0 + (i0 < in-1); l0 = 1 - l1.  Equal sizes copy.
// =========================================================================================================
struct lea_axis_lerp { int i0, i1; float l0, l1; };

LEA_HD lea_axis_lerp lea_axis_ac(int dst, int in_n, int out_n) {
    lea_axis_lerp r;
    if (in_n == out_n) { r.i0 = r.i1 = dst; r.l0 = 1.0f; r.l1 = 0.0f; return r; }
    const float scale = out_n > 1 ? (float)(in_n - 1) / (float)(out_n - 1) : 0.0f;
    const float src = scale * (float)dst;
    int i0 = (int)floorf(src);
    if (i0 > in_n - 1) i0 = in_n - 1;
    float l1 = src - (float)i0;
    l1 = l1 < 0.0f ? 0.0f : (l1 > 1.0f ? 1.0f : l1);
    r.i0 = i0; r.i1 = i0 + (i0 < in_n - 1 ? 1 : 0); r.l0 = 1.0f - l1; r.l1 = l1;
    return r;
}

__global__ void __launch_bounds__(128)
lea_trilinear_ac_kernel(lea_vol src, int src_c0, lea_vol dst, int dst_c0, int c,
                        const float* __restrict__ bn_scale, const float* __restrict__ bn_shift, int relu) {
    const int w = blockIdx.x * 128 + threadIdx.x;
    if (w >= dst.W) return;
    const int h = blockIdx.y % dst.H, d = blockIdx.y / dst.H;
    const int cbn = c >> 3;
    const int b = blockIdx.z / cbn, cb = blockIdx.z - b * cbn;
    const lea_axis_lerp ad = lea_axis_ac(d, src.D, dst.D);
    const lea_axis_lerp ah = lea_axis_ac(h, src.H, dst.H);
    const lea_axis_lerp aw = lea_axis_ac(w, src.W, dst.W);
    // one base pointer + small offsets instead of a 64-bit index chain per load (the kernel is instruction-bound)
    const int64_t sHW = (int64_t)src.H * src.W, sPS = sHW * src.D;
    const lea_u4* sbase = (const lea_u4*)src.data + ((int64_t)b * (src.C >> 3) + (src_c0 >> 3) + cb) * src.P * sPS;
    float out[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) out[j] = 0.0f;
#pragma unroll
    for (int zd = 0; zd < 2; ++zd) {
        const float wd = zd ? ad.l1 : ad.l0;
        const lea_u4* sd = sbase + (int64_t)(zd ? ad.i1 : ad.i0) * sHW;
        float accd[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) accd[j] = 0.0f;
#pragma unroll
        for (int zh = 0; zh < 2; ++zh) {
            const float wh = zh ? ah.l1 : ah.l0;
            const lea_u4* sh = sd + (zh ? ah.i1 : ah.i0) * src.W;
            float a[8], bb[8];
            lea_load8_at(sh + aw.i0, sPS, src.P, a);
            lea_load8_at(sh + aw.i1, sPS, src.P, bb);
#pragma unroll
            for (int j = 0; j < 8; ++j) accd[j] += wh * (aw.l0 * a[j] + aw.l1 * bb[j]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) out[j] += wd * accd[j];
    }
    if (bn_scale) {
#pragma unroll
        for (int j = 0; j < 8; ++j) out[j] = out[j] * __ldg(bn_scale + cb * 8 + j) + __ldg(bn_shift + cb * 8 + j);
    }
    if (relu) {
#pragma unroll
        for (int j = 0; j < 8; ++j) out[j] = out[j] > 0.0f ? out[j] : 0.0f;
    }
    const int64_t dHW = (int64_t)dst.H * dst.W, dPS = dHW * dst.D;
    lea_store8_at((lea_u4*)dst.data + ((int64_t)b * (dst.C >> 3) + (dst_c0 >> 3) + cb) * dst.P * dPS + d * dHW +
                  (int64_t)h * dst.W + w, dPS, dst.P, out);
}

// K5c  down-sampling resample fused with the 1x1x1 ConvBR(s) that consume it (retrain/skip_model_3d.py:44-53: a cell
//      resamples s0 / s1 to its own size, then applies pre_preprocess / preprocess = Conv3d 1x1x1 + BN + ReLU).  The
//      resampled tensor is often needed by two cells (s1 of cell i = s0 of cell i+1) with different 1x1x1 convs; one pass
//      over the big source volume interpolates each 8-channel group once and feeds it to every consumer's matrix:
//      the resampled intermediate (and its conv launches) disappear.  One thread per output voxel; weights of all
//      consumers transposed in shared memory [c_in][n_total]; fp32 FMA.
template <int NT>              // NT = total output channels of all consumers, padded to 16 / 32 / 64
__global__ void __launch_bounds__(128)
lea_resample_conv1_kernel(lea_vol src, int src_c0, int c_in, lea_rc_out o0, lea_rc_out o1, int n_out) {
    LEA_DYN_SMEM(float, w_s);                       // [c_in][NT]
    const int tid = threadIdx.x;
    const int n0 = o0.c_out, n1 = n_out > 1 ? o1.c_out : 0;
    for (int e = tid; e < c_in * NT; e += 128) {
        const int ci = e / NT, n = e - ci * NT;
        float wv = 0.0f;
        if (n < n0) wv = __ldg(o0.weight + (int64_t)n * c_in + ci);
        else if (n < n0 + n1) wv = __ldg(o1.weight + (int64_t)(n - n0) * c_in + ci);
        w_s[e] = wv;
    }
    __syncthreads();
    const lea_vol& dst = o0.dst;                    // all consumers share the target geometry
    const int w = blockIdx.x * 128 + tid;
    if (w >= dst.W) return;
    const int h = blockIdx.y % dst.H, d = blockIdx.y / dst.H, b = blockIdx.z;
    const lea_axis_lerp ad = lea_axis_ac(d, src.D, dst.D);
    const lea_axis_lerp ah = lea_axis_ac(h, src.H, dst.H);
    const lea_axis_lerp aw = lea_axis_ac(w, src.W, dst.W);
    const int64_t sHW = (int64_t)src.H * src.W, sPS = sHW * src.D;
    const lea_u4* sb0 = (const lea_u4*)src.data + ((int64_t)b * (src.C >> 3) + (src_c0 >> 3)) * src.P * sPS;
    float acc[NT];
#pragma unroll
    for (int n = 0; n < NT; ++n) acc[n] = 0.0f;
    for (int cb = 0; cb < (c_in >> 3); ++cb) {
        const lea_u4* sbase = sb0 + (int64_t)cb * src.P * sPS;
        float val[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) val[j] = 0.0f;
#pragma unroll
        for (int zd = 0; zd < 2; ++zd) {
            const float wd = zd ? ad.l1 : ad.l0;
            const lea_u4* sd = sbase + (int64_t)(zd ? ad.i1 : ad.i0) * sHW;
            float accd[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) accd[j] = 0.0f;
#pragma unroll
            for (int zh = 0; zh < 2; ++zh) {
                const float wh = zh ? ah.l1 : ah.l0;
                const lea_u4* sh = sd + (zh ? ah.i1 : ah.i0) * src.W;
                float a[8], bb[8];
                lea_load8_at(sh + aw.i0, sPS, src.P, a);
                lea_load8_at(sh + aw.i1, sPS, src.P, bb);
#pragma unroll
                for (int j = 0; j < 8; ++j) accd[j] += wh * (aw.l0 * a[j] + aw.l1 * bb[j]);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) val[j] += wd * accd[j];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4* wr = reinterpret_cast<const float4*>(w_s + (cb * 8 + j) * NT);
#pragma unroll
            for (int n4 = 0; n4 < NT / 4; ++n4) {
                const float4 wv = wr[n4];
                acc[n4 * 4 + 0] += val[j] * wv.x; acc[n4 * 4 + 1] += val[j] * wv.y;
                acc[n4 * 4 + 2] += val[j] * wv.z; acc[n4 * 4 + 3] += val[j] * wv.w;
            }
        }
    }
    const int64_t dHW = (int64_t)dst.H * dst.W, dPS = dHW * dst.D;
#pragma unroll
    for (int g8 = 0; g8 < NT / 8; ++g8) {
        const int n = g8 * 8;
        if (n >= n0 + n1) break;
        const bool first = n < n0;                   // consumer channel counts are multiples of 8: a group never straddles
        const lea_rc_out& o = first ? o0 : o1;
        const int c = first ? n : n - n0;
        float out[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float v = acc[n + j];
            if (o.bn_scale) v = v * __ldg(o.bn_scale + c + j) + __ldg(o.bn_shift + c + j);
            if (o.relu) v = v > 0.0f ? v : 0.0f;
            out[j] = v;
        }
        lea_store8_at((lea_u4*)o.dst.data + ((int64_t)b * (o.dst.C >> 3) + ((o.dst_c0 + c) >> 3)) * o.dst.P * dPS + d * dHW +
                      (int64_t)h * dst.W + w, dPS, o.dst.P, out);
    }
}

// Up-sampling variant: one thread owns an output column (h, w) of a chunk of LEA_UP_DCH depths and marches along d.
// The (h, w)-interpolated values of the two low-res depth slices in use stay in registers; a new slice costs 4 corner
// loads, and with out/in ~ 2 along d only every second output needs one -> ~5 loads per output instead of 16.
// Same blend order (w, then h, then d) as the kernel above.
#define LEA_UP_DCH 16
struct lea_up_ctx {
    const lea_u4* sbase;          // (b, channel block, plane 0, depth 0) of the source
    int64_t sHW, sPS;             // source slice / plane strides in 16-byte groups
    int o00, o01, o10, o11;       // the four (h, w) corners inside a slice
    float wh0, wh1, ww0, ww1;
    int P;
};
// PP = plane count known at compile time (0: taken from the volume at run time).  The run-time form predicates every
// plane access (a third of the kernel's instructions, which is what bounds it); the launcher picks PP = 2 / 3 instances.
template <int PP>
LEA_D void lea_up_slice(const lea_up_ctx& c, int id, float* o /*[8]*/) {
    const lea_u4* sp = c.sbase + (int64_t)id * c.sHW;
    const int P = PP ? PP : c.P;
    float a[8], bb[8], e[8], g[8];
    lea_load8_at(sp + c.o00, c.sPS, P, a);
    lea_load8_at(sp + c.o01, c.sPS, P, bb);
    lea_load8_at(sp + c.o10, c.sPS, P, e);
    lea_load8_at(sp + c.o11, c.sPS, P, g);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float v = 0.0f;
        v += c.wh0 * (c.ww0 * a[j] + c.ww1 * bb[j]);
        v += c.wh1 * (c.ww0 * e[j] + c.ww1 * g[j]);
        o[j] = v;
    }
}

template <int PP>
__global__ void __launch_bounds__(128)
lea_trilinear_ac_up_kernel(lea_vol src, int src_c0, lea_vol dst, int dst_c0, int c,
                           const float* __restrict__ bn_scale, const float* __restrict__ bn_shift, int relu) {
    const int w = blockIdx.x * 128 + threadIdx.x;
    if (w >= dst.W) return;
    const int h = blockIdx.y % dst.H, dch = blockIdx.y / dst.H;
    const int cbn = c >> 3;
    const int b = blockIdx.z / cbn, cb = blockIdx.z - b * cbn;
    const lea_axis_lerp ah = lea_axis_ac(h, src.H, dst.H);
    const lea_axis_lerp aw = lea_axis_ac(w, src.W, dst.W);
    lea_up_ctx cx;
    cx.sHW = (int64_t)src.H * src.W;
    cx.sPS = cx.sHW * src.D;
    cx.P = src.P;
    cx.sbase = (const lea_u4*)src.data + ((int64_t)b * (src.C >> 3) + (src_c0 >> 3) + cb) * src.P * cx.sPS;
    cx.o00 = ah.i0 * src.W + aw.i0; cx.o01 = ah.i0 * src.W + aw.i1;
    cx.o10 = ah.i1 * src.W + aw.i0; cx.o11 = ah.i1 * src.W + aw.i1;
    cx.wh0 = ah.l0; cx.wh1 = ah.l1; cx.ww0 = aw.l0; cx.ww1 = aw.l1;
    const int64_t dHW = (int64_t)dst.H * dst.W, dPS = dHW * dst.D;
    lea_u4* dbase = (lea_u4*)dst.data + ((int64_t)b * (dst.C >> 3) + (dst_c0 >> 3) + cb) * dst.P * dPS +
                    (int64_t)h * dst.W + w;
    float sc[8], sh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        sc[j] = bn_scale ? __ldg(bn_scale + cb * 8 + j) : 1.0f;
        sh[j] = bn_scale ? __ldg(bn_shift + cb * 8 + j) : 0.0f;
    }
    const float dscale = dst.D > 1 ? (float)(src.D - 1) / (float)(dst.D - 1) : 0.0f;      // as in lea_axis_ac
    float A[8], Bv[8];
    int cur0 = -1, cur1 = -1;
    const int d_end = min(dst.D, (dch + 1) * LEA_UP_DCH);
    for (int d = dch * LEA_UP_DCH; d < d_end; ++d) {
        const float sd = dscale * (float)d;
        int i0 = (int)floorf(sd);
        if (i0 > src.D - 1) i0 = src.D - 1;
        float l1 = sd - (float)i0;
        l1 = l1 < 0.0f ? 0.0f : (l1 > 1.0f ? 1.0f : l1);
        const int i1 = i0 + (i0 < src.D - 1 ? 1 : 0);
        const float l0 = 1.0f - l1;
        if (i0 != cur0) {
            if (i0 == cur1) {
#pragma unroll
                for (int j = 0; j < 8; ++j) A[j] = Bv[j];
            } else {
                lea_up_slice<PP>(cx, i0, A);
            }
            cur0 = i0;
        }
        if (i1 != cur1) {
            if (i1 == cur0) {
#pragma unroll
                for (int j = 0; j < 8; ++j) Bv[j] = A[j];
            } else {
                lea_up_slice<PP>(cx, i1, Bv);
            }
            cur1 = i1;
        }
        float out[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float v = 0.0f;
            v += l0 * A[j];
            v += l1 * Bv[j];
            if (bn_scale) v = v * sc[j] + sh[j];
            if (relu) v = v > 0.0f ? v : 0.0f;
            out[j] = v;
        }
        lea_store8_at(dbase + (int64_t)d * dHW, dPS, PP ? PP : dst.P, out);
    }
}

// =========================================================================================================
// K2s  ConvBR on CUDA cores, fp32 accumulate (models/operations_3d.py:31-47).  The exact-arithmetic conv of the
//      engine ("simt" mode), the checker the tcgen05 kernel is validated against on the GPU, and the kernel for
//      shapes tensor cores do not take.  Tile = 8 (w) x 16 (h) output voxels of one depth slice per 128-thread CTA;
//      each thread owns one voxel and all NPAD output channels.
// =========================================================================================================
#define LEA_TW 8
#define LEA_TH 16

template <int NPAD>
LEA_D void lea_conv_epilogue(const lea_conv& p, int b, int d, int h, int w, float* acc) {
    if (h >= p.src.H || w >= p.src.W) return;
#pragma unroll
    for (int n = 0; n < NPAD; ++n) {
        if (n < p.c_out) {
            float v = acc[n];
            if (p.bn_scale) v = v * __ldg(p.bn_scale + n) + __ldg(p.bn_shift + n);
            if (p.relu) v = v > 0.0f ? v : 0.0f;
            acc[n] = v;
        }
    }
    if (p.dst_f32) {
        const int64_t sp = (int64_t)p.src.D * p.src.H * p.src.W;
        float* o = p.dst_f32 + (int64_t)b * p.c_out * sp + ((int64_t)d * p.src.H + h) * p.src.W + w;
        for (int n = 0; n < p.c_out; ++n) o[n * sp] = acc[n];
        return;
    }
#pragma unroll
    for (int cb = 0; cb < NPAD / 8; ++cb) {
        if (cb * 8 < p.c_out) {
            if (p.has_res) {
                float r[8];
                lea_vol_load8(p.res, b, (p.res_c0 >> 3) + cb, d, h, w, r);
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[cb * 8 + j] += r[j];
            }
            lea_vol_store8(p.dst, b, (p.dst_c0 >> 3) + cb, d, h, w, acc + cb * 8);
        }
    }
}

template <int NPAD>
__global__ void __launch_bounds__(128)
lea_conv3_simt_kernel(lea_conv p, const float* __restrict__ weight) {
    // smem: in_s[8][3][TH+2][TW+2] floats, then w_s[8][27][NPAD] floats
    LEA_DYN_SMEM(float, smem);
    constexpr int HH = LEA_TH + 2, WW = LEA_TW + 2, SLAB = HH * WW, HALO = 3 * SLAB;
    float* in_s = smem;
    float* w_s = smem + 8 * HALO;
    const int tiles_w = (p.src.W + LEA_TW - 1) / LEA_TW;
    const int tw = blockIdx.x % tiles_w, th = blockIdx.x / tiles_w;
    const int d = blockIdx.y, b = blockIdx.z;
    const int w0 = tw * LEA_TW, h0 = th * LEA_TH;
    const int tid = threadIdx.x;
    const int lw = tid % LEA_TW, lh = tid / LEA_TW;
    float acc[NPAD];
#pragma unroll
    for (int n = 0; n < NPAD; ++n) acc[n] = 0.0f;

    const int ncb = p.c_in >> 3;
    for (int cb = 0; cb < ncb; ++cb) {
        // ---- stage the input halo (8 channels, 3 depth slabs) as fp32 ----
        for (int v = tid; v < HALO; v += 128) {
            const int kd = v / SLAB, r = v - kd * SLAB;
            const int hh = r / WW, ww = r - hh * WW;
            const int gd = d + kd - 1, gh = h0 + hh - 1, gw = w0 + ww - 1;
            float f[8];
            if (gd >= 0 && gd < p.src.D && gh >= 0 && gh < p.src.H && gw >= 0 && gw < p.src.W) {
                lea_vol_load8(p.src, b, (p.src_c0 >> 3) + cb, gd, gh, gw, f);
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = 0.0f;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) in_s[j * HALO + v] = f[j];
        }
        // ---- stage the weights of these 8 input channels: w_s[(j*27+tap)*NPAD + n] ----
        for (int e = tid; e < 216 * NPAD; e += 128) {
            const int n = e / 216, r = e - n * 216;          // r = j*27 + tap, contiguous in the PyTorch layout
            float wv = 0.0f;
            if (n < p.c_out) wv = __ldg(weight + ((int64_t)n * p.c_in + cb * 8) * 27 + r);
            w_s[r * NPAD + n] = wv;
        }
        __syncthreads();
#pragma unroll 1
        for (int j = 0; j < 8; ++j) {
#pragma unroll 1
            for (int kd = 0; kd < 3; ++kd) {
#pragma unroll
                for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw) {
                        const float a = in_s[j * HALO + kd * SLAB + (lh + kh) * WW + (lw + kw)];
                        const float4* wr = reinterpret_cast<const float4*>(
                            w_s + (j * 27 + kd * 9 + kh * 3 + kw) * NPAD);
#pragma unroll
                        for (int n4 = 0; n4 < NPAD / 4; ++n4) {
                            const float4 wv = wr[n4];
                            acc[n4 * 4 + 0] += a * wv.x;
                            acc[n4 * 4 + 1] += a * wv.y;
                            acc[n4 * 4 + 2] += a * wv.z;
                            acc[n4 * 4 + 3] += a * wv.w;
                        }
                    }
                }
            }
        }
        __syncthreads();
    }
    lea_conv_epilogue<NPAD>(p, b, d, h0 + lh, w0 + lw, acc);
}

template <int NPAD>
__global__ void __launch_bounds__(128)
lea_conv1_simt_kernel(lea_conv p, const float* __restrict__ weight) {
    // smem: w_s[c_in][NPAD]
    LEA_DYN_SMEM(float, w_s);
    const int tid = threadIdx.x;
    for (int e = tid; e < p.c_in * NPAD; e += 128) {
        const int n = e / p.c_in, ci = e - n * p.c_in;
        w_s[ci * NPAD + n] = (n < p.c_out) ? __ldg(weight + (int64_t)n * p.c_in + ci) : 0.0f;
    }
    __syncthreads();
    const int tiles_w = (p.src.W + LEA_TW - 1) / LEA_TW;
    const int tw = blockIdx.x % tiles_w, th = blockIdx.x / tiles_w;
    const int d = blockIdx.y, b = blockIdx.z;
    const int w = tw * LEA_TW + tid % LEA_TW, h = th * LEA_TH + tid / LEA_TW;
    if (h >= p.src.H || w >= p.src.W) return;
    float acc[NPAD];
#pragma unroll
    for (int n = 0; n < NPAD; ++n) acc[n] = 0.0f;
    const int ncb = p.c_in >> 3;
    for (int cb = 0; cb < ncb; ++cb) {
        float f[8];
        lea_vol_load8(p.src, b, (p.src_c0 >> 3) + cb, d, h, w, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4* wr = reinterpret_cast<const float4*>(w_s + (cb * 8 + j) * NPAD);
#pragma unroll
            for (int n4 = 0; n4 < NPAD / 4; ++n4) {
                const float4 wv = wr[n4];
                acc[n4 * 4 + 0] += f[j] * wv.x;
                acc[n4 * 4 + 1] += f[j] * wv.y;
                acc[n4 * 4 + 2] += f[j] * wv.z;
                acc[n4 * 4 + 3] += f[j] * wv.w;
            }
        }
    }
    lea_conv_epilogue<NPAD>(p, b, d, h, w, acc);
}

// =========================================================================================================
// K6  fused disparity head (models/build_model_2d.py:45-57 + :27-42):
//       F.interpolate(x, [maxdisp, 3*H3, 3*W3], 'trilinear', align_corners=False) -> Softmin(dim=1) -> sum_d p*d
//     in one kernel; the maxdisp x 3H3 x 3W3 probability volume is never written.  align_corners=False index rule:
//     scale = in/out; src = max(scale*(dst+0.5)-0.5, 0); i0 = floor(src) (<= in-1); l1 = src-i0; i1 = min(i0+1, in-1).
//     One thread owns one low-res cell (h3, w3) and produces its 3x3 output pixels: with the exact x3 scale every one
//     of them blends only the 3x3 low-res neighbourhood, so per disparity sample k the thread loads 9 values
//     (coalesced along w3, neighbours hit L1) and forms the 9 blended logits u_p[k] with per-axis 3-tap weights.
//     (Measured alternatives: thread-per-pixel with a shared-memory tile 168 us, 3 threads per cell 146 us, this 105 us.)
//     Pass 1 finds m_p = min_k u_p[k] (a valid softmin stabiliser: every up-sampled logit is a convex combination of
//     u_p); pass 2 streams the maxdisp samples with a sliding (k0, k1) window.  Cost per pair: 2*9*D3 loads per cell
//     and maxdisp exps per pixel - the kernel is bound by the exp (MUFU) rate, not by HBM.
// =========================================================================================================
LEA_HD lea_axis_lerp lea_axis_half_pixel(int dst, int in_n, int out_n) {
    lea_axis_lerp r;
    // ATen (UpSampleKernel.cpp, area_pixel_compute_source_index) evaluates scale * (dst + 0.5) - 0.5 as ONE fused
    // multiply-add in fp32 (probed: F.interpolate's weights equal this formula bit for bit for every BASELINE size, and
    // differ from the two-rounding form by up to 1.5e-5) - e.g. output 1 of an exact x3 up-sample gets l1 = 1.49e-8, not
    // 0.  With un-normalised logits (~1e8) those weight bits decide the soft-argmin, so they are reproduced exactly.
    const float scale = (float)in_n / (float)out_n;
    float src = fmaf(scale, (float)dst + 0.5f, -0.5f);
    if (src < 0.0f) src = 0.0f;
    int i0 = (int)floorf(src);
    if (i0 > in_n - 1) i0 = in_n - 1;
    float l1 = src - (float)i0;
    l1 = l1 < 0.0f ? 0.0f : (l1 > 1.0f ? 1.0f : l1);
    r.i0 = i0; r.i1 = i0 + (i0 < in_n - 1 ? 1 : 0); r.l0 = 1.0f - l1; r.l1 = l1;
    return r;
}

// weights of output index 3*c + r on the three taps (c-1, c, c+1); returns false if a tap falls outside them
LEA_HD bool lea_three_tap(int c, int r, int in_n, float* t /*[3]*/) {
    const lea_axis_lerp a = lea_axis_half_pixel(3 * c + r, in_n, 3 * in_n);
    const int j0 = a.i0 - c + 1, j1 = a.i1 - c + 1;
#pragma unroll
    for (int j = 0; j < 3; ++j) t[j] = (j0 == j ? a.l0 : 0.0f) + (j1 == j ? a.l1 : 0.0f);   // no dynamic indexing
    return !(j0 < 0 || j0 > 2 || j1 < 0 || j1 > 2);
}

#define LEA_DH_BX 32             // backward kernel (lea_train_kernels.cuh): one thread per cell, 32 x 4 cells per block
#define LEA_DH_BY 4
#define LEA_DH_CELLS 32          // low-res cells per 128-thread block (along w3)
#define LEA_DH_PARTS 4           // lanes per cell: each owns a quarter of the disparity range

// raw 3x3 neighbourhood of a cell at disparity sample k, and the 9 blended logits
//   u[r*3+c] = sum_{i,j} th[r][i] * tw[c][j] * s[k][i][j]
#define LEA_DH_LOAD(raw, k)                                                                           \
    {                                                                                                 \
        const float* __restrict__ pk = mb + (int64_t)(k) * H3 * W3;                                   \
        _Pragma("unroll") for (int i = 0; i < 3; ++i)                                                 \
            _Pragma("unroll") for (int j = 0; j < 3; ++j) raw[i * 3 + j] = __ldg(pk + ro[i] + wo[j]); \
    }
#define LEA_DH_COMBINE(u, raw)                                                                        \
    {                                                                                                 \
        float row[3][3];                                                                              \
        _Pragma("unroll") for (int i = 0; i < 3; ++i)                                                 \
            _Pragma("unroll") for (int c = 0; c < 3; ++c)                                             \
                row[i][c] = tw[c][0] * raw[i * 3] + tw[c][1] * raw[i * 3 + 1] + tw[c][2] * raw[i * 3 + 2]; \
        _Pragma("unroll") for (int r = 0; r < 3; ++r)                                                 \
            _Pragma("unroll") for (int c = 0; c < 3; ++c)                                             \
                u[r * 3 + c] = th[r][0] * row[0][c] + th[r][1] * row[1][c] + th[r][2] * row[2][c];    \
    }

// Four adjacent lanes share one low-res cell (its 3x3 output pixels): each takes a quarter of the disparity range in
// both passes, and the softmin is completed with warp shuffles (min of the stabiliser, sums of numerator/denominator).
// 4x the threads of the one-thread-per-cell formulation: the kernel is latency-bound on its FMA/MUFU chains, so the
// extra warps in flight are what buys time (measured 105 -> see DESIGN.md).  No early exit: every lane takes part in
// the shuffles; lanes past the right edge work on the last column and do not store.
// `cache` != 0: dynamic shared memory holds kchunk x 9 x blockDim floats and pass 1 keeps its blended logits there
// ([k][pixel][thread]: conflict-free), so that pass 2 of the exact x3 path reads them back (9 LDS per sample) instead of
// loading and blending the 3x3 neighbourhood a second time (9 loads + 54 FMAs per sample).
__global__ void __launch_bounds__(LEA_DH_CELLS * LEA_DH_PARTS)
lea_disp_head_kernel(const float* __restrict__ mat, float* __restrict__ disp,
                     int D3, int H3, int W3, int maxdisp, int cache) {
    LEA_DYN_SMEM(float, ucache);
    constexpr int kThreads = LEA_DH_CELLS * LEA_DH_PARTS;
    const int part = threadIdx.x & (LEA_DH_PARTS - 1);
    const int w3r = blockIdx.x * LEA_DH_CELLS + (threadIdx.x >> 2);
    const int h3 = blockIdx.y;
    const int b = blockIdx.z;
    const bool live = w3r < W3;
    const int w3 = live ? w3r : W3 - 1;
    const float* __restrict__ mb = mat + (int64_t)b * D3 * H3 * W3;
    float th[3][3], tw[3][3];
#pragma unroll
    for (int r = 0; r < 3; ++r) { lea_three_tap(h3, r, H3, th[r]); lea_three_tap(w3, r, W3, tw[r]); }
    // clamped neighbour offsets (edge taps carry weight 0)
    int ro[3], wo[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        int hh = h3 - 1 + i, ww = w3 - 1 + i;
        hh = hh < 0 ? 0 : (hh > H3 - 1 ? H3 - 1 : hh);
        ww = ww < 0 ? 0 : (ww > W3 - 1 ? W3 - 1 : ww);
        ro[i] = hh * W3; wo[i] = ww;
    }
    // pass 1: per-pixel minimum of the blended column; this lane's quarter of the samples, then min over the 4 lanes
    float m[9];
    float amax = 0.0f;                       // largest |logit| of the cell: decides whether the x3 shortcut below is safe
#pragma unroll
    for (int q = 0; q < 9; ++q) m[q] = 3.0e38f;
    {
        const int kchunk = (D3 + LEA_DH_PARTS - 1) / LEA_DH_PARTS;
        const int ka = min(D3, part * kchunk), kb = min(D3, ka + kchunk);
        float u[9], cur[9], nxt[9];
        if (ka < kb) LEA_DH_LOAD(cur, ka);
        for (int k = ka; k < kb; ++k) {
            if (k + 1 < kb) LEA_DH_LOAD(nxt, k + 1);
            LEA_DH_COMBINE(u, cur);
#pragma unroll
            for (int q = 0; q < 9; ++q) {
                m[q] = u[q] < m[q] ? u[q] : m[q];
                amax = fmaxf(amax, fabsf(u[q]));
                cur[q] = nxt[q];
                if (cache) ucache[((k - ka) * 9 + q) * kThreads + threadIdx.x] = u[q];
            }
        }
#pragma unroll
        for (int q = 0; q < 9; ++q) {
            float o = __shfl_xor_sync(0xffffffffu, m[q], 1);
            m[q] = o < m[q] ? o : m[q];
            o = __shfl_xor_sync(0xffffffffu, m[q], 2);
            m[q] = o < m[q] ? o : m[q];
        }
        amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1));
        amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 2));
    }
    float den[9], num[9], u0[9], u1[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) { den[q] = 0.0f; num[q] = 0.0f; u0[q] = 0.0f; u1[q] = 0.0f; }
    // The shortcut below replaces the reference's fp32 interpolation weights along disparity (lea_axis_half_pixel:
    // 2/3 and 1/3 carry rounding errors up to ~2e-6, output 1 has l1 = 1.49e-8) by exact thirds.  That changes a logit
    // by at most 2e-6 * |u_k+1 - u_k|: harmless for normalised logits, decisive for the soft-argmin of un-normalised ones
    // (random-init weights with identity BN give |u| ~ 1e8; measured at 288x576: 0.2 % of pixels off by up to 0.4 px
    // against the reference, all of them ties that the reference's weight bits break).  So it is taken only where every
    // logit of the cell is below 64 (error <= 2.6e-4 in the exponent); other cells use the reference's own weights.
    if (cache) __syncthreads();              // pass 2 also reads the first sample of the next lane's chunk
    if (maxdisp == 3 * D3 && amax <= 64.0f) {
        // pass 2, exact x3 scale (every BASELINE config): output 0 is sample 0, 3k+1 is sample k, 3k+2 and 3k+3 blend
        // samples k and k+1 with weights (2/3, 1/3) and (1/3, 2/3) (sample D3 := sample D3-1).  With
        // T_k = exp((m - u_k) / 3) the three softmin terms are T_k^3, T_k^2 T_{k+1}, T_k T_{k+1}^2: ONE exp per sample
        // instead of three and no per-output blends.  This lane owns samples [ka, kb).
        const int kchunk = (D3 + LEA_DH_PARTS - 1) / LEA_DH_PARTS;
        const int ka = min(D3, part * kchunk), kb = min(D3, ka + kchunk);
        const float c3 = 1.4426950408889634f / 3.0f;                   // log2(e) / 3
        float T[9], raw[9], u[9];
        if (ka < kb) {
            if (cache) {
#pragma unroll
                for (int q = 0; q < 9; ++q) u[q] = ucache[q * kThreads + threadIdx.x];
            } else {
                LEA_DH_LOAD(raw, ka);
                LEA_DH_COMBINE(u, raw);
            }
#pragma unroll
            for (int q = 0; q < 9; ++q) {
                T[q] = exp2f((m[q] - u[q]) * c3);              // subtract first: logits can be ~1e8
                if (ka == 0) den[q] += T[q] * T[q] * T[q];             // output 0 (weight 0 in the numerator)
            }
        }
        for (int k = ka; k < kb; ++k) {
            float Tn[9];
            if (k + 1 < D3) {
                if (cache) {
                    // sample k+1: this lane's own chunk, or the first sample of the next lane (same cell, thread + 1)
                    const int own = (k + 1 < kb) ? 1 : 0;
                    const int base = own ? (k + 1 - ka) * 9 * kThreads + (int)threadIdx.x : (int)threadIdx.x + 1;
#pragma unroll
                    for (int q = 0; q < 9; ++q) u[q] = ucache[base + q * kThreads];
                } else {
                    LEA_DH_LOAD(raw, k + 1);
                    LEA_DH_COMBINE(u, raw);
                }
#pragma unroll
                for (int q = 0; q < 9; ++q) Tn[q] = exp2f((m[q] - u[q]) * c3);
            } else {
#pragma unroll
                for (int q = 0; q < 9; ++q) Tn[q] = T[q];
            }
            const float i1 = (float)(3 * k + 1), i2 = (float)(3 * k + 2), i3 = (float)(3 * k + 3);
            const bool has3 = (k + 1 < D3);                             // output 3k+3 exists up to k = D3-2
#pragma unroll
            for (int q = 0; q < 9; ++q) {
                const float t2 = T[q] * T[q];
                const float e1 = t2 * T[q], e2 = t2 * Tn[q], e3 = has3 ? T[q] * Tn[q] * Tn[q] : 0.0f;
                den[q] += e1 + e2 + e3;
                num[q] += e1 * i1 + e2 * i2 + e3 * i3;
                T[q] = Tn[q];
            }
        }
    } else {
    // pass 2, general scale: this lane's quarter of the maxdisp output samples; (k0, k1) window along disparity
    int kc = -1, k1c = -1;
    const int kchunk_all = (D3 + LEA_DH_PARTS - 1) / LEA_DH_PARTS;     // owner lane / slot of a cached sample
    const int ichunk = (maxdisp + LEA_DH_PARTS - 1) / LEA_DH_PARTS;
    const int ia = min(maxdisp, part * ichunk), ib = min(maxdisp, ia + ichunk);
    for (int i = ia; i < ib; ++i) {
        const lea_axis_lerp ad = lea_axis_half_pixel(i, D3, maxdisp);
        if (ad.i0 != kc) {
            if (ad.i0 == k1c) {
#pragma unroll
                for (int q = 0; q < 9; ++q) u0[q] = u1[q];
            } else if (cache) {
                const int pp = ad.i0 / kchunk_all, base = ((ad.i0 - pp * kchunk_all) * 9) * kThreads + (int)threadIdx.x - part + pp;
#pragma unroll
                for (int q = 0; q < 9; ++q) u0[q] = ucache[base + q * kThreads];
            } else {
                float raw[9];
                LEA_DH_LOAD(raw, ad.i0);
                LEA_DH_COMBINE(u0, raw);
            }
            kc = ad.i0; k1c = -1;
        }
        if (ad.i1 != k1c) {
            if (ad.i1 == kc) {
#pragma unroll
                for (int q = 0; q < 9; ++q) u1[q] = u0[q];
            } else if (cache) {
                const int pp = ad.i1 / kchunk_all, base = ((ad.i1 - pp * kchunk_all) * 9) * kThreads + (int)threadIdx.x - part + pp;
#pragma unroll
                for (int q = 0; q < 9; ++q) u1[q] = ucache[base + q * kThreads];
            } else {
                float raw[9];
                LEA_DH_LOAD(raw, ad.i1);
                LEA_DH_COMBINE(u1, raw);
            }
            k1c = ad.i1;
        }
        const float fi = (float)i;
#pragma unroll
        for (int q = 0; q < 9; ++q) {
            const float v = ad.l0 * u0[q] + ad.l1 * u1[q];
            const float e = __expf(m[q] - v);                  // softmin: exp(-(v - m)), v >= m up to rounding
            den[q] += e;
            num[q] += e * fi;
        }
    }
    }
#pragma unroll
    for (int q = 0; q < 9; ++q) {
        den[q] += __shfl_xor_sync(0xffffffffu, den[q], 1);
        num[q] += __shfl_xor_sync(0xffffffffu, num[q], 1);
        den[q] += __shfl_xor_sync(0xffffffffu, den[q], 2);
        num[q] += __shfl_xor_sync(0xffffffffu, num[q], 2);
    }
    // lanes 0..2 of the cell write output rows 0..2
    if (live && part < 3) {
        float* __restrict__ o = disp + ((int64_t)b * 3 * H3 + 3 * h3 + part) * (3 * W3) + 3 * w3;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float nn = part == 0 ? num[c] : (part == 1 ? num[3 + c] : num[6 + c]);
            const float dd = part == 0 ? den[c] : (part == 1 ? den[3 + c] : den[6 + c]);
            o[c] = nn / dd;
        }
    }
}
#undef LEA_DH_LOAD
#undef LEA_DH_COMBINE

// =========================================================================================================
// K7  head without the up-sampled volume (retrain/skip_model_3d.py:162-169, :132):
//       mat = last_3( upsample_6( x ) ),  upsample_6 = trilinear align_corners=True to (D, H, W),
//       last_3 = Conv3d(C -> 1, 3x3x3, pad 1, no bias, no BN, no ReLU).
//     Both operators are linear and the interpolation acts on space only, so the channel contraction of last_3 can
//     run BEFORE the interpolation, on the small volume:
//       q[t](i,j,k) = sum_c w3[c, t] * x[c](i,j,k)              (t = kd*9 + kh*3 + kw; a 1x1x1 conv C -> 27)
//       mat(d,h,w)  = sum_t [v+t-1 inside] * up(q[t])(d+kd-1, h+kh-1, w+kw-1)
//     and because up = Ud x Uh x Uw is separable the tap sum collapses axis by axis:
//       R[kd,kh](i,j,w) = sum_kw [0<=w'<W] sum_e lw_e(w') q[kd,kh,kw](i, j, iw_e(w')),     w' = w+kw-1
//       S[kd](i,h,w)    = sum_kh [0<=h'<H] sum_e lh_e(h') R[kd,kh](i, ih_e(h'), w),        h' = h+kh-1
//       mat(d,h,w)      = sum_kd [0<=d'<D] sum_e ld_e(d') S[kd](id_e(d'), h, w),           d' = d+kd-1
//     ~100 M FMAs and ~170 MB of traffic per KITTI pair instead of a 436 MB up-sampled volume written, read back
//     and contracted with 5.9 GFLOP.  Needs out >= 2*in-1 per axis (then the three positions x-1..x+1 touch at most
//     three consecutive low-res samples), which every scale_dimension(.,2) up-sample satisfies.
// =========================================================================================================
struct lea_up3 { int base; float c[3][3]; };     // c[k][m]: weight of low-res sample base+m for position x+k-1

LEA_HD lea_up3 lea_up3_weights(int x, int in_n, int out_n) {
    lea_up3 r;
    r.base = lea_axis_ac(x > 0 ? x - 1 : 0, in_n, out_n).i0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int xx = x + k - 1;
        const bool inside = xx >= 0 && xx < out_n;                 // outside = the conv's zero padding
        const lea_axis_lerp a = lea_axis_ac(inside ? xx : x, in_n, out_n);
        const int m0 = a.i0 - r.base, m1 = a.i1 - r.base;
#pragma unroll
        for (int m = 0; m < 3; ++m)
            r.c[k][m] = inside ? ((m0 == m ? a.l0 : 0.0f) + (m1 == m ? a.l1 : 0.0f)) : 0.0f;
    }
    return r;
}

// stage w: q = channels [q_c0, q_c0+27) of a planes volume (B, ., D1, H1, W1)  ->  R fp32 (B, 9, D1, H1, W)
__global__ void __launch_bounds__(128)
lea_head_taps_w_kernel(lea_vol q, int q_c0, float* __restrict__ R, int W) {
    const int w = blockIdx.x * 128 + threadIdx.x;
    if (w >= W) return;
    const int j = blockIdx.y % q.H, i = blockIdx.y / q.H, b = blockIdx.z;
    const lea_up3 u = lea_up3_weights(w, q.W, W);
    const int64_t PS = (int64_t)q.D * q.H * q.W, CBS = PS * q.P;           // plane / channel-block strides (groups)
    const lea_u4* row = (const lea_u4*)q.data + ((int64_t)b * (q.C >> 3) + (q_c0 >> 3)) * CBS + ((int64_t)i * q.H + j) * q.W;
    float acc[9];
#pragma unroll
    for (int c = 0; c < 9; ++c) acc[c] = 0.0f;
#pragma unroll
    for (int m = 0; m < 3; ++m) {
        if (u.c[0][m] == 0.0f && u.c[1][m] == 0.0f && u.c[2][m] == 0.0f) continue;
        const int k = min(u.base + m, q.W - 1);
        float t[32];
#pragma unroll
        for (int cb = 0; cb < 4; ++cb) lea_load8_at(row + cb * CBS + k, PS, q.P, t + cb * 8);
#pragma unroll
        for (int c = 0; c < 9; ++c)
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) acc[c] += u.c[kw][m] * t[c * 3 + kw];
    }
    const int64_t plane = (int64_t)q.D * q.H * W;
    float* __restrict__ o = R + (int64_t)b * 9 * plane + ((int64_t)i * q.H + j) * W + w;
#pragma unroll
    for (int c = 0; c < 9; ++c) o[c * plane] = acc[c];
}

// stages h and d work on plain fp32 rows; VEC = 4 processes four consecutive w per thread with 128-bit accesses
// (W % 4 == 0), VEC = 1 is the ragged-width variant.  Both were instruction-bound on address arithmetic, hence the
// single base pointer + compile-time multiples of the plane stride.
template <int VEC> struct lea_fvec;
template <> struct lea_fvec<1> { float v[1]; };
template <> struct __attribute__((aligned(16))) lea_fvec<4> { float v[4]; };

// stage h: R (B, 9, D1, H1, W) -> S (B, 3, D1, H, W)
template <int VEC>
__global__ void __launch_bounds__(128)
lea_head_taps_h_kernel(const float* __restrict__ R, float* __restrict__ S, int D1, int H1, int H, int W) {
    const int w = (blockIdx.x * 128 + threadIdx.x) * VEC;
    if (w >= W) return;
    const int h = blockIdx.y % H, i = blockIdx.y / H, b = blockIdx.z;
    const lea_up3 u = lea_up3_weights(h, H1, H);
    const int64_t plane = (int64_t)D1 * H1 * W;                         // one (kd, kh) channel of R
    const float* __restrict__ rb = R + (int64_t)b * 9 * plane + (int64_t)i * H1 * W + w;
    float acc[3][VEC];
#pragma unroll
    for (int kd = 0; kd < 3; ++kd)
#pragma unroll
        for (int e = 0; e < VEC; ++e) acc[kd][e] = 0.0f;
#pragma unroll
    for (int m = 0; m < 3; ++m) {
        if (u.c[0][m] == 0.0f && u.c[1][m] == 0.0f && u.c[2][m] == 0.0f) continue;
        const float* __restrict__ rr = rb + (int64_t)min(u.base + m, H1 - 1) * W;
#pragma unroll
        for (int kd = 0; kd < 3; ++kd)
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
                const lea_fvec<VEC> x = *reinterpret_cast<const lea_fvec<VEC>*>(rr + (kd * 3 + kh) * plane);
#pragma unroll
                for (int e = 0; e < VEC; ++e) acc[kd][e] += u.c[kh][m] * x.v[e];
            }
    }
    const int64_t splane = (int64_t)D1 * H * W;
    float* __restrict__ so = S + (int64_t)b * 3 * splane + ((int64_t)i * H + h) * W + w;
#pragma unroll
    for (int kd = 0; kd < 3; ++kd) {
        lea_fvec<VEC> o;
#pragma unroll
        for (int e = 0; e < VEC; ++e) o.v[e] = acc[kd][e];
        *reinterpret_cast<lea_fvec<VEC>*>(so + kd * splane) = o;
    }
}

// stage d: S (B, 3, D1, H, W) -> mat (B, 1, D, H, W)
template <int VEC>
__global__ void __launch_bounds__(128)
lea_head_taps_d_kernel(const float* __restrict__ S, float* __restrict__ mat, int D1, int D, int H, int W) {
    const int w = (blockIdx.x * 128 + threadIdx.x) * VEC;
    if (w >= W) return;
    const int h = blockIdx.y % H, d = blockIdx.y / H, b = blockIdx.z;
    const lea_up3 u = lea_up3_weights(d, D1, D);
    const int64_t HW = (int64_t)H * W, splane = HW * D1;
    const float* __restrict__ sb = S + (int64_t)b * 3 * splane + (int64_t)h * W + w;
    float acc[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) acc[e] = 0.0f;
#pragma unroll
    for (int m = 0; m < 3; ++m) {
        if (u.c[0][m] == 0.0f && u.c[1][m] == 0.0f && u.c[2][m] == 0.0f) continue;
        const float* __restrict__ sr = sb + (int64_t)min(u.base + m, D1 - 1) * HW;
#pragma unroll
        for (int kd = 0; kd < 3; ++kd) {
            const lea_fvec<VEC> x = *reinterpret_cast<const lea_fvec<VEC>*>(sr + kd * splane);
#pragma unroll
            for (int e = 0; e < VEC; ++e) acc[e] += u.c[kd][m] * x.v[e];
        }
    }
    lea_fvec<VEC> o;
#pragma unroll
    for (int e = 0; e < VEC; ++e) o.v[e] = acc[e];
    *reinterpret_cast<lea_fvec<VEC>*>(mat + (((int64_t)b * D + d) * H + h) * W + w) = o;
}

// =========================================================================================================
// K8  collapsed stem0 (retrain/LEAStereo.py:34-48 + matching.stem0, skip_model_3d.py:92,141).  The cost volume is
//     cost[c] = x[c,h,w], cost[C+c] = y[c,h,w-d] wherever w >= d.  For an output voxel whose whole 3x3x3 window is
//     un-masked (lea_cv_interior) the conv separates exactly:
//         sum_{c,kd,kh,kw} W[co,c,kd,kh,kw] x[c,h+kh-1,w+kw-1]            = L[co,h,w]      (3x3 kernel  sum_kd W)
//         sum_{c,kd,kh,kw} W[co,C+c,kd,kh,kw] y[c,h+kh-1,(w-d)+(kw-kd)]   = R[co,h,w-d]    (3x5 kernel over kw-kd)
//     and R(u) = A(u-1) + B(u+1) with two ordinary 3x3 kernels (taps kw-kd in {-2,-1,0} and {1,2}).  L, A, B are 2-D
//     convs of the feature maps, computed ONCE per pair instead of once per disparity (377 GFLOP -> 3 GFLOP); this
//     kernel adds them, applies BN+ReLU and writes stem0's output.  The remaining voxels (first/last depth, the band
//     w <= d+1, the last column tile) come from the tensor-core kernel with the fused loader (lea_tc_opts.cv_skip).
//     lmap (B, C_out, 1, H, W), abmap (B, 2*C_out, 1, H, W): planes volumes (3 planes = fp32 exact).
// =========================================================================================================
#define LEA_AS_DCH 16        // depths per CTA: the A/B rows it needs (128 + 16 positions) are staged once in shared memory
__global__ void __launch_bounds__(128)
lea_stem0_assemble_kernel(lea_vol lmap, lea_vol abmap, lea_vol dst, int dst_c0, int c_out,
                          const float* __restrict__ bn_scale, const float* __restrict__ bn_shift, int relu) {
    __shared__ float sA[(128 + LEA_AS_DCH) * 8];
    __shared__ float sB[(128 + LEA_AS_DCH) * 8];
    const int tid = threadIdx.x;
    const int w0 = blockIdx.x * 128, w = w0 + tid;
    const int h = blockIdx.y % dst.H, dch = blockIdx.y / dst.H;
    const int d0 = dch * LEA_AS_DCH, d1 = min(dst.D, d0 + LEA_AS_DCH);
    const int cbn = c_out >> 3;
    const int b = blockIdx.z / cbn, cb = blockIdx.z - b * cbn;
    const int64_t hw = (int64_t)lmap.H * lmap.W;
    const lea_u4* lb = (const lea_u4*)lmap.data + ((int64_t)b * (lmap.C >> 3) + cb) * lmap.P * hw + (int64_t)h * lmap.W;
    const lea_u4* ab = (const lea_u4*)abmap.data + ((int64_t)b * (abmap.C >> 3) + cb) * abmap.P * hw + (int64_t)h * lmap.W;
    const lea_u4* bb = ab + (int64_t)cbn * abmap.P * hw;                   // B = channels [c_out, 2*c_out)
    // A is read at w-d-1, B at w-d+1 = that + 2: position k of both staging rows belongs to A index ia0 + k
    const int ia0 = w0 - d1;
    const int nstage = 127 + (d1 - d0);
    for (int k = tid; k < nstage; k += 128) {
        // clamped positions are only ever used by voxels outside the interior (which this kernel does not write)
        const int ia = min(max(ia0 + k, 0), lmap.W - 1), ib = min(max(ia0 + k + 2, 0), lmap.W - 1);
        lea_load8_at(ab + ia, hw, abmap.P, sA + k * 8);
        lea_load8_at(bb + ib, hw, abmap.P, sB + k * 8);
    }
    __syncthreads();
    if (w >= dst.W) return;
    float l[8], sc[8], sh[8];
    lea_load8_at(lb + w, hw, lmap.P, l);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        sc[j] = bn_scale ? __ldg(bn_scale + cb * 8 + j) : 1.0f;
        sh[j] = bn_scale ? __ldg(bn_shift + cb * 8 + j) : 0.0f;
    }
    const int64_t dHW = (int64_t)dst.H * dst.W, dPS = dHW * dst.D;
    lea_u4* ob = (lea_u4*)dst.data + ((int64_t)b * (dst.C >> 3) + (dst_c0 >> 3) + cb) * dst.P * dPS + (int64_t)h * dst.W + w;
    float cst[8];                                                          // relu(bn(0)): the fully masked voxels
#pragma unroll
    for (int j = 0; j < 8; ++j) cst[j] = (relu && sh[j] < 0.0f) ? 0.0f : sh[j];
    const int tw = w >> 3;
    for (int d = d0; d < d1; ++d) {
        if (lea_cv_masked(d, tw)) {
            lea_store8_at(ob + (int64_t)d * dHW, dPS, dst.P, cst);
            continue;
        }
        if (!lea_cv_interior(d, tw, dst.D, dst.W)) continue;
        const int k = (w - d - 1) - ia0;                                   // = tid + d1 - d - 1
        float out[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float v = l[j] + (sA[k * 8 + j] + sB[k * 8 + j]);
            v = v * sc[j] + sh[j];
            if (relu) v = v > 0.0f ? v : 0.0f;
            out[j] = v;
        }
        lea_store8_at(ob + (int64_t)d * dHW, dPS, dst.P, out);
    }
}

// DisparityRegression alone (models/build_model_2d.py:36-41)
__global__ void __launch_bounds__(256)
lea_disparity_regression_kernel(const float* __restrict__ p, float* __restrict__ out, int maxdisp, int HW) {
    const int q = blockIdx.x * 256 + threadIdx.x;
    if (q >= HW) return;
    const int b = blockIdx.y;
    const float* __restrict__ pb = p + (int64_t)b * maxdisp * HW + q;
    float acc = 0.0f;
    for (int d = 0; d < maxdisp; ++d) acc += __ldg(pb + (int64_t)d * HW) * (float)d;
    out[(int64_t)b * HW + q] = acc;
}


// =========================================================================================================
// F1  fused feature stems (retrain/new_model_2d.py:93-94, 130-131): stem0 = ConvBR2d(3 -> c_mid, 3x3, s1, p1) on the
//     full-resolution image, stem1 = ConvBR2d(c_mid -> c_out, 3x3, stride 3, p1).  stem1's 3x3 windows do not overlap
//     (stride 3 = kernel 3), so one thread per stem1 output pixel recomputes exactly the 9 stem0 pixels it needs from
//     a 5x5x3 image patch; the c_mid x H x W stem0 activation (61 MB per KITTI image) is never written.
//     Output: 1/3-resolution planes volume (depth 1).  fp32 FMA, eval-mode BN folded to scale/shift by the host.
// =========================================================================================================
#define LEA_FS_MAXMID 16
#define LEA_FS_MAXOUT 32
__global__ void __launch_bounds__(128)
lea_feature_stem_kernel(const float* __restrict__ img, int H, int W,
                        const float* __restrict__ w0, const float* __restrict__ sc0, const float* __restrict__ sh0,
                        int c_mid, const float* __restrict__ w1, const float* __restrict__ sc1,
                        const float* __restrict__ sh1, int c_out, lea_vol dst, int dst_c0) {
    __shared__ float w0s[27 * LEA_FS_MAXMID];                    // [c*9 + a*3 + b][m]
    __shared__ float w1s[9 * LEA_FS_MAXMID * LEA_FS_MAXOUT];     // [(i*3+j)*c_mid + m][o]
    __shared__ float bn0[2 * LEA_FS_MAXMID], bn1[2 * LEA_FS_MAXOUT];
    const int tid = threadIdx.x;
    for (int e = tid; e < 27 * LEA_FS_MAXMID; e += 128) {
        const int t = e / LEA_FS_MAXMID, m = e % LEA_FS_MAXMID;
        w0s[e] = m < c_mid ? __ldg(w0 + m * 27 + t) : 0.0f;
    }
    for (int e = tid; e < 9 * LEA_FS_MAXMID * LEA_FS_MAXOUT; e += 128) {
        const int o = e % LEA_FS_MAXOUT, r = e / LEA_FS_MAXOUT;
        const int m = r % LEA_FS_MAXMID, pos = r / LEA_FS_MAXMID;
        w1s[e] = (o < c_out && m < c_mid) ? __ldg(w1 + ((int64_t)o * c_mid + m) * 9 + pos) : 0.0f;
    }
    if (tid < LEA_FS_MAXMID) { bn0[tid] = tid < c_mid ? sc0[tid] : 0.0f; bn0[LEA_FS_MAXMID + tid] = tid < c_mid ? sh0[tid] : 0.0f; }
    if (tid < LEA_FS_MAXOUT) { bn1[tid] = tid < c_out ? sc1[tid] : 0.0f; bn1[LEA_FS_MAXOUT + tid] = tid < c_out ? sh1[tid] : 0.0f; }
    __syncthreads();
    // Two horizontally adjacent stem1 pixels per thread: every weight vector fetched from shared memory (one LDS.128
    // per 4 FMAs in the one-pixel version, which made the kernel shared-memory-issue bound at 12 TFLOP/s) feeds both.
    // The image is walked one stem0 row (= 3 image rows x 8 columns x 3 channels in registers) at a time.
    const int w3 = (blockIdx.x * 128 + tid) * 2;
    const int h3 = blockIdx.y, b = blockIdx.z;
    if (w3 >= dst.W) return;
    const bool two = w3 + 1 < dst.W;
    const float* __restrict__ ib = img + (int64_t)b * 3 * H * W;
    float acc[2][LEA_FS_MAXOUT];
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
        for (int o = 0; o < LEA_FS_MAXOUT; ++o) acc[p][o] = 0.0f;
#pragma unroll 1
    for (int ij = 0; ij < 9; ++ij) {             // one stem1 tap at a time (kept rolled: the unrolled body spilled)
        const int i = ij / 3, j = ij - 3 * i;
        const int py = 3 * h3 - 1 + i;                                   // stem0 row feeding taps (i, .) of stem1
        if (py < 0 || py >= H) continue;                                 // stem1's zero padding (uniform per block)
        {
            float im[3][3][6];                   // [channel][image row py-1+a][column 3*w3-2+j+x]: both pixels' 3x3 windows
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int x = 0; x < 6; ++x) {
                        const int y = py - 1 + a, xx = 3 * w3 - 2 + j + x;
                        im[c][a][x] = (y >= 0 && y < H && xx >= 0 && xx < W) ? __ldg(ib + ((int64_t)c * H + y) * W + xx) : 0.0f;
                    }
            const int px0 = 3 * w3 - 1 + j, px1 = px0 + 3;               // stem0 columns of the two pixels
            const bool ok0 = px0 >= 0 && px0 < W, ok1 = two && px1 < W;
            float mid[2][LEA_FS_MAXMID];
#pragma unroll
            for (int m = 0; m < LEA_FS_MAXMID; ++m) { mid[0][m] = 0.0f; mid[1][m] = 0.0f; }
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int e = 0; e < 3; ++e) {
                        const float v0 = im[c][a][e], v1 = im[c][a][3 + e];
                        const float4* wr = reinterpret_cast<const float4*>(w0s + (c * 9 + a * 3 + e) * LEA_FS_MAXMID);
#pragma unroll
                        for (int m4 = 0; m4 < LEA_FS_MAXMID / 4; ++m4) {
                            const float4 q = wr[m4];
                            mid[0][m4 * 4 + 0] += v0 * q.x; mid[0][m4 * 4 + 1] += v0 * q.y;
                            mid[0][m4 * 4 + 2] += v0 * q.z; mid[0][m4 * 4 + 3] += v0 * q.w;
                            mid[1][m4 * 4 + 0] += v1 * q.x; mid[1][m4 * 4 + 1] += v1 * q.y;
                            mid[1][m4 * 4 + 2] += v1 * q.z; mid[1][m4 * 4 + 3] += v1 * q.w;
                        }
                    }
#pragma unroll
            for (int m = 0; m < LEA_FS_MAXMID; ++m) {
                float t0 = mid[0][m] * bn0[m] + bn0[LEA_FS_MAXMID + m];
                float t1 = mid[1][m] * bn0[m] + bn0[LEA_FS_MAXMID + m];
                t0 = (ok0 && t0 > 0.0f) ? t0 : 0.0f;                     // outside the image: stem1's zero padding
                t1 = (ok1 && t1 > 0.0f) ? t1 : 0.0f;
                const float4* wr = reinterpret_cast<const float4*>(w1s + ((i * 3 + j) * LEA_FS_MAXMID + m) * LEA_FS_MAXOUT);
#pragma unroll
                for (int o4 = 0; o4 < LEA_FS_MAXOUT / 4; ++o4) {
                    const float4 q = wr[o4];
                    acc[0][o4 * 4 + 0] += t0 * q.x; acc[0][o4 * 4 + 1] += t0 * q.y;
                    acc[0][o4 * 4 + 2] += t0 * q.z; acc[0][o4 * 4 + 3] += t0 * q.w;
                    acc[1][o4 * 4 + 0] += t1 * q.x; acc[1][o4 * 4 + 1] += t1 * q.y;
                    acc[1][o4 * 4 + 2] += t1 * q.z; acc[1][o4 * 4 + 3] += t1 * q.w;
                }
            }
        }
    }
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        if (p == 0 || two) {
#pragma unroll
            for (int cb = 0; cb < LEA_FS_MAXOUT / 8; ++cb) {
                float out[8];
#pragma unroll
                for (int o = 0; o < 8; ++o) {
                    const float t = acc[p][cb * 8 + o] * bn1[cb * 8 + o] + bn1[LEA_FS_MAXOUT + cb * 8 + o];
                    out[o] = t > 0.0f ? t : 0.0f;
                }
                if (cb * 8 < c_out) lea_vol_store8(dst, b, (dst_c0 >> 3) + cb, 0, h3, w3 + p, out);
            }
        }
    }
}
