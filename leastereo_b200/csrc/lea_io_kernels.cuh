// Kernels of the callers on either side of the hot path: input normalisation + pad/crop
// (predict.py:144-184, dataloaders/datasets/common.py:94-131), masked smooth-L1 loss and its gradient
// (train.py:116-118, :157), fused Adam over a flat parameter buffer (train.py:76), disparity metrics
// (utils/metrics.py:6-46, train.py:162).  All HBM-bound elementwise / reduction work: coalesced grid-stride loops,
// block reduction in shared memory, one atomic per block and quantity.  Compiles for sm_100a and for the CPU emulator.
#pragma once
#include "lea_common.h"

#define LEA_IO_THREADS 256

// block-wide sum of NQ per-thread quantities; result valid in thread 0
template <typename T, int NQ>
LEA_D void lea_block_sum(T (&v)[NQ], T* scratch /*[NQ * LEA_IO_THREADS]*/) {
    const int tid = threadIdx.x;
#pragma unroll
    for (int q = 0; q < NQ; ++q) scratch[q * LEA_IO_THREADS + tid] = v[q];
    __syncthreads();
    for (int s = LEA_IO_THREADS / 2; s > 0; s >>= 1) {
        if (tid < s) {
#pragma unroll
            for (int q = 0; q < NQ; ++q) scratch[q * LEA_IO_THREADS + tid] += scratch[q * LEA_IO_THREADS + tid + s];
        }
        __syncthreads();
    }
#pragma unroll
    for (int q = 0; q < NQ; ++q) v[q] = scratch[q * LEA_IO_THREADS];
}

// ---------------------------------------------------------------------------------------------------------
// I1  per-channel statistics of an 8-bit RGB image in HWC order (what PIL hands to predict.py:158-163):
//     sums[c] = sum x, sums[3+c] = sum x^2 over all pixels, exact in 64-bit integers, so mean and the population
//     standard deviation of np.mean / np.std (predict.py:167-169) follow in fp64 on either side.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(LEA_IO_THREADS)
lea_image_stats_u8_kernel(const uint8_t* __restrict__ img, int64_t npix, unsigned long long* __restrict__ sums) {
    __shared__ unsigned long long scratch[6 * LEA_IO_THREADS];
    unsigned long long v[6] = {0, 0, 0, 0, 0, 0};
    for (int64_t i = (int64_t)blockIdx.x * LEA_IO_THREADS + threadIdx.x; i < npix; i += (int64_t)gridDim.x * LEA_IO_THREADS) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const unsigned long long x = img[i * 3 + c];
            v[c] += x; v[3 + c] += x * x;
        }
    }
    lea_block_sum<unsigned long long, 6>(v, scratch);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int q = 0; q < 6; ++q) atomicAdd(sums + q, v[q]);
    }
}

// I2  z-normalisation + test_transform (predict.py:144-156, :167-175): out (3, ch, cw) fp32, channel-planar.
//     If the image fits (h <= ch and w <= cw) it is placed bottom-right on zeros, otherwise the centre is cropped
//     (start = int((w - cw) / 2), int((h - ch) / 2)).  out[c] = float((double)x - mean_c) / std_c) in fp64 like numpy.
__global__ void __launch_bounds__(LEA_IO_THREADS)
lea_normalize_pad_u8_kernel(const uint8_t* __restrict__ img, int h, int w, const unsigned long long* __restrict__ sums,
                            float* __restrict__ out, int ch, int cw) {
    const int x = blockIdx.x * LEA_IO_THREADS + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= cw) return;
    const double n = (double)h * (double)w;
    const bool fits = (h <= ch) && (w <= cw);
    int sy, sx;
    bool inside;
    if (fits) { sy = y - (ch - h); sx = x - (cw - w); inside = (sy >= 0) && (sx >= 0); }
    else      { sy = y + (h - ch) / 2; sx = x + (w - cw) / 2; inside = (sy >= 0) && (sy < h) && (sx >= 0) && (sx < w); }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float r = 0.0f;
        if (inside) {
            const double mean = (double)sums[c] / n;
            double var = (double)sums[3 + c] / n - mean * mean;
            var = var > 0.0 ? var : 0.0;
            r = (float)(((double)img[((int64_t)sy * w + sx) * 3 + c] - mean) / sqrt(var));
        }
        out[((int64_t)c * ch + y) * cw + x] = r;
    }
}

// ---------------------------------------------------------------------------------------------------------
// L1  masked smooth-L1 (train.py:116-118, :157; F.smooth_l1_loss beta = 1, reduction 'mean') and the mean absolute
//     error of train.py:162.  acc[0] += sum of per-pixel loss, acc[1] += sum |d|, acc[2] += number of valid pixels
//     (mask = 0.001 < target < maxdisp); the caller divides.  fp64 accumulation.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(LEA_IO_THREADS)
lea_masked_smooth_l1_kernel(const float* __restrict__ disp, const float* __restrict__ target, int64_t n, float maxdisp,
                            double* __restrict__ acc) {
    __shared__ double scratch[3 * LEA_IO_THREADS];
    double v[3] = {0.0, 0.0, 0.0};
    for (int64_t i = (int64_t)blockIdx.x * LEA_IO_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * LEA_IO_THREADS) {
        const float t = target[i];
        if (t < maxdisp && t > 0.001f) {
            const float d = disp[i] - t;
            const float a = d < 0.0f ? -d : d;
            v[0] += (double)(a < 1.0f ? 0.5f * d * d : a - 0.5f);
            v[1] += (double)a;
            v[2] += 1.0;
        }
    }
    lea_block_sum<double, 3>(v, scratch);
    if (threadIdx.x == 0) {
        atomicAdd(acc + 0, v[0]); atomicAdd(acc + 1, v[1]); atomicAdd(acc + 2, v[2]);
    }
}

// gradient of the mean loss w.r.t. disp: mask * clamp(d, -1, 1) * upstream / count  (count read from acc[2] on the device)
__global__ void __launch_bounds__(LEA_IO_THREADS)
lea_masked_smooth_l1_bwd_kernel(const float* __restrict__ disp, const float* __restrict__ target, int64_t n,
                                float maxdisp, const double* __restrict__ acc, float upstream, float* __restrict__ grad) {
    const double cnt = acc[2];
    const float k = cnt > 0.0 ? (float)((double)upstream / cnt) : 0.0f;
    for (int64_t i = (int64_t)blockIdx.x * LEA_IO_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * LEA_IO_THREADS) {
        const float t = target[i];
        float g = 0.0f;
        if (t < maxdisp && t > 0.001f) {
            const float d = disp[i] - t;
            g = (d < -1.0f ? -1.0f : (d > 1.0f ? 1.0f : d)) * k;
        }
        grad[i] = g;
    }
}

// ---------------------------------------------------------------------------------------------------------
// A1  Adam step over a flat fp32 buffer (torch.optim.Adam defaults of train.py:76: no weight decay, no amsgrad):
//       m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  p -= (lr / (1-b1^t)) * m / (sqrt(v) / sqrt(1-b2^t) + eps)
//     bias corrections are computed by the host in fp64 and passed as step_size and inv_sqrt_bc2.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(LEA_IO_THREADS)
lea_adam_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                     int64_t n, float beta1, float beta2, float eps, float step_size, float inv_sqrt_bc2) {
    for (int64_t i = (int64_t)blockIdx.x * LEA_IO_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * LEA_IO_THREADS) {
        const float gi = g[i];
        const float mi = beta1 * m[i] + (1.0f - beta1) * gi;
        const float vi = beta2 * v[i] + (1.0f - beta2) * gi * gi;
        m[i] = mi; v[i] = vi;
        p[i] -= step_size * (mi / (sqrtf(vi) * inv_sqrt_bc2 + eps));
    }
}

// ---------------------------------------------------------------------------------------------------------
// M1  disparity metrics in one pass, REFERENCE-IDENTICAL by default (utils/metrics.py:6-46, evaluation.py:290-307,
//     train.py:116-118,203).  The reference fills `abs_diff = np.full(shape, 10000)` - an int64 array - and assigns
//     `abs_diff[mask] = |true - pred|` into it, which truncates the error toward zero; its tests are therefore
//     `trunc|d| < 3 or trunc|d| < float32(target * 0.05)` and `trunc|d| <= thr`, evaluated over EVERY pixel with the
//     invalid ones standing at 10000 (so an invalid pixel with target * 0.05 > 10000, or a threshold >= 10000, counts
//     as correct - reproduced).  evaluation.py:290-291 takes its EPE over the INCLUSIVE mask 0.001 <= t <= maxdisp,
//     train.py:203 over the strict one.  flags bit 0 (LEA_METRICS_FLOAT_DIFF) compares the un-truncated float |d|
//     instead (not what the reference computes; opt-in only).
//       acc[0] = # valid (0.001 < t < maxdisp)          acc[1] = sum |d| over 0.001 <= t <= maxdisp
//       acc[2] = # "3-px correct"                      acc[3 + k] = # pixels passing threshold k (k < 4)
//       acc[7] = # pixels with 0.001 <= t <= maxdisp    acc[8] = sum |d| over the strict mask
// ---------------------------------------------------------------------------------------------------------
#define LEA_METRIC_THR 4
#define LEA_METRIC_ACC (5 + LEA_METRIC_THR)
#define LEA_METRICS_FLOAT_DIFF 1
__global__ void __launch_bounds__(LEA_IO_THREADS)
lea_disparity_metrics_kernel(const float* __restrict__ pred, const float* __restrict__ target, int64_t n, float maxdisp,
                             float t0, float t1, float t2, float t3, int flags, double* __restrict__ acc) {
    __shared__ double scratch[LEA_METRIC_ACC * LEA_IO_THREADS];
    double v[LEA_METRIC_ACC];
#pragma unroll
    for (int q = 0; q < LEA_METRIC_ACC; ++q) v[q] = 0.0;
    const bool float_diff = (flags & LEA_METRICS_FLOAT_DIFF) != 0;
    const double thr[LEA_METRIC_THR] = {(double)t0, (double)t1, (double)t2, (double)t3};
    for (int64_t i = (int64_t)blockIdx.x * LEA_IO_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * LEA_IO_THREADS) {
        const float t = target[i];
        const float d = t - pred[i];                         // float32 like the numpy expression
        const float a = d < 0.0f ? -d : d;
        const bool valid = (t < maxdisp) && (t > 0.001f);
        double e = 10000.0;                                  // the reference's fill value for invalid pixels
        if (valid) {
            e = float_diff ? (double)a : (double)truncf(a);  // float -> int64 assignment truncates toward zero
            v[0] += 1.0; v[8] += (double)a;
        }
        if (t >= 0.001f && t <= maxdisp) { v[1] += (double)a; v[7] += 1.0; }
        if (e < 3.0 || e < (double)(t * 0.05f)) v[2] += 1.0;
#pragma unroll
        for (int k = 0; k < LEA_METRIC_THR; ++k)
            if (e <= thr[k]) v[3 + k] += 1.0;
    }
    lea_block_sum<double, LEA_METRIC_ACC>(v, scratch);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int q = 0; q < LEA_METRIC_ACC; ++q) atomicAdd(acc + q, v[q]);
    }
}
