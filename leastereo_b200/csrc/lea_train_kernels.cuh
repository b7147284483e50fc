// Training-side kernels of the hot path (the reference gets all of this from autograd:
// train.py:156-160 -> loss.backward()).  Correctness-first CUDA-core kernels on the planes layout:
//   BatchNorm3d in train mode (batch statistics, models/operations_3d.py:38,44): statistics + affine/ReLU apply,
//   its backward (two per-channel reductions + apply), conv weight gradient, trilinear-resample backward,
//   cost-volume backward and the disparity-head backward.
// The data gradient of a convolution is a convolution with transposed, tap-flipped weights and reuses the forward
// kernels (tcgen05 or SIMT).  Compiles with nvcc and, for the no-GPU tests, with g++ -DLEA_CPU_EMU.
#pragma once
#include "lea_simt_kernels.cuh"

// ---------------------------------------------------------------------------------------------------------
// per-channel reductions over a channel slice of a planes volume.
//   mode 0: out[0] = sum x,            out[1] = sum x^2                       (BN statistics)
//   mode 1: g = dy * [x*scale+shift > 0 or !relu];  xh = (x-mean)*invstd;  out[0] = sum g, out[1] = sum g*xh
// partial[(chunk*2 + which)*c + ch]; the host adds the chunks (in fp64).  grid (chunks, c/8), block 256.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
lea_channel_reduce_kernel(lea_vol x, int x_c0, lea_vol dy, int dy_c0, int c, int mode, int relu,
                          const float* __restrict__ scale, const float* __restrict__ shift,
                          const float* __restrict__ mean, const float* __restrict__ invstd,
                          float* __restrict__ partial) {
    __shared__ float red[2][8][256];
    const int cb = blockIdx.y;
    const int64_t vox = (int64_t)x.B * x.D * x.H * x.W;
    const int64_t sp = (int64_t)x.D * x.H * x.W;
    float s0[8], s1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s0[j] = 0.0f; s1[j] = 0.0f; }
    // a channel block of one batch element is contiguous over (d, h, w): walk it linearly (no per-voxel divisions;
    // the kernel was instruction-bound on index arithmetic)
    (void)vox;
    for (int b = 0; b < x.B; ++b) {
        const lea_u4* xb = (const lea_u4*)x.data + ((int64_t)b * (x.C >> 3) + (x_c0 >> 3) + cb) * x.P * sp;
        const lea_u4* gb = (const lea_u4*)dy.data + ((int64_t)b * (dy.C >> 3) + (dy_c0 >> 3) + cb) * dy.P * sp;
        for (int64_t r = (int64_t)blockIdx.x * 256 + threadIdx.x; r < sp; r += (int64_t)gridDim.x * 256) {
            float f[8];
            lea_load8_at(xb + r, sp, x.P, f);
            if (mode == 0) {
#pragma unroll
                for (int j = 0; j < 8; ++j) { s0[j] += f[j]; s1[j] += f[j] * f[j]; }
            } else {
                float g[8];
                lea_load8_at(gb + r, sp, dy.P, g);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int ch = cb * 8 + j;
                    const float y = scale ? f[j] * scale[ch] + shift[ch] : f[j];
                    const float gg = (relu && !(y > 0.0f)) ? 0.0f : g[j];
                    const float xh = mean ? (f[j] - mean[ch]) * invstd[ch] : f[j];
                    s0[j] += gg; s1[j] += gg * xh;
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) { red[0][j][threadIdx.x] = s0[j]; red[1][j][threadIdx.x] = s1[j]; }
    __syncthreads();
    for (int stride = 128; stride > 0; stride >>= 1) {
        if ((int)threadIdx.x < stride) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                red[0][j][threadIdx.x] += red[0][j][threadIdx.x + stride];
                red[1][j][threadIdx.x] += red[1][j][threadIdx.x + stride];
            }
        }
        __syncthreads();
    }
    if (threadIdx.x < 16) {
        const int which = threadIdx.x >> 3, j = threadIdx.x & 7;
        partial[((int64_t)blockIdx.x * 2 + which) * c + cb * 8 + j] = red[which][j][0];
    }
}

// ---------------------------------------------------------------------------------------------------------
// elementwise over a channel slice:   t = relu?(x*scale[ch] + shift[ch]);   dst = accumulate ? dst + t : t
// (BN apply in train mode, also the state sum: the second summand is accumulated into the first one's slot)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
lea_affine_relu_kernel(lea_vol x, int x_c0, lea_vol dst, int dst_c0, int c, const float* __restrict__ scale,
                       const float* __restrict__ shift, int relu, int accumulate) {
    const int w = blockIdx.x * 256 + threadIdx.x;
    if (w >= x.W) return;
    const int h = blockIdx.y % x.H, d = blockIdx.y / x.H;
    const int cbn = c >> 3;
    const int b = blockIdx.z / cbn, cb = blockIdx.z - b * cbn;
    float f[8];
    lea_vol_load8(x, b, (x_c0 >> 3) + cb, d, h, w, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float t = scale ? f[j] * scale[cb * 8 + j] + shift[cb * 8 + j] : f[j];
        if (relu) t = t > 0.0f ? t : 0.0f;
        f[j] = t;
    }
    if (accumulate) {
        float o[8];
        lea_vol_load8(dst, b, (dst_c0 >> 3) + cb, d, h, w, o);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] += o[j];
    }
    lea_vol_store8(dst, b, (dst_c0 >> 3) + cb, d, h, w, f);
}

// ---------------------------------------------------------------------------------------------------------
// BN(train)+ReLU backward apply:  g = dy*[relu mask];  xh = (x-mean)*invstd;  dx = ka[ch]*g - kb[ch] - xh*kc[ch]
// with ka = gamma*invstd, kb = ka*sum(g)/n, kc = ka*sum(g*xh)/n computed by the host from the reductions above.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
lea_bn_relu_bwd_kernel(lea_vol x, int x_c0, lea_vol dy, int dy_c0, lea_vol dx, int dx_c0, int c, int relu,
                       const float* __restrict__ scale, const float* __restrict__ shift,
                       const float* __restrict__ mean, const float* __restrict__ invstd,
                       const float* __restrict__ ka, const float* __restrict__ kb, const float* __restrict__ kc) {
    const int w = blockIdx.x * 256 + threadIdx.x;
    if (w >= x.W) return;
    const int h = blockIdx.y % x.H, d = blockIdx.y / x.H;
    const int cbn = c >> 3;
    const int b = blockIdx.z / cbn, cb = blockIdx.z - b * cbn;
    float f[8], g[8];
    lea_vol_load8(x, b, (x_c0 >> 3) + cb, d, h, w, f);
    lea_vol_load8(dy, b, (dy_c0 >> 3) + cb, d, h, w, g);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int ch = cb * 8 + j;
        const float y = f[j] * scale[ch] + shift[ch];
        const float gg = (relu && !(y > 0.0f)) ? 0.0f : g[j];
        const float xh = (f[j] - mean[ch]) * invstd[ch];
        g[j] = ka[ch] * gg - kb[ch] - xh * kc[ch];
    }
    lea_vol_store8(dx, b, (dx_c0 >> 3) + cb, d, h, w, g);
}

// ---------------------------------------------------------------------------------------------------------
// conv weight gradient:  dW[co][ci][tap] += sum_{b,d,h,w} dO[b,co,d,h,w] * In[b,ci,(d,h,w)+tap-pad]
// One CTA owns an 8(w) x 16(h) column (all depths of one batch element) and one block of 8 input channels; per depth
// it stages the input halo and the dO tile in shared memory, every thread accumulates the k^3 taps of its (co, ci)
// pairs in registers, and the CTA flushes once with atomics.  grid (tiles, c_in/8, B), block 128.
// ---------------------------------------------------------------------------------------------------------
template <int KS>
__global__ void __launch_bounds__(128)
lea_conv_wgrad_kernel(lea_vol in, int in_c0, int c_in, lea_vol dout, int dout_c0, int c_out,
                      float* __restrict__ dw /* [c_out][c_in][KS^3] */) {
    constexpr int HALO = (KS == 3) ? 1 : 0;
    constexpr int HH = LEA_TH + 2 * HALO, WW = LEA_TW + 2 * HALO, SLAB = HH * WW, NSLAB = KS, TAPS = KS * KS * KS;
    LEA_DYN_SMEM(float, smem);
    float* in_s = smem;                                  // [8][NSLAB][HH][WW]
    float* do_s = smem + 8 * NSLAB * SLAB;               // [c_out_pad8][128]
    const int tiles_w = (in.W + LEA_TW - 1) / LEA_TW;
    const int tw = blockIdx.x % tiles_w, th = blockIdx.x / tiles_w;
    const int cb = blockIdx.y, b = blockIdx.z;
    const int w0 = tw * LEA_TW, h0 = th * LEA_TH;
    const int tid = threadIdx.x;
    const int npairs = c_out * 8;
    const int cpad = (c_out + 7) & ~7;
    for (int p0 = 0; p0 < npairs; p0 += 256) {           // each thread: pairs p0+tid and p0+128+tid
        float acc[2][TAPS];
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
            for (int t = 0; t < TAPS; ++t) acc[q][t] = 0.0f;
        for (int d = 0; d < in.D; ++d) {
            __syncthreads();
            for (int v = tid; v < NSLAB * SLAB; v += 128) {
                const int kd = v / SLAB, r = v - kd * SLAB;
                const int hh = r / WW, ww = r - hh * WW;
                const int gd = d + kd - HALO, gh = h0 + hh - HALO, gw = w0 + ww - HALO;
                float f[8];
                if (gd >= 0 && gd < in.D && gh >= 0 && gh < in.H && gw >= 0 && gw < in.W) {
                    lea_vol_load8(in, b, (in_c0 >> 3) + cb, gd, gh, gw, f);
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) f[j] = 0.0f;
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) in_s[j * NSLAB * SLAB + v] = f[j];
            }
            {
                const int lh = tid / LEA_TW, lw = tid % LEA_TW;
                const int gh = h0 + lh, gw = w0 + lw;
                const bool ok = gh < in.H && gw < in.W;
                for (int ocb = 0; ocb < (cpad >> 3); ++ocb) {
                    float f[8];
                    if (ok && ocb * 8 < c_out) lea_vol_load8(dout, b, (dout_c0 >> 3) + ocb, d, gh, gw, f);
                    else {
#pragma unroll
                        for (int j = 0; j < 8; ++j) f[j] = 0.0f;
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) do_s[(ocb * 8 + j) * 128 + tid] = f[j];
                }
            }
            __syncthreads();
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int pr = p0 + q * 128 + tid;
                if (pr < npairs) {
                    const int co = pr >> 3, ci = pr & 7;
                    const float* dq = do_s + co * 128;
                    const float* iq = in_s + ci * NSLAB * SLAB;
                    // one tile row (8 voxels along w) at a time: the 8 dO values and, per (kd, kh), the WW input values
                    // of the row are fetched once with 128/64-bit shared-memory loads and reused by all kw and all 8
                    // voxels (0.2 loads per FMA instead of 1: the kernel was shared-memory-issue bound)
                    for (int lh = 0; lh < LEA_TH; ++lh) {
                        float a[LEA_TW];
                        {
                            const float4 a0 = *reinterpret_cast<const float4*>(dq + lh * LEA_TW);
                            const float4 a1 = *reinterpret_cast<const float4*>(dq + lh * LEA_TW + 4);
                            a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
                            a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
                        }
#pragma unroll
                        for (int kd = 0; kd < KS; ++kd)
#pragma unroll
                            for (int kh = 0; kh < KS; ++kh) {
                                const float* rp = iq + kd * SLAB + (lh + kh) * WW;
                                float r[WW];
#pragma unroll
                                for (int x = 0; x < WW; x += 2) {
                                    const float2 t = *reinterpret_cast<const float2*>(rp + x);
                                    r[x] = t.x; r[x + 1] = t.y;
                                }
#pragma unroll
                                for (int kw = 0; kw < KS; ++kw) {
                                    float sum = acc[q][(kd * KS + kh) * KS + kw];
#pragma unroll
                                    for (int lw = 0; lw < LEA_TW; ++lw) sum += a[lw] * r[lw + kw];
                                    acc[q][(kd * KS + kh) * KS + kw] = sum;
                                }
                            }
                    }
                }
            }
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int pr = p0 + q * 128 + tid;
            if (pr < npairs) {
                const int co = pr >> 3, ci = cb * 8 + (pr & 7);
                if (ci < c_in) {
#pragma unroll
                    for (int t = 0; t < TAPS; ++t) atomicAdd(dw + ((int64_t)co * c_in + ci) * TAPS + t, acc[q][t]);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// 1x1x1 weight gradient: dW[co][ci] += sum_v dO[v][co] * In[v][ci] - a skinny GEMM (K = voxels), HBM-bound on reading
// In and dO once.  One CTA walks LEA_W1_TILES tiles of 64 consecutive voxels of one batch element; a tile is staged in
// shared memory as fp32 [voxel][channel] and every thread accumulates a 4 x 4 block of (co, ci) pairs in registers
// (2 vector loads per 16 FMAs); one atomic flush per CTA.  grid (voxel chunks, B), block 256.
// ---------------------------------------------------------------------------------------------------------
#define LEA_W1_VOX 64
#define LEA_W1_TILES 64
__global__ void __launch_bounds__(256)
lea_conv1_wgrad_kernel(lea_vol in, int in_c0, int c_in, lea_vol dout, int dout_c0, int c_out, int cpad,
                       float* __restrict__ dw /* [c_out][c_in] */) {
    LEA_DYN_SMEM(float, smem);
    float* in_s = smem;                                  // [LEA_W1_VOX][c_in]
    float* do_s = smem + LEA_W1_VOX * c_in;              // [LEA_W1_VOX][cpad]
    const int tid = threadIdx.x;
    const int b = blockIdx.y;
    const int64_t sp = (int64_t)in.D * in.H * in.W;
    const int64_t v_begin = (int64_t)blockIdx.x * LEA_W1_VOX * LEA_W1_TILES;
    const int64_t v_end = v_begin + LEA_W1_VOX * LEA_W1_TILES < sp ? v_begin + LEA_W1_VOX * LEA_W1_TILES : sp;
    const lea_u4* ib = (const lea_u4*)in.data + ((int64_t)b * (in.C >> 3) + (in_c0 >> 3)) * in.P * sp;
    const lea_u4* gb = (const lea_u4*)dout.data + ((int64_t)b * (dout.C >> 3) + (dout_c0 >> 3)) * dout.P * sp;
    const int nci4 = c_in >> 2, nco4 = cpad >> 2;
    const int nblk = nci4 * nco4;                        // 4 x 4 blocks of (co, ci); each thread owns blocks tid, tid+256
    float acc[2][16];
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int t = 0; t < 16; ++t) acc[q][t] = 0.0f;
    for (int64_t v0 = v_begin; v0 < v_end; v0 += LEA_W1_VOX) {
        __syncthreads();
        // stage: one (voxel, channel block) group per thread iteration
        for (int e = tid; e < LEA_W1_VOX * (c_in >> 3); e += 256) {
            const int vx = e % LEA_W1_VOX, cb = e / LEA_W1_VOX;
            float f[8];
            if (v0 + vx < v_end) lea_load8_at(ib + (int64_t)cb * in.P * sp + v0 + vx, sp, in.P, f);
            else {
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = 0.0f;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) in_s[vx * c_in + cb * 8 + j] = f[j];
        }
        for (int e = tid; e < LEA_W1_VOX * (cpad >> 3); e += 256) {
            const int vx = e % LEA_W1_VOX, cb = e / LEA_W1_VOX;
            float f[8];
            if (v0 + vx < v_end && cb * 8 < c_out) lea_load8_at(gb + (int64_t)cb * dout.P * sp + v0 + vx, sp, dout.P, f);
            else {
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = 0.0f;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) do_s[vx * cpad + cb * 8 + j] = f[j];
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int blk = tid + q * 256;
            if (blk < nblk) {
                const int co4 = blk / nci4, ci4 = blk - co4 * nci4;
                for (int vx = 0; vx < LEA_W1_VOX; ++vx) {
                    const float4 a = *reinterpret_cast<const float4*>(do_s + vx * cpad + co4 * 4);
                    const float4 x = *reinterpret_cast<const float4*>(in_s + vx * c_in + ci4 * 4);
                    acc[q][0] += a.x * x.x; acc[q][1] += a.x * x.y; acc[q][2] += a.x * x.z; acc[q][3] += a.x * x.w;
                    acc[q][4] += a.y * x.x; acc[q][5] += a.y * x.y; acc[q][6] += a.y * x.z; acc[q][7] += a.y * x.w;
                    acc[q][8] += a.z * x.x; acc[q][9] += a.z * x.y; acc[q][10] += a.z * x.z; acc[q][11] += a.z * x.w;
                    acc[q][12] += a.w * x.x; acc[q][13] += a.w * x.y; acc[q][14] += a.w * x.z; acc[q][15] += a.w * x.w;
                }
            }
        }
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int blk = tid + q * 256;
        if (blk < nblk) {
            const int co4 = blk / nci4, ci4 = blk - co4 * nci4;
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int co = co4 * 4 + i, ci = ci4 * 4 + j;
                    if (co < c_out) atomicAdd(dw + (int64_t)co * c_in + ci, acc[q][i * 4 + j]);
                }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// BatchNorm3d (train mode) per-channel glue, one thread per channel - what the host did with ~35 tiny tensor ops per
// ConvBR (operations_3d.py:38,44; F.batch_norm semantics):
//   finalize:  from the chunk partials of (sum x, sum x^2): mean, invstd = 1/sqrt(var_biased + eps),
//              scale = gamma*invstd, shift = beta - mean*scale, and the running statistics
//              r = (1-momentum)*r + momentum*{mean, var_unbiased} (fp32 like torch), num_batches_tracked += 1;
//   bwd coeffs: from the partials of (sum g, sum g*xh): ka = gamma*invstd, kb = ka*sum_g/n, kc = ka*sum_gx/n
//              (the coefficients of lea_bn_relu_bwd) and the parameter gradients dgamma = sum g*xh, dbeta = sum g.
// Chunk partials are added in fp64.
// ---------------------------------------------------------------------------------------------------------
// Chunk partials of channel ch are added by the 32 threads that share threadIdx.x / 32 (lane l takes chunks l, l+32, ...)
// and then in a fixed order by lane 0 - deterministic, and ~300 dependent fp64 adds per channel become ~10.
// block 256 = 8 channels x 32 lanes, grid ceil(c / 8).
LEA_D void lea_bn_chunk_sums(const float* __restrict__ partial, int chunks, int c, int ch, double (*red)[8][32],
                               double& s0, double& s1) {
    const int slot = threadIdx.x >> 5, l = threadIdx.x & 31;
    double a = 0.0, b = 0.0;
    if (ch < c) {
        for (int k = l; k < chunks; k += 32) {
            a += (double)partial[((int64_t)k * 2 + 0) * c + ch];
            b += (double)partial[((int64_t)k * 2 + 1) * c + ch];
        }
    }
    red[0][slot][l] = a; red[1][slot][l] = b;
    __syncthreads();
    s0 = 0.0; s1 = 0.0;
    if (l == 0) {
        for (int k = 0; k < 32; ++k) { s0 += red[0][slot][k]; s1 += red[1][slot][k]; }
    }
}

__global__ void __launch_bounds__(256)
lea_bn_finalize_kernel(const float* __restrict__ partial, int chunks, int c, double n, const float* __restrict__ gamma,
                       const float* __restrict__ beta, double eps, double momentum, float* __restrict__ running_mean,
                       float* __restrict__ running_var, long long* __restrict__ num_batches_tracked,
                       float* __restrict__ mean, float* __restrict__ invstd, float* __restrict__ scale,
                       float* __restrict__ shift) {
    __shared__ double red[2][8][32];
    const int ch = blockIdx.x * 8 + (threadIdx.x >> 5);
    double s, q;
    lea_bn_chunk_sums(partial, chunks, c, ch, red, s, q);
    if (ch >= c || (threadIdx.x & 31) != 0) return;
    const double m = s / n;
    double var = q / n - m * m;
    var = var > 0.0 ? var : 0.0;
    const double is = 1.0 / sqrt(var + eps);
    const double g = gamma ? (double)gamma[ch] : 1.0, bt = beta ? (double)beta[ch] : 0.0;
    mean[ch] = (float)m; invstd[ch] = (float)is;
    scale[ch] = (float)(g * is); shift[ch] = (float)(bt - m * g * is);
    if (running_mean) {
        const float keep = (float)(1.0 - momentum), mom = (float)momentum;
        const float unb = (float)(var * (n / (n - 1.0 > 1.0 ? n - 1.0 : 1.0)));
        const float rm = running_mean[ch] * keep, rv = running_var[ch] * keep;      // separately rounded, as the two
        const float am = mom * (float)m, av = mom * unb;                            // tensor ops of torch would
        running_mean[ch] = rm + am;
        running_var[ch] = rv + av;
    }
    if (ch == 0 && num_batches_tracked) *num_batches_tracked += 1;
}

// dgamma / dbeta: written, or ADDED when accumulate != 0 (the flat gradient bucket, zeroed once per step)
__global__ void __launch_bounds__(256)
lea_bn_bwd_coeffs_kernel(const float* __restrict__ partial, int chunks, int c, double n, const float* __restrict__ gamma,
                         const float* __restrict__ invstd, float* __restrict__ ka, float* __restrict__ kb,
                         float* __restrict__ kc, float* __restrict__ dgamma, float* __restrict__ dbeta, int accumulate) {
    __shared__ double red[2][8][32];
    const int ch = blockIdx.x * 8 + (threadIdx.x >> 5);
    double sg, sgx;
    lea_bn_chunk_sums(partial, chunks, c, ch, red, sg, sgx);
    if (ch >= c || (threadIdx.x & 31) != 0) return;
    const double a = (gamma ? (double)gamma[ch] : 1.0) * (double)invstd[ch];
    ka[ch] = (float)a; kb[ch] = (float)(a * sg / n); kc[ch] = (float)(a * sgx / n);
    if (dgamma) dgamma[ch] = (accumulate ? dgamma[ch] : 0.0f) + (float)sgx;
    if (dbeta) dbeta[ch] = (accumulate ? dbeta[ch] : 0.0f) + (float)sg;
}

// ---------------------------------------------------------------------------------------------------------
// trilinear (align_corners=True) backward, gather form: every SOURCE voxel sums the destination gradients that
// referenced it, dsrc += up^T(ddst).  Destination index j touches source i iff floor(j*s) is i-1 or i
// (s = (in-1)/(out-1)), i.e. j in [(i-1)/s, (i+1)/s): a short range per axis, scanned with one step of margin.
// ---------------------------------------------------------------------------------------------------------
struct lea_axis_range { int lo, hi; };
LEA_HD lea_axis_range lea_axis_ac_sources(int i, int in_n, int out_n) {
    lea_axis_range r;
    if (in_n == out_n) { r.lo = r.hi = i; return r; }
    if (out_n == 1) { r.lo = 0; r.hi = 0; return r; }
    const float s = (float)(in_n - 1) / (float)(out_n - 1);
    if (s <= 0.0f) { r.lo = 0; r.hi = out_n - 1; return r; }           // in_n == 1: every destination reads source 0
    int lo = (int)floorf((float)(i - 1) / s) - 1, hi = (int)floorf((float)(i + 1) / s) + 1;
    r.lo = lo < 0 ? 0 : lo;
    r.hi = hi > out_n - 1 ? out_n - 1 : hi;
    return r;
}
LEA_HD float lea_axis_ac_weight(int j, int i, int in_n, int out_n) {
    const lea_axis_lerp a = lea_axis_ac(j, in_n, out_n);
    return (a.i0 == i ? a.l0 : 0.0f) + (a.i1 == i ? a.l1 : 0.0f);
}

// Along w (the per-thread axis) the destinations with a non-zero weight are compacted into registers once
// (LEA_TB_TAPS entries: enough for up-sampling factors up to ~3; longer ranges fall back to the scanning loop), so the
// inner loop is loads + FMAs only: the first version evaluated a weight - a float division - for every candidate of
// a 7 x 7 x 7 range and was 10x off the traffic it needs.  accumulate == 0 overwrites dsrc (first writer).
#define LEA_TB_TAPS 6
__global__ void __launch_bounds__(128)
lea_trilinear_ac_bwd_kernel(lea_vol ddst, int ddst_c0, lea_vol dsrc, int dsrc_c0, int c, int accumulate) {
    const int w = blockIdx.x * 128 + threadIdx.x;
    if (w >= dsrc.W) return;
    const int h = blockIdx.y % dsrc.H, d = blockIdx.y / dsrc.H;
    const int cbn = c >> 3;
    const int b = blockIdx.z / cbn, cb = blockIdx.z - b * cbn;
    const lea_axis_range rd = lea_axis_ac_sources(d, dsrc.D, ddst.D);
    const lea_axis_range rh = lea_axis_ac_sources(h, dsrc.H, ddst.H);
    const lea_axis_range rw = lea_axis_ac_sources(w, dsrc.W, ddst.W);
    float acc[8];
    if (accumulate) lea_vol_load8(dsrc, b, (dsrc_c0 >> 3) + cb, d, h, w, acc);
    else {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.0f;
    }
    int tj[LEA_TB_TAPS];
    float tw[LEA_TB_TAPS];
    int nt = 0;
    bool compact = true;
#pragma unroll
    for (int q = 0; q < LEA_TB_TAPS; ++q) { tj[q] = 0; tw[q] = 0.0f; }
    for (int jw = rw.lo; jw <= rw.hi; ++jw) {
        const float ww = lea_axis_ac_weight(jw, w, dsrc.W, ddst.W);
        if (ww == 0.0f) continue;
        if (nt >= LEA_TB_TAPS) { compact = false; break; }
#pragma unroll
        for (int q = 0; q < LEA_TB_TAPS; ++q)
            if (q == nt) { tj[q] = jw; tw[q] = ww; }
        ++nt;
    }
    const int64_t dHW = (int64_t)ddst.H * ddst.W, dPS = dHW * ddst.D;
    const lea_u4* gbase = (const lea_u4*)ddst.data + ((int64_t)b * (ddst.C >> 3) + (ddst_c0 >> 3) + cb) * ddst.P * dPS;
    for (int jd = rd.lo; jd <= rd.hi; ++jd) {
        const float wd = lea_axis_ac_weight(jd, d, dsrc.D, ddst.D);
        if (wd == 0.0f) continue;
        for (int jh = rh.lo; jh <= rh.hi; ++jh) {
            const float wh = lea_axis_ac_weight(jh, h, dsrc.H, ddst.H);
            if (wh == 0.0f) continue;
            const float wdh = wd * wh;
            const lea_u4* grow = gbase + (int64_t)jd * dHW + (int64_t)jh * ddst.W;
            if (compact) {
#pragma unroll
                for (int q = 0; q < LEA_TB_TAPS; ++q) {
                    if (q < nt) {
                        float g[8];
                        lea_load8_at(grow + tj[q], dPS, ddst.P, g);
                        const float k = wdh * tw[q];
#pragma unroll
                        for (int j = 0; j < 8; ++j) acc[j] += k * g[j];
                    }
                }
            } else {
                for (int jw = rw.lo; jw <= rw.hi; ++jw) {
                    const float ww = lea_axis_ac_weight(jw, w, dsrc.W, ddst.W);
                    if (ww == 0.0f) continue;
                    float g[8];
                    lea_load8_at(grow + jw, dPS, ddst.P, g);
                    const float k = wdh * ww;
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[j] += k * g[j];
                }
            }
        }
    }
    lea_vol_store8(dsrc, b, (dsrc_c0 >> 3) + cb, d, h, w, acc);
}

// ---------------------------------------------------------------------------------------------------------
// cost-volume backward (retrain/LEAStereo.py:42-48 transposed):
//   dx[b,c,h,w] = sum_{d <= w} dcost[b,c,d,h,w]          dy[b,c,h,w] = sum_{d: w+d < W} dcost[b,C+c,d,h,w+d]
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
lea_cost_volume_bwd_kernel(lea_vol dcost, int C, float* __restrict__ dx, float* __restrict__ dy) {
    const int W = dcost.W, H = dcost.H, D = dcost.D;
    const int w = blockIdx.x * 256 + threadIdx.x;
    if (w >= W) return;
    const int h = blockIdx.y;
    const int cbn = dcost.C >> 3;
    const int b = blockIdx.z / cbn, cb = blockIdx.z - b * cbn;
    const bool left = cb < (C >> 3);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.0f;
    for (int d = 0; d < D; ++d) {
        const int ww = left ? w : w + d;
        if (left ? (d > w) : (ww >= W)) break;
        float g[8];
        lea_vol_load8(dcost, b, cb, d, h, ww, g);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += g[j];
    }
    float* __restrict__ o = (left ? dx : dy) + (((int64_t)b * C + (left ? cb : cb - (C >> 3)) * 8) * H + h) * W + w;
#pragma unroll
    for (int j = 0; j < 8; ++j) o[(int64_t)j * H * W] = acc[j];
}

// ---------------------------------------------------------------------------------------------------------
// disparity-head backward.  disp = sum_i p_i * i, p = softmax(-v)  =>  d disp / d v_j = -p_j * (j - disp);
// v_j = l0*u[k0] + l1*u[k1] along disparity, u = 3x3-tap bilinear blend of the low-res column.  One thread per
// low-res cell (same geometry as the forward kernel): it recomputes the forward, then pushes the 9 pixels' gradients
// back through the blends and adds them to the 3x3 neighbouring cells of dmat with atomics (dmat must be zeroed).
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(LEA_DH_BX * LEA_DH_BY)
lea_disp_head_bwd_kernel(const float* __restrict__ mat, const float* __restrict__ gout, float* __restrict__ dmat,
                         int D3, int H3, int W3, int maxdisp) {
    const int w3 = blockIdx.x * LEA_DH_BX + threadIdx.x % LEA_DH_BX;
    const int h3 = blockIdx.y * LEA_DH_BY + threadIdx.x / LEA_DH_BX;
    const int b = blockIdx.z;
    if (w3 >= W3 || h3 >= H3) return;
    const float* __restrict__ mb = mat + (int64_t)b * D3 * H3 * W3;
    float* __restrict__ db = dmat + (int64_t)b * D3 * H3 * W3;
    float th[3][3], tw[3][3];
#pragma unroll
    for (int r = 0; r < 3; ++r) { lea_three_tap(h3, r, H3, th[r]); lea_three_tap(w3, r, W3, tw[r]); }
    int ro[3], wo[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        int hh = h3 - 1 + i, ww = w3 - 1 + i;
        hh = hh < 0 ? 0 : (hh > H3 - 1 ? H3 - 1 : hh);
        ww = ww < 0 ? 0 : (ww > W3 - 1 ? W3 - 1 : ww);
        ro[i] = hh * W3; wo[i] = ww;
    }
#define LEA_DHB_BLEND(u, k)                                                                           \
    {                                                                                                 \
        const float* __restrict__ pk = mb + (int64_t)(k) * H3 * W3;                                   \
        float row[3][3];                                                                              \
        _Pragma("unroll") for (int i = 0; i < 3; ++i) {                                               \
            const float v0 = __ldg(pk + ro[i] + wo[0]), v1 = __ldg(pk + ro[i] + wo[1]),               \
                        v2 = __ldg(pk + ro[i] + wo[2]);                                               \
            _Pragma("unroll") for (int c = 0; c < 3; ++c)                                             \
                row[i][c] = tw[c][0] * v0 + tw[c][1] * v1 + tw[c][2] * v2;                            \
        }                                                                                             \
        _Pragma("unroll") for (int r = 0; r < 3; ++r)                                                 \
            _Pragma("unroll") for (int c = 0; c < 3; ++c)                                             \
                u[r * 3 + c] = th[r][0] * row[0][c] + th[r][1] * row[1][c] + th[r][2] * row[2][c];    \
    }
    // du[q] for sample k -> the 3x3 cells (weights th[r][i]*tw[c][j]); clamped edge taps have weight 0
#define LEA_DHB_FLUSH(du, k)                                                                          \
    {                                                                                                 \
        float cell[3][3];                                                                             \
        _Pragma("unroll") for (int i = 0; i < 3; ++i)                                                 \
            _Pragma("unroll") for (int j = 0; j < 3; ++j) {                                           \
                float a = 0.0f;                                                                       \
                _Pragma("unroll") for (int r = 0; r < 3; ++r)                                         \
                    _Pragma("unroll") for (int c = 0; c < 3; ++c) a += th[r][i] * tw[c][j] * du[r * 3 + c]; \
                cell[i][j] = a;                                                                       \
            }                                                                                         \
        float* __restrict__ dk = db + (int64_t)(k) * H3 * W3;                                         \
        _Pragma("unroll") for (int i = 0; i < 3; ++i)                                                 \
            _Pragma("unroll") for (int j = 0; j < 3; ++j)                                             \
                if (cell[i][j] != 0.0f) atomicAdd(dk + ro[i] + wo[j], cell[i][j]);                    \
    }
    float m[9], u[9];
    LEA_DHB_BLEND(u, 0);
#pragma unroll
    for (int q = 0; q < 9; ++q) m[q] = u[q];
    for (int k = 1; k < D3; ++k) {
        LEA_DHB_BLEND(u, k);
#pragma unroll
        for (int q = 0; q < 9; ++q) m[q] = u[q] < m[q] ? u[q] : m[q];
    }
    float den[9], num[9], u0[9], u1[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) { den[q] = 0.0f; num[q] = 0.0f; u0[q] = 0.0f; u1[q] = 0.0f; }
    int kc = -1, k1c = -1;
    for (int i = 0; i < maxdisp; ++i) {
        const lea_axis_lerp ad = lea_axis_half_pixel(i, D3, maxdisp);
        if (ad.i0 != kc) { LEA_DHB_BLEND(u0, ad.i0); kc = ad.i0; }
        if (ad.i1 != k1c) { LEA_DHB_BLEND(u1, ad.i1); k1c = ad.i1; }
#pragma unroll
        for (int q = 0; q < 9; ++q) {
            const float e = __expf(m[q] - (ad.l0 * u0[q] + ad.l1 * u1[q]));
            den[q] += e; num[q] += e * (float)i;
        }
    }
    float go[9], dsp[9];
    {
        const float* __restrict__ gp = gout + ((int64_t)b * 3 * H3 + 3 * h3) * (3 * W3) + 3 * w3;
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) { go[r * 3 + c] = gp[(int64_t)r * 3 * W3 + c]; dsp[r * 3 + c] = num[r * 3 + c] / den[r * 3 + c]; }
    }
    // pass 3, gather over the low-res disparity index k: the output samples i that reference k have
    // floor(src(i)) in {k-1, k}, src(i) = max(scale*(i+0.5)-0.5, 0), i.e. i+0.5 in [(k-0.5)/scale, (k+1.5)/scale);
    // they only involve the blended columns at k-1, k, k+1, which slide along with k.
    const float dscale = (float)D3 / (float)maxdisp;
    float up[9], uc[9], un[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) up[q] = 0.0f;
    LEA_DHB_BLEND(uc, 0);
    if (D3 > 1) { LEA_DHB_BLEND(un, 1); }
    else {
#pragma unroll
        for (int q = 0; q < 9; ++q) un[q] = uc[q];
    }
    for (int k = 0; k < D3; ++k) {
        int ilo = (int)floorf(((float)k - 0.5f) / dscale - 0.5f) - 2;
        int ihi = (int)floorf(((float)k + 1.5f) / dscale - 0.5f) + 2;
        ilo = ilo < 0 ? 0 : ilo;
        ihi = ihi > maxdisp - 1 ? maxdisp - 1 : ihi;
        float du[9];
#pragma unroll
        for (int q = 0; q < 9; ++q) du[q] = 0.0f;
        for (int i = ilo; i <= ihi; ++i) {
            const lea_axis_lerp ad = lea_axis_half_pixel(i, D3, maxdisp);
            const float wk = (ad.i0 == k ? ad.l0 : 0.0f) + (ad.i1 == k ? ad.l1 : 0.0f);
            if (wk == 0.0f) continue;
#pragma unroll
            for (int q = 0; q < 9; ++q) {
                const float x0 = ad.i0 == k - 1 ? up[q] : (ad.i0 == k ? uc[q] : un[q]);
                const float x1 = ad.i1 == k - 1 ? up[q] : (ad.i1 == k ? uc[q] : un[q]);
                const float p = __expf(m[q] - (ad.l0 * x0 + ad.l1 * x1)) / den[q];
                du[q] += wk * (-go[q] * p * ((float)i - dsp[q]));
            }
        }
        LEA_DHB_FLUSH(du, k);
#pragma unroll
        for (int q = 0; q < 9; ++q) { up[q] = uc[q]; uc[q] = un[q]; }
        if (k + 2 < D3) LEA_DHB_BLEND(un, k + 2);
    }
#undef LEA_DHB_BLEND
#undef LEA_DHB_FLUSH
}
