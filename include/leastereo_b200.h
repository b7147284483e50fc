/*
 * leastereo_b200 - C ABI of the B200-native LEAStereo hot path (cost volume -> 3D MatchingNet -> disparity head).
 *
 * The reference (devmentality/LEAStereo) is pure Python/PyTorch and has no FFI of its own: its only
 * boundary is the nn.Module API.  The entry points below are what a binding for that path would call; each cites
 * the reference lines whose arithmetic it replaces.  INTEGRATION.md shows the ctypes stub a reference maintainer
 * would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer into memory owned by the caller (PyTorch allocates, keeps alive until the
 *     stream has passed); nothing is allocated, freed or synchronised inside the library;
 *   - `stream` is a cudaStream_t passed as void*; kernels are launched on it, on the current device;
 *   - return value: 0 = ok, non-zero = error (message via lea_last_error(), thread-local);
 *   - no torch types, no C++ types, no ownership transfer.
 *
 * "Planes volume" (lea_vol): the activation layout of the 3D net.  A logical fp32 tensor (B, C, D, H, W) is stored
 * as P bf16 planes (hi, lo[, lo2]) with value = sum of planes (P=2: 16 significant bits, the bf16x3 split-precision
 * operand; P=3: exact fp32), channel-blocked by 8 so that one (voxel, 8-channel, plane) group is a 16-byte UMMA
 * core-matrix row:
 *        element (b, c, p, d, h, w) lives at  (((((b*(C/8) + c/8)*P + p)*D + d)*H + h)*W + w)*8 + c%8      [bf16]
 */
#ifndef LEASTEREO_B200_H_
#define LEASTEREO_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LEA_ABI_VERSION 3

typedef struct lea_vol {
    void*   data;          /* bf16 planes volume, 16-byte aligned                                   */
    int32_t B, C, P;       /* batch, TOTAL channels of the buffer (multiple of 8), planes (1..3)    */
    int32_t D, H, W;       /* spatial extent                                                        */
} lea_vol;

/* One ConvBR (models/operations_3d.py:31-47): Conv3d(k in {1,3}, stride 1, pad (k-1)/2, no bias) -> optional
 * BatchNorm (eval: y = x*scale + shift, scale/shift precomputed by the host from gamma/beta/running stats, eps 1e-5)
 * -> optional ReLU -> optional "+ res" (the state sum of retrain/skip_model_3d.py:70) -> written into a channel
 * slice of a planes volume (the concat of skip_model_3d.py:74 is a slice write) or as plain fp32 NCDHW. */
typedef struct lea_conv {
    lea_vol      src;      int32_t src_c0;   /* input volume and first input channel                         */
    int32_t      c_in, c_out, ksize;         /* c_in % 8 == 0; c_out % 8 == 0 unless dst_f32 != NULL         */
    const float* bn_scale;                   /* [c_out] or NULL                                               */
    const float* bn_shift;                   /* [c_out] or NULL                                               */
    int32_t      relu;
    int32_t      has_res;  lea_vol res;      int32_t res_c0;   /* summand added after BN/ReLU                 */
    lea_vol      dst;      int32_t dst_c0;   /* planes output (used when dst_f32 == NULL)                     */
    float*       dst_f32;                    /* fp32 (B, c_out, D, H, W) output when non-NULL                 */
} lea_conv;

int         lea_abi_version(void);
const char* lea_last_error(void);
/* 1 if this build contains device code (0 for the CPU emulation used by the no-GPU unit tests). */
int         lea_is_device_build(void);

/* ---- cost volume: retrain/LEAStereo.py:34-48 --------------------------------------------------------------- */
/* fp32 reference layout, bit-exact: cost (B, 2C, D3, H, W); x, y (B, C, H, W). */
int lea_cost_volume_f32(const float* x, const float* y, float* cost,
                        int32_t B, int32_t C, int32_t H, int32_t W, int32_t D3, void* stream);
/* same volume written straight into the planes layout that stem0 consumes (vol.C == 2C). */
int lea_cost_volume_planes(const float* x, const float* y, const lea_vol* vol,
                           int32_t C, void* stream);
/* feature maps (B, C, H, W) fp32 -> 2-D planes volume (D == 1) used by the fused-cost-volume stem0 loader. */
int lea_pack_planes(const float* src, const lea_vol* dst, int32_t dst_c0, int32_t c, void* stream);
int lea_unpack_planes(const lea_vol* src, int32_t src_c0, int32_t c, float* dst, void* stream);

/* ---- trilinear resample, align_corners=True: retrain/skip_model_3d.py:44-51, :162-169 ----------------------- */
/* Optional epilogue y = relu(x*bn_scale[ch] + bn_shift[ch]) (bn_* may be NULL): a 1x1x1 convolution commutes with
 * the interpolation, so for UP-sampling cells the engine runs the cell's 1x1x1 conv on the small volume and applies
 * its BatchNorm+ReLU here, after interpolating (same value in exact arithmetic, ~8x less traffic). */
int lea_trilinear_ac(const lea_vol* src, int32_t src_c0, const lea_vol* dst, int32_t dst_c0, int32_t c,
                     const float* bn_scale, const float* bn_shift, int32_t relu, void* stream);

/* Down-sampling resample fused with the 1x1x1 ConvBR(s) that consume it (skip_model_3d.py:44-53: resample s0 / s1 to the
 * cell's size, then pre_preprocess / preprocess).  Up to two consumers (s1 of cell i is s0 of cell i+1) are computed in
 * ONE pass over the source; weight = PyTorch layout fp32 (c_out, c_in[,1,1,1]); all consumers share the target size. */
typedef struct lea_rc_out {
    lea_vol      dst;       int32_t dst_c0, c_out;
    const float* weight;    const float* bn_scale;   const float* bn_shift;   int32_t relu;
} lea_rc_out;
int lea_resample_conv1x1(const lea_vol* src, int32_t src_c0, int32_t c_in, const lea_rc_out* outs, int32_t n_out,
                         void* stream);

/* ---- ConvBR -------------------------------------------------------------------------------------------------- */
/* fp32-FMA kernel; weight = PyTorch layout fp32 (c_out, c_in, k, k, k). */
int lea_conv3d_simt(const lea_conv* p, const float* weight, void* stream);

/* tcgen05/TMEM implicit-GEMM kernel (sm_100a).  wimg = image produced by lea_pack_weights_tc. */
int64_t lea_tc_weight_image_bytes(int32_t c_in, int32_t c_out, int32_t ksize, int32_t planes);
int lea_pack_weights_tc(const float* weight, void* wimg, int32_t c_in, int32_t c_out, int32_t ksize,
                        int32_t planes, void* stream);
/* Image of the DATA-GRADIENT conv of a forward weight (fwd_c_out = c_in here, fwd_c_in >= c_out), read in place:
 * transposed in the channel pair and tap-flipped, W'[co][ci][tap] = weight[ci][fwd_ci0 + co][k^3-1-tap]
 * (the slice [fwd_ci0, fwd_ci0 + c_out) of the forward input channels: a launch takes at most 64; train.py:156-160). */
int lea_pack_weights_tc_dgrad(const float* weight, void* wimg, int32_t c_in, int32_t c_out, int32_t ksize,
                              int32_t planes, int32_t fwd_c_in, int32_t fwd_ci0, void* stream);
/* mma_terms: 1 = single-pass bf16 (hi*hi), 0 = full triangular split for the volume's P (P=2: bf16x3, P=3: bf16x6),
 * 2 (P = 3 volumes with c_in %% 16 == 0) = only the product terms of order < 2 (a0 w0, a1 w0, a0 w1): bf16x3 products on
 * 3-plane (exactly stored) operands.
 * fused_cv: when non-zero, src is ignored and the operand loader builds the cost volume on the fly from the two
 * 2-D planes volumes fx/fy (retrain/LEAStereo.py:34-48 fused into stem0; the volume is never materialised). */
typedef struct lea_tc_opts {
    int32_t mma_terms;
    int32_t fused_cv;  lea_vol fx, fy;  int32_t d3;
    int32_t num_sms;   /* 0 = query */
    int32_t accum_split;   /* 0 = default (1): hi*hi term in its own TMEM accumulator, correction terms in a second one
                              (the tensor core loses ~1 ulp per fp32 accumulation step; this cuts the steps of the
                              dominant accumulator 3x); 2 = all terms share one accumulator (more depth per work item);
                              3 = two accumulators only when the reduction spans more than one 16-channel group */
    int32_t acc_sets;      /* 0 = auto, 1 or 2 TMEM accumulator sets (2 = epilogue overlaps the next item's MMAs) */
    const void* cv_maps;   /* fused_cv: device array built by lea_build_fused_cv_maps for these fx/fy/d3 */
    int32_t resident_weights;  /* 0 = auto (all channel groups' weights stay in shared memory when they fit), 2 = never */
    int32_t cv_skip;           /* fused_cv only: 1 = leave the voxels lea_stem0_assemble writes (collapsed stem0) untouched */
    int32_t debug;             /* bits 0-3: timing-ablation switches (only in -DLEA_TC_ABLATION builds); bit 4: no flat mode for depth-1 volumes (A/B) */
    int32_t depth_chunk;       /* 0 = auto, n = depth slices per work item, clamped to what the schedule allows (test knob) */
    int32_t tile_w_log2;       /* 1x1x1 convs only: 0 = auto, 3..7 = tile of 2^n voxels along w by 128/2^n along h */
} lea_tc_opts;
int lea_conv3d_tc(const lea_conv* p, const void* wimg, const lea_tc_opts* opts, void* stream);
/* Fused cost volume (retrain/LEAStereo.py:34-48 inside stem0's operand loader).  fx, fy: the two feature maps as
 * 2-D planes volumes (D == 1, C channels each).  For every disparity d the loader needs x[h,w]*[w>=d] and y[h,w-d]*[w>=d];
 * both are plain TMA box loads through a per-disparity tensor map whose base is shifted by d voxels (x) and whose
 * width is W-d, so the out-of-bounds zero fill of TMA produces exactly the zeros of the reference loop.  The maps
 * (2*d3 descriptors) depend only on the buffers' addresses and shapes: build once per plan, pass via cv_maps.
 * With fused_cv the conv's `src` is ignored except for src.P; c_in must equal 2*fx.C and ksize 3. */
int64_t lea_fused_cv_maps_bytes(int32_t d3);
int lea_build_fused_cv_maps(const lea_vol* fx, const lea_vol* fy, int32_t d3, void* maps_dev, void* stream);
/* self-test of the tcgen05 path on a synthetic GEMM-shaped conv; returns 0 when it matches the SIMT kernel. */
int lea_tc_selftest(int32_t verbose, void* stream);

/* ---- collapsed stem0: retrain/LEAStereo.py:34-48 + matching.stem0 (skip_model_3d.py:141) --------------------------- */
/* Where the 3x3x3 window of an output voxel lies entirely in the un-masked part of the cost volume (first/last depth,
 * the band w <= d+1 and the last 8-column tile excluded), stem0's convolution separates exactly into 2-D convolutions
 * of the two feature maps: out = L[h,w] + A[h,w-d-1] + B[h,w-d+1] (L: x with sum_kd W; A/B: y with the taps kw-kd in
 * {-2,-1,0} / {1,2}).  The maps are computed once per pair by any ConvBR kernel above on depth-1 volumes; this entry
 * point adds them, applies BN+ReLU and writes those voxels of dst.  The excluded voxels are written by lea_conv3d_tc
 * with fused_cv and cv_skip = 1 (it skips exactly the voxels written here). */
int lea_stem0_assemble(const lea_vol* lmap, const lea_vol* abmap, const lea_vol* dst, int32_t dst_c0, int32_t c_out,
                       const float* bn_scale, const float* bn_shift, int32_t relu, void* stream);

/* ---- matching-net head without the up-sampled volume: retrain/skip_model_3d.py:162-169 (upsample_6 -> last_3) ---- */
/* last_3 (Conv3d C->1, 3x3x3, no BN/ReLU, skip_model_3d.py:132) and the trilinear align_corners=True up-sample are both
 * linear, so last_3's channel contraction runs first on the SMALL volume (a 1x1x1 conv C -> 27 "tap" channels,
 * channel t = kd*9+kh*3+kw, padded to 32; any ConvBR kernel above does it) and this entry point then evaluates
 *     mat(d,h,w) = sum_t [tap inside the volume] * upsample(q[t])(d+kd-1, h+kh-1, w+kw-1)
 * separably (three small kernels; workspace = lea_head_taps_workspace_bytes).  q: planes volume (B, >=32, D1, H1, W1)
 * with the tap channels at q_c0; mat: fp32 (B, 1, D, H, W).  Each axis must satisfy out >= 2*in-1 or out == in. */
int64_t lea_head_taps_workspace_bytes(int32_t B, int32_t D1, int32_t H1, int32_t D, int32_t H, int32_t W);
int lea_head_taps(const lea_vol* q, int32_t q_c0, float* mat, int32_t D, int32_t H, int32_t W, float* workspace,
                  void* stream);

/* ---- disparity head: models/build_model_2d.py:27-57 ------------------------------------------------------------ */
/* mat (B, D3, H3, W3) fp32 -> disp (B, 3*H3, 3*W3) fp32; upsample + softmin + regression in one kernel. */
int lea_disp_head(const float* mat, float* disp, int32_t B, int32_t D3, int32_t H3, int32_t W3,
                  int32_t maxdisp, void* stream);
/* DisparityRegression alone (build_model_2d.py:36-41): p (B, maxdisp, H, W) -> (B, H, W). */
int lea_disparity_regression(const float* p, float* out, int32_t B, int32_t maxdisp, int32_t H, int32_t W,
                             void* stream);

/* ---- feature-net producer (retrain/new_model_2d.py:93-94,130-131) ---------------------------- */
/* stem0 (3 -> c_mid, 3x3, stride 1) + BN + ReLU fused with stem1 (c_mid -> c_out, 3x3, stride 3) + BN + ReLU:
 * img (B, 3, H, W) fp32 -> channel slice of a depth-1 planes volume at 1/3 resolution.  w0 (c_mid, 3, 3, 3),
 * w1 (c_out, c_mid, 3, 3) in PyTorch layout; scale/shift = eval-mode BatchNorm folded by the host. */
int lea_feature_stem(const float* img, int32_t B, int32_t H, int32_t W, const float* w0, const float* scale0,
                     const float* shift0, int32_t c_mid, const float* w1, const float* scale1, const float* shift1,
                     int32_t c_out, const lea_vol* dst, int32_t dst_c0, void* stream);

/* ---- training side (what the reference gets from autograd, train.py:156-160) --------------------- */
/* Per-channel reductions over a channel slice, written as `chunks` partial rows partial[(chunk*2+which)*c + ch]
 * (the caller adds the chunks).  mode 0: sum x, sum x^2 (BatchNorm3d batch statistics, operations_3d.py:38,44).
 * mode 1: with g = dy*[relu mask of x*scale+shift], xh = (x-mean)*invstd:  sum g, sum g*xh  (BN backward). */
int lea_channel_reduce(const lea_vol* x, int32_t x_c0, const lea_vol* dy, int32_t dy_c0, int32_t c, int32_t mode,
                       int32_t relu, const float* scale, const float* shift, const float* mean, const float* invstd,
                       float* partial, int32_t chunks, void* stream);
/* BatchNorm3d train-mode glue on the device (operations_3d.py:38,44), one launch each instead of dozens of tiny host
 * tensor ops per ConvBR.  lea_bn_finalize: chunk partials of (sum x, sum x^2) over n voxels -> mean, invstd, scale =
 * gamma*invstd, shift = beta - mean*scale, running statistics updated in place with `momentum` (unbiased variance),
 * *num_batches_tracked += 1 (gamma/beta/running/num_batches_tracked may be NULL).  lea_bn_bwd_coeffs: chunk partials of
 * (sum g, sum g*xh) -> the coefficients ka, kb, kc of lea_bn_relu_bwd and dgamma = sum g*xh, dbeta = sum g
 * (written, or added to what the vectors hold when accumulate != 0). */
int lea_bn_finalize(const float* partial, int32_t chunks, int32_t c, double n, const float* gamma, const float* beta,
                    double eps, double momentum, float* running_mean, float* running_var, int64_t* num_batches_tracked,
                    float* mean, float* invstd, float* scale, float* shift, void* stream);
int lea_bn_bwd_coeffs(const float* partial, int32_t chunks, int32_t c, double n, const float* gamma, const float* invstd,
                      float* ka, float* kb, float* kc, float* dgamma, float* dbeta, int32_t accumulate, void* stream);
/* dst = [dst +] relu?(x*scale[ch] + shift[ch])  - BN apply in train mode and the state sums (skip_model_3d.py:70). */
int lea_affine_relu(const lea_vol* x, int32_t x_c0, const lea_vol* dst, int32_t dst_c0, int32_t c, const float* scale,
                    const float* shift, int32_t relu, int32_t accumulate, void* stream);
/* dx = ka[ch]*g - kb[ch] - xh*kc[ch]  (BN(train)+ReLU backward; ka = gamma*invstd, kb = ka*mean(g), kc = ka*mean(g*xh)). */
int lea_bn_relu_bwd(const lea_vol* x, int32_t x_c0, const lea_vol* dy, int32_t dy_c0, const lea_vol* dx, int32_t dx_c0,
                    int32_t c, int32_t relu, const float* scale, const float* shift, const float* mean,
                    const float* invstd, const float* ka, const float* kb, const float* kc, void* stream);
/* dw[c_out][c_in][k^3] += sum_voxels dout * in(shifted)   (fp32, atomically accumulated: zero dw first). */
int lea_conv3d_wgrad(const lea_vol* in, int32_t in_c0, int32_t c_in, const lea_vol* dout, int32_t dout_c0, int32_t c_out,
                     int32_t ksize, float* dw, void* stream);
/* The same weight gradient on the tensor cores for the shapes that carry the training step (k = 3, two-plane volumes,
 * channel counts tiling by 8, 16 or 32): warp-level mma.sync.m16n8k16 on bf16 planes with the bf16x3 split product
 * (hi*hi + hi*lo + lo*hi, fp32 accumulate), operands straight from the planes layout via TMA + ldmatrix.trans.  The
 * contraction runs over voxels and the output is tiny per tap, which fits 16 x 8 warp tiles and not tcgen05's
 * 128 x N x 16 instruction (lea_wgrad_mma.cu).  dw is accumulated atomically (zero it first).
 * lea_conv3d_wgrad_tc_supported returns 1 for the shapes taken; others go through lea_conv3d_wgrad. */
int lea_conv3d_wgrad_tc_supported(int32_t c_in, int32_t c_out, int32_t ksize, int32_t planes);
int lea_conv3d_wgrad_tc(const lea_vol* in, int32_t in_c0, int32_t c_in, const lea_vol* dout, int32_t dout_c0,
                        int32_t c_out, int32_t ksize, float* dw, void* stream);
/* dsrc (+)= transpose of the align_corners=True trilinear resample applied to ddst (accumulate == 0 overwrites). */
int lea_trilinear_ac_bwd(const lea_vol* ddst, int32_t ddst_c0, const lea_vol* dsrc, int32_t dsrc_c0, int32_t c,
                         int32_t accumulate, void* stream);
/* transpose of the cost-volume construction (LEAStereo.py:42-48): dcost (2C channels) -> dx, dy (B, C, H, W) fp32. */
int lea_cost_volume_bwd(const lea_vol* dcost, int32_t C, float* dx, float* dy, void* stream);
/* backward of lea_disp_head: dmat (B, D3, H3, W3) += J^T gout; dmat must be zero-initialised by the caller. */
int lea_disp_head_bwd(const float* mat, const float* gout, float* dmat, int32_t B, int32_t D3, int32_t H3, int32_t W3,
                      int32_t maxdisp, void* stream);

/* ---- callers on either side of the path -------------------------------------------------- */
/* Input side, predict.py:144-184 / dataloaders/datasets/common.py:119-131: per-channel z-normalisation of an 8-bit RGB
 * image in HWC order (PIL's layout) and test_transform's pad (image bottom-right on zeros) or centre crop to
 * crop_h x crop_w.  Two launches: exact 64-bit sums (sums[0..2] = sum x, sums[3..5] = sum x^2 per channel; zero them
 * first), then out (3, crop_h, crop_w) fp32 = ((double)x - mean) / std with the population std, like np.std.
 * The 8-bit image is what crosses PCIe (4x fewer bytes than the fp32 tensor the reference uploads). */
int lea_image_stats_u8(const uint8_t* img_hwc, int32_t H, int32_t W, uint64_t* sums6, void* stream);
int lea_normalize_pad_u8(const uint8_t* img_hwc, int32_t H, int32_t W, const uint64_t* sums6, float* out,
                         int32_t crop_h, int32_t crop_w, void* stream);
/* Training loss, train.py:116-118,157,162: F.smooth_l1_loss(disp[mask], target[mask], 'mean') with
 * mask = 0.001 < target < maxdisp.  acc3 (zeroed by the caller) += {sum loss, sum |disp - target|, #valid};
 * the backward writes grad = mask * clamp(disp - target, -1, 1) * upstream / #valid reading #valid from acc3 on the
 * device (no host synchronisation between forward and backward). */
int lea_masked_smooth_l1(const float* disp, const float* target, int64_t n, float maxdisp, double* acc3, void* stream);
int lea_masked_smooth_l1_bwd(const float* disp, const float* target, int64_t n, float maxdisp, const double* acc3,
                             float upstream, float* grad, void* stream);
/* optim.Adam(lr, betas=(0.9, 0.999)) of train.py:76 as one launch over a flat fp32 parameter buffer; step >= 1. */
int lea_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                  float beta2, float eps, int32_t step, void* stream);
/* Evaluation metrics in one pass, reference-identical by default: utils/metrics.py:6-46 (validity mask
 * 0.001 < t < maxdisp; the reference stores |true - pred| into an int64 array, so its 3-px and bad-N tests see the
 * error TRUNCATED toward zero, and invalid pixels stand at 10000 in those tests), evaluation.py:290-291 (EPE over the
 * inclusive mask 0.001 <= t <= maxdisp) and train.py:203 (EPE over the strict mask).  acc9 (zeroed by the caller) +=
 * {#valid, sum |d| inclusive mask, #3-px-correct (e < 3 or e < float32(0.05 t)), #(e <= thresholds4[0..3]),
 *  #inclusive-valid, sum |d| strict mask};  3-px error = 1 - acc[2]/acc[0], bad-N = 1 - acc[3+k]/acc[0],
 * EPE(evaluation.py) = acc[1]/acc[7], EPE(train.py) = acc[8]/acc[0].
 * flags: LEA_METRICS_FLOAT_DIFF (1) compares the un-truncated float error instead (NOT the reference's numbers). */
#define LEA_METRICS_FLOAT_DIFF 1
int lea_disparity_metrics(const float* pred, const float* target, int64_t n, float maxdisp, const float* thresholds4,
                          int32_t flags, double* acc9, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LEASTEREO_B200_H_ */
