// C-ABI entry points of the I/O-side kernels (lea_io_kernels.cuh); included by leastereo_b200.cu and the emulator.

static int lea_io_grid(int64_t n) {
    int64_t g = (n + LEA_IO_THREADS - 1) / LEA_IO_THREADS;
    if (g > 148 * 8) g = 148 * 8;                 // a few CTAs per SM, grid-stride inside
    return g < 1 ? 1 : (int)g;
}

extern "C" int lea_image_stats_u8(const uint8_t* img_hwc, int32_t H, int32_t W, uint64_t* sums, void* stream) {
    LEA_CHECK(img_hwc && sums && H > 0 && W > 0, "image_stats_u8: bad argument");
    LEA_LAUNCH(lea_image_stats_u8_kernel, dim3(lea_io_grid((int64_t)H * W)), dim3(LEA_IO_THREADS), 0, stream,
               img_hwc, (int64_t)H * W, reinterpret_cast<unsigned long long*>(sums));
    return LEA_POST_LAUNCH();
}

extern "C" int lea_normalize_pad_u8(const uint8_t* img_hwc, int32_t H, int32_t W, const uint64_t* sums, float* out,
                                    int32_t crop_h, int32_t crop_w, void* stream) {
    LEA_CHECK(img_hwc && sums && out && H > 0 && W > 0 && crop_h > 0 && crop_w > 0 && crop_h <= 65535,
              "normalize_pad_u8: bad argument");
    LEA_LAUNCH(lea_normalize_pad_u8_kernel, dim3((crop_w + LEA_IO_THREADS - 1) / LEA_IO_THREADS, crop_h),
               dim3(LEA_IO_THREADS), 0, stream, img_hwc, H, W, reinterpret_cast<const unsigned long long*>(sums), out,
               crop_h, crop_w);
    return LEA_POST_LAUNCH();
}

extern "C" int lea_masked_smooth_l1(const float* disp, const float* target, int64_t n, float maxdisp, double* acc3,
                                    void* stream) {
    LEA_CHECK(disp && target && acc3 && n > 0, "masked_smooth_l1: bad argument");
    LEA_LAUNCH(lea_masked_smooth_l1_kernel, dim3(lea_io_grid(n)), dim3(LEA_IO_THREADS), 0, stream, disp, target, n,
               maxdisp, acc3);
    return LEA_POST_LAUNCH();
}

extern "C" int lea_masked_smooth_l1_bwd(const float* disp, const float* target, int64_t n, float maxdisp,
                                        const double* acc3, float upstream, float* grad, void* stream) {
    LEA_CHECK(disp && target && acc3 && grad && n > 0, "masked_smooth_l1_bwd: bad argument");
    LEA_LAUNCH(lea_masked_smooth_l1_bwd_kernel, dim3(lea_io_grid(n)), dim3(LEA_IO_THREADS), 0, stream, disp, target, n,
               maxdisp, acc3, upstream, grad);
    return LEA_POST_LAUNCH();
}

extern "C" int lea_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                             float beta1, float beta2, float eps, int32_t step, void* stream) {
    LEA_CHECK(param && grad && exp_avg && exp_avg_sq && n > 0 && step >= 1, "adam_step: bad argument");
    const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    LEA_LAUNCH(lea_adam_step_kernel, dim3(lea_io_grid(n)), dim3(LEA_IO_THREADS), 0, stream, param, grad, exp_avg,
               exp_avg_sq, n, beta1, beta2, eps, (float)((double)lr / bc1), (float)(1.0 / sqrt(bc2)));
    return LEA_POST_LAUNCH();
}

extern "C" int lea_disparity_metrics(const float* pred, const float* target, int64_t n, float maxdisp,
                                     const float* thresholds4, int32_t flags, double* acc9, void* stream) {
    LEA_CHECK(pred && target && thresholds4 && acc9 && n > 0, "disparity_metrics: bad argument");
    LEA_CHECK((flags & ~LEA_METRICS_FLOAT_DIFF) == 0, "disparity_metrics: unknown flag");
    LEA_LAUNCH(lea_disparity_metrics_kernel, dim3(lea_io_grid(n)), dim3(LEA_IO_THREADS), 0, stream, pred, target, n,
               maxdisp, thresholds4[0], thresholds4[1], thresholds4[2], thresholds4[3], (int)flags, acc9);
    return LEA_POST_LAUNCH();
}
