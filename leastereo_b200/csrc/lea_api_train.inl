// C-ABI entry points of the training-side kernels (included by leastereo_b200.cu and tests/emu/emu_lib.cpp).

extern "C" int lea_channel_reduce(const lea_vol* x, int32_t x_c0, const lea_vol* dy, int32_t dy_c0, int32_t c,
                                  int32_t mode, int32_t relu, const float* scale, const float* shift,
                                  const float* mean, const float* invstd, float* partial, int32_t chunks,
                                  void* stream) {
    if (lea_check_vol(x, "channel_reduce") || lea_check_slice(x, x_c0, c, "channel_reduce")) return 1;
    LEA_CHECK(partial != nullptr && chunks >= 1 && chunks <= 65535, "channel_reduce: bad partial buffer");
    lea_vol dyv = *x;
    if (mode == 1) {
        if (lea_check_vol(dy, "channel_reduce dy") || lea_check_slice(dy, dy_c0, c, "channel_reduce dy")) return 1;
        LEA_CHECK(lea_same_space(x, dy), "channel_reduce: x/dy shapes differ");
        dyv = *dy;
    }
    LEA_LAUNCH(lea_channel_reduce_kernel, dim3(chunks, c >> 3), dim3(256), 0, stream,
               *x, x_c0, dyv, dy_c0, c, mode, relu, scale, shift, mean, invstd, partial);
    return LEA_POST_LAUNCH();
}

extern "C" int lea_bn_finalize(const float* partial, int32_t chunks, int32_t c, double n, const float* gamma,
                               const float* beta, double eps, double momentum, float* running_mean, float* running_var,
                               int64_t* num_batches_tracked, float* mean, float* invstd, float* scale, float* shift,
                               void* stream) {
    LEA_CHECK(partial && mean && invstd && scale && shift && chunks >= 1 && c >= 1 && n >= 1.0, "bn_finalize: bad argument");
    LEA_CHECK((running_mean == nullptr) == (running_var == nullptr), "bn_finalize: running statistics must come together");
    LEA_LAUNCH(lea_bn_finalize_kernel, dim3((c + 7) / 8), dim3(256), 0, stream, partial, chunks, c, n, gamma, beta, eps,
               momentum, running_mean, running_var, (long long*)num_batches_tracked, mean, invstd, scale, shift);
    return LEA_POST_LAUNCH();
}

extern "C" int lea_bn_bwd_coeffs(const float* partial, int32_t chunks, int32_t c, double n, const float* gamma,
                                 const float* invstd, float* ka, float* kb, float* kc, float* dgamma, float* dbeta,
                                 int32_t accumulate, void* stream) {
    LEA_CHECK(partial && invstd && ka && kb && kc && chunks >= 1 && c >= 1 && n >= 1.0, "bn_bwd_coeffs: bad argument");
    LEA_LAUNCH(lea_bn_bwd_coeffs_kernel, dim3((c + 7) / 8), dim3(256), 0, stream, partial, chunks, c, n, gamma, invstd,
               ka, kb, kc, dgamma, dbeta, accumulate);
    return LEA_POST_LAUNCH();
}

extern "C" int lea_affine_relu(const lea_vol* x, int32_t x_c0, const lea_vol* dst, int32_t dst_c0, int32_t c,
                               const float* scale, const float* shift, int32_t relu, int32_t accumulate, void* stream) {
    if (lea_check_vol(x, "affine_relu") || lea_check_slice(x, x_c0, c, "affine_relu")) return 1;
    if (lea_check_vol(dst, "affine_relu dst") || lea_check_slice(dst, dst_c0, c, "affine_relu dst")) return 1;
    LEA_CHECK(lea_same_space(x, dst), "affine_relu: shapes differ");
    LEA_CHECK((scale == nullptr) == (shift == nullptr), "affine_relu: scale/shift must come together");
    LEA_CHECK((int64_t)x->B * (c >> 3) <= 65535, "affine_relu: grid too large");
    LEA_LAUNCH(lea_affine_relu_kernel, dim3((x->W + 255) / 256, x->D * x->H, x->B * (c >> 3)), dim3(256), 0, stream,
               *x, x_c0, *dst, dst_c0, c, scale, shift, relu, accumulate);
    return LEA_POST_LAUNCH();
}

extern "C" int lea_bn_relu_bwd(const lea_vol* x, int32_t x_c0, const lea_vol* dy, int32_t dy_c0, const lea_vol* dx,
                               int32_t dx_c0, int32_t c, int32_t relu, const float* scale, const float* shift,
                               const float* mean, const float* invstd, const float* ka, const float* kb,
                               const float* kc, void* stream) {
    if (lea_check_vol(x, "bn_relu_bwd") || lea_check_slice(x, x_c0, c, "bn_relu_bwd")) return 1;
    if (lea_check_vol(dy, "bn_relu_bwd dy") || lea_check_slice(dy, dy_c0, c, "bn_relu_bwd dy")) return 1;
    if (lea_check_vol(dx, "bn_relu_bwd dx") || lea_check_slice(dx, dx_c0, c, "bn_relu_bwd dx")) return 1;
    LEA_CHECK(lea_same_space(x, dy) && lea_same_space(x, dx), "bn_relu_bwd: shapes differ");
    LEA_CHECK(scale && shift && mean && invstd && ka && kb && kc, "bn_relu_bwd: null parameter vector");
    LEA_CHECK((int64_t)x->B * (c >> 3) <= 65535, "bn_relu_bwd: grid too large");
    LEA_LAUNCH(lea_bn_relu_bwd_kernel, dim3((x->W + 255) / 256, x->D * x->H, x->B * (c >> 3)), dim3(256), 0, stream,
               *x, x_c0, *dy, dy_c0, *dx, dx_c0, c, relu, scale, shift, mean, invstd, ka, kb, kc);
    return LEA_POST_LAUNCH();
}

extern "C" int lea_conv3d_wgrad(const lea_vol* in, int32_t in_c0, int32_t c_in, const lea_vol* dout, int32_t dout_c0,
                                int32_t c_out, int32_t ksize, float* dw, void* stream) {
    if (lea_check_vol(in, "conv3d_wgrad") || lea_check_slice(in, in_c0, c_in, "conv3d_wgrad")) return 1;
    if (lea_check_vol(dout, "conv3d_wgrad dout")) return 1;
    LEA_CHECK(c_out >= 1 && c_out <= 64 && dout_c0 % 8 == 0 && dout_c0 + ((c_out + 7) & ~7) <= dout->C,
              "conv3d_wgrad: bad output-gradient slice");
    LEA_CHECK(lea_same_space(in, dout), "conv3d_wgrad: shapes differ");
    LEA_CHECK(ksize == 1 || ksize == 3, "conv3d_wgrad: kernel size %d not supported", ksize);
    LEA_CHECK(dw != nullptr && in->B <= 65535, "conv3d_wgrad: bad argument");
    const int tiles = ((in->W + LEA_TW - 1) / LEA_TW) * ((in->H + LEA_TH - 1) / LEA_TH);
    const int cpad = (c_out + 7) & ~7;
    const dim3 grid(tiles, c_in >> 3, in->B);
    if (ksize == 3) {
        const size_t smem = (size_t)(8 * 3 * (LEA_TH + 2) * (LEA_TW + 2) + cpad * 128) * sizeof(float);
#ifndef LEA_CPU_EMU
        if (smem > 48 * 1024)
            cudaFuncSetAttribute(lea_conv_wgrad_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
#endif
        LEA_LAUNCH(lea_conv_wgrad_kernel<3>, grid, dim3(128), smem, stream, *in, in_c0, c_in, *dout, dout_c0, c_out, dw);
    } else if ((c_in >> 2) * (cpad >> 2) <= 512 && (size_t)LEA_W1_VOX * (c_in + cpad) * sizeof(float) <= 96 * 1024) {
        // skinny-GEMM kernel: (c_out/4) x (c_in/4) register blocks fit 256 threads x 2
        const size_t smem = (size_t)LEA_W1_VOX * (c_in + cpad) * sizeof(float);
        const int64_t sp = (int64_t)in->D * in->H * in->W;
        const int chunks = (int)((sp + (int64_t)LEA_W1_VOX * LEA_W1_TILES - 1) / ((int64_t)LEA_W1_VOX * LEA_W1_TILES));
#ifndef LEA_CPU_EMU
        if (smem > 48 * 1024)
            cudaFuncSetAttribute(lea_conv1_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
#endif
        LEA_LAUNCH(lea_conv1_wgrad_kernel, dim3(chunks, in->B), dim3(256), smem, stream, *in, in_c0, c_in, *dout, dout_c0,
                   c_out, cpad, dw);
    } else {
        const size_t smem = (size_t)(8 * LEA_TH * LEA_TW + cpad * 128) * sizeof(float);
        LEA_LAUNCH(lea_conv_wgrad_kernel<1>, grid, dim3(128), smem, stream, *in, in_c0, c_in, *dout, dout_c0, c_out, dw);
    }
    return LEA_POST_LAUNCH();
}

extern "C" int lea_trilinear_ac_bwd(const lea_vol* ddst, int32_t ddst_c0, const lea_vol* dsrc, int32_t dsrc_c0,
                                    int32_t c, int32_t accumulate, void* stream) {
    if (lea_check_vol(ddst, "trilinear_ac_bwd ddst") || lea_check_vol(dsrc, "trilinear_ac_bwd dsrc")) return 1;
    if (lea_check_slice(ddst, ddst_c0, c, "trilinear_ac_bwd ddst") || lea_check_slice(dsrc, dsrc_c0, c, "trilinear_ac_bwd dsrc"))
        return 1;
    LEA_CHECK(ddst->B == dsrc->B && (int64_t)dsrc->B * (c >> 3) <= 65535, "trilinear_ac_bwd: bad batch");
    LEA_LAUNCH(lea_trilinear_ac_bwd_kernel, dim3((dsrc->W + 127) / 128, dsrc->D * dsrc->H, dsrc->B * (c >> 3)), dim3(128),
               0, stream, *ddst, ddst_c0, *dsrc, dsrc_c0, c, accumulate);
    return LEA_POST_LAUNCH();
}

extern "C" int lea_cost_volume_bwd(const lea_vol* dcost, int32_t C, float* dx, float* dy, void* stream) {
    if (lea_check_vol(dcost, "cost_volume_bwd")) return 1;
    LEA_CHECK(dx && dy && C > 0 && (C & 7) == 0 && dcost->C == 2 * C, "cost_volume_bwd: bad arguments");
    LEA_CHECK(dcost->H <= 65535 && (int64_t)dcost->B * (dcost->C >> 3) <= 65535, "cost_volume_bwd: grid too large");
    LEA_LAUNCH(lea_cost_volume_bwd_kernel, dim3((dcost->W + 255) / 256, dcost->H, dcost->B * (dcost->C >> 3)), dim3(256),
               0, stream, *dcost, C, dx, dy);
    return LEA_POST_LAUNCH();
}

extern "C" int lea_disp_head_bwd(const float* mat, const float* gout, float* dmat, int32_t B, int32_t D3, int32_t H3,
                                 int32_t W3, int32_t maxdisp, void* stream) {
    LEA_CHECK(mat && gout && dmat, "disp_head_bwd: null pointer");
    LEA_CHECK(B > 0 && D3 > 0 && H3 > 0 && W3 > 0 && maxdisp > 0 && B <= 65535, "disp_head_bwd: bad shape");
    LEA_LAUNCH(lea_disp_head_bwd_kernel, dim3((W3 + LEA_DH_BX - 1) / LEA_DH_BX, (H3 + LEA_DH_BY - 1) / LEA_DH_BY, B),
               dim3(LEA_DH_BX * LEA_DH_BY), 0, stream, mat, gout, dmat, D3, H3, W3, maxdisp);
    return LEA_POST_LAUNCH();
}
